timeout 300 python -m pytest tests -m gpu -q --timeout 120 2>&1 | tail -2
python bench.py > gpurun_out/s26_ours.json 2> gpurun_out/s26_ours.err; echo "ours rc $?"
python bench.py --impl reference > gpurun_out/s26_ref.json 2> gpurun_out/s26_ref.err; echo "ref rc $?"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fit_|pose_" -c 400 --csv --log-file gpurun_out/s26_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/s26_ncu.log 2>&1; echo "ncu rc $?"
