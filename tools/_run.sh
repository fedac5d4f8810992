timeout 300 python -m pytest tests -m gpu -q --timeout 120 2>&1 | tail -2
timeout 100 python tools/fuzz_parity.py --cases 300 --seconds 50 --seed 13 > gpurun_out/s34_fuzz.log 2>&1; echo "fuzz rc $?"; tail -1 gpurun_out/s34_fuzz.log
python bench.py > gpurun_out/s34_ours.json 2> gpurun_out/s34_ours.err; echo "ours rc $?"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fit_|pose_" -c 400 --csv --log-file gpurun_out/s34_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/s34_ncu.log 2>&1; echo "ncu rc $?"
