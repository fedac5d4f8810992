timeout 300 python -m pytest tests -m gpu -q --timeout 120 2>&1 | tail -2
timeout 150 python tools/fuzz_parity.py --cases 400 --seconds 100 --seed 11 > gpurun_out/s31_fuzz.log 2>&1; echo "fuzz rc $?"; tail -1 gpurun_out/s31_fuzz.log
python bench.py > gpurun_out/s31_ours.json 2> gpurun_out/s31_ours.err; echo "ours rc $?"
python bench.py --impl reference > gpurun_out/s31_ref.json 2> gpurun_out/s31_ref.err; echo "ref rc $?"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fit_|pose_" -c 400 --csv --log-file gpurun_out/s31_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/s31_ncu.log 2>&1; echo "ncu rc $?"
