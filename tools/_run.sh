# one gpurun call: GPU tests, randomised parity run, both bench arms, smoke, ncu launch list of the bench command
timeout 300 python -m pytest tests -m gpu -q --timeout 120 2>&1 | tail -2
timeout 150 python tools/fuzz_parity.py --cases 400 --seconds 100 > gpurun_out/fuzz.log 2>&1; echo "fuzz rc $?"; tail -1 gpurun_out/fuzz.log
python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "ours rc $?"
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc $?"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fit_|pose_" -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/ncu.log 2>&1; echo "ncu rc $?"
