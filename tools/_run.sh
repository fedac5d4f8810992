timeout 300 python -m pytest tests -m gpu -q --timeout 120 2>&1 | tail -2
python bench.py > gpurun_out/s28_ours.json 2> gpurun_out/s28_ours.err; echo "ours rc $?"
python bench.py --impl reference > gpurun_out/s28_ref.json 2> gpurun_out/s28_ref.err; echo "ref rc $?"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
