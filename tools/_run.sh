timeout 300 python -m pytest tests -m gpu -q --timeout 120 2>&1 | tail -2
python bench.py > gpurun_out/s23_ours.json 2> gpurun_out/s23_ours.err; echo "ours rc $?"
python bench.py --impl reference > gpurun_out/s23_ref.json 2> gpurun_out/s23_ref.err; echo "ref rc $?"
python -c "import __graft_entry__ as g; g.smoke()"
