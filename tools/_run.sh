python bench.py > gpurun_out/s17_ours.json 2> gpurun_out/s17_ours.err; echo "ours rc $?"
python bench.py --impl reference > gpurun_out/s17_ref.json 2> gpurun_out/s17_ref.err; echo "ref rc $?"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fit_|pose_" -c 400 --csv --log-file gpurun_out/s17_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/s17_ncu.log 2>&1; echo "ncu rc $?"
python tools/prof_case.py c3 1 > /dev/null && ncu --set full --import-source on --clock-control none -k regex:fit_ransac -c 1 -o gpurun_out/s17_ransac_c3 python tools/prof_case.py c3 1 > gpurun_out/s17_r.log 2>&1; echo "ncu2 rc $?"
