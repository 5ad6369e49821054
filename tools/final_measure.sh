cd $GRAFT_REPO_ROOT
timeout 150 python -m pytest tests -m gpu -x -q > gpurun_out/r02f_gputests2.log 2>&1; echo "pytest rc=$?" 
timeout 200 python bench.py > gpurun_out/r02f_b1.json 2> gpurun_out/r02f_b1.err; echo "bench rc=$?"
timeout 90 python bench.py --steps 5 --warmup 3 --no-e2e --no-extra --cpu-budget-s 0.3 > gpurun_out/r02_v10_plain.log 2>&1; echo "plain rc=$?"
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_v10_launches.csv python bench.py --steps 5 --warmup 3 --no-e2e --no-extra --cpu-budget-s 0.3 > gpurun_out/r02_v10_ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k "regex:rk45_init_kernel|rk45_attempt_kernel|head_kernel" --launch-skip 21 --launch-count 3 -f -o gpurun_out/r02_v10_step python bench.py --steps 5 --warmup 3 --no-e2e --no-extra --cpu-budget-s 0.3 > gpurun_out/r02_v10_ncu2.log 2>&1; echo "ncu2 rc=$?"
tail -2 gpurun_out/r02f_gputests2.log
