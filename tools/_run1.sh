timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest14.log 2>&1; tail -n 3 gpurun_out/r2_pytest14.log; grep "differential run" gpurun_out/r2_pytest14.log
timeout 300 python -m pytest tests -m gpu -q -s -k "long_differential" 2>&1 | grep -E "differential|passed|failed" | tee gpurun_out/r2_eq_differential.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fused -s 6 -c 2 -o gpurun_out/r2_prof_fused python bench.py --workload audio --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/ncu_fused.log 2>&1
tail -n 2 gpurun_out/ncu_fused.log | cut -c1-300
