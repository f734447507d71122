timeout 600 python -m pytest tests -m gpu -x -q -k "fused or graph or session or full_chain or concurrent" > gpurun_out/r2_pytest15.log 2>&1; tail -n 3 gpurun_out/r2_pytest15.log
python bench.py --workload audio --steps 50 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json
r=json.loads(sys.stdin.readline()); print('audio T128 ticks/s', r['value'], 'ms/step', r['ms_per_step'])"
python bench.py --workload audio --ticks-per-step 1024 --steps 50 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json
r=json.loads(sys.stdin.readline()); print('audio T1024 ticks/s', r['value'], 'ms/step', r['ms_per_step'])"
python bench.py --workload audio --ticks-per-step 1 --steps 2000 --warmup 50 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json
r=json.loads(sys.stdin.readline()); print('audio T1 ticks/s', r['value'], 'ms/step', r['ms_per_step'])"
