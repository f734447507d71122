python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -12 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for i in 1 2; do python bench.py --no-cpu-baseline --steps 50 > gpurun_out/bench_i$i.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_i$i.json').readline())
print(round(d['value']), round(d['roofline']['frac'],3), {k:(round(v,2) if isinstance(v,float) else v) for k,v in d['e2e'].items() if k in ('value','h2d_gbs','d2h_gbs','ms_per_step')})
"; done
python bench.py --no-cpu-baseline --steps 30 --e2e-mode serial > gpurun_out/bench_serial.json 2>> gpurun_out/bench.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_serial.json').readline())
print('serial e2e', {k:(round(v,2) if isinstance(v,float) else v) for k,v in d['e2e'].items() if k in ('value','h2d_gbs','d2h_gbs','ms_per_step')})
"
