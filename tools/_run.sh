python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -12 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --no-cpu-baseline --steps 50 > gpurun_out/bench_h.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_h.json').readline())
print(round(d['value']), {k:(round(v,4) if isinstance(v,float) else v) for k,v in d['roofline'].items() if k not in ('timing','peak_source')})
print({k:round(v['avg_launch_ms'],4) for k,v in d['kernels'].items()})
"
