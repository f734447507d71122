python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -12 gpurun_out/pytest_gpu.log
python bench.py --no-cpu-baseline --steps 50 > gpurun_out/bench_g.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_g.json').readline())
print(round(d['value']), d['roofline'])
print(d['kernels'])
print(d['e2e'])
"
