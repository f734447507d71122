(time python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err); tail -n 5 gpurun_out/r2f_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2f_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'sess',d['e2e_session']['value'])
for p in d['sub']['config5_mixer_bus']['points']: print(p)
PY
