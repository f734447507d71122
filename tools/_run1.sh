python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
for lc in 32 64 0; do echo "LC=$lc"; if [ $lc = 0 ]; then unset MXL_EQ_STREAM_CHUNK; else export MXL_EQ_STREAM_CHUNK=$lc; fi; python tools/kernel_roofline.py --only EqThree | cut -c1-200; python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --workload audio | python -c "import sys,json; r=json.loads(sys.stdin.readline()); print({k:round(v['ms'],4) for k,v in r['stages'].items()}, round(r['ms_per_step'],4))"; done
unset MXL_EQ_STREAM_CHUNK
python tools/kernel_roofline.py --only Meter | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:eq_stream -s 3 -c 1 -o gpurun_out/prof_eqstream_big_r1b python tools/kernel_roofline.py --only EqThree --reps 2 > gpurun_out/ncu_eqstream_big.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:eq_stream -s 6 -c 1 -o gpurun_out/prof_eqstream_small_r1b python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --workload audio > gpurun_out/ncu_eqstream_small.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:meter -s 3 -c 1 -o gpurun_out/prof_meter_r1 python tools/kernel_roofline.py --only Meter --reps 2 > gpurun_out/ncu_meter.log 2>&1
