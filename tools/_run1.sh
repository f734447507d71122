python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest9.log 2>&1; tail -3 gpurun_out/r2_pytest9.log
for mb in 8 12 16; do
  a=$(MXL_ENV_MINB=$mb python tools/kernel_roofline.py --only Envelope | python -c "import sys,json; r=json.loads(sys.stdin.readline()); print(round(r['ms'],4))")
  echo "pt=16 minb=$mb ms=$a"
done | tee gpurun_out/r2_env_tune.txt
python tools/kernel_roofline.py --only NOAUDIO 2>/dev/null | grep scale_tiled | tee gpurun_out/r2_scale_tune.txt
ncu --set full --clock-control none --import-source on -k regex:envelope -s 3 -c 1 -o gpurun_out/r2_prof_env python tools/kernel_roofline.py --only Envelope --reps 1 > gpurun_out/ncu_env.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scale_tiled -c 2 -o gpurun_out/r2_prof_scale python tools/kernel_roofline.py --only NOAUDIO --reps 1 > gpurun_out/ncu_scale.log 2>&1
