python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
python tools/kernel_roofline.py > gpurun_out/kernel_roofline_r1d.jsonl 2>gpurun_out/kr.err; cut -c1-175 gpurun_out/kernel_roofline_r1d.jsonl
ncu --set full --clock-control none --import-source on -k regex:envelope -s 3 -c 1 -o gpurun_out/prof_envelope_r1c python tools/kernel_roofline.py --only Envelope --reps 2 > gpurun_out/ncu_env.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:oscillator -s 3 -c 1 -o gpurun_out/prof_osc_r1c python tools/kernel_roofline.py --only "Oscillator(sine)" --reps 2 > gpurun_out/ncu_osc.log 2>&1
