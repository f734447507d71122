python -m pytest tests/test_parity_audio.py -m gpu -x -q -k "oscillator or fm_sine" 2>&1 | tail -2
python tools/kernel_roofline.py --only Osc 2>/dev/null | cut -c1-170
python tools/kernel_roofline.py --only FmSine 2>/dev/null | cut -c1-170
python tools/kernel_roofline.py --only Trigger 2>/dev/null | cut -c1-170
ncu --set full --clock-control none --import-source on -k regex:oscillator -s 7 -c 1 -o gpurun_out/prof_osc_saw python tools/kernel_roofline.py --only "Oscillator" --reps 1 > gpurun_out/ncu_osc.log 2>&1
