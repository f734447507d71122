MXL_EQ_PERSIST=1 timeout 600 python -m pytest tests -m gpu -x -q -s -k "eq_three" 2>&1 | grep -E "differential|passed|failed|Error" | tail -n 8
timeout 300 python -m pytest tests -m gpu -x -q -k "eq_three or fused or graph" 2>&1 | tail -n 2
for pe in 0 1; do for lc in 32 64; do
  a=$(MXL_EQ_PERSIST=$pe MXL_EQ_STREAM_CHUNK=$lc timeout 120 python tools/kernel_roofline.py --only EqThree | python -c "import sys,json; r=json.loads(sys.stdin.readline()); print(round(r['ms'],4))")
  echo "persist=$pe lc=$lc ms=$a"
done; done | tee gpurun_out/r2_eq_persist.txt
MXL_EQ_PERSIST=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:eq_stream -s 3 -c 1 -o gpurun_out/r2_prof_eq4 python tools/kernel_roofline.py --only EqThree --reps 1 > gpurun_out/ncu_eq.log 2>&1
