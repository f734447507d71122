# EqThree chunk-length A/B on the GPU box: the kernel alone on a 2^25-sample line and inside the audio bench step
for lc in 16 32 64; do
  a=$(MXL_EQ_STREAM_CHUNK=$lc python tools/kernel_roofline.py --only EqThree | python -c "import sys,json; r=json.loads(sys.stdin.readline()); print(round(r['ms'],3))")
  b=$(MXL_EQ_STREAM_CHUNK=$lc python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --workload audio | python -c "import sys,json; r=json.loads(sys.stdin.readline()); print(round(r['stages']['EqThree']['ms'],4), round(r['ms_per_step'],4))")
  echo "lc=$lc big_ms=$a audio128: $b"
done
