timeout 600 python -m pytest tests -m gpu -x -q -k "fused or graph or session or full_chain or concurrent" > gpurun_out/r2_pytest16.log 2>&1; tail -n 3 gpurun_out/r2_pytest16.log
b() { python bench.py --workload audio --ticks-per-step $1 --steps $2 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json
r=json.loads(sys.stdin.readline()); print('ticks/s', round(r['value']), 'ms/step', round(r['ms_per_step'],5))"; }
echo "chain default T128"; b 128 50
echo "chain off T128"; MXL_FUSED_CHAIN=0 b 128 50
echo "chain lc32 T128"; MXL_FUSED_CHUNK=32 MXL_FUSED_CHAIN=1 b 128 50
echo "chain lc16 u256 T128"; MXL_FUSED_CHUNK=16 MXL_FUSED_CHAIN=1 MXL_FUSED_OWNED=256 b 128 50
echo "chain lc16 u192 T128"; MXL_FUSED_CHUNK=16 MXL_FUSED_CHAIN=1 MXL_FUSED_OWNED=192 b 128 50
echo "chain default T1024"; b 1024 30
echo "chain off T1024"; MXL_FUSED_CHAIN=0 b 1024 30
echo "chain lc32 T1024"; MXL_FUSED_CHUNK=32 MXL_FUSED_CHAIN=1 b 1024 30
echo "chain lc64 T1024"; MXL_FUSED_CHUNK=64 MXL_FUSED_CHAIN=1 b 1024 30
echo "T1"; b 1 2000
