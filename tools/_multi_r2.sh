# under gpurun --gpus 8: final bench lines at N = 8, 4, 2 (+ reference arm at 8), shared-source test
mkdir -p gpurun_out
python -m pytest tests/test_shared_source.py -m gpu -q -s 2>&1 | tail -n 4
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $1 --steps 20 --warmup 3 $3 > gpurun_out/r2f_bench_n$1$4.json 2> gpurun_out/r2f_bench_n$1$4.err; tail -c 300 gpurun_out/r2f_bench_n$1$4.err; }
run 8 29511 "" ""
run 8 29512 "--impl reference" "_ref"
run 4 29513 "" ""
run 2 29514 "" ""
python - <<PY
import json
for f in ("gpurun_out/r2f_bench_n8.json", "gpurun_out/r2f_bench_n8_ref.json", "gpurun_out/r2f_bench_n4.json", "gpurun_out/r2f_bench_n2.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "e2e_session", round(d["e2e_session"]["value"]) if d.get("e2e_session") else None,
              "h2d_gbs", d["e2e"].get("h2d_gbs"), "shared", d.get("shared_source"))
    except Exception as e:
        print(f, "unreadable", e)
PY
