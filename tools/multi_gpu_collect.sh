#!/bin/bash
# usage (under gpurun --gpus N): tools/multi_gpu_collect.sh N   -- host-fed ceilings, shared-source test, bench at N GPUs
N=${1:-2}
mkdir -p gpurun_out
echo "== topology"; nvidia-smi topo -m 2>/dev/null | head -14; ls /sys/devices/system/node/ | grep node; nproc
for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ] && [ "$(cat $d/class 2>/dev/null | cut -c1-4)" = "0x03" ]; then echo "$(basename $d) numa_node=$(cat $d/numa_node)"; fi; done
echo "== pcie_multi"
: > gpurun_out/r2_pcie_multi.jsonl
for n in 1 2 4 8; do
  if [ $n -le $N ]; then
    for mode in "" "affinity" "affinity streams=2"; do tools/pcie_multi.bin $n $mode | tee -a gpurun_out/r2_pcie_multi.jsonl | cut -c1-330; done
  fi
done
echo "== shared source test"
python -m pytest tests/test_shared_source.py -m gpu -q -s 2>&1 | tail -5
echo "== bench N=$N"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; tail -c 400 gpurun_out/r2_bench_n$N.err
MXL_NO_NUMA_BIND=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_bench_n${N}_nobind.json 2>> gpurun_out/r2_bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_bench_n${N}_ref.json 2>> gpurun_out/r2_bench_n$N.err
python - <<PY
import json
for f in ("gpurun_out/r2_bench_n$N.json", "gpurun_out/r2_bench_n${N}_nobind.json", "gpurun_out/r2_bench_n${N}_ref.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "e2e_session", round(d["e2e_session"]["value"]) if d.get("e2e_session") else None,
              "h2d_gbs", d["e2e"].get("h2d_gbs"), "shared", d.get("shared_source"), "numa", d.get("config", {}).get("host_numa"))
    except Exception as e:
        print(f, "unreadable", e)
PY
