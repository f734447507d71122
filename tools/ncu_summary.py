#!/usr/bin/env python
"""Condenses .ncu-rep captures (ncu --set full --import-source on) into the small JSON summaries kept under
profiles/: per kernel launch the duration, DRAM traffic, pipe utilisation, occupancy limits, and the top
stall reasons from the source page.  Runs here (no GPU needed): ncu -i <rep> --page raw/source --csv."""
import csv
import io
import json
import subprocess
import sys

RAW_KEYS = {
    "gpu__time_duration.sum": "duration_ns",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "dram__bytes.sum.per_second": "dram_bytes_per_second",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct_of_peak",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed": "fp64_pipe_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed": "xu_pipe_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pipe_pct_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pipe_pct_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid_size",
    "launch__block_size": "block_size",
    "launch__occupancy_limit_registers": "occ_limit_regs",
    "launch__occupancy_limit_shared_mem": "occ_limit_smem",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
}


def run(args):
    return subprocess.run(["ncu"] + args, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def summarize(path):
    rows = list(csv.reader(io.StringIO(run(["-i", path, "--page", "raw", "--csv"]))))
    if len(rows) < 3:
        return []
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")].split("(")[0].replace("unnamed>::", "")}
        for h, u, v in zip(hdr, units, r):
            if h in RAW_KEYS:
                try:
                    x = float(v)
                except ValueError:
                    continue
                if h.startswith("dram__bytes"):
                    x *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1}.get(u, 1)
                if h == "gpu__time_duration.sum":
                    x *= {"us": 1e3, "ms": 1e6, "ns": 1, "second": 1e9}.get(u, 1)
                d[RAW_KEYS[h]] = x
        if "dram_read" in d and "dram_write" in d:
            d["dram_bytes"] = d["dram_read"] + d["dram_write"]
        out.append(d)
    # stall reasons over the whole kernel (source page is per launch in capture order; aggregate = first launch)
    src = list(csv.reader(io.StringIO(run(["-i", path, "--page", "source", "--csv"]))))
    if len(src) > 2:
        hdr = src[1]
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        agg = {h: 0 for h in stalls}
        for r in src[2:]:
            if len(r) < len(hdr):
                continue
            for h in stalls:
                try:
                    agg[h] += int(r[hdr.index(h)] or 0)
                except ValueError:
                    pass
        tot = sum(agg.values()) or 1
        top = sorted(agg.items(), key=lambda kv: -kv[1])[:5]
        if out:
            out[0]["top_stalls_pct"] = {k: round(100.0 * v / tot, 1) for k, v in top}
    return out


if __name__ == "__main__":
    res = {}
    for p in sys.argv[1:]:
        res[p.split("/")[-1]] = summarize(p)
    print(json.dumps(res, indent=1))
