#!/usr/bin/env python
"""Opcode histogram of every kernel in libmixlab_b200.so (cuobjdump -sass): what DESIGN.md says about the instruction mix
-- no tensor-core path (no UTC*MMA / HMMA / LDTM), cp.async staging (LDGSTS), FP64 arithmetic (DADD/DMUL/DFMA), integer
dot products in the scaler (IDP) -- as a file.  Runs anywhere the library has been built (no GPU needed)."""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mixlab_b200", "libmixlab_b200.so")
WATCH = ["UTMALDG", "UTMASTG", "UTCHMMA", "UTCIMMA", "UTCQMMA", "HMMA", "IMMA", "DMMA", "LDTM", "STTM", "LDGSTS", "LDSM",
         "DFMA", "DADD", "DMUL", "F2F", "I2F", "F2I", "FRND", "IDP", "PRMT", "SHFL", "BAR", "UCGABAR_ARV", "ATOM", "ATOMG", "RED",
         "LDG", "STG", "LDS", "STS", "LDC", "LDCU", "MUFU"]


def demangle(names):
    try:
        out = subprocess.run(["c++filt"], input="\n".join(names), stdout=subprocess.PIPE, text=True, check=True).stdout.splitlines()
        return dict(zip(names, out))
    except Exception:
        return {n: n for n in names}


def main():
    text = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    kernels, arch, cur = collections.OrderedDict(), set(), None
    for line in text.splitlines():
        m = re.match(r"\s*arch = (\S+)", line)
        if m:
            arch.add(m.group(1))
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            kernels[cur][m.group(1)] += 1
    names = demangle(list(kernels))
    out = {"library": os.path.relpath(LIB, ROOT), "arch": sorted(arch), "kernels": {}, "totals": collections.Counter()}
    for k, c in kernels.items():
        full = names[k].replace("(anonymous namespace)::", "").replace("mxl::k::", "").replace("mxl::", "")
        short = re.sub(r"^void ", "", full)
        short = re.sub(r"\((?:[^()]|\([^()]*\))*\)\s*$", "", short)             # drop the argument list, keep template arguments
        entry = {"instructions": sum(c.values()), "watch": {w: c[w] for w in WATCH if c[w]},
                 "top": dict(c.most_common(12))}
        out["kernels"][short] = entry
        out["totals"].update(c)
    t = out["totals"]
    out["summary"] = {"kernels": len(kernels), "instructions": sum(t.values()),
                      "tensor_core_opcodes": {w: t[w] for w in ("UTCHMMA", "UTCIMMA", "UTCQMMA", "HMMA", "IMMA", "DMMA", "LDTM", "STTM")},
                      "tma_opcodes": {w: t[w] for w in ("UTMALDG", "UTMASTG")},
                      "cp_async_LDGSTS": t["LDGSTS"], "fp64_DADD_DMUL_DFMA": t["DADD"] + t["DMUL"] + t["DFMA"], "IDP": t["IDP"],
                      "cluster_barrier_UCGABAR": t["UCGABAR_ARV"]}
    out["totals"] = dict(t.most_common(40))
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
