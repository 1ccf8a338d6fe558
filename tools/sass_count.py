#!/usr/bin/env python
"""Instruction mix of the pair loop of a kernel, from the SASS of the built library (runs without a GPU).

    python tools/sass_count.py [--write]        # --write: profiles/sass_fp64.json, read by bench.py for the FP64-pipe roofline

The pair loop is the backward branch with the longest span in the function.  FP64-pipe instructions are those the half-rate FP64 pipe
of sm_100 issues (measured in profiles/r01/fp64_pipe_microbench.txt: DFMA/DMUL/DADD one warp-instruction per 2 cycles per SMSP, MUFU.RSQ64H
one per 8): DFMA DMUL DADD DSETP DMNMX and the F64 conversions; MUFU.RSQ64H runs on the XU pipe and is listed separately.
"""
import collections
import hashlib
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ndspmhd_b200", "libndspmhd_b200.so")
KERNELS = {
    "rates_pair_kernel<3,MHD,FAST=2>": "rates_pair_kernelILi3ELb1ELb0ELi2ELb0EE",
    "rates_pair_kernel<3,hydro,FAST=2>": "rates_pair_kernelILi3ELb0ELb0ELi2ELb0EE",
    "rates_pair_kernel<2,MHD,FAST=2>": "rates_pair_kernelILi2ELb1ELb0ELi2ELb0EE",
    "rates_pair_kernel<3,hydro,DRAG>": "rates_pair_kernelILi3ELb0ELb1ELi0ELb0EE",
    "density_round_kernel<3,FIRST,LIGHT>": "density_round_kernelILi3ELb1ELb0ELb1EE",
    "density_round_kernel<3,PARTIAL,LIGHT>": "density_round_kernelILi3ELb0ELb0ELb1EE",
}
FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX", "F2F.F64", "I2F.F64", "F2I.F64", "F2I.S64.F64", "F2I.U32.F64", "F2I.F64.TRUNC", "I2F.F64.S32")


def source_sha() -> str:
    """Hash of the sources the kernels are built from: ties profiles/*.json to the build bench.py times."""
    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "ndspmhd_b200", "csrc")
    for f in sorted(os.listdir(csrc)):
        if f.endswith((".cu", ".cuh", ".h")):
            h.update(open(os.path.join(csrc, f), "rb").read())
    return h.hexdigest()[:16]


def functions(lib=LIB):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    name, cur = None, []
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                yield name, cur
            name, cur = m.group(1), []
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and name:
            cur.append((int(m.group(1), 16), m.group(2).strip()))
    if name:
        yield name, cur


def _is_fp64(op: str) -> bool:
    return op.startswith(("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")) or (".F64" in op and op.startswith(("F2F", "I2F", "F2I")))


def _op(text: str) -> str:
    return re.sub(r"^@!?U?P\d+\s+", "", text).split()[0]


def loop_mix(instrs):
    # backward branches = loops; the pair loop is the SHORTEST span that still holds most of the function's FP64 work (the persistent
    # outer loop around it holds all of it, the list-batch loops hold none)
    spans = []
    for addr, text in instrs:
        m = re.search(r"\bBRA\S*\s+(?:\S+,\s*)?(0x[0-9a-f]+)", text)
        if m and int(m.group(1), 16) < addr:
            tgt = int(m.group(1), 16)
            spans.append((addr - tgt, tgt, addr, sum(1 for a, t in instrs if tgt <= a <= addr and _is_fp64(_op(t)))))
    if not spans:
        return None
    most = max(sp[3] for sp in spans)
    best = min((sp for sp in spans if sp[3] >= 0.6 * most), key=lambda sp: sp[0])[1:3]
    mix = collections.Counter()
    n = 0
    for addr, text in instrs:
        if best[0] <= addr <= best[1]:
            mix[_op(text)] += 1
            n += 1
    fp64 = sum(v for k, v in mix.items() if _is_fp64(k))
    return {"loop_instructions": n, "fp64_pipe_instructions": fp64, "mufu_rsq64h": sum(v for k, v in mix.items() if k.startswith("MUFU.RSQ64H")),
            "ldg": sum(v for k, v in mix.items() if k.startswith("LDG")), "lds": sum(v for k, v in mix.items() if k.startswith("LDS")),
            "by_op": {k: v for k, v in sorted(mix.items(), key=lambda kv: -kv[1]) if v >= 2}}


def main():
    res = {"source_sha": source_sha(), "how": "cuobjdump -sass of ndspmhd_b200/libndspmhd_b200.so; loop = longest backward branch span (tools/sass_count.py)", "kernels": {}}
    for fname, instrs in functions():
        for label, key in KERNELS.items():
            if key in fname:
                res["kernels"][label] = loop_mix(instrs)
    for k, v in res["kernels"].items():
        print(f"{k:42s} loop {v['loop_instructions']:4d}  FP64-pipe {v['fp64_pipe_instructions']:4d}  RSQ64H {v['mufu_rsq64h']}  LDG {v['ldg']}  LDS {v['lds']}")
    if "--write" in sys.argv:
        json.dump(res, open(os.path.join(ROOT, "profiles", "sass_fp64.json"), "w"), indent=1)
    return res


if __name__ == "__main__":
    main()
