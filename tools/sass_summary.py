#!/usr/bin/env python3
"""SASS evidence for the shipped library: per kernel, the counts of the Blackwell-specific mnemonics (UTCHMMA = tcgen05.mma,
LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor load / store (cp.async.bulk.tensor), ACQBULK / PREEXIT =
griddepcontrol.wait / launch_dependents (programmatic dependent launch), UBLKCP = TMA bulk copy, UBLKRED = TMA bulk reduce-add, LDGSTS = cp.async, SYNCS = mbarrier, UTCBAR = tcgen05.commit) and of
FFMA2 / FMUL2 = packed fp32 pairs (paired tensor-product kernels), and of atomics (ATOMS = shared memory, ATOMG / REDG = global: the CSR / cell-list cursors, the layer-norm weight
gradient, and the d/dx reductions of the decoupled tensor-product backward kernels -- evaluation mode only, REDG.E.ADD.F32 / .F32x2), plus a short excerpt around the first tensor-core instruction of the GEMM kernels.
  python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "equivariant-nn-zoo_b200", "lib", "libe3b200.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kernels, cur = {}, None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = []
    elif cur is not None:
        kernels[cur].append(line)
demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
names = dict(zip(kernels, demangle))
KEYS = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "UBLKRED", "LDGSTS", "SYNCS", "ACQBULK", "PREEXIT",
        "FFMA2", "FMUL2", "ATOMS", "ATOMG", "REDG"]
PAT = {"ATOMS": r"\bATOMS", "ATOMG": r"\bATOM(G|\.)", "REDG": r"\bRED(G|\.)"}
print("# cuobjdump -sass equivariant-nn-zoo_b200/lib/libe3b200.so -- mnemonic counts per kernel (sm_100a)")
print("# %-78s %s" % ("kernel", " ".join("%7s" % k for k in KEYS)))
tot = {k: 0 for k in KEYS}
for k, lines in sorted(kernels.items(), key=lambda kv: names[kv[0]]):
    cnt = {key: sum(1 for ln in lines if re.search(PAT.get(key, r"\b" + re.escape(key)), ln)) for key in KEYS}
    if not any(cnt.values()):
        continue
    for key in KEYS:
        tot[key] += cnt[key]
    short = re.sub(r"\(anonymous namespace\)::", "", names[k])
    short = re.sub(r"\(.*", "", short)
    print("  %-78s %s" % (short[:78], " ".join("%7d" % cnt[key] for key in KEYS)))
print("  %-78s %s" % ("TOTAL", " ".join("%7d" % tot[key] for key in KEYS)))
for pat in ("gemm_tf32x3_kernel<128", "wgrad_tf32x3_kernel", "tpfp_S3<64>", "tpfp2_S3"):
    for k, lines in kernels.items():
        if pat in names[k]:
            idx = next((i for i, ln in enumerate(lines) if "UTCHMMA" in ln or "UBLKCP" in ln), None)
            if idx is None:
                continue
            print(f"\n# excerpt: {pat}... around its first {'UTCHMMA' if 'UTCHMMA' in lines[idx] else 'UBLKCP'}")
            for ln in lines[max(0, idx - 6): idx + 10]:
                ln = re.sub(r"/\* 0x[0-9a-f]+ \*/", "", ln).rstrip()
                if ln.strip():
                    print("   " + ln)
            break
