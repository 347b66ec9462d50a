#!/usr/bin/env python
"""Per-kernel SASS summary of libvllm_ltr_b200.so (cuobjdump -sass, runs without a GPU): the
mnemonics that prove what a kernel is built on - tcgen05 (UTCHMMA / UTCBAR / LDTM), TMA (UTMALDG /
UBLKCP), warp-level tensor cores (HMMA), async copies (LDGSTS), NVLS multimem.ld_reduce (LDGMC;
multimem.st is a plain STG.STRONG.SYS to the multicast address), barriers.

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vllm_ltr_b200", "libvllm_ltr_b200.so")
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UBLKCP", "HMMA", "LDSM", "MOVM", "LDGSTS",
        "SYNCS", "REDUX", "LDGMC", "LDG", "STG", "LDS", "STS", "BAR", "ATOM", "RED", "MEMBAR",
        "FENCE", "ACQBULK", "NANOSLEEP"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            base = op.split(".")[0]
            if base in KEYS:
                kernels[cur][base] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True,
                              text=True).stdout.splitlines()
    groups = collections.OrderedDict()
    for mangled, name in zip(kernels, demangle):
        short = re.sub(r"\(.*", "", name)
        short = re.sub(r"^void ", "", short)
        fam = re.sub(r"<.*", "", short)
        groups.setdefault(fam, []).append((short, kernels[mangled]))
    print(f"SASS summary of {os.path.relpath(LIB, ROOT)} (sm_100a), {len(kernels)} kernel "
          f"instantiations in {len(groups)} families.  Per family: instantiations, instruction "
          "count range, and the summed counts of the mnemonics of interest.\n")
    for fam, items in groups.items():
        tot = collections.Counter()
        for _, c in items:
            tot.update(c)
        lo = min(c["_total"] for _, c in items)
        hi = max(c["_total"] for _, c in items)
        marks = "  ".join(f"{k}x{tot[k]}" for k in KEYS if tot[k])
        print(f"{fam}\n    {len(items)} instantiation(s), {lo}..{hi} instructions each\n    {marks}\n")


if __name__ == "__main__":
    main()
