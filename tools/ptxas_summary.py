#!/usr/bin/env python3
"""Summarises natrium_b200/csrc/_build/*.ptxas.log (nvcc -Xptxas -v) into a tracked table: registers, stack frame,
spill bytes and static shared memory per kernel.  Usage: python tools/ptxas_summary.py > profiles/rNN_ptxas_summary.md"""
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    short = []
    for s in out:
        s = re.sub(r"\(int\)", "", s)
        s = re.sub(r"^void ", "", s)
        s = re.sub(r"\(StreamArgs.*$|\(long.*$|\(int.*$|\(double.*$|\(NbHaloSeg.*$|\(const.*$", "", s)
        short.append(s)
    return short


def main():
    rows = []
    for log in sorted(glob.glob(os.path.join(ROOT, "natrium_b200", "csrc", "_build", "*.ptxas.log"))):
        unit = os.path.basename(log).replace(".ptxas.log", "")
        txt = open(log).read()
        for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n"
                             r"ptxas info\s+: Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes cumulative stack size)?(?:, (\d+) bytes smem)?", txt):
            rows.append((unit, m.group(1), int(m.group(5)), int(m.group(2)), int(m.group(3)), int(m.group(4)), int(m.group(8) or 0)))
    names = demangle([r[1] for r in rows])
    print("| unit | kernel | registers | stack B | spill stores B | spill loads B | static smem B |")
    print("|---|---|---|---|---|---|---|")
    for (unit, _, regs, stack, ss, sl, smem), name in zip(rows, names):
        print(f"| {unit} | `{name}` | {regs} | {stack} | {ss} | {sl} | {smem} |")
    spilled = [(n, r[4], r[5]) for r, n in zip(rows, names) if r[4] or r[5]]
    print(f"\n{len(rows)} kernels, {len(spilled)} with spills.")


if __name__ == "__main__":
    main()
