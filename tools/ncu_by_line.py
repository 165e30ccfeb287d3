#!/usr/bin/env python3
"""Per-source-line instruction counts of one profiled kernel (developer tool).

Joins the SASS page of an ncu report (`ncu -i rep --page source --csv`) with the line table of the cubin the report
was taken from (`nvdisasm -g -c`), by instruction offset, and prints executed warp-instructions, the pipe-weighted
share and the stall samples per line of csrc/footsies_kernels.cu (inlined callees are attributed to their own line).

    python tools/ncu_by_line.py gpurun_out/prof_step_X.ncu-rep [--kernel 'step_kernelILb0ELb0ELb1ELb1E'] [--top 60]
"""
import argparse
import collections
import csv
import io
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--lib", default=os.path.join(ROOT, "footsies_gym_b200", "libfootsies_b200.so"))
    ap.add_argument("--kernel", default="step_kernelILb0ELb0ELb1ELb1E")
    ap.add_argument("--kernel-id", default=":::1")
    ap.add_argument("--top", type=int, default=60)
    ap.add_argument("--sass", action="store_true", help="also dump the joined SASS listing")
    a = ap.parse_args()

    sass = subprocess.run(["ncu", "-i", a.report, "--page", "source", "--csv", "--kernel-id", a.kernel_id],
                          capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(sass)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    col = {n: hdr.index(n) for n in ("Address", "Source", "Instructions Executed", "Thread Instructions Executed",
                                     "# Samples", "Warp Stall Sampling (All Samples)")}
    inst = []
    for r in rows[hi + 1:]:
        if not r or not r[0].startswith("0x"):
            break                                   # next kernel's section
        inst.append(r)
    base = int(inst[0][col["Address"]], 16)

    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(a.lib)], cwd=td, capture_output=True)
        cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
        dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cubin)], capture_output=True, text=True).stdout
    # walk the function's listing: "//## File "...", line N" markers (possibly with inlined-at), then instructions
    lines = dis.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and a.kernel in l)
    line_of = {}
    cur = None
    off_re = re.compile(r"^\s*/\*([0-9a-f]{4,})\*/\s+(.*?);")
    file_re = re.compile(r'//## File "([^"]+)", line (\d+)')
    for l in lines[start + 1:]:
        if l.startswith(".text.") or l.startswith("//--------------------- .text"):
            break
        m = file_re.search(l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = off_re.match(l)
        if m:
            line_of[int(m.group(1), 16)] = cur

    agg = collections.defaultdict(lambda: [0, 0, 0, 0])
    total = 0
    listing = []
    for r in inst:
        off = int(r[col["Address"]], 16) - base
        n = int(r[col["Instructions Executed"]] or 0)
        t = int(r[col["Thread Instructions Executed"]] or 0)
        s = int(r[col["# Samples"]] or 0)
        key = line_of.get(off)
        ent = agg[key]
        ent[0] += n
        ent[1] += t
        ent[2] += s
        ent[3] += 1
        total += n
        listing.append((off, key, n, t, s, r[col["Source"]]))
    csrc = os.path.join(ROOT, "footsies_gym_b200", "csrc")
    srcs = {f: open(os.path.join(csrc, f)).read().splitlines() for f in os.listdir(csrc)}
    print(f"total warp-instructions executed: {total}")
    print(f"{'line':>16} {'winst':>11} {'%':>6} {'thr/inst':>8} {'samples':>8} {'#sass':>5}  source")
    for key, (n, t, s, k) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:a.top]:
        text = ""
        if key and key[0] in srcs and key[1] <= len(srcs[key[0]]):
            text = srcs[key[0]][key[1] - 1].strip()[:110]
        name = f"{key[0][:9]}:{key[1]}" if key else "?"
        print(f"{name:>16} {n:>11} {100.0 * n / total:>6.2f} {t / max(n, 1):>8.1f} {s:>8} {k:>5}  {text}")
    if a.sass:
        for off, key, n, t, s, text in listing:
            where = f"{key[0][:6]}:{key[1]}" if key else "?"
            print(f"{off:06x} {where:>11} {n:>9} {t / max(n, 1):>5.1f} {s:>5}  {text}")


if __name__ == "__main__":
    main()
