#!/usr/bin/env python3
"""Disassembly evidence for profiles/: `cuobjdump -sass` of the in-tree library, per kernel an opcode histogram and the
excerpts that show HOW the kernel talks to the hardware (TMA bulk copies + mbarriers, programmatic dependent launch,
packed FFMA2, no tensor-core instructions where nothing is a contraction).

    python tools/sass_report.py step    > profiles/r02_sass_step_kernel.txt
    python tools/sass_report.py rollout > profiles/r02_sass_rollout_kernel.txt
    python tools/sass_report.py policy  > profiles/r02_sass_policy_kernel.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "footsies_gym_b200", "libfootsies_b200.so")

# which instantiation stands for the family (c++filt'ed name must contain every piece)
TARGETS = {
    "step": ("step_kernel", ["StepShape<768, 256, 3, 1>", "false, false, true, true, false"],
             "K = 1, P1 = agent, P2 = BattleAI, dense reward, unmasked: the benchmark kernel (bench.py roofline)"),
    "rollout": ("rollout_mma_kernel", ["<64, 4, 1, true, false, false>"],
                "hidden 64, 4 warps x 16 battles, dense reward, P2 = BattleAI: BASELINE configs[4] at 16 384 battles per GPU"),
    "policy": ("policy_mma_sample_kernel", ["<64, false>"], "hidden 64: the per-step policy kernel"),
}
INTERESTING = [
    ("TMA bulk copy (cp.async.bulk, 1-D)", r"\bUBLKCP\b|\bUBLKRED\b"),
    ("mbarrier (SYNCS)", r"\bSYNCS\b"),
    ("programmatic dependent launch (griddepcontrol)", r"\bACQBULK\b|\bPDL\b|\bDEPBAR\b.*SB|\bBAR\.ARV\b|PREEXIT|\bACQ"),
    ("packed fp32 pairs (FFMA2 / FADD2 / FMUL2)", r"\bFFMA2\b|\bFADD2\b|\bFMUL2\b"),
    ("tensor-core MMA (step kernel: none -- nothing there is a contraction; rollout / policy kernels: HMMA.1688.F32.TF32 in layer 1, HMMA.16816.F32 on FP16 operands in layers 2 and 3)", r"\bHMMA\b|\bIMMA\b|\bUTCHMMA\b|\bUTCMMA\b|\bUTC[A-Z]*MMA\b|\bQGMMA\b|\bHGMMA\b"),
    ("128-bit global stores / loads", r"\bSTG\.E\.128\b|\bLDG\.E\.128\b"),
    ("warp reductions (REDUX) / votes", r"\bREDUX\b|\bVOTE\b"),
    ("local memory (spills; expected: none)", r"\bSTL\b|\bLDL\b"),
]


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "step"
    base, pieces, what = TARGETS[which]
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)[1:]
    chosen = None
    for f in funcs:
        mangled = f.split("\n", 1)[0].strip()
        name = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
        if base in name and all(p in name for p in pieces):
            chosen = (name, f)
            break
    if chosen is None:
        raise SystemExit(f"no instantiation of {base} with {pieces} in {LIB}")
    name, body = chosen
    lines = [ln for ln in body.split("\n") if re.search(r"/\*[0-9a-f]{4}\*/", ln)]
    ops = collections.Counter()
    insts = []
    for ln in lines:
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m:
            ops[m.group(1).split(".")[0]] += 1
            insts.append(ln.strip())
    print(f"# {which}: {what}")
    print(f"# {name}")
    print(f"# source: cuobjdump -sass footsies_gym_b200/libfootsies_b200.so (sm_100a), regenerate with tools/sass_report.py {which}")
    hdr = re.search(r"\.headerflags.*", body)
    print(f"# static instructions: {len(insts)}")
    print("\n## opcode histogram (static count)")
    for op, n in ops.most_common():
        print(f"{op:14s} {n}")
    for title, pat in INTERESTING:
        hits = [ln for ln in insts if re.search(pat, ln)]
        print(f"\n## {title}: {len(hits)} instruction(s)")
        if "MMA" in title and hits:          # by full mnemonic: which operand types the tensor cores are fed
            kinds = collections.Counter(re.search(r"([A-Z]*MMA[A-Z0-9_.]*)", ln).group(1) for ln in hits)
            print("    by variant: " + ", ".join(f"{k} x {n}" for k, n in kinds.most_common()))
        for ln in hits[:12]:
            print("   ", re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", ln))
        if len(hits) > 12:
            print(f"    ... {len(hits) - 12} more")


if __name__ == "__main__":
    main()
