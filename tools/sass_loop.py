"""Opcode histogram of the unrolled chunk loop of one kernel in a cubin/.o/.so (cuobjdump -sass).

    python tools/sass_loop.py <file> <mangled-kernel-name-substring> [steps_per_trip]

The loop is taken as the longest run of instructions between two backward branches that contains FFMA2s; prints the
per-step opcode mix, the number of operands flagged .reuse and a register-file read estimate per step
(distinct source registers per instruction, minus .reuse hits; 64-bit operands count two)."""
import collections
import re
import subprocess
import sys


def sass(path, name):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    blocks = out.split("Function : ")
    for b in blocks[1:]:
        if name in b.split("\n", 1)[0]:
            ins = []
            for line in b.split("\n"):
                m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
                if m:
                    ins.append((int(m.group(1), 16), m.group(2).strip()))
            return ins
    raise SystemExit(f"kernel {name} not found")


def main():
    path, name = sys.argv[1], sys.argv[2]
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    ins = sass(path, name)
    # basic blocks delimited by branch targets; find the straight-line region with the most FFMA2
    best = (0, 0, 0)
    start = 0
    cnt = 0
    for k, (pc, t) in enumerate(ins):
        if "FFMA2" in t:
            cnt += 1
        if re.search(r"\bBRA\b|\bEXIT\b|SYNCS.PHASECHK", t):
            if cnt > best[0]:
                best = (cnt, start, k)
            start, cnt = k + 1, 0
    _, a, b = best
    body = ins[a:b + 1]
    hist = collections.Counter()
    reuse = 0
    rf_reads = 0
    for pc, t in body:
        t2 = re.sub(r"^@!?U?P\d+\s+", "", t)
        op = t2.split()[0].split(".")[0]
        hist[op] += 1
        reuse += t2.count(".reuse")
        ops = t2.split(None, 1)[1] if " " in t2 else ""
        srcs = ops.split(",")[1:]
        regs = set()
        for sreg in srcs:
            m = re.search(r"\bR(\d+)(\.reuse)?", sreg)
            if m and not m.group(2):
                r = int(m.group(1))
                regs.add(r)
                if "F32x2" in sreg or ".64" in sreg:
                    regs.add(r + 1)
        rf_reads += len(regs)
    n = len(body)
    print(f"{name}: straight-line body {n} instructions = {n / steps:.2f} per step over {steps} steps "
          f"(pc {body[0][0]:#x}..{body[-1][0]:#x})")
    for op, c in hist.most_common():
        print(f"  {op:10s} {c:5d}  ({c / steps:6.2f} per step)")
    print(f"  .reuse operands {reuse} ({reuse / steps:.2f} per step); register-file source reads ~{rf_reads / steps:.1f} per step")


if __name__ == "__main__":
    main()
