#!/usr/bin/env python
"""Per-SOURCE-LINE executed warp-instructions and stall samples of one kernel.

    python scripts/ncu_lines.py <report.ncu-rep> <kernel-regex> <object .o/.so/.cubin> [per-unit divisor] [top-N]

ncu's source page (csv) lists SASS instructions with their execution counts and stall samples but not the
source line; `nvdisasm -g` of the same binary (built with -lineinfo) lists SASS offsets with file:line.  The
two are joined on the instruction offset inside the kernel.  `per-unit divisor` (e.g. warps x scenes) turns the
counts into instructions per unit of work.
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def sass_lines(obj, kern_rx):
    with tempfile.TemporaryDirectory() as tmp:
        if obj.endswith(".cubin"):
            cubins = [os.path.abspath(obj)]
        else:
            subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
            cubins = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")]
        out = {}
        for cb in cubins:
            txt = subprocess.run(["nvdisasm", "-g", "-c", cb], capture_output=True, text=True, check=True).stdout
            cur, line, take = None, None, False
            for l in txt.splitlines():
                m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
                if m:
                    cur = m.group(1)
                    take = re.search(kern_rx, subprocess.run(["c++filt", cur], capture_output=True, text=True).stdout) is not None
                    continue
                if not take:
                    continue
                m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
                if m:
                    line = (os.path.basename(m.group(1)), int(m.group(2)))
                    continue
                m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
                if m:
                    out.setdefault(cur, {})[int(m.group(1), 16)] = (line, m.group(2).strip())
        return out


def main(rep, kern, obj, per="1", top="40"):
    per, top = float(per), int(top)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h0 = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    name = rows[h0 - 1][1]
    hdr = rows[h0]
    body = []
    for r in rows[h0 + 1:]:
        if len(r) != len(hdr):
            break
        body.append(r)
    xi, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
    base = int(body[0][0], 16)
    # pick the disassembled function whose demangled name matches the profiled one best (same template args)
    funcs = sass_lines(obj, kern)
    want = re.sub(r"[^0-9A-Za-z]", "", name)
    best = max(funcs, key=lambda f: len(os.path.commonprefix([
        re.sub(r"[^0-9A-Za-z]", "", subprocess.run(["c++filt", f], capture_output=True, text=True).stdout), want])))
    sl = funcs[best]
    agg = collections.defaultdict(lambda: [0, 0, 0])
    tot = tot_s = 0
    for r in body:
        off = int(r[0], 16) - base
        line = sl.get(off, (("?", 0), ""))[0] or ("?", 0)
        c, s = int(r[xi] or 0), int(r[si] or 0)
        a = agg[line]
        a[0] += c
        a[1] += s
        a[2] += 1
        tot += c
        tot_s += s
    print(name)
    print(f"executed warp-instructions {tot} ({tot / per:.1f} per unit), stall samples {tot_s}")
    print(f"{'file:line':28s} {'instr/unit':>10s} {'% instr':>8s} {'% samples':>9s}  SASS instrs")
    for line, (c, s, n) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{line[0]}:{line[1]:<6d}".ljust(28) + f" {c / per:10.1f} {100.0 * c / tot:8.1f} {100.0 * s / max(tot_s, 1):9.1f}  {n}")


if __name__ == "__main__":
    main(*sys.argv[1:])
