#!/usr/bin/env python
"""Aggregate the per-instruction warp-stall samples of one kernel from an .ncu-rep (source page).

    python scripts/ncu_stalls.py gpurun_out/prof.ncu-rep <kernel-regex> [top-N instructions]
"""
import csv
import subprocess
import sys


def main(rep, kern, top=12):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    # several kernel instances may follow each other: take the first
    hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    h0 = hdr_idx[0]
    end = hdr_idx[1] - 1 if len(hdr_idx) > 1 else len(rows)
    hdr, body = rows[h0], [r for r in rows[h0 + 1:end] if len(r) == len(rows[h0])]
    print(rows[h0 - 1][:2] if h0 else "")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = {h: sum(int(r[i] or 0) for r in body) for i, h in stall_cols}
    all_s = sum(tot.values())
    print("total samples", all_s)
    for h, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        if v:
            print(f"  {h:28s} {v:8d}  {100.0 * v / all_s:5.1f}%")
    si, xi = hdr.index("# Samples"), hdr.index("Instructions Executed")
    print("hottest instructions:")
    for r in sorted(body, key=lambda r: -int(r[si] or 0))[:int(top)]:
        reasons = sorted(((int(r[i] or 0), h) for i, h in stall_cols), reverse=True)[:2]
        print(f"  {int(r[si]):6d} x{r[xi]:>9s}  {r[1].strip()[:70]:70s} {reasons}")
    ops = {}
    for r in body:
        op = r[1].strip().split()[0] if r[1].strip() else "?"
        if op.startswith("@"):
            op = r[1].strip().split()[1]
        op = op.split(".")[0]
        ops[op] = ops.get(op, 0) + int(r[xi] or 0)
    tot_i = sum(ops.values())
    print("executed warp-instructions by opcode:", tot_i)
    for op, v in sorted(ops.items(), key=lambda kv: -kv[1])[:12]:
        print(f"  {op:10s} {v:12d} {100.0 * v / tot_i:5.1f}%")


if __name__ == "__main__":
    main(*sys.argv[1:])
