#!/usr/bin/env python
"""Read an .ncu-rep captured with `--set full --import-source on` (no GPU needed) and print, per kernel instance:
duration, instructions, issue / warp occupancy, DRAM traffic, the stall breakdown, the opcode histogram and the source
lines that execute the most instructions.

    python profiles/hotspots.py gpurun_out/prof.ncu-rep [kernel-name-regex] [top-N]

This is how the round-1 K2 rewrite was steered: barrier stalls + 7-way shared-atomic bank conflicts in k_mask_items,
then the per-line instruction counts of k_mask_bytes (profiles/README.md)."""
import collections
import csv
import re
import subprocess
import sys


def ncu(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def raw_page(rep, pat):
    rows = ncu(rep, "raw")
    hdr = rows[0]
    want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread",
            "launch__grid_size", "launch__block_size"]
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        name = r[hdr.index("Kernel Name")]
        if pat and not re.search(pat, name):
            continue
        print("==", name[:100])
        for w in want:
            if w in hdr:
                print("   %-62s %s %s" % (w, r[hdr.index(w)], rows[1][hdr.index(w)]))
        stalls = [(float(r[i]), h) for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and r[i]]
        for v, h in sorted(stalls, reverse=True)[:8]:
            print("   stall %-55s %.2f" % (h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""), v))


def source_page(rep, pat, top):
    extra = ["--print-source", "sass"] + (["--kernel-name", "regex:" + pat] if pat else [])
    rows = ncu(rep, "source", extra)
    try:
        hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
    except StopIteration:
        print("(no source page in this report)")
        return
    hdr = rows[hi]
    ia, isrc = hdr.index("Instructions Executed"), hdr.index("Source")
    body = []
    for r in rows[hi + 1:]:
        if len(r) <= ia or not r[ia].isdigit():
            if body:
                break  # first kernel instance only
            continue
        body.append(r)
    total = sum(int(r[ia]) for r in body)
    print("   SASS lines %d, warp instructions %d" % (len(body), total))
    ops = collections.Counter()
    for r in body:
        w = r[isrc].split()
        op = w[1] if w and w[0].startswith("@") and len(w) > 1 else (w[0] if w else "?")
        ops[op.split(".")[0]] += int(r[ia])
    print("   opcodes:", ", ".join("%s %.1f%%" % (k, 100.0 * v / max(total, 1)) for k, v in ops.most_common(12)))
    for r in sorted(body, key=lambda r: -int(r[ia]))[:top]:
        print("   %10s  %s" % (r[ia], r[isrc][:110]))


if __name__ == "__main__":
    rep = sys.argv[1]
    pat = sys.argv[2] if len(sys.argv) > 2 else ""
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    raw_page(rep, pat)
    source_page(rep, pat, top)
