"""Reduce `ncu --page source --csv` output (one block per kernel launch) to the columns that matter for an instruction diet:
address, SASS, stall samples, instructions executed, shared-memory wavefronts.  Usage: python scripts/ncu_source_summary.py in.csv out.csv"""
import csv, sys
csv.field_size_limit(10 ** 9)
keep = ["Address", "Source", "# Samples", "Instructions Executed", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_mio", "stall_lg", "stall_barrier", "stall_math", "stall_not_selected"]
with open(sys.argv[1]) as f, open(sys.argv[2], "w", newline="") as g:
    w = csv.writer(g)
    idx = None
    for row in csv.reader(f):
        if not row:
            continue
        if row[0] == "Kernel Name":
            w.writerow(row[:2])
            continue
        if row[0] == "Address":
            idx = [row.index(k) for k in keep if k in row]
            w.writerow([row[i] for i in idx])
            continue
        if idx and (row[idx[2]] not in ("0", "") or row[idx[3]] not in ("0", "")):
            w.writerow([row[i] for i in idx])
