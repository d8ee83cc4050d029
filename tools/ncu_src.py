"""Summarise an `ncu --page source --csv` dump (SASS view): per instruction the
executed count (millions), average active threads, stall samples and the top stall reason.
usage: python tools/ncu_src.py src.csv [min_exec_millions]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot_inst = sum(int(r[ix["Instructions Executed"]]) for r in rows[2:])
tot_samp = sum(int(r[ix["# Samples"]]) for r in rows[2:])
print("total warp-instructions %.1f M, samples %d" % (tot_inst / 1e6, tot_samp))
for n, r in enumerate(rows[2:]):
    ex = int(r[ix["Instructions Executed"]])
    if ex / 1e6 < thr:
        continue
    st = sorted(((int(r[i]), hdr[i]) for i in stall_cols), reverse=True)[:2]
    print("%4d %-58s ex=%7.2fM thr=%5s samp=%6s  %s" % (
        n, r[ix["Source"]].strip()[:58], ex / 1e6, r[ix["Avg. Threads Executed"]],
        r[ix["# Samples"]], " ".join("%s:%d" % (h[6:], v) for v, h in st if v)))
