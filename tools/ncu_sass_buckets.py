"""Bucket the SASS-level warp-stall samples of an ncu report by code position.
    ncu -i rep --page source --csv --launch-count 1 > src.csv ; python tools/ncu_sass_buckets.py src.csv [buckets]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
B = int(sys.argv[2]) if len(sys.argv) > 2 else 60
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
idx = {k: i for i, k in enumerate(hdr)}
data = [r for r in rows[h + 1:] if len(r) == len(hdr) and r[0] != "Address"]


def f(r, c):
    try:
        return float(r[idx[c]] or 0)
    except ValueError:
        return 0.0


S = [f(r, "# Samples") for r in data]
I = [f(r, "Instructions Executed") for r in data]
tot, ti = sum(S), sum(I)
print("sass rows", len(data), "samples", tot, "warp-inst", ti)
stc = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
n = len(data)
for b in range(B):
    lo, hi = b * n // B, (b + 1) * n // B
    s, i = sum(S[lo:hi]), sum(I[lo:hi])
    ops = {}
    for r in data[lo:hi]:
        t = r[1].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        ops[op] = ops.get(op, 0) + 1
    top = sorted(ops.items(), key=lambda x: -x[1])[:4]
    st = {c: sum(f(r, c) for r in data[lo:hi]) for c in stc}
    tops = sorted(st.items(), key=lambda x: -x[1])[:3]
    print("%3d [%5d-%5d] samp %5.1f%% inst %5.1f%% %-40s | %s" % (
        b, lo, hi, 100 * s / tot, 100 * i / ti, " ".join("%s:%d" % t for t in top),
        " ".join("%s=%.0f" % (k[6:], v) for k, v in tops)))
