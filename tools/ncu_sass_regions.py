"""Group the SASS of an ncu source-page CSV into regions of equal execution count and print their share of
executed instructions and of stall samples.  python tools/ncu_sass_regions.py src.csv [min_share]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
idx = {k: i for i, k in enumerate(hdr)}
data = [r for r in rows[h + 1:] if len(r) == len(hdr) and r[0] != "Address"]
# the page is printed once per launch in the file: keep the first copy
first = data[0][0]
for k in range(1, len(data)):
    if data[k][0] == first:
        data = data[:k]
        break


def f(r, c):
    try:
        return float(r[idx[c]] or 0)
    except ValueError:
        return 0.0


stc = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
out, prev, start = [], None, 0
for k, r in enumerate(data):
    c = int(f(r, "Instructions Executed"))
    if c != prev:
        if prev is not None:
            out.append((start, k - 1, prev))
        prev, start = c, k
out.append((start, len(data) - 1, prev))
tot = sum((b - a + 1) * c for a, b, c in out)
ts = sum(f(r, "# Samples") for r in data)
print("sass", len(data), "warp-inst", tot, "samples", ts)
for a, b, c in out:
    if (b - a + 1) * c > thr * tot:
        s = sum(f(data[k], "# Samples") for k in range(a, b + 1))
        st = {x: sum(f(data[k], x) for k in range(a, b + 1)) for x in stc}
        tops = sorted(st.items(), key=lambda x: -x[1])[:3]
        ops = {}
        for k in range(a, b + 1):
            t = data[k][1].split()
            op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
            ops[op] = ops.get(op, 0) + 1
        top = sorted(ops.items(), key=lambda x: -x[1])[:4]
        print("%5d-%5d len %4d exec %9d inst %5.1f%% samp %5.1f%% | %-36s | %s" % (
            a, b, b - a + 1, c, 100 * (b - a + 1) * c / tot, 100 * s / ts,
            " ".join("%s:%d" % t for t in top), " ".join("%s=%.0f" % (k[6:], v) for k, v in tops)))
