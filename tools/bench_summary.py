import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print("value %.0f ms/step %.2f frac %.3f" % (d["value"], d["ms_per_step"], d["roofline"]["frac"]), [ (k["kernel"], round(k["ms_total"]/k["launches"],3)) for k in d["kernels"]])
