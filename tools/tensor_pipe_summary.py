"""Time-weighted tensor-pipe utilisation of the GEMM kernels of one step from an ncu metrics pass:

    ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --csv ...

Forward = every launch before the first weight-gradient kernel.  usage: python tools/tensor_pipe_summary.py <csv>"""
import collections
import csv
import re
import sys

rows = collections.OrderedDict()
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    d = rows.setdefault(r["ID"], {"name": re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")})
    v = float(r["Metric Value"].replace(",", ""))
    if r["Metric Name"] == "gpu__time_duration.sum":
        d["us"] = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(r["Metric Unit"], 1e-3)
    elif "pipe_tensor" in r["Metric Name"]:
        d["tp"] = v
launches = [d for d in rows.values() if "us" in d and "tp" in d]
first_wgrad = next((i for i, d in enumerate(launches) if "wgrad" in d["name"]), len(launches))


def agg(ls):
    t = sum(d["us"] for d in ls)
    return t, (sum(d["us"] * d["tp"] for d in ls) / t if t else 0.0)


def table(ls, title):
    by = collections.OrderedDict()
    for d in ls:
        e = by.setdefault(d["name"], [0, 0.0, 0.0])
        e[0] += 1
        e[1] += d["us"]
        e[2] += d["us"] * d["tp"]
    t, tp = agg(ls)
    print(f"{title}: {len(ls)} launches, {t / 1e3:.3f} ms (cold-cache, serialised), time-weighted tensor pipe {tp:.1f} %")
    for k, (n, us, w) in sorted(by.items(), key=lambda kv: -kv[1][1]):
        print(f"   {k[:74]:74s} {n:4d} launches {us:9.1f} us   tensor pipe {w / us:5.1f} %")


gemm = [d for d in launches if "gemm" in d["name"]]
fwd = [d for d in launches[:first_wgrad] if "gemm" in d["name"]]
table(fwd, "FORWARD GEMM kernels (backbone + head + Patch-PnP)")
print()
table([d for d in launches[first_wgrad:] if "gemm" in d["name"]], "BACKWARD GEMM kernels (dgrad + wgrad)")
print()
table(gemm, "ALL GEMM kernels of the step")
