"""profiles/r02_launches.md from the one-pass launch list of tools/measure_r02.sh: device time AND DRAM bytes of every
launch of the LAST timed step, taken with --cache-control none in the step's natural order (so the bytes are the in-step
traffic: earlier kernels' data and the previous calls' output planes are what L2 holds), plus profiles/roofline_traffic.json.

    python tools/summarise_step.py gpurun_out/r02_launches.csv gpurun_out/r02_bench.json
"""
import collections
import csv
import json
import re
import sys

src, bench = sys.argv[1], sys.argv[2]
rows = list(csv.reader(l for l in open(src) if not l.startswith("==")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h, data = rows[hi], rows[hi + 1:]
ci = {n: i for i, n in enumerate(h)}
L = collections.OrderedDict()
for r in data:
    if len(r) < len(h):
        continue
    d = L.setdefault(int(r[ci["ID"]]), {"name": re.sub(r"\(.*", "", r[ci["Kernel Name"]]).replace("void ", "").replace("tcs::", "")})
    v, u = float(r[ci["Metric Value"]].replace(",", "")), r[ci["Metric Unit"]]
    m = r[ci["Metric Name"]]
    if m.startswith("gpu__time"):
        v = v / 1e3 if u.startswith("n") else v * 1e3 if u.startswith("m") else v          # -> us
    else:
        v = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u] / 1e6             # -> MB
    d[m] = v
ids = list(L)
# the last step: from the last fused build (or prepass) to the last lookup
last_build = max(i for i, k in enumerate(ids) if "corr_build" in L[k]["name"])
step = [L[k] for k in ids[last_build:] if not L[k]["name"].startswith("at::")]
b = json.load(open(bench))
B = b["config"]["seqs_per_gpu"]
H, W = b["config"]["feature_hw"]
alg_lookup = 308 * B * H * W / 1e6
agg = collections.OrderedDict()
for d in step:
    a = agg.setdefault(d["name"], {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0})
    a["n"] += 1; a["us"] += d["gpu__time_duration.sum"]; a["rd"] += d["dram__bytes_read.sum"]; a["wr"] += d["dram__bytes_write.sum"]
tot = sum(a["us"] for a in agg.values())
out = ["# Round 2: every launch of one step with its in-step DRAM traffic", "",
       "`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none`",
       "on `python bench.py --steps 3 --warmup 3 --skip-e2e --skip-cpu --skip-gpu-reference` (tools/measure_r02.sh): one pass, so",
       "every kernel ran once, in the step's own order, on the caches its predecessors left.  Times are serialised (compare",
       "SHARES with the bench's phases); bytes are what the step really moves through DRAM.  Last timed step, %d sequences of" % B,
       "%dx%d features:" % (H, W), "",
       "| kernel | launches | mean us | total us | share | DRAM read MB / launch | DRAM write MB / launch |", "|---|---|---|---|---|---|---|"]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    out.append("| `%s` | %d | %.1f | %.1f | %.1f%% | %.1f | %.1f |" % (k, a["n"], a["us"] / a["n"], a["us"], 100 * a["us"] / tot, a["rd"] / a["n"], a["wr"] / a["n"]))
lk = [d for d in step if "corr_lookup" in d["name"]]
steady = lk[1:]
rd = sum(d["dram__bytes_read.sum"] for d in steady) / len(steady)
wr = sum(d["dram__bytes_write.sum"] for d in steady) / len(steady)
out += ["", "Lookup, calls 2..32 of the step (the first one also pays for the warp phase's leftovers: %.1f + %.1f MB):" % (lk[0]["dram__bytes_read.sum"], lk[0]["dram__bytes_write.sum"]),
        "**%.1f MB read + %.1f MB written = %.1f MB per launch** against %.1f MB algorithmic (308 B/pixel): %.2fx." % (rd, wr, rd + wr, alg_lookup, (rd + wr) / alg_lookup),
        "At the bench's in-step %.2f us per launch that is %.0f GB/s of real DRAM traffic (%.0f %% of the measured 6538 GB/s):" % (
            1e3 * b["phases"]["lookups_ms"] / 32, (rd + wr) / (b["phases"]["lookups_ms"] / 32) , 100 * (rd + wr) / (b["phases"]["lookups_ms"] / 32) / 6538),
        "the kernel is not at the DRAM limit; what it waits for is the latency of its scattered 32/16-byte window loads."]
open("profiles/r02_launches.md", "w").write("\n".join(out) + "\n")
json.dump({"source": "profiles/r02_launches.md: ncu one-pass launch list of bench.py (--cache-control none), mean over lookup calls 2..32 of the last timed step",
           "corr_lookup_in_step_bytes_per_launch_B%d" % B: round((rd + wr) * 1e6),
           "corr_lookup_in_step_read_bytes_B%d" % B: round(rd * 1e6), "corr_lookup_in_step_write_bytes_B%d" % B: round(wr * 1e6),
           "corr_lookup_bytes_per_launch_B8_isolated_r01": 81800000}, open("profiles/roofline_traffic.json", "w"), indent=1)
print("\n".join(out))
