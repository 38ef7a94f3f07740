"""Print the last step's launches (name, us, DRAM MB read / written) from a one-pass ncu launch list csv."""
import collections, csv, re, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]; ci = {n: i for i, n in enumerate(h)}
L = collections.OrderedDict()
U = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}
for r in rows[hi + 1:]:
    if len(r) < len(h):
        continue
    d = L.setdefault(int(r[ci["ID"]]), {"name": re.sub(r"\(.*", "", r[ci["Kernel Name"]]).replace("void ", "").replace("tcs::", "")})
    d[r[ci["Metric Name"]]] = float(r[ci["Metric Value"]].replace(",", "")) * U[r[ci["Metric Unit"]]]
ids = list(L)
last = max(i for i, k in enumerate(ids) if "corr_build" in L[k]["name"] or "alt_tc" in L[k]["name"])
first = last
while first > 0 and "prepass" in L[ids[first - 1]]["name"]:
    first -= 1
seen = 0
for k in ids[first:]:
    d = L[k]
    if "corr_lookup" in d["name"] or "alt_tc" in d["name"]:
        seen += 1
        if seen > 3:
            continue
    print("%-45s %8.1f us  %8.1f MB read %8.1f MB written" % (d["name"][-45:], d["gpu__time_duration.sum"], d["dram__bytes_read.sum"], d["dram__bytes_write.sum"]))
