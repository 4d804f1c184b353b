"""ncu `--metrics gpu__time_duration.sum[,dram__bytes_*] --csv` launch list -> compact per-kernel summary
(markdown).  usage: python tools/summarize_launches.py launches.csv > profiles/rNN_launches.md"""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    recs = collections.OrderedDict()
    for row in csv.DictReader(lines):
        d = recs.setdefault(row["ID"], {"name": re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
                                        .replace("wmk::<unnamed>::", "").replace("wmk::", ""), "by": 0.0, "us": 0.0})
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        if "time" in row["Metric Name"]:
            d["us"] = v / 1e3 if u == "ns" else (v if u in ("us", "usecond") else v * 1e3)
        elif "dram__bytes" in row["Metric Name"]:
            d["by"] += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    agg = collections.OrderedDict()
    for d in recs.values():
        a = agg.setdefault(d["name"], [0, 0.0, 0.0])
        a[0] += 1
        a[1] += d["us"]
        a[2] += d["by"]
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total us | share | avg us | DRAM GB/s |")
    print("|---|---:|---:|---:|---:|---:|")
    for k, (c, us, by) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("| `%s` | %d | %.1f | %.1f%% | %.1f | %s |" % (k[:70], c, us, 100 * us / tot, us / c,
                                                             "%.0f" % (by / us / 1e3) if by else "-"))
    print("\ntotal %.1f us over %d launches (cold-cache, serialised by ncu: compare shares, not absolutes)" % (tot, len(recs)))


if __name__ == "__main__":
    main(sys.argv[1])
