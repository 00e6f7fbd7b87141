"""Aggregate an ncu launch list (`--metrics gpu__time_duration.sum --csv`) per kernel: count, total, average, share."""
import collections
import csv
import json
import sys


def summarize(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    scale = {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"].split("(")[0]
        agg[name][0] += 1
        agg[name][1] += float(row["Metric Value"].replace(",", "")) * scale[row["Metric Unit"]]
    total = sum(v[1] for v in agg.values())
    out = {"file": path, "launches": sum(v[0] for v in agg.values()), "total_ms": total / 1e6, "kernels": []}
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out["kernels"].append({"kernel": name, "launches": n, "total_ms": round(t / 1e6, 3),
                               "avg_us": round(t / n / 1e3, 2), "share": round(t / total, 4)})
    return out


if __name__ == "__main__":
    print(json.dumps(summarize(sys.argv[1]), indent=1))
