"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: total time per kernel name and share."""
import csv
import collections
import re
import sys


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        rows.append((r["Kernel Name"], val * scale))
    tot = sum(v for _, v in rows)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for name, v in rows:
        short = name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
        short = re.sub(r"^void ", "", short)
        short = re.sub(r"\(.*", "", short)                  # drop the argument list
        short = re.sub(r"<(?!\d+(, *\w+)*>).*", "", short)  # keep small integer template arguments only
        if "elementwise_kernel" in short:                    # PyTorch's generic kernels: add the functor name
            inner = re.findall(r"native::(?:<unnamed>::)?(\w+)", name.replace("(anonymous namespace)::", ""))
            if len(inner) > 1:
                short += "[" + inner[1] + "]"
        short = short[:70]
        agg[short][0] += 1
        agg[short][1] += v
    print(f"{len(rows)} launches, {tot/1e3:.3f} ms total device time (serialised, cold-cache)")
    print(f"{'kernel':72s} {'n':>6s} {'total us':>12s} {'avg us':>10s} {'share':>7s}")
    for name, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
        print(f"{name:72s} {n:6d} {v:12.1f} {v/n:10.2f} {100*v/tot:6.2f}%")


if __name__ == "__main__":
    main(sys.argv[1])
