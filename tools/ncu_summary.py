"""Turns an .ncu-rep (--set full) into the per-kernel text summary committed under profiles/."""
import collections
import csv
import io
import subprocess
import sys

KEEP = ["Duration", "SM Frequency", "DRAM Throughput", "L2 Cache Throughput", "Compute (SM) Throughput", "Memory Throughput",
        "Registers Per Thread", "Achieved Occupancy", "Executed Ipc Active", "Dynamic Shared Memory Per Block"]
RAW = ["dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
       "lts__t_bytes.sum", "smsp__inst_executed.sum"]


def main(path):
    det = subprocess.run(["ncu", "-i", path, "--page", "details", "--csv"], capture_output=True, text=True).stdout
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    out = collections.OrderedDict()
    for r in csv.DictReader(io.StringIO(det)):
        name = r["Kernel Name"].split("(")[0].split("::")[-1]
        k = (int(r["ID"]), name, r["Grid Size"], r["Block Size"])
        if r["Metric Name"] in KEEP:
            out.setdefault(k, collections.OrderedDict())[r["Metric Name"]] = f'{r["Metric Value"]} {r["Metric Unit"]}'.strip()
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    rawd = {}
    for r in rows[2:]:
        rawd[int(r[hdr.index("ID")])] = {m: f"{r[hdr.index(m)]} {units[hdr.index(m)]}" for m in RAW if m in hdr}
    print(f"# ncu --set full --clock-control none summary of {path}")
    for (i, name, grid, block), m in out.items():
        print(f"\n[{i}] {name}  grid={grid} block={block}")
        for a, b in m.items():
            print(f"    {a:34s} {b}")
        for a, b in rawd.get(i, {}).items():
            print(f"    {a:70s} {b}")


def traffic(path, key, out_json, pattern="head_fwd_kernel"):
    """Records dram__bytes_read.sum + dram__bytes_write.sum of the first `pattern` launch of the capture under `key`
    (e.g. mini_L256_B160) in out_json -- bench.py reads roofline.traffic from that file."""
    import json
    import os
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows[2:]:
        if pattern in r[hdr.index("Kernel Name")]:
            total = 0.0
            for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                total += float(r[hdr.index(m)]) * scale[units[hdr.index(m)]]
            table = {}
            if os.path.exists(out_json):
                with open(out_json) as f:
                    table = json.load(f)
            table[key] = {"dram_bytes": total, "source": f"ncu --set full capture {os.path.basename(path)} "
                                                         f"(summary: profiles/r02_ncu_full_kernels_c2.txt), kernel {pattern}"}
            with open(out_json, "w") as f:
                json.dump(table, f, indent=1)
            print(key, total)
            return
    raise SystemExit(f"no {pattern} launch in {path}")


if __name__ == "__main__":
    if len(sys.argv) >= 5 and sys.argv[2] == "--traffic":
        traffic(sys.argv[1], sys.argv[3], sys.argv[4])
    else:
        main(sys.argv[1])
