"""Key counters + top stalled SASS instructions of the first kernel in an .ncu-rep."""
import csv, io, subprocess, sys
path = sys.argv[1]
raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, r = rows[0], rows[1], rows[2]
keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__cycles_active.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_pipe_alu.sum", "smsp__inst_executed_pipe_fma.sum"]
for k in keys:
    if k in hdr:
        print(f"{k:75s} {r[hdr.index(k)]} {units[hdr.index(k)]}")
src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h, data = rows[1], rows[2:]
isrc, isamp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
stall = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(int(x[isamp] or 0) for x in data)
print("total samples", tot)
for x in sorted(data, key=lambda x: -int(x[isamp] or 0))[:int(sys.argv[2]) if len(sys.argv) > 2 else 14]:
    st = {h[i]: int(x[i]) for i in stall if x[i] and int(x[i]) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:2])
    print(f"{int(x[isamp]):7d} {100*int(x[isamp])/tot:5.1f}% ex={x[iex]:>9s} {x[isrc][:58]:58s} {st}")
