"""Summarise an .ncu-rep (read on the CPU box): key raw metrics + dynamic SASS opcode mix + stall reasons."""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
kid = sys.argv[2] if len(sys.argv) > 2 else "::regex:advect:1"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__cycles_elapsed.max", "launch__shared_mem_per_block_dynamic"]
for w in want:
    for i, h in enumerate(hdr):
        if h == w:
            print(f"{w}: {[r[i] for r in rows[1:]]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", kid], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
secs = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hdr = rows[secs[0]]
data = rows[secs[0] + 1 : (secs[1] - 1 if len(secs) > 1 else len(rows))]
ix = {h: i for i, h in enumerate(hdr)}
tot = 0; byop = collections.Counter(); samp = collections.Counter(); stalls = collections.Counter()
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in data:
    if len(r) < len(hdr) or r[0] == "Address": continue
    parts = r[ix["Source"]].split()
    op = parts[0] if not parts[0].startswith("@") else parts[1]
    op = op.split(".")[0] if not op.startswith("F2F") else op
    n = int(r[ix["Instructions Executed"]]); s = int(r[ix["# Samples"]])
    tot += n; byop[op] += n; samp[op] += s
    for c in stall_cols: stalls[c] += int(r[ix[c]])
print("SASS instructions in kernel:", len(data), " dynamic warp-instr:", tot)
for op, n in byop.most_common(22):
    print(f"  {op:14s} {n / tot * 100:6.2f}%  samples {samp[op] / max(1, sum(samp.values())) * 100:6.2f}%")
print("stalls:", {k: v for k, v in stalls.most_common(8)})
