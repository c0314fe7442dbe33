"""Summarise an ncu report: key metrics per launch, then per kernel the executed-instruction histogram by opcode and
the SASS lines with the most stall samples.  usage: ncu_top.py rep [n]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg"]
for vals in rows[2:]:
    print("=" * 100)
    for h, u, v in zip(hdr, units, vals):
        if h in want:
            print(f"  {h} [{u}] = {v[:110]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
kernels = []
cur = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        kernels.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None and len(r) >= len(cur["hdr"]) - 2:
        cur["data"].append(r)
for k in kernels:
    h = k["hdr"]
    isrc, iss, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    data = k["data"]
    tot = sum(int(r[iss] or 0) for r in data)
    totx = sum(int(r[iex] or 0) for r in data)
    print("=" * 100)
    print(k["name"][:120])
    print("total samples", tot, "warp instructions", totx)
    ops = collections.Counter()
    for r in data:
        toks = r[isrc].split()
        op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
        ops[op.split(".")[0]] += int(r[iex] or 0)
    print("opcode histogram:", ", ".join(f"{o}={100 * c / max(totx, 1):.1f}%" for o, c in ops.most_common(22)))
    for r in sorted(data, key=lambda r: -int(r[iss] or 0))[:n]:
        print(f"{int(r[iss] or 0):7d} {100 * int(r[iss] or 0) / max(tot, 1):5.1f}% exec={r[iex]:>9}  {r[isrc].strip()[:100]}")
