"""Developer aid: prints the headline metrics, stall breakdown and dynamic opcode mix of an .ncu-rep (first kernel)."""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
per = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0   # normalise instruction counts (e.g. number of samples)
det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
keys = ("Duration", "Registers Per", "Executed Ipc Active", "Issue Slots Busy", "L1/TEX Hit", "L2 Hit", "Warp Cycles Per Issued",
        "No Eligible", "Eligible Warps", "Active Warps Per", "Achieved Occupancy", "Dynamic Shared", "Grid Size", "Block Size",
        "DRAM Throughput", "Mem Busy", "Max Bandwidth", "L2 Cache Throughput", "Theoretical Occ")
for ln in det.splitlines():
    if any(k in ln for k in keys):
        print(ln.rstrip())
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, v = rows[0], rows[2]
m = dict(zip(h, v))
for k in ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum", "smsp__inst_executed.sum",
          "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
          "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.max", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum"):
    if k in m:
        print(f"{k:70s} {m[k]}")
st = sorted(((float(val), k) for k, val in m.items() if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio") and val), reverse=True)
print("stalls (cycles per issued instruction):", ", ".join(f"{k.split('stalled_')[1].split('_per_issue')[0]} {a:.2f}" for a, k in st[:9]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
iS, iE = hdr.index("Source"), hdr.index("Instructions Executed")
cnt = collections.Counter()
for r in rows[2:]:
    if len(r) <= iE:
        continue
    toks = r[iS].split()
    op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
    cnt[op] += int(r[iE])
tot = sum(cnt.values())
print(f"warp instructions: {tot} ({tot / per:.1f} per unit)")
print("  ".join(f"{op} {100 * c / tot:.1f}%" for op, c in cnt.most_common(16)))
