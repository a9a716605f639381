"""tools/ncu_summary.py — summarise an .ncu-rep (read offline, no GPU): headline metrics, stall mix,
instruction mix per input sample and the hottest source lines. Usage:
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [samples_per_launch] > profiles/<name>.txt"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
nsamp = float(sys.argv[2]) if len(sys.argv) > 2 else float(1 << 28)


def ncu(*args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(ncu("--page", "raw", "--csv"))))
hdr, units, vals = raw[0], raw[1], raw[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg"]
print("kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w:70s} {vals[i]:>16s} {units[i]}")
print()
for i, h in enumerate(hdr):
    if h.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in h:
        print(f"{h:70s} {vals[i]:>16s}")
print()

rows = list(csv.reader(io.StringIO(ncu("--page", "source", "--csv", "--print-source", "cuda,sass"))))
sections, cur, curfile = [], None, "?"
for r in rows:
    if r and r[0] == "File Path":
        curfile = r[1]
    elif r and r[0] == "Line No":
        cur = []
        sections.append((curfile, r, cur))
    elif cur is not None and r and r[0] != "Function Name":
        cur.append(r)
seen, ops, tot = set(), collections.Counter(), 0
lines = collections.Counter()
lstall = collections.defaultdict(collections.Counter)
for f, h, data in sections:
    ai, ii, si, li = h.index("Address"), h.index("Instructions Executed"), h.index("# Samples"), h.index("Line No")
    srci = [i for i, x in enumerate(h) if x == "Source"]
    stallcols = {x: i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x}
    for r in data:
        if r[ai]:
            if r[ai] in seen:
                continue
            seen.add(r[ai])
            try:
                n = int(r[ii])
            except ValueError:
                continue
            tot += n
            toks = r[srci[1]].split()
            op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
            ops[op.split(".")[0]] += n
        elif r[li]:
            try:
                s = int(r[si])
            except ValueError:
                continue
            if s:
                key = (f.split("/")[-1], r[li], r[srci[0]].strip()[:80])
                lines[key] += s
                for x, i in stallcols.items():
                    try:
                        lstall[key][x] += int(r[i])
                    except ValueError:
                        pass
print(f"warp instructions executed: {tot}   thread-instructions per input sample: {tot * 32 / nsamp:.2f}")
for k, v in ops.most_common(24):
    print(f"  {k:10s} {v:12d}  {v * 32 / nsamp:7.2f} /sample")
print("\nhottest source lines (stall samples, top reasons):")
for k, v in lines.most_common(22):
    print(f"  {v:7d} {k[0]}:{k[1]:>4s} {k[2]:80s} {lstall[k].most_common(3)}")
