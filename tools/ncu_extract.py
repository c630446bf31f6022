"""Summarise an .ncu-rep (read on the CPU box with `ncu -i`) into a small markdown file for profiles/.
usage: python tools/ncu_extract.py gpurun_out/X.ncu-rep profiles/X.md [cells_per_launch]"""
import collections
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
cells = float(sys.argv[3]) if len(sys.argv) > 3 else 8192.0 * 8192.0
METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2 % of peak"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe % active"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots % active"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_shared_mem", "CTAs/SM (smem limit)"),
    ("launch__occupancy_limit_registers", "CTAs/SM (reg limit)"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts"),
    ("sm__cycles_elapsed.avg.per_second", "SM clock"),
]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
lines = [f"# {rep.split('/')[-1]} — ncu --set full --clock-control none\n"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    lines.append(f"## {d['Kernel Name'][:110]}\n")
    lines.append("| metric | value | unit |\n|---|---|---|")
    for m, label in METRICS:
        if m in d:
            lines.append(f"| {label} (`{m}`) | {d[m]} | {units[hdr.index(m)]} |")
    try:
        tr = float(d["dram__bytes_read.sum"]) + float(d["dram__bytes_write.sum"])
        u = units[hdr.index("dram__bytes_read.sum")]
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u]
        lines.append(f"| **traffic = dram read + write per launch** | {tr * scale / 1e9:.4f} | GB ({tr * scale / cells:.1f} B/cell) |")
    except Exception:
        pass
    lines.append("")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
idx = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
names = [rows[i - 1][1] if i > 0 and len(rows[i - 1]) > 1 else "?" for i in idx]
for n, i0 in enumerate(idx):
    h = rows[i0]
    end = idx[n + 1] - 1 if n + 1 < len(idx) else len(rows)
    ia, ie = h.index("Source"), h.index("Instructions Executed")
    stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    agg, st, tot, nstat = collections.Counter(), collections.Counter(), 0, 0
    for r in rows[i0 + 1:end]:
        if len(r) <= ie:
            continue
        toks = r[ia].split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        op = op.split(".")[0].rstrip(";")
        c = int(r[ie])
        agg[op] += c
        tot += c
        nstat += 1
        for i, cname in stall_cols:
            try:
                st[cname] += int(r[i])
            except ValueError:
                pass
    lines.append(f"## SASS mix: {names[n][:100]}\n")
    lines.append(f"static instructions {nstat}; executed warp instructions {tot} = {tot * 32 / cells:.1f} thread-instr/cell\n")
    lines.append("| opcode | per cell | share |\n|---|---|---|")
    for op, c in agg.most_common(14):
        lines.append(f"| {op} | {c * 32 / cells:.1f} | {100 * c / tot:.1f}% |")
    ssum = sum(st.values()) or 1
    lines.append("\n| stall reason (warp samples) | share |\n|---|---|")
    for cname, v in st.most_common(10):
        lines.append(f"| {cname} | {100 * v / ssum:.1f}% |")
    lines.append("")
open(out, "w").write("\n".join(lines) + "\n")
print("wrote", out)
