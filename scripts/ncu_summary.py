"""Summarise an .ncu-rep (raw page + SASS hot spots) into text.  Usage: ncu_summary.py rep [units_per_launch]"""
import csv, io, subprocess, sys
from collections import Counter
rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, unit = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__t_bytes.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "local_load", "local_store", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum"]
for r in rows[2:]:
    print("KERNEL", r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "")
    for h, u, v in zip(hdr, unit, r):
        if any(h == k or (k in h and k in ("local_load", "local_store")) for k in KEYS) or "average_warps_issue_stalled" in h and "per_issue_active" in h:
            print(f"  {h:80s} {u:14s} {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
allrows = list(csv.reader(io.StringIO(src)))
sections, cur_sec = [], None
for r in allrows:
    if r and r[0] == "Kernel Name":
        cur_sec = {"name": r[1] if len(r) > 1 else "", "hdr": None, "rows": []}; sections.append(cur_sec)
    elif cur_sec is not None and cur_sec["hdr"] is None:
        cur_sec["hdr"] = r
    elif cur_sec is not None and len(r) == len(cur_sec["hdr"]):
        cur_sec["rows"].append(r)
for sec in sections:
    hdr, data = sec["hdr"], sec["rows"]
    ia, isrc, isamp, ithr = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Avg. Threads Executed")
    tot = sum(int(r[ia]) for r in data); stot = sum(int(r[isamp]) for r in data)
    print(f"SASS of {sec['name'][:60]}: {len(data)} instructions, {tot} warp-inst executed = {tot/units:.1f} per unit, {stot} samples")
    blocks, cur = [], None
    for n, r in enumerate(data):
        c = int(r[ia]); op = r[isrc].split()
        op = op[1] if op and op[0].startswith("@") and len(op) > 1 else (op[0] if op else "")
        if cur and abs(c - cur["c"]) <= 0.02 * max(c, cur["c"]) + 1:
            cur["n1"] = n; cur["tot"] += c; cur["samp"] += int(r[isamp]); cur["ops"].append(op); cur["thr"] += float(r[ithr]) * c
        else:
            cur = dict(n0=n, n1=n, c=c, tot=c, samp=int(r[isamp]), thr=float(r[ithr]) * c, ops=[op]); blocks.append(cur)
    print("  [sass range]  exec/unit  ninst  share_inst share_samples avg_threads  top opcodes")
    for b in blocks:
        if b["tot"] > 0.004 * tot or b["samp"] > 0.004 * max(stot, 1):
            print(f"  [{b['n0']:4d}-{b['n1']:4d}] {b['c']/units:9.3f} {b['n1']-b['n0']+1:5d} {100*b['tot']/max(tot,1):8.1f}% {100*b['samp']/max(stot,1):8.1f}% {b['thr']/max(b['tot'],1):8.1f}   {Counter(b['ops']).most_common(6)}")
