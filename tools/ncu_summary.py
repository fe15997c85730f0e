import csv, sys
raw, src = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","lts__t_sectors.sum","l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum","l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum","l1tex__t_sector_hit_rate.pct","lts__t_sector_hit_rate.pct",
"sm__warps_active.avg.pct_of_peak_sustained_active","launch__registers_per_thread","smsp__inst_executed.sum","sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
"smsp__issue_active.avg.pct_of_peak_sustained_active","lts__throughput.avg.pct_of_peak_sustained_elapsed","l1tex__throughput.avg.pct_of_peak_sustained_elapsed","sm__cycles_elapsed.max","l1tex__data_pipe_lsu_wavefronts.sum","l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum","l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
for w in want:
    if w in hdr:
        i = hdr.index(w); print(w, units[i], [r[i] for r in data])
for i,h in enumerate(hdr):
    if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
        v=float(data[0][i])
        if v>0.15: print("  stall", h.replace("smsp__average_warps_issue_stalled_","").replace("_per_issue_active.ratio",""), round(v,2))
rows = list(csv.reader(open(src)))
hdr = rows[1]
i_src, i_s, i_ex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
cols = {k: hdr.index(k) for k in ("stall_long_sb","stall_short_sb","stall_wait","stall_barrier","stall_lg","stall_mio","stall_math")}
ins = []
for r in rows[2:]:
    try: ins.append((int(r[i_s]), r[i_src].strip(), int(r[i_ex]), {k:int(r[c]) for k,c in cols.items()}))
    except: break
tot = sum(x[0] for x in ins); print("n instr", len(ins), "samples", tot)
W=int(sys.argv[3]) if len(sys.argv)>3 else 300
for k in range(0, len(ins), W):
    w = ins[k:k+W]; s = sum(x[0] for x in w)
    if s < tot*0.004: continue
    ops = {}
    for x in w:
        t = x[1].split(); op = t[1] if t[0].startswith("@") else t[0]; op = op.split(".")[0]
        ops[op] = ops.get(op,0)+x[2]
    top = sorted(ops.items(), key=lambda t:-t[1])[:5]
    st = {kk: sum(x[3][kk] for x in w) for kk in cols}
    print(f"{k:5d} {100*s/tot:5.1f}% exec {sum(x[2] for x in w):9d}", {kk.replace('stall_',''):v for kk,v in st.items() if v>s*0.08}, [t[0] for t in top])
