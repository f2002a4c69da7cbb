#!/usr/bin/env python
"""Condenses `ncu --page raw/source --csv` dumps into the short text summaries kept under profiles/.
usage: ncu_summary.py raw.csv src_sass.csv [src_cuda_sass.csv]"""
import collections, csv, re, sys

def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    keep = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__block_size',
            'launch__grid_size', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
            'launch__occupancy_limit_registers', 'launch__shared_mem_per_block_dynamic',
            'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
            'smsp__issue_active.avg.pct_of_peak_sustained_active',
            'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
            'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
            'smsp__average_warp_latency_per_inst_issued.ratio', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
            'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'lts__t_bytes.sum',
            'l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum']
    for h, u, v in zip(hdr, units, vals):
        if h in keep or (h.startswith('smsp__average_warps_issue_stalled') and float(v or 0) > 0.02):
            print(f"{h} [{u}] = {v}")

def sass(path, nwarps=None):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]; data = rows[2:]
    iS = hdr.index('Source'); iE = hdr.index('Instructions Executed')
    ops = collections.Counter(); tot = 0
    for r in data:
        try: n = int(r[iE])
        except Exception: continue
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[iS])
        op = (m.group(2) if m else r[iS][:10]).split('.')[0]
        ops[op] += n; tot += n
    print(f"warp-level instructions executed: {tot}   static SASS instructions: {len(data)}")
    for op, n in ops.most_common(22):
        print(f"  {op:8s} {n:12d} {100*n/tot:5.1f}%")

def cuda(path, top=40):
    rows = list(csv.reader(open(path)))
    cur = None; out = []; tot = 0
    for r in rows:
        if r and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
        if len(r) < 8 or r[0] in ('Line No', 'Function Name', ''): continue
        try: ln = int(r[0]); n = int(r[7]); s = int(r[6])
        except Exception: continue
        out.append((n, s, cur, ln, r[1].strip()[:100])); tot += n
    out.sort(reverse=True)
    print("top source lines by executed warp instructions (inlined lines are attributed to every level):")
    for n, s, f, ln, src in out[:top]:
        print(f"  {100*n/tot:5.2f}%  samples {s:6d}  {f}:{ln}  {src}")

if __name__ == "__main__":
    raw(sys.argv[1]); sass(sys.argv[2])
    if len(sys.argv) > 3: cuda(sys.argv[3])
