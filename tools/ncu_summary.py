#!/usr/bin/env python
"""Development helper: condense an `ncu --set full` report of surface_warp_kernel into the handful of numbers
DESIGN.md / profiles/ quote.   python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv, io, subprocess, sys, collections

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "launch__grid_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_write.sum.per_second", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tex.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


def main(rep):
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    hdr, units = rows[0], rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("kernel:", r[ci["Kernel Name"]], "launch id", r[ci["ID"]])
        for k in KEYS:
            if k in ci:
                print(f"  {k:75s} {r[ci[k]]:>16s} {units[ci[k]]}")
        st = sorted(((float(r[i] or 0), h) for h, i in ci.items() if "issue_stalled" in h and h.endswith("per_issue_active.ratio")), reverse=True)
        print("  warp-cycles per issued instruction by stall reason:", ", ".join(f"{h.split('issue_stalled_')[1].split('_per_issue')[0]} {v:.2f}" for v, h in st[:8]))
    # where the samples fall: loops of the first kernel
    src = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--print-source", "sass"))))
    hdr = src[1]
    ci = {h: i for i, h in enumerate(hdr)}
    data = []
    for r in src[2:]:
        if len(r) < len(hdr) // 2:
            break
        data.append(r)
    A, S, N, IE = ci["Address"], ci["Source"], ci["# Samples"], ci["Instructions Executed"]
    base = int(data[0][A], 16)
    tot_s = sum(int(r[N] or 0) for r in data); tot_i = sum(int(r[IE] or 0) for r in data)
    print(f"source page (first launch): {len(data)} SASS instructions, {tot_s} samples, {tot_i} warp-instructions executed")
    for r in data:
        s = r[S].strip()
        if "BRA" in s:
            try:
                t = int(s.split()[-1].rstrip(";"), 16)
            except ValueError:
                continue
            t -= base if t >= base else 0
            off = int(r[A], 16) - base
            if t < off:
                sel = [q for q in data if t <= int(q[A], 16) - base <= off]
                ss = sum(int(q[N] or 0) for q in sel); ii = sum(int(q[IE] or 0) for q in sel)
                ops = collections.Counter(q[S].split()[0] if not q[S].strip().startswith("@") else q[S].split()[1] for q in sel)
                print(f"  loop [{t:#x},{off:#x}] {len(sel)} instr, {int(r[IE] or 0)} trips: {100*ss/tot_s:.1f}% of samples, {100*ii/tot_i:.1f}% of executed instructions; "
                      + " ".join(f"{k}x{v}" for k, v in ops.most_common(8)))


if __name__ == "__main__":
    main(sys.argv[1])
