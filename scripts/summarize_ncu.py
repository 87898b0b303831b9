"""Developer tool: compact, committed summary of an .ncu-rep (key metrics, stall reasons, hottest source lines).
usage: summarize_ncu.py <report.ncu-rep> <out.txt>"""
import csv, io, subprocess, sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__shared_mem_per_block',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__cycles_active.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main(rep, out):
    lines = []
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
        lines.append(f"=== {d['Kernel Name'][1]}  grid {d['Grid Size'][1]} block {d['Block Size'][1]}")
        for k in KEYS:
            if k in d and d[k][1] != "":
                lines.append(f"  {k:72s} {d[k][1]:>16s} {d[k][0]}")
        st = [(h, float(v[1])) for h, v in d.items() if 'smsp__average_warps_issue_stalled' in h and h.endswith('_per_issue_active.ratio') and v[1]]
        lines.append("  stall reasons (warps per issue-active cycle): " + ", ".join(
            f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}={v:.2f}" for h, v in sorted(st, key=lambda x: -x[1])[:8]))
    src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"]))))
    cur, hdr2, agg, kernel = None, None, {}, None
    for r in src:
        if not r:
            continue
        if r[0] == 'File Path':
            cur = r[1].split('/')[-1]; continue
        if r[0] == 'Function Name':
            kernel = r[1]; continue
        if r[0] == 'Line No':
            hdr2 = r; iex = hdr2.index('Instructions Executed'); continue
        if hdr2 and len(r) == len(hdr2):
            try:
                ln, ex = int(r[0]), int(r[iex] or 0)
            except ValueError:
                continue
            key = (kernel, cur, ln)
            agg[key] = (agg.get(key, (0, ''))[0] + ex, r[1])
    by_kernel = {}
    for (kn, f, l), (e, s) in agg.items():
        by_kernel.setdefault(kn, []).append((e, f, l, s))
    for kn, items in by_kernel.items():
        tot = sum(e for e, *_ in items)
        if tot == 0:
            continue
        lines.append(f"--- hottest source lines of {kn[:80]} (share of executed warp instructions)")
        for e, f, l, s in sorted(items, reverse=True)[:25]:
            lines.append(f"  {100 * e / tot:5.2f}%  {f}:{l:<4d} {s.strip()[:110]}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:60]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
