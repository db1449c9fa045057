"""Turns a `.ncu-rep` into the markdown summary kept under profiles/.

    python bench/ncu_summary.py gpurun_out/prof_allreduce.ncu-rep profiles/ncu_allreduce_kernel.md

Reads the report with `ncu -i <rep> --page raw --csv` (works without a GPU), keeps the roofline-relevant metrics of
B200_PROFILING.md (duration, DRAM bytes and %-of-peak, SM / L2 throughput, occupancy and its limiter, registers, warp
stall break-down) and lists the hottest source lines from the source page when the report carries source.
"""
import csv
import io
import subprocess
import sys

KEEP = [
    ('gpu__time_duration.sum', 'duration'),
    ('launch__grid_size', 'grid'), ('launch__block_size', 'block'), ('launch__registers_per_thread', 'registers / thread'),
    ('launch__shared_mem_per_block_static', 'static smem / block'), ('launch__waves_per_multiprocessor', 'waves / SM'),
    ('launch__occupancy_limit_registers', 'occupancy limit: registers (blocks)'),
    ('launch__occupancy_limit_warps', 'occupancy limit: warps (blocks)'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy'),
    ('dram__bytes_read.sum', 'DRAM read'), ('dram__bytes_write.sum', 'DRAM written'),
    ('dram__bytes.sum.per_second', 'DRAM bandwidth'),
    ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput, % of peak'),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 throughput, % of peak'),
    ('l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'L1/TEX throughput, % of peak'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput, % of peak'),
    ('sm__inst_executed.sum', 'instructions executed'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy'),
    ('smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'stall: long scoreboard (cycles / issue)'),
    ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'stall: long scoreboard (warps / issue)'),
    ('smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'stall: LG throttle (warps / issue)'),
    ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'stall: barrier (warps / issue)'),
    ('smsp__average_warps_issue_stalled_membar_per_issue_active.ratio', 'stall: membar (warps / issue)'),
    ('smsp__average_warps_issue_stalled_drain_per_issue_active.ratio', 'stall: drain (warps / issue)'),
    ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'stall: wait (warps / issue)'),
    ('l1tex__t_sector_hit_rate.pct', 'L1 hit rate'), ('lts__t_sector_hit_rate.pct', 'L2 hit rate'),
]


def ncu(rep, page):
    r = subprocess.run(['ncu', '-i', rep, '--page', page, '--csv'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    return list(csv.reader(io.StringIO(r.stdout)))


def main():
    rep, out = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else rep
    rows = ncu(rep, 'raw')
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    lines = [f'# ncu summary: {title}', '',
             f'Source report: `{rep}` (`ncu --set full --clock-control none --import-source on`, one GPU, under `gpurun`). '
             'Numbers in this file come from a profiler replay: they explain the kernel, they are never bench values.', '']
    for k, r in enumerate(data):
        name = r[col['Kernel Name']]
        lines += [f'## launch {k}: `{name}`', '', '| metric | value |', '|---|---|']
        for m, label in KEEP:
            if m in col and r[col[m]] != '':
                v = r[col[m]]
                try:
                    v = f'{float(v):,.2f}'.rstrip('0').rstrip('.')
                except ValueError:
                    pass
                lines.append(f'| {label} (`{m}`) | {v} {units[col[m]]} |')
        lines.append('')
    src = ncu(rep, 'source')
    start = next((i for i, r in enumerate(src) if r and r[0] == 'Address'), None)
    if start is not None:
        end = next((i for i in range(start + 1, len(src)) if src[i] and src[i][0] == 'Kernel Name'), len(src))
        src = src[start:end]
        h = {x: i for i, x in enumerate(src[0])}
        keycol = next((c for c in ('Warp Stall Sampling (All Samples)', 'Warp Stall Sampling (All Cycles)', '# Samples') if c in h), None)
        srccol = next((c for c in ('Source', 'SASS', 'Instruction') if c in h), None)
        if keycol and srccol:
            def val(r):
                try:
                    return float(r[h[keycol]])
                except (ValueError, IndexError):
                    return 0.0
            top = sorted(src[1:], key=val, reverse=True)[:14]
            total = sum(val(r) for r in src[1:]) or 1.0
            lines += ['## hottest lines (first launch in the report, by warp-stall samples)', '', f'| {keycol} | share | {srccol} |', '|---|---|---|']
            for r in top:
                if val(r) > 0:
                    lines.append(f'| {val(r):.0f} | {100 * val(r) / total:.1f}% | `{r[h[srccol]].strip()[:150]}` |')
            lines.append('')
    open(out, 'w').write('\n'.join(lines) + '\n')
    print('wrote', out, len(data), 'launches')


if __name__ == '__main__':
    main()
