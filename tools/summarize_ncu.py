#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.
    python tools/summarize_ncu.py launches gpurun_out/launches.csv profiles/r01_launches.md "<command>"
    python tools/summarize_ncu.py full gpurun_out/prof.ncu-rep profiles/r01_fan_lse_tc.md
"""
import collections, csv, subprocess, sys


def launches(src, dst, cmd):
    lines = open(src).read().splitlines()
    i = [k for k, l in enumerate(lines) if l.startswith('"ID"')][0]
    rows = list(csv.DictReader(lines[i:]))
    agg = collections.OrderedDict()
    for r in rows:
        k = r['Kernel Name'].split('(')[0][:90] + " grid" + r['Grid Size'] + " block" + r['Block Size']
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r['Metric Value']) / 1e3
    tot = sum(v[1] for v in agg.values())
    with open(dst, 'w') as f:
        f.write(f"# ncu launch list (gpu__time_duration.sum, --clock-control none)\n\ncommand: `{cmd}`\n\n"
                f"{len(rows)} launches, {tot:.1f} us in total (cold-cache, serialised: compare SHARES, not absolutes)\n\n"
                "| us total | launches | share | kernel |\n|---:|---:|---:|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {v[1]:.1f} | {v[0]} | {100 * v[1] / tot:.1f}% | `{k}` |\n")
    print(open(dst).read())


KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_xu.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'sm__cycles_elapsed.avg',
        'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct', 'smsp__warp_issue_stalled_wait_per_warp_active.pct',
        'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct', 'smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.avg.per_second',
        'smsp__warp_issue_stalled_barrier_per_warp_active.pct', 'smsp__warp_issue_stalled_tex_throttle_per_warp_active.pct']


def full(src, dst):
    # a .ncu-rep, or the `ncu -i ... --page raw --csv` export made on the GPU box (reports over 64 MiB do not travel)
    raw = open(src).read() if src.endswith('.csv') else \
        subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, 'w') as f:
        f.write(f"# ncu --set full summary of `{src}`\n\n")
        for r in rows[2:]:
            f.write(f"## {r[hdr.index('Kernel Name')]}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                cand = [i for i, h in enumerate(hdr) if h == k]
                if cand:
                    f.write(f"| {k} | {r[cand[0]]} | {units[cand[0]]} |\n")
            f.write("\n")
    print(open(dst).read())


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else '')
    else:
        full(sys.argv[2], sys.argv[3])
