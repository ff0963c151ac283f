#!/usr/bin/env python3
"""Turns the raw ncu outputs of a gpurun call (gpurun_out/) into the summaries committed under profiles/.
Usage: python tools/profile_summaries.py [--round r01]
  gpurun_out/<round>_pipe_full.ncu-rep  -> profiles/<round>_ncu_pipe_kernel_raw.csv, _summary.txt
  gpurun_out/<round>_launches.csv       -> profiles/<round>_launches.txt
"""
import argparse
import collections
import csv
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sass__inst_executed_local_loads',
        'sass__inst_executed_local_stores', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_elapsed.max',
        'smsp__average_warp_latency_per_inst_issued.ratio'] + [
    'smsp__average_warps_issue_stalled_%s_per_issue_active.ratio' % k for k in
    ('barrier', 'wait', 'short_scoreboard', 'long_scoreboard', 'not_selected', 'math_pipe_throttle', 'mio_throttle',
     'no_instruction', 'branch_resolving', 'dispatch_stall')]


def full(rnd):
    rep = os.path.join(ROOT, "gpurun_out", rnd + "_pipe_full.ncu-rep")
    raw = os.path.join(ROOT, "profiles", rnd + "_ncu_pipe_kernel_raw.csv")
    with open(raw, "w") as f:
        subprocess.check_call(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=f, stderr=subprocess.DEVNULL)
    rows = list(csv.reader(open(raw)))
    hdr, units, d = rows[0], rows[1], rows[2]
    out = ['# ncu --set full --clock-control none --import-source on -k regex:hsl_gait_pipe_kernel -s 4 -c 1  python bench.py --steps 3 --warmup 3 --no-cpu',
           '# default cost-only kernel of the six-limbed models (hexapod, 4096 candidates x 256 frames = 1 048 576 frame solves per launch; persistent, one block per SM)',
           '# full metric table: profiles/%s_ncu_pipe_kernel_raw.csv' % rnd, '']
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            out.append('%-95s %-16s %s' % (w, units[i], d[i]))
    open(os.path.join(ROOT, "profiles", rnd + "_ncu_pipe_kernel_summary.txt"), "w").write("\n".join(out) + "\n")
    print("\n".join(out))


def launches(rnd):
    rows = [r for r in csv.reader(open(os.path.join(ROOT, "gpurun_out", rnd + "_launches.csv"))) if len(r) > 5]
    hdr, data = rows[0], rows[1:]
    ik, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in data:
        try:
            v = float(r[iv].replace(',', ''))
        except ValueError:
            continue
        u = r[iu]
        us = v / 1000 if u in ('ns', 'nsecond') else v * 1000 if u in ('ms', 'msecond') else v
        a = agg.setdefault(re.sub(r'\(.*', '', r[ik])[:74], [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    lines = ['# ncu launch list (final default kernel): `ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv python bench.py --steps 5 --warmup 3 --no-cpu`',
             '# per-launch device times are cold-cache and serialised: compare SHARES, not absolutes.', '',
             '%-76s %5s %12s %8s %10s' % ('kernel', 'n', 'total_us', 'share', 'avg_us')]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append('%-76s %5d %12.1f %8.3f %10.1f' % (k, n, t, t / tot, t / n))
    lib = [(k, v) for k, v in agg.items() if 'hsl_' in k and 'probe' not in k]
    lt = sum(v[1] for _, v in lib)
    lines += ['', '# library kernels only (a step = setup + frames + finish + selection):']
    for k, (n, t) in sorted(lib, key=lambda kv: -kv[1][1]):
        lines.append('%-76s %5d %12.1f %8.3f %10.1f' % (k, n, t, t / lt, t / n))
    open(os.path.join(ROOT, "profiles", rnd + "_launches.txt"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[-7:]))


ENTRY_WANT = WANT[:8] + ['sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
                         'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
                         'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
                         'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
                         'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores', 'smsp__average_warp_latency_per_inst_issued.ratio',
                         'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
                         'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio']


def entries(rnd, stem, header, keep=None):
    """gpurun_out/<round>_<stem>.ncu-rep -> profiles/<round>_ncu_<stem>_summary.txt: the first capture of every distinct kernel."""
    rep = os.path.join(ROOT, "gpurun_out", "%s_%s.ncu-rep" % (rnd, stem))
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ik = hdr.index('Kernel Name')
    seen, out = set(), ['# ' + h for h in header] + ['']
    for r in rows[2:]:
        name = re.sub(r'\(.*', '', r[ik])
        key = (name, r[hdr.index('launch__grid_size')])
        if key in seen or (keep and not any(k in name for k in keep)):
            continue
        seen.add(key)
        for w in ENTRY_WANT:
            if w in hdr:
                i = hdr.index(w)
                out.append('%-95s %-16s %s' % (w, units[i], r[i]))
        out.append('')
    open(os.path.join(ROOT, "profiles", "%s_ncu_%s_summary.txt" % (rnd, stem)), "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--round", default="r01")
    ap.add_argument("--entries", default="", help="stem of an extra capture (gpurun_out/<round>_<stem>.ncu-rep) to summarise per kernel")
    a = ap.parse_args()
    if a.entries:
        entries(a.round, a.entries, ["ncu --set full --clock-control none --import-source on ... python tools/profile_entries.py (see tools/profile_r02.sh)"])
    else:
        full(a.round)
        launches(a.round)
