"""Summarise an .ncu-rep (read on the CPU box): key raw metrics per kernel + top source lines by stall samples."""
import collections, csv, subprocess, sys

KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__waves_per_multiprocessor', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum']


def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index('Kernel Name')
    for r in rows[2:]:
        print('\n##', r[ki][:80])
        for i, h in enumerate(hdr):
            if h in KEEP:
                print('  %-66s %-14s %s' % (h, units[i], r[i]))


def source(rep, top=22):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur_file, cur_fn = None, None
    agg = collections.OrderedDict()
    for r in rows:
        if len(r) >= 2 and r[0] == 'File Path':
            cur_file = r[1].split('/')[-1]; continue
        if len(r) >= 2 and r[0] == 'Function Name':
            cur_fn = r[1].split('(')[0]; continue
        if len(r) < 9 or r[0] in ('Line No', ''):
            continue
        try:
            ln = int(r[0]); inst = int(r[7]); smp = int(r[6]); tinst = int(r[8])
        except ValueError:
            continue
        a = agg.setdefault((cur_fn, cur_file, ln, r[1][:86]), [0, 0, 0])
        a[0] += inst; a[1] += smp; a[2] += tinst
    fns = collections.OrderedDict()
    for k, v in agg.items():
        fns.setdefault(k[0], []).append((k, v))
    for fn, items in fns.items():
        tot = sum(v[0] for _, v in items) or 1; tots = sum(v[1] for _, v in items) or 1
        print('\n## %s: %d warp-inst, %d samples' % (fn, tot, tots))
        for k, v in sorted(items, key=lambda kv: -kv[1][1])[:top]:
            print('  %5.1f%% smp %5.1f%% inst thr/inst %4.1f  %s:%d  %s' % (100 * v[1] / tots, 100 * v[0] / tot, v[2] / max(v[0], 1), k[1], k[2], k[3]))
        print('  -- by executed instructions --')
        for k, v in sorted(items, key=lambda kv: -kv[1][0])[:top]:
            print('  %5.1f%% inst %5.1f%% smp thr/inst %4.1f  %s:%d  %s' % (100 * v[0] / tot, 100 * v[1] / tots, v[2] / max(v[0], 1), k[1], k[2], k[3]))


if __name__ == '__main__':
    raw(sys.argv[1])
    if len(sys.argv) < 3 or sys.argv[2] != '--raw-only':
        source(sys.argv[1])
