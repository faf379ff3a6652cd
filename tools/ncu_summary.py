"""Condense ncu CSV exports into the small tables committed under profiles/.

    python tools/ncu_summary.py raw   gpurun_out/step_r01_raw.csv   profiles/r01_duet_cfg2_step_full.csv
    python tools/ncu_summary.py list  gpurun_out/launches_r01.csv   profiles/r01_duet_cfg2_step_launches.csv

``raw``  : one row per profiled launch of an `ncu --set full` report (`--page raw --csv`), selected metrics only.
``list`` : one row per launch of a `--metrics gpu__time_duration.sum` run, plus a per-kernel share table on stdout.
"""
import collections
import csv
import sys

RAW_COLS = [
    ('Kernel Name', 'kernel'), ('launch__grid_size', 'grid'), ('launch__block_size', 'block'),
    ('launch__registers_per_thread', 'regs'), ('launch__shared_mem_per_block_dynamic', 'dyn_smem_B'),
    ('gpu__time_duration.sum', 'duration_ns'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor_pipe_pct_active'),
    ('TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 'tensor_pipe_pct_elapsed'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm_throughput_pct'),
    ('gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed', 'mem_throughput_pct'),
    ('dram__bytes_read.sum', 'dram_read_B'), ('dram__bytes_write.sum', 'dram_write_B'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram_pct'),
    ('lts__t_sector_hit_rate.pct', 'l2_hit_pct'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved_occupancy_pct'),
    ('sm__cycles_elapsed.avg.per_second', 'sm_hz'),
]


SCALE = {'us': 1e3, 'usecond': 1e3, 'ms': 1e6, 'msecond': 1e6, 'ns': 1, 'nsecond': 1, 'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6,
         'Gbyte': 1e9, 'Ghz': 1e9, 'Mhz': 1e6, 'hz': 1}


def short(name):
    name = name.replace('void ', '').replace('<unnamed>::', '')
    return name.split('(')[0]


def raw(src, dst):
    rows = list(csv.reader(open(src)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    idx = [(hdr.index(c), out) for c, out in RAW_COLS if c in hdr]
    with open(dst, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow([o for _, o in idx])
        for r in body:
            vals = []
            for i, o in idx:
                v = r[i].replace(',', '')
                u = units[i]
                if o == 'kernel':
                    v = short(r[i])
                elif u in SCALE and v:
                    v = '%.0f' % (float(v) * SCALE[u])
                vals.append(v)
            w.writerow(vals)
    print('wrote', dst, len(body), 'launches; duration unit in the export:', units[hdr.index('gpu__time_duration.sum')])


def launch_list(src, dst):
    rows = list(csv.reader(open(src)))
    hdr = None
    out = []
    for r in rows:
        if len(r) > 10 and r[0] == 'ID':
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            out.append((short(d['Kernel Name']), d['Block Size'], d['Grid Size'], float(d['Metric Value'].replace(',', '')), d['Metric Unit']))
    with open(dst, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['kernel', 'block', 'grid', 'duration', 'unit'])
        for o in out:
            w.writerow(o)
    cnt, tot = collections.Counter(), collections.Counter()
    for k, _, _, t, _ in out:
        cnt[k] += 1
        tot[k] += t
    total = sum(tot.values())
    print('%-64s %6s %12s %7s' % ('kernel', 'n', 'sum', 'share'))
    for k, v in tot.most_common():
        print('%-64s %6d %12.1f %6.1f%%' % (k[:64], cnt[k], v, 100 * v / total))
    print('total', total, out[0][4] if out else '')


if __name__ == '__main__':
    {'raw': raw, 'list': launch_list}[sys.argv[1]](sys.argv[2], sys.argv[3])
