"""Turns ncu outputs in gpurun_out/ into the small tracked summaries under profiles/.

    python tools/summarise_ncu.py launches gpurun_out/launches_r01.csv profiles/r01_launches.md
    python tools/summarise_ncu.py raw      gpurun_out/prof_r01.ncu-rep  profiles/r01_kernels.md
"""
import collections
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sector_hit_rate.pct']


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith('==')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v = float(row['Metric Value'].replace(',', ''))
        u = row['Metric Unit']
        v = v / 1e3 if u.startswith('n') else v * 1e3 if u.startswith('m') else v
        agg.setdefault(row['Kernel Name'].split('(')[0][-70:], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    with open(dst, 'w') as f:
        f.write("# ncu launch list summary (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare shares)\n\n")
        f.write("source: %s\n\n| kernel | launches | mean us | total us | share |\n|---|---|---|---|---|\n" % src)
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write("| `%s` | %d | %.1f | %.1f | %.1f%% |\n" % (k, len(v), sum(v) / len(v), sum(v), 100 * sum(v) / tot))
    print(open(dst).read())


def raw(src, dst):
    out = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, 'w') as f:
        f.write("# ncu --set full summary (per launch)\n\nsource: %s\n\n" % src)
        for d in data:
            f.write("## %s\n\n| metric | value | unit |\n|---|---|---|\n" % d[idx['Kernel Name']].split('(')[0])
            for k in KEYS:
                if k in idx:
                    f.write("| %s | %s | %s |\n" % (k, d[idx[k]], units[idx[k]]))
            f.write("\n")
    print(open(dst).read()[:3000])


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2], sys.argv[3])
