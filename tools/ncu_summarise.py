#!/usr/bin/env python
"""Turn the raw ncu CSV exports of a bench.py run into the summaries kept under profiles/.

    python tools/ncu_summarise.py launches <launch_list.csv> <out.csv> "<header note>"
    python tools/ncu_summarise.py full <raw_page.csv> <out.csv> "<header note>"

`launches`: per-kernel launch count, total time and share from `ncu --metrics gpu__time_duration.sum --csv`.
`full`: one row per launch (layer order of one Unet forward) with time, DRAM bytes, tensor-pipe utilisation, from
`ncu -i <rep> --page raw --csv` of a `--set full` capture of the network kernels.
"""
import csv
import sys
from collections import OrderedDict

LAYERS = ['encode1(first)', 'encode2+pool', 'encode3', 'encode4+pool', 'encode5', 'encode6+pool', 'encode7',
          'encode8+pool', 'middle_conv1', 'middle_conv2', 'up1', 'decode1', 'decode2', 'up2', 'decode3', 'decode4',
          'up3', 'decode5', 'decode6', 'up4', 'decode7', 'decode8+head']


def rows_of(path):
    with open(path, newline='') as f:
        lines = [ln for ln in f if not ln.startswith('==')]
    return list(csv.reader(lines))


def launches(src, dst, note):
    rows = rows_of(src)
    hdr = rows[0]
    ki, vi, mi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name')
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi or r[mi] != 'gpu__time_duration.sum':
            continue
        name = r[ki].split('(')[0][:70]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(',', '')) / 1e3        # ns -> us
    total = sum(v[1] for v in agg.values())
    with open(dst, 'w') as f:
        f.write(f'# {note}\n# cold-cache, serialised per-launch times: compare SHARES, not absolutes\n')
        f.write('kernel,launches,total_us,share\n')
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f'{k},{n},{us:.1f},{us / total:.3f}\n')


def full(src, dst, note):
    rows = rows_of(src)
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    want = [('kernel', 'Kernel Name'), ('time', 'gpu__time_duration.sum'), ('grid', 'launch__grid_size'),
            ('dram_rd', 'dram__bytes_read.sum'), ('dram_wr', 'dram__bytes_write.sum'),
            ('dram_pct', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
            ('l2_hit_pct', 'lts__t_sector_hit_rate.pct'),
            ('tensor_math_pct_of_peak', 'sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed'),
            ('sm_cycles', 'sm__cycles_elapsed.max'), ('regs', 'launch__registers_per_thread')]
    want = [(a, b) for a, b in want if b in col]
    data = rows[2:]
    to_gb = {'byte': 1e-9, 'Kbyte': 1e-6, 'Mbyte': 1e-3, 'Gbyte': 1.0}
    to_ms = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}
    total_dram = 0.0
    out = []
    for li, r in enumerate(data):
        vals = []
        for a, b in want:
            v, u = r[col[b]], units[col[b]]
            if a == 'kernel':
                v = v.split('(')[0].replace('biu::', '').replace(' ', '').replace(',', '_')
            elif a in ('dram_rd', 'dram_wr'):
                v = f'{float(v.replace(",", "")) * to_gb.get(u, 1.0):.6f}'
                if li > 0:
                    total_dram += float(v)
            elif a == 'time':
                v = f'{float(v.replace(",", "")) * to_ms.get(u, 1.0):.6f}'
            else:
                v = v.replace(',', '')
            vals.append(v)
        out.append((LAYERS[li] if li < len(LAYERS) else f'launch{li}', vals))
    with open(dst, 'w') as f:
        f.write(f'# {note}\n')
        f.write('# tensor_math_pct_of_peak = sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed\n')
        f.write(f'# DRAM traffic of the 21 tcgen05 launches: {total_dram:.2f} GB per forward\n')
        f.write('# units: time=ms, dram_rd=Gbyte, dram_wr=Gbyte, dram_pct=%, l2_hit_pct=%, tensor_math_pct_of_peak=%\n')
        f.write('layer,' + ','.join(a for a, _ in want) + '\n')
        for name, vals in out:
            f.write(name + ',' + ','.join(vals) + '\n')
    print(f'DRAM traffic of the tcgen05 launches: {total_dram:.2f} GB')


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](*sys.argv[2:5])
