#!/usr/bin/env python
"""Turn the raw ncu CSV exports of a bench.py run into the summaries kept under profiles/.

    python tools/ncu_summarise.py launches <launch_list.csv> <out.csv> "<header note>"
    python tools/ncu_summarise.py full <raw_page.csv> <out.csv> "<header note>"
    python tools/ncu_summarise.py metrics <metrics.csv> <out.csv> "<header note>" [first_id last_id]

`launches`: per-kernel launch count, total time and share from `ncu --metrics gpu__time_duration.sum --csv`.
`metrics`: one row per launch of a `tools/ncu_step.sh` capture (ncu --metrics ... --csv, long format): time, DRAM bytes
and achieved GB/s against the measured HBM peak, DRAM %, tensor-pipe %, L2 hit rate; optional launch-id window.
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


def metrics(src, dst, note, first=None, last=None, hbm_peak=6549.1):
    rows = rows_of(src)
    hdr = rows[0]
    ci = {h: i for i, h in enumerate(hdr)}
    per = OrderedDict()
    for r in rows[1:]:
        if len(r) < len(hdr):
            continue
        k = (int(r[ci['ID']]), r[ci['Kernel Name']].split('(')[0].replace('void ', '').replace('biu::', '')[:64])
        per.setdefault(k, {})[r[ci['Metric Name']]] = (float(r[ci['Metric Value']].replace(',', '') or 0)
                                                       if r[ci['Metric Value']] not in ('n/a', '') else float('nan'),
                                                       r[ci['Metric Unit']])
    to_gb = {'byte': 1e-9, 'Kbyte': 1e-6, 'Mbyte': 1e-3, 'Gbyte': 1.0}
    to_ms = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}
    T = 'sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed'
    tot_ms = tot_gb = tc_gb = tc_ms = 0.0
    lines = []
    for (i, name), m in per.items():
        if (first is not None and i < first) or (last is not None and i > last):
            continue
        t = m['gpu__time_duration.sum'][0] * to_ms[m['gpu__time_duration.sum'][1]]
        rd = m['dram__bytes_read.sum'][0] * to_gb[m['dram__bytes_read.sum'][1]]
        wr = m['dram__bytes_write.sum'][0] * to_gb[m['dram__bytes_write.sum'][1]]
        gbs = (rd + wr) / t * 1e3 if t > 0 else 0.0
        tp = m.get(T, (float('nan'), ''))[0]
        tot_ms += t
        tot_gb += rd + wr
        if 'conv_rows' in name or 'conv_halo' in name or 'conv_tc' in name:
            tc_gb += rd + wr
            tc_ms += t
        qname = '"' + name + '"'
        lines.append(f"{i},{qname},{t:.4f},{rd:.4f},{wr:.4f},{gbs:.0f},{gbs / hbm_peak:.3f},"
                     f"{m['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'][0]:.1f},{tp:.1f},"
                     f"{m['lts__t_sector_hit_rate.pct'][0]:.1f},{int(m['launch__registers_per_thread'][0])},"
                     f"{int(m['launch__grid_size'][0])}")
    with open(dst, 'w') as f:
        f.write(f'# {note}\n')
        f.write('# ncu --metrics (tools/ncu_step.sh), --clock-control none: cold-cache, serialised launches - compare SHARES and\n')
        f.write(f'# per-kernel rates, not absolute step time. hbm_frac = achieved GB/s / {hbm_peak} (MEASURED_PEAKS.json hbm_gbs).\n')
        f.write(f'# all launches: {tot_ms:.3f} ms, {tot_gb:.2f} GB DRAM; tcgen05 conv launches: {tc_ms:.3f} ms, {tc_gb:.2f} GB DRAM\n')
        f.write('id,kernel,time_ms,dram_rd_gb,dram_wr_gb,dram_gbs,hbm_frac,dram_pct,tensor_math_pct,l2_hit_pct,regs,grid\n')
        f.write('\n'.join(lines) + '\n')
    return tc_gb, tc_ms, tot_gb, tot_ms


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
    if sys.argv[1] == 'metrics':
        extra = sys.argv[5:]
        print(metrics(sys.argv[2], sys.argv[3], sys.argv[4], int(extra[0]) if extra else None,
                      int(extra[1]) if len(extra) > 1 else None))
    else:
        {'launches': launches, 'full': full}[sys.argv[1]](*sys.argv[2:5])
