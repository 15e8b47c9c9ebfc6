"""Helpers to read ncu reports offline (no GPU): `python profiles/ncu_tools.py raw|lines <report.ncu-rep>`."""
import csv
import subprocess
import sys
from collections import defaultdict

RAW_KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum ', 'dram__bytes_write.sum ', 'gpu__dram_throughput.avg.pct',
            'sm__warps_active.avg.pct', 'launch__registers_per_thread ', 'launch__occupancy_limit', 'smsp__inst_executed.sum ',
            'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__average_warps_issue_stalled',
            'launch__shared_mem_per_block_dynamic', 'sm__throughput.avg.pct', 'smsp__issue_active.avg.pct',
            'launch__grid_size', 'launch__block_size', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ',
            'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'lts__t_sectors_srcunit_tex_op_read.sum ']


def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print('---', vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '')
        for h, u, v in zip(hdr, units, vals):
            if any(w in h + ' ' for w in RAW_KEYS):
                print(f'{h} [{u}] = {v}')


def lines(rep, top=40):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass,cuda'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur, hdr = None, None
    agg = defaultdict(lambda: [0, 0, ''])
    for r in rows:
        if len(r) >= 2 and r[0] == 'File Path':
            cur, hdr = r[1].split('/')[-1], None
            continue
        if r and r[0] == 'Line No':
            hdr = r
            ie, sm = hdr.index('Instructions Executed'), hdr.index('# Samples')
            continue
        if hdr and len(r) > ie and r[0].isdigit():
            try:
                n, s = int(r[ie] or 0), int(r[sm] or 0)
            except ValueError:
                continue
            k = (cur, int(r[0]))
            agg[k][0] += n
            agg[k][1] += s
            agg[k][2] = r[1][:100]
    tot = sum(v[0] for v in agg.values()) or 1
    tots = sum(v[1] for v in agg.values()) or 1
    print('total warp instructions', tot, 'samples', tots)
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f'{k[0]}:{k[1]:4d} {100 * v[0] / tot:5.1f}% inst {100 * v[1] / tots:5.1f}% smp | {v[2]}')


if __name__ == '__main__':
    {'raw': raw, 'lines': lines}[sys.argv[1]](sys.argv[2])
