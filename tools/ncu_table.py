"""Markdown table of the kernels of an .ncu-rep (`ncu --set full` capture): grid x block, registers, dynamic shared memory, time,
DRAM bytes read + written, DRAM GB/s and its fraction of the measured peak, issue-active %, top warp stalls per issue.

    python tools/ncu_table.py gpurun_out/r02f_full.ncu-rep [peak_gbs]
"""
import csv, json, subprocess, sys

rep = sys.argv[1]
peak = float(sys.argv[2]) if len(sys.argv) > 2 else json.load(open('MEASURED_PEAKS.json')).get('hbm_gbs', 6549.4)
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units, data = rows[0], rows[1], rows[2:]
col = {c: i for i, c in reversed(list(enumerate(h)))}
stall = {c.split('issue_stalled_')[1].split('_per_issue')[0]: i for i, c in enumerate(h)
         if c.startswith('smsp__average_warps_issue_stalled_') and c.endswith('_per_issue_active.ratio')}
short = {'long_scoreboard': 'long_sb', 'short_scoreboard': 'short_sb', 'not_selected': 'not_sel', 'mio_throttle': 'mio',
         'math_pipe_throttle': 'math', 'lg_throttle': 'lg', 'no_instruction': 'no_inst'}


def num(r, name):
    v = r[col[name]].replace(',', '')
    return float(v) if v else 0.0


def scaled(r, name, to):
    """value of a metric converted to unit `to` (ncu picks ns/us/ms, byte/Kbyte/Mbyte/Gbyte per column)"""
    u = units[col[name]]
    f = {'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 's': 1.0, 'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
    g = {'ms': 1e-3, 'GB': 1e9}[to]
    return num(r, name) * f / g


print('| kernel | grid x block | regs | dyn smem | time (ms) | DRAM read + write (GB) | DRAM GB/s (frac of peak) | issue active % | top stalls (warps per issue) |')
print('|---|---|---|---|---|---|---|---|---|')
for r in data:
    name = r[col['Kernel Name']].replace('void ', '').split('(')[0]
    t = scaled(r, 'gpu__time_duration.sum', 'ms')
    rd, wr = scaled(r, 'dram__bytes_read.sum', 'GB'), scaled(r, 'dram__bytes_write.sum', 'GB')
    gbs = (rd + wr) / (t * 1e-3) if t else 0.0
    st = sorted(((float(r[i].replace(',', '') or 0), short.get(k, k)) for k, i in stall.items()), reverse=True)[:4]
    smem = r[col['launch__shared_mem_per_block_dynamic']] + ' ' + units[col['launch__shared_mem_per_block_dynamic']]
    print(f"| `{name}` | {int(num(r, 'launch__grid_size'))} x {int(num(r, 'launch__block_size'))} | "
          f"{int(num(r, 'launch__registers_per_thread'))} | {smem} | {t:.3f} | {rd:.3f} + {wr:.3f} | "
          f"{gbs:.0f} ({gbs / peak:.2f}) | {num(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.0f} | "
          + ', '.join(f'{n} {v:.1f}' for v, n in st) + ' |')
