"""Turn an .ncu-rep (ncu --set full) into a short per-kernel text summary (run on the CPU box)."""
import csv
import subprocess
import sys

WANT = [
    ('gpu__time_duration.sum', 'duration'),
    ('launch__grid_size', 'grid'), ('launch__block_size', 'block'),
    ('launch__registers_per_thread', 'regs/thread'),
    ('launch__occupancy_limit_shared_mem', 'occupancy limit (smem, blocks)'),
    ('launch__occupancy_limit_registers', 'occupancy limit (regs, blocks)'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %'),
    ('dram__bytes_read.sum', 'dram read'), ('dram__bytes_write.sum', 'dram write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram throughput %'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe active %'),
    ('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'xu (MUFU) pipe %'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue active %'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm throughput %'),
    ('smsp__inst_executed.sum', 'warp instructions'),
    ('sm__cycles_elapsed.avg', 'sm cycles'),
]


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    seen = set()
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')]
        short = name.split('(')[0][-70:]
        if short in seen:
            continue
        seen.add(short)
        print(f'## {short}')
        for key, label in WANT:
            if key in hdr:
                i = hdr.index(key)
                print(f'  {label:34s} {r[i]} {units[i]}')
        stalls = {h: float(r[i]) for i, h in enumerate(hdr)
                  if h.startswith('smsp__pcsamp_warps_issue_stalled') and not h.endswith('_not_issued') and r[i]}
        tot = sum(stalls.values()) or 1
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:6]
        print('  top stall reasons (pc samples):   ' + ', '.join(
            f"{k.replace('smsp__pcsamp_warps_issue_stalled_', '')} {100 * v / tot:.0f}%" for k, v in top))
        print()


if __name__ == '__main__':
    main(sys.argv[1])
