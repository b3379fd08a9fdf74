"""Developer helper: per-source-line instruction / stall-sample shares of one kernel from an .ncu-rep.
Usage: python tools/ncu_lines.py <report.ncu-rep> <kernel-regex> [top]"""
import collections
import csv
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
if len(rows) > 2:
    hdr, units = rows[0], rows[1]
    keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
            'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
            'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__grid_size', 'launch__block_size',
            'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'smsp__inst_executed.sum',
            'smsp__issue_active.avg.pct_of_peak_sustained_active',
            'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__shared_mem_per_block_dynamic',
            'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
            'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum']
    r = rows[2]
    print(r[hdr.index('Kernel Name')][:100])
    for k in keys:
        if k in hdr:
            print("  %s: %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + rx], capture_output=True, text=True).stdout
hdr = None
cur = None
agg = collections.defaultdict(lambda: [0, 0, ''])
stall = collections.Counter()
done_first = False
for r in csv.reader(src.splitlines()):
    if len(r) >= 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        continue
    if len(r) >= 2 and r[0] == 'Line No':
        hdr = r
        iS, iI = hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
        sc = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
        continue
    if hdr is None or len(r) <= max(iS, iI):
        continue
    if r[0] != '' and r[2] == '-':
        try:
            a = agg[(cur, int(r[0]))]
            a[0] += int(r[iS])
            a[1] += int(r[iI])
            a[2] = r[1].strip()[:95]
        except ValueError:
            pass
    elif r[0] == '' and r[2] not in ('-', '...'):
        for i in sc:
            try:
                stall[hdr[i]] += int(r[i])
            except (ValueError, IndexError):
                pass
ts = sum(a[0] for a in agg.values()) or 1
ti = sum(a[1] for a in agg.values()) or 1
print("total samples", ts, "warp inst", ti)
print("stalls:", ", ".join("%s=%.1f%%" % (k, 100 * v / max(1, sum(stall.values()))) for k, v in stall.most_common(8)))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-14s:%4d inst=%5.1f%% samp=%5.1f%%  %s" % (k[0][:14], k[1], 100 * a[1] / ti, 100 * a[0] / ts, a[2]))
