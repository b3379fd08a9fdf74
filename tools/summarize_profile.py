"""Turn the scratch outputs of tools/gpu_profile.sh (gpurun_out/<tag>_*) into tracked summaries under
profiles/: the bench JSON lines, the per-kernel launch list (count, total device time, share) and the key
ncu --set full metrics of the hot kernels.  Usage: python tools/summarize_profile.py <tag>"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "memory_l1_wavefronts_shared_ideal",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed"]


def launches(tag, lines):
    path = os.path.join(OUT, tag + "_launches.csv")
    if not os.path.exists(path):
        return
    with open(path) as f:
        rows = [ln for ln in f if not ln.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(rows):
        k = row["Kernel Name"].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        v = float(row["Metric Value"].replace(",", ""))
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    lines.append("## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, bench.py --steps 2 --warmup 1)\n")
    lines.append("Per-launch times under ncu are cold-cache and serialised: compare shares, not absolutes.\n")
    lines.append("| kernel | launches | total ms | ms/launch | share |\n|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("| `%s` | %d | %.3f | %.4f | %.1f%% |" % (k, a[0], a[1] / 1e6, a[1] / 1e6 / a[0], 100 * a[1] / tot))
    lines.append("")


def stalls(rep, kernel_regex):
    """warp-stall shares of one kernel from the source page (sampling)"""
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name",
                          "regex:" + kernel_regex], capture_output=True, text=True).stdout
    hdr, tot = None, collections.Counter()
    for r in csv.reader(src.splitlines()):
        if len(r) > 3 and r[0] == "Address":
            hdr = r
            cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
            continue
        if hdr and r and r[0].startswith("0x"):
            for i in cols:
                try:
                    tot[hdr[i]] += int(r[i])
                except (ValueError, IndexError):
                    pass
    n = sum(tot.values()) or 1
    return ", ".join("%s %.0f%%" % (k.replace("stall_", ""), 100.0 * v / n) for k, v in tot.most_common(6))


def full(tag, lines, suffix="_full"):
    rep = os.path.join(OUT, tag + suffix + ".ncu-rep")
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        return
    hdr, units = rows[0], rows[1]
    lines.append("## `ncu --set full --clock-control none` (%s%s.ncu-rep, per launch)\n" % (tag, suffix))
    seen = set()
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        if name in seen:
            continue
        seen.add(name)
        lines.append("### `%s`\n" % name)
        lines.append("| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                lines.append("| %s | %s | %s |" % (k, r[i], units[i]))
        lines.append("| warp stall samples | %s | |" % stalls(rep, name.split("<")[0]))
        lines.append("")


def main():
    tag = sys.argv[1]
    os.makedirs(PROF, exist_ok=True)
    lines = ["# Profile summary `%s`\n" % tag]
    for suffix, title in (("_bench.json", "bench.py (this repo)"), ("_bench_ref.json", "bench.py --impl reference")):
        p = os.path.join(OUT, tag + suffix)
        if os.path.exists(p):
            for ln in open(p):
                ln = ln.strip()
                if ln.startswith("{"):
                    lines.append("## %s\n\n```json\n%s\n```\n" % (title, json.dumps(json.loads(ln), indent=1)))
    g = os.path.join(OUT, tag + "_gpu.csv")
    if os.path.exists(g):
        lines.append("## Box\n\n```\n%s```\n" % open(g).read())
    launches(tag, lines)
    for suffix in ("_full", "_join", "_fillsel", "_spmvshort"):
        full(tag, lines, suffix)
    for suffix, title in (("_write_bw.log", "tools/write_bw.cu: write bandwidth of the CSR row layout (1 002 001 rows x 2221 entries)"),
                          ("_spmv_chunk.log", "tools/spmv_chunk.py: short-row SpMV, rows per CTA range (PYCI_B200_SPMV_CHUNK)"),
                          ("_hci_grow.json", "tools/hci_grow.py: config-5-style space grown by add_hci on the device")):
        p = os.path.join(OUT, tag + suffix)
        if os.path.exists(p):
            lines.append("## %s\n\n```\n%s\n```\n" % (title, open(p).read().strip()))
    with open(os.path.join(PROF, tag + "_summary.md"), "w") as f:
        f.write("\n".join(lines) + "\n")
    src = os.path.join(OUT, tag + "_launches.csv")
    if os.path.exists(src):
        import shutil
        shutil.copy(src, os.path.join(PROF, tag + "_launches.csv"))
    print("wrote", os.path.join(PROF, tag + "_summary.md"))


if __name__ == "__main__":
    main()
