"""Print the metrics that matter from an .ncu-rep (raw page) and, optionally, the hottest source lines.
usage: python tools/ncu_summary.py report.ncu-rep [--source N]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio" ]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    print("==", name[:90])
    for k in hdr:
        if k in KEYS or ("tensor" in k and k.endswith("avg.pct_of_peak_sustained_elapsed") and r[hdr.index(k)] not in ("0", "")):
            i = hdr.index(k)
            print(f"   {k:90s} {r[i]:>16s} {units[i]}")
    stalls = [(float(r[i].replace(",", "")), hdr[i]) for i in range(len(hdr)) if "warp_issue_stalled" in hdr[i] and hdr[i].endswith("_per_warp_active.pct") and r[i]]
    for v, k in sorted(stalls, reverse=True)[:8]:
        print(f"   stall {k.split('issue_stalled_')[1].split('_per_warp')[0]:40s} {v:8.2f} %")
if "--source" in sys.argv:
    n = int(sys.argv[sys.argv.index("--source") + 1])
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h = rows[0]
    print(h)
    try:
        ci = h.index("# Samples") if "# Samples" in h else [i for i, x in enumerate(h) if "Sampling" in x][0]
    except Exception:
        ci = None
    if ci is not None:
        data = [r for r in rows[1:] if len(r) > ci and r[ci].replace(",", "").isdigit()]
        data.sort(key=lambda r: -int(r[ci].replace(",", "")))
        for r in data[:n]:
            print(r[ci], "|", r[h.index("Source")][:110] if "Source" in h else r[:3])
