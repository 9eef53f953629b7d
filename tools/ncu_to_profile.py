"""Summarise an .ncu-rep (raw page) into a small text file for profiles/.
usage: python tools/ncu_to_profile.py report.ncu-rep out.txt "title line(s)" """
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__cluster_size", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum"]
rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u = rows[0], rows[1]
with open(out, "w") as f:
    f.write(title.replace("\\n", "\n") + "\n")
    for r in rows[2:]:
        f.write("== " + r[h.index("Kernel Name")] + "\n")
        for k in KEYS:
            if k in h and r[h.index(k)] != "":
                f.write(f"   {k:95s} {r[h.index(k)]:>18s} {u[h.index(k)]}\n")
        st = [(float(r[i].replace(",", "")), h[i]) for i in range(len(h))
              if "issue_stalled" in h[i] and h[i].endswith("_per_warp_active.pct") and r[i]]
        for v, k in sorted(st, reverse=True)[:6]:
            f.write(f"   stall {k.split('issue_stalled_')[1].split('_per_warp')[0]:45s} {v:8.2f} % of active warps\n")
print(open(out).read())
