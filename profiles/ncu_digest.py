"""Digest of an ncu report: one column per captured launch, the metrics the roofline discussion uses.

    python profiles/ncu_digest.py gpurun_out/final_pointwise.ncu-rep profiles/r2_pointwise_ncu "title line"

writes <out>.md (table) and <out>_raw.csv (the same metrics, machine readable).  Needs `ncu` on PATH (reads the report, no GPU)."""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def main(rep, out, title):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    head, units, data = rows[0], rows[1], rows[2:]
    names = []
    for r in data:
        n = r[head.index("Kernel Name")]
        names.append(n.replace("void ", "").replace("vsb::", "").split("(")[0][:48])
    have = [m for m in METRICS if m in head]
    with open(out + "_raw.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + names)
        for m in have:
            i = head.index(m)
            w.writerow([m, units[i]] + [r[i] for r in data])
    with open(out + ".md", "w") as f:
        f.write(f"# {title}\n\nSource report: `{rep}` (scratch, not tracked); raw metrics: `{out.split('/')[-1]}_raw.csv`.  Cold-cache, serialised launches.\n\n")
        f.write("| metric | " + " | ".join(f"`{n}`" for n in names) + " |\n|---|" + "---|" * len(names) + "\n")
        for m in have:
            i = head.index(m)
            u = f" [{units[i]}]" if units[i] else ""
            f.write(f"| `{m}`{u} | " + " | ".join(r[i] for r in data) + " |\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else sys.argv[1])
