"""Summarise `ncu --set full` reports into a markdown table (run here, no GPU needed).

  python profiles/summarize_ncu.py gpurun_out/prof_x.ncu-rep [more.ncu-rep ...] > profiles/rNN_x_summary.md
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "lts__t_sectors_srcunit_tex_op_red.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
]


def load(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
        res.append(d)
    return res


def main():
    for path in sys.argv[1:]:
        launches = load(path)
        print(f"### `{path.split('/')[-1]}`\n")
        names = [l["Kernel Name"][0].replace("<unnamed>::", "").split("(")[0][:40] for l in launches]
        print("| metric | " + " | ".join(f"`{n}`" for n in names) + " |")
        print("|---|" + "---|" * len(names))
        for k in KEYS:
            if not any(k in l for l in launches):
                continue
            unit = next((l[k][1] for l in launches if k in l), "")
            cells = []
            for l in launches:
                v = l.get(k, ("", ""))[0]
                try:
                    cells.append(f"{float(v):.4g}")
                except ValueError:
                    cells.append(v)
            print(f"| {k} [{unit}] | " + " | ".join(cells) + " |")
        print()


if __name__ == "__main__":
    main()
