"""Developer tool: condense an ncu report (--set full) into a small markdown summary for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_xxx.md
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of ncu peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("launch__occupancy_limit_shared_mem", "CTAs/SM (smem limit)"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_sb / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu summary of `{path.split('/')[-1]}` (`ncu --set full --clock-control none`)\n")
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0]
        print(f"## {name}\n\n| metric | value |\n|---|---|")
        for k, label in KEYS:
            if k in idx and r[idx[k]] != "":
                print(f"| {label} (`{k}`) | {r[idx[k]]} {units[idx[k]]} |")
        try:
            rd = float(r[idx["dram__bytes_read.sum"]].replace(",", ""))
            wr = float(r[idx["dram__bytes_write.sum"]].replace(",", ""))
            dur = float(r[idx["gpu__time_duration.sum"]].replace(",", ""))
            ru, wu, du = units[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_write.sum"]], units[idx["gpu__time_duration.sum"]]
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            tsc = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}
            tot = rd * scale[ru] + wr * scale[wu]
            print(f"| **DRAM traffic (read+write)** | {tot / 1e9:.4f} GB -> {tot / (dur * tsc[du]) / 1e9:.0f} GB/s under ncu |")
        except Exception:
            pass
        print()


if __name__ == "__main__":
    main(sys.argv[1])
