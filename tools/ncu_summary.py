"""Summarise .ncu-rep captures into the few numbers the roofline needs (reads with `ncu -i ... --page raw --csv`)."""
import csv, io, subprocess, sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs"),
    ("sm__cycles_active.avg", "sm_cycles_active"),
    ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor_pipe_pct(realtime,elapsed)"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "hmma_cycles_active"),
    ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "hmma_inst_pct"),
    ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tmem_cycles_pct"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_pct"),
    ("smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "stall_long_sb"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
]


def summarise(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        return None
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"}
        for key, name in WANT:
            if key in hdr:
                i = hdr.index(key)
                d[name] = f"{r[i]} {units[i]}".strip()
        res.append(d)
    return res


if __name__ == "__main__":
    for p in sys.argv[1:]:
        print("==", p)
        for d in summarise(p) or []:
            for k, v in d.items():
                print(f"   {k:36s} {v}")
