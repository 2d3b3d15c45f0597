"""One line per kernel launch of a .ncu-rep (--set full): duration, achieved HBM GB/s, tensor-pipe % of peak.
tensor pipe % = (imma + hmma sub-pipe active cycles, summed over the SM's 4 sub-cores by ncu) / 4 / elapsed SM cycles,
i.e. the fraction of cycles the tensor pipe was busy (for the bf16 capture this reproduces the 40 % that
sm__ops_path_tensor_src_bf16_dst_fp32 / 8192 ops/clk/SM gives, profiles/r01_notes.md).
usage: python tools/ncu_kernels.py file.ncu-rep [more.ncu-rep ...]"""
import csv, io, subprocess, sys

def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return float("nan")

SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    def col(name):
        return hdr.index(name) if name in hdr else -1
    c = {k: col(k) for k in [
        "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "TPC.TriageCompute.sm__pipe_tensor_subpipe_imma_cycles_active_realtime.avg",
        "TPC.TriageCompute.sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
        "sm__cycles_elapsed.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.sum", "launch__grid_size",
        "smsp__inst_executed_op_popc.sum", "sm__inst_executed_pipe_xu.sum"]}
    print(f"== {path}")
    print(f"{'#':>3s} {'kernel':34s} {'grid':>6s} {'us':>8s} {'HBM rd MB':>10s} {'HBM wr MB':>10s} {'HBM GB/s':>9s} {'dram %':>7s} "
          f"{'tensor pipe %':>13s} {'imma cyc':>9s} {'hmma cyc':>9s} {'SM thr %':>8s}")
    for k, r in enumerate(rows[2:]):
        def val(name, scale=True):
            i = c[name]
            if i < 0 or r[i] == "":
                return float("nan")
            return num(r[i]) * (SCALE.get(units[i], 1.0) if scale else 1.0)
        name = r[c["Kernel Name"]].replace("void yb::<unnamed>::", "").replace("void yb::", "")
        name = name.split("(")[0][:34]
        t = val("gpu__time_duration.sum")
        imma = val('TPC.TriageCompute.sm__pipe_tensor_subpipe_imma_cycles_active_realtime.avg', False)
        hmma = val('TPC.TriageCompute.sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg', False)
        rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
        print(f"{k:3d} {name:34s} {r[c['launch__grid_size']]:>6s} {t*1e6:8.1f} {rd/1e6:10.2f} {wr/1e6:10.2f} {(rd+wr)/t/1e9:9.0f} "
              f"{val('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', False):7.1f} "
              f"{100.0 * ((0 if imma != imma else imma) + (0 if hmma != hmma else hmma)) / 4.0 / val('sm__cycles_elapsed.avg', False):13.1f} "
              f"{val('TPC.TriageCompute.sm__pipe_tensor_subpipe_imma_cycles_active_realtime.avg', False):9.0f} "
              f"{val('TPC.TriageCompute.sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg', False):9.0f} "
              f"{val('sm__throughput.avg.pct_of_peak_sustained_elapsed', False):8.1f}")
