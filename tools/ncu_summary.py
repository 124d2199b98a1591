"""Condenses an .ncu-rep (read here, no GPU needed) into the per-kernel summary
kept under profiles/: python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_warps",
    "launch__waves_per_multiprocessor",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
    "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
]
STALL = "smsp__average_warps_issue_stalled_"


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE,
                         text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, body = rows[0], rows[1], rows[2:]
    ki = head.index("Kernel Name")
    for r in body:
        print("== %s  (id %s)" % (r[ki], r[0]))
        for w in WANT:
            if w in head:
                i = head.index(w)
                print("  %-62s %s %s" % (w, r[i], units[i]))
        stalls = []
        for i, name in enumerate(head):
            if name.startswith(STALL) and name.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(r[i]), name[len(STALL):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print("  stall cycles per issued instruction (top): " +
              ", ".join("%s %.2f" % (n, v) for v, n in stalls[:7]))
        rd = float(r[head.index("dram__bytes_read.sum")])
        wr = float(r[head.index("dram__bytes_write.sum")])
        ui = units[head.index("dram__bytes_read.sum")]
        t = float(r[head.index("gpu__time_duration.sum")])
        tu = units[head.index("gpu__time_duration.sum")]
        print("  dram traffic %.1f + %.1f %s in %.1f %s" % (rd, wr, ui, t, tu))


if __name__ == "__main__":
    main(sys.argv[1])
