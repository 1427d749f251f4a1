"""One line per profiled launch of an `ncu --set full` report: duration, DRAM bytes, tensor-pipe %, DRAM %, registers,
grid.  Reads the report with `ncu -i <rep> --page raw --csv` (no GPU needed).
python tools/ncu_rep_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_<kernel>_ncu_full.txt"""
import csv
import io
import subprocess
import sys

COLS = [
    ("gpu__time_duration.sum", "dur"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
    ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "hmma%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dsmem"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units, body = rows[0], rows[1], rows[2:]
    name_i = head.index("Kernel Name")
    idx = [(head.index(c), lab) for c, lab in COLS if c in head]
    print("report %s: %d launches" % (path, len(body)))
    print("units: " + ", ".join("%s[%s]" % (lab, units[i]) for i, lab in idx))
    for r in body:
        print("%-70s " % r[name_i][:70] + " ".join("%s=%s" % (lab, r[i]) for i, lab in idx))


if __name__ == "__main__":
    main(sys.argv[1])
