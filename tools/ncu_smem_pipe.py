"""Shared-memory data-pipe accounting of an `ncu --set full` report: LSU load / store wavefronts and tensor-core operand
wavefronts per SM against the kernel's elapsed cycles (one wavefront = one pipe cycle).  No GPU needed.
python tools/ncu_smem_pipe.py gpurun_out/prof.ncu-rep [steps_per_launch] >> profiles/rNN_ncu_attention_smem_pipe.txt"""
import csv
import io
import subprocess
import sys

LD = "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum"
ST = "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum"
TC = "l1tex__data_pipe_tc_wavefronts_mem_shared.sum"
CONF_LD = "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum"
CYC = "sm__cycles_elapsed.avg"
XU = "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"
TENSOR = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
ISSUE = "sm__issue_active.avg.pct_of_peak_sustained_elapsed"
SMS = 148


def num(d, k):
    return float(d[k].replace(",", ""))


def main(path, steps=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head = rows[0]
    print("report %s" % path)
    for r in rows[2:]:
        d = dict(zip(head, r))
        ld, st, tc, cyc = num(d, LD), num(d, ST), num(d, TC), num(d, CYC)
        if ld + st + tc == 0:
            continue
        line = ("%-28s dur %7.1f us  cycles/SM %8.0f  smem wavefronts/SM: ld %8.0f (bank-conflict part %8.0f) st %8.0f "
                "tensor-operand %8.0f  sum %8.0f = %5.1f %% of cycles | xu %4.1f %% tensor %4.1f %% issue %4.1f %%" % (
                    d["Kernel Name"][:28], num(d, "gpu__time_duration.sum"), cyc, ld / SMS, num(d, CONF_LD) / SMS, st / SMS,
                    tc / SMS, (ld + st + tc) / SMS, 100.0 * (ld + st + tc) / SMS / cyc, num(d, XU), num(d, TENSOR),
                    num(d, ISSUE)))
        print(line)
        if steps:
            print("%-28s per inner step (%d steps per launch): ld %.0f st %.0f tensor-operand %.0f" % (
                "", steps, ld / steps, st / steps, tc / steps))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else None)
