#!/bin/bash
# Box-side profile recipe of round 2 (run under gpurun, ONE GPU): every ncu capture follows a plain run of the same command
# that exited 0, and the multi-MB .ncu-rep files are reduced to text on the box (gpurun_out/ brings back at most 64 MiB).
set -u
O=gpurun_out
mkdir -p $O
# 1. launch list of three eager steps (the last one is the steady-state step)
timeout 300 python tools/ncu_step.py 3 > $O/r2_ncu_plain.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 14000 --csv --log-file $O/r2_launches_3steps.csv \
    python tools/ncu_step.py 3 > $O/r2_ncu_run.log 2>&1
echo "launch list rc=$?"
python tools/summarize_ncu.py $O/r2_launches_3steps.csv --last-step > $O/r2_ncu_launches_steady_step.txt 2>&1
gzip -f $O/r2_launches_3steps.csv
# 2. the dominant GEMM signatures, ncu --set full
SH="qkv ffn_up_gelu_grad ffn_up_dgrad_mul out_proj_resid ffn_down_resid wgrad_ffn_up"
export TAVK_PROBE_REPS=1
timeout 100 python tools/gemm_probe.py $SH > $O/r2_probe_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none -k regex:gemm_bf16 -c 18 -o $O/r2_prof_gemm \
    python tools/gemm_probe.py $SH > $O/r2_ncu_gemm.log 2>&1
echo "gemm full rc=$?"
python tools/ncu_rep_summary.py $O/r2_prof_gemm.ncu-rep > $O/r2_ncu_full_gemm.txt 2>&1
python tools/ncu_smem_pipe.py $O/r2_prof_gemm.ncu-rep 1 > $O/r2_ncu_gemm_smem_pipe.txt 2>&1
rm -f $O/r2_prof_gemm.ncu-rep
# 3. attention kernels at the VideoMAE shape, ncu --set full with source
export TAVK_NO_KINETO=1
timeout 100 python tools/attn_only.py 16 1464 1 > $O/r2_attn_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 4 -c 4 -o $O/r2_prof_attn \
    python tools/attn_only.py 16 1464 1 > $O/r2_ncu_attn.log 2>&1
echo "attention full rc=$?"
python tools/ncu_rep_summary.py $O/r2_prof_attn.ncu-rep > $O/r2_ncu_full_attention.txt 2>&1
python tools/ncu_smem_pipe.py $O/r2_prof_attn.ncu-rep 1 > $O/r2_ncu_attention_smem_pipe.txt 2>&1
for i in 1 2 3 4; do python tools/ncu_src_hot.py $O/r2_prof_attn.ncu-rep $i 25 > $O/r2_ncu_attention_src_hot_$i.txt 2>&1; done
rm -f $O/r2_prof_attn.ncu-rep
ls -la $O | head -40
