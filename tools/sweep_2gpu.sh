# 2-GPU A/B runs of bench.py (gpurun --gpus 2 -- 'bash tools/sweep_2gpu.sh name ENV=.. -- name2 ENV=..'); default set below
O=gpurun_out
PORT=29561
run() { name=$1; shift; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 2 --steps 10 --warmup 3 --no-roofline > $O/r2h_2gpu_$name.json 2> $O/r2h_2gpu_$name.err; PORT=$((PORT+1)); python - <<PY
import json
for l in open("$O/r2h_2gpu_$name.json"):
    if l.startswith("{"):
        d=json.loads(l); print("$name", round(d["value"],1), round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1))
PY
}
timeout 300 python -m pytest tests/test_dp_nccl_gpu.py -x -q -m gpu -s 2>&1 | grep -E "global loss|whole-model|passed|failed|Error" | head
run rowsparse A=1
run dense TAVK_ROW_SPARSE=0
run rowsparse_b TAVK_X=1
