O=gpurun_out
run() { name=$1; shift; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 2 --steps 10 --warmup 3 --no-roofline > $O/r2c_2gpu_$name.json 2> $O/r2c_2gpu_$name.err; PORT=$((PORT+1)); python - <<PY
import json
for l in open("$O/r2c_2gpu_$name.json"):
    if l.startswith("{"):
        d=json.loads(l); print("$name", round(d["value"],1), round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1))
PY
}
PORT=29531
timeout 300 python -m pytest tests/test_dp_nccl_gpu.py -x -q -m gpu -s > $O/r2c_pytest_nccl2.log 2>&1; tail -3 $O/r2c_pytest_nccl2.log
run default A=1
run nopair TAVK_GEMM_PAIR=0
run comm8 TAVK_COMM_SMS=8
run comm4 TAVK_COMM_SMS=4
run comm16 TAVK_COMM_SMS=16
run ncclmax8 NCCL_MAX_CTAS=8
