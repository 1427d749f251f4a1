O=gpurun_out
PORT=29571
for mb in 32 64 256; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 2 --steps 10 --warmup 3 --no-roofline --bucket-mb $mb > $O/r2i_2gpu_b$mb.json 2> $O/r2i_2gpu_b$mb.err; PORT=$((PORT+1))
python - <<PY
import json
for l in open("$O/r2i_2gpu_b$mb.json"):
    if l.startswith("{"):
        d=json.loads(l); print("bucket_mb $mb", round(d["value"],1), round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1))
PY
done
