"""Per-kernel GPU time of one eager TAV training step via torch.profiler (CUPTI): the quick breakdown used to decide
what to optimise next (the ncu launch list under profiles/ is the judged evidence).
python tools/profile_step.py [variant] [cfg] [B] > gpurun_out/kernels.txt"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_modal_emotion_b200 import dp, synthetic as syn, tav  # noqa: E402
from multi_modal_emotion_b200.losses import NewCrossEntropyLoss  # noqa: E402
from multi_modal_emotion_b200.optim import FusedAdamW  # noqa: E402

variant = sys.argv[1] if len(sys.argv) > 1 else "baseline"
cfg = sys.argv[2] if len(sys.argv) > 2 else "C2"
B = int(sys.argv[3]) if len(sys.argv) > 3 else None
tav.set_encoder_variant(variant)
C = syn.CONFIGS[cfg]["C"]
model = tav.TAVForMAE({"output_dim": C, "dropout": 0.4, "learn_PosEmbeddings": True, "num_layers": 12}).cuda().train()
pre = tav.PreFormer().cuda().train()
crit = NewCrossEntropyLoss(torch.tensor(syn.MELD_CLASS_WEIGHTS if C == 7 else [0.5, 0.5]))
params = [p for p in model.parameters() if p.requires_grad] + [p for p in pre.parameters() if p.requires_grad]
runner = dp.DataParallelTAV(model, pre, crit, FusedAdamW(params, lr=1e-5, weight_decay=1e-4), clip=1.0)
inputs, labels = syn.make_batch(cfg, B=B)
inputs = [{k: v.cuda() for k, v in d.items()} for d in inputs]
labels = labels.cuda()
pre.static_keep_count = 104
model.static_keep_count = 1568 - 104
for _ in range(3):
    runner.train_step(inputs, labels, 1, "train")
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    runner.train_step(inputs, labels, 1, "train")
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_time_total > 0]
ev.sort(key=lambda e: -e.device_time_total)
tot = sum(e.self_device_time_total for e in ev)
print("total device time of one step: %.2f ms over %d kernel launches" % (tot / 1e3, sum(e.count for e in ev if e.self_device_time_total > 0)))
print("%-90s %8s %10s %7s" % ("kernel", "count", "total_us", "share"))
for e in ev[:60]:
    if e.self_device_time_total <= 0:
        continue
    print("%-90s %8d %10.0f %6.1f%%" % (e.key[:90], e.count, e.self_device_time_total, 100 * e.self_device_time_total / tot))
