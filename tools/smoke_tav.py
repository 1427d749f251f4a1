"""Quick end-to-end check + timing of the full TAV step on the GPU (not a test).  python tools/smoke_tav.py [variant] [cfg] [B]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_modal_emotion_b200 import _lib as L, synthetic as syn, tav  # noqa: E402
from multi_modal_emotion_b200.losses import NewCrossEntropyLoss  # noqa: E402
from multi_modal_emotion_b200.tav_train import get_statistics  # noqa: E402

variant = sys.argv[1] if len(sys.argv) > 1 else "baseline"
cfg = sys.argv[2] if len(sys.argv) > 2 else "C2"
B = int(sys.argv[3]) if len(sys.argv) > 3 else None
tav.set_encoder_variant(variant)
t0 = time.time()
model = tav.TAVForMAE({"output_dim": syn.CONFIGS[cfg]["C"], "dropout": 0.4, "learn_PosEmbeddings": True, "num_layers": 12}).cuda()
pre = tav.PreFormer().cuda()
print("built in %.1fs" % (time.time() - t0), flush=True)
inputs, labels = syn.make_batch(cfg, B=B)
for d in inputs:
    for k in d:
        d[k] = d[k].pin_memory()
crit = NewCrossEntropyLoss(torch.tensor(syn.MELD_CLASS_WEIGHTS if syn.CONFIGS[cfg]["C"] == 7 else [0.5, 0.5]))


def step():
    loss = get_statistics(inputs, labels, model, pre, crit, None, check="val", epoch=1)
    loss.backward()
    model.zero_grad(set_to_none=True)
    pre.zero_grad(set_to_none=True)
    return loss


for i in range(3):
    t = time.time()
    l = step()
    torch.cuda.synchronize()
    print("warmup %d: %.1f ms loss %.4f" % (i, 1e3 * (time.time() - t), l.item()), flush=True)
n0 = L.launch_count
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t = time.time()
e0.record()
for _ in range(5):
    step()
e1.record()
host = (time.time() - t) / 5
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("TAV %s %s B=%d: %.2f ms/step (host enqueue %.2f ms)  %.1f samples/s; %d library calls/step; peak mem %.1f GB" % (
    variant, cfg, len(labels), ms, host * 1e3, len(labels) / ms * 1e3, (L.launch_count - n0) // 5, torch.cuda.max_memory_allocated() / 2**30))
