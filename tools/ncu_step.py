"""Minimal eager driver for ncu: N training steps of the TAV workload, nothing else.  python tools/ncu_step.py [steps] [B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_modal_emotion_b200 import dp, synthetic as syn, tav  # noqa: E402
from multi_modal_emotion_b200.losses import NewCrossEntropyLoss  # noqa: E402
from multi_modal_emotion_b200.optim import FusedAdamW  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
tav.set_encoder_variant("baseline")
model = tav.TAVForMAE({"output_dim": 7, "dropout": 0.4, "learn_PosEmbeddings": True, "num_layers": 12}).cuda().train()
pre = tav.PreFormer().cuda().train()
crit = NewCrossEntropyLoss(torch.tensor(syn.MELD_CLASS_WEIGHTS))
params = [p for p in model.parameters() if p.requires_grad] + [p for p in pre.parameters() if p.requires_grad]
runner = dp.DataParallelTAV(model, pre, crit, FusedAdamW(params, lr=1e-5, weight_decay=1e-4), clip=1.0)
inputs, labels = syn.make_batch("C2", B=B)
inputs = [{k: v.cuda() for k, v in d.items()} for d in inputs]
labels = labels.cuda()
pre.static_keep_count, model.static_keep_count = 104, 1568 - 104
for i in range(steps):
    loss = runner.train_step(inputs, labels, 1, "train")
torch.cuda.synchronize()
print("ok loss %.4f" % loss.item())
