"""Device timeline of ONE captured training step (the graph bench.py replays) from CUPTI kernel records: busy time per
kernel family, idle gaps, and concurrency (records on different streams overlap when the step forks).
Usage: python tools/step_timeline.py [B=16] [cfg=C2] [out.txt]"""
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multi_modal_emotion_b200 import dp, synthetic as syn, tav  # noqa: E402
from multi_modal_emotion_b200.losses import NewCrossEntropyLoss  # noqa: E402
from multi_modal_emotion_b200.optim import FusedAdamW  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    cfg = sys.argv[2] if len(sys.argv) > 2 else "C2"
    out_path = sys.argv[3] if len(sys.argv) > 3 else None
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:      # torchrun: one rank per GPU; rank 0 reports its own timeline incl. every ncclAllReduce bucket
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    C = syn.CONFIGS[cfg]["C"]
    tav.set_encoder_variant("baseline")
    torch.manual_seed(0)
    model = tav.TAVForMAE({"output_dim": C, "dropout": 0.4, "learn_PosEmbeddings": True, "num_layers": 12})
    pre = tav.PreFormer()
    pre.load_state_dict(syn.synth_state_dict(pre, seed=1))
    model.load_state_dict(syn.synth_state_dict(model, seed=2))
    model, pre = model.cuda().train(), pre.cuda().train()
    crit = NewCrossEntropyLoss(torch.tensor(syn.MELD_CLASS_WEIGHTS if C == 7 else [0.5, 0.5]), epoch_switch=2)
    params = [p for p in model.parameters() if p.requires_grad] + [p for p in pre.parameters() if p.requires_grad]
    runner = dp.DataParallelTAV(model, pre, crit, FusedAdamW(params, lr=1e-5, weight_decay=1e-4), clip=1.0, use_cuda_graph=True)
    inputs, labels = syn.make_batch(cfg, seed=1234 + rank, B=B)
    inputs = [{k: v.cuda() for k, v in d.items()} for d in inputs]
    labels = labels.cuda()
    runner.train_step(inputs, labels, 1, "train")
    inputs, labels = runner.static_inputs()
    for _ in range(3):
        runner.train_step(inputs, labels, 1, "train")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        runner.train_step(inputs, labels, 1, "train")
    e1.record()
    torch.cuda.synchronize()
    ms_plain = e0.elapsed_time(e1) / 10
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        runner.train_step(inputs, labels, 1, "train")
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
    ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in evs), key=lambda t: t[0])
    span = ks[-1][1] - ks[0][0]
    busy, cnt = collections.defaultdict(float), collections.Counter()
    union, cover_end = 0.0, ks[0][0]
    for s, e, n in ks:
        n = n.replace("tavk::", "").replace("void ", "")[:80]
        busy[n] += e - s
        cnt[n] += 1
        union += max(0.0, e - max(s, cover_end))
        cover_end = max(cover_end, e)
    lines = ["# one captured training step, B=%d %s: plain %.3f ms/step (%.1f samples/s); profiled span %.1f us, GPU busy (union of "
             "records) %.1f us = %.1f%%, sum of kernel durations %.1f us (> busy when streams overlap)" % (
                 B, cfg, ms_plain, B / ms_plain * 1e3, span, union, 100 * union / span, sum(busy.values())),
             "%-82s %6s %10s %7s %9s" % ("kernel", "count", "busy_us", "share", "avg_us")]
    for n in sorted(busy, key=lambda k: -busy[k])[:60]:
        lines.append("%-82s %6d %10.1f %6.1f%% %9.2f" % (n, cnt[n], busy[n], 100 * busy[n] / span, busy[n] / cnt[n]))
    fam = collections.defaultdict(float)
    for n, v in busy.items():
        key = ("gemm" if n.startswith("gemm_bf16") else "attention" if n.startswith("attn") else "layernorm" if n.startswith("layernorm")
               else "adamw+sqnorm" if ("adamw" in n or "sqnorm" in n) else "nccl" if "nccl" in n.lower() else
               "other tavk" if not (n.startswith("at::") or "cutlass" in n or "cudnn" in n or "Memcpy" in n or "Memset" in n) else "non-tavk (ATen/memcpy/library)")
        fam[key] += v
    lines.append("families (share of sum of durations): " + ", ".join("%s %.1f%%" % (k, 100 * v / sum(busy.values())) for k, v in sorted(fam.items(), key=lambda t: -t[1])))
    if world > 1:
        # every collective kernel of the step: when it ran relative to the step and to the last compute kernel of backward
        t0 = ks[0][0]
        nccl = [(s, e, n) for s, e, n in ks if "nccl" in n.lower()]
        comp_before_adam = [e for s, e, n in ks if "nccl" not in n.lower() and "adamw" not in n and "sqnorm" not in n and "adamw_prep" not in n]
        adam = [s for s, e, n in ks if "sqnorm" in n or "adamw_kernel" in n]
        last_bwd = max(x for x in comp_before_adam if not adam or x <= adam[0]) if comp_before_adam else t0
        lines.append("# %d ranks: %d collective kernels on rank 0; last backward compute kernel ends at %.1f us, clip+AdamW starts at %.1f us"
                     % (world, len(nccl), last_bwd - t0, (adam[0] - t0) if adam else -1.0))
        lines.append("%10s %10s %10s  %s" % ("start_us", "dur_us", "end_us", "collective (exposed = past the last backward kernel)"))
        for s, e, n in nccl:
            lines.append("%10.1f %10.1f %10.1f  %s%s" % (s - t0, e - s, e - t0, n[:60], "   EXPOSED %.1f us" % (e - last_bwd) if e > last_bwd else ""))
        lines.append("# sum of collective kernel time %.1f us; exposed tail (last collective end - last backward end) %.1f us" % (
            sum(e - s for s, e, n in nccl), max([e for s, e, n in nccl] + [last_bwd]) - last_bwd))
    text = "\n".join(lines)
    if rank == 0:
        print(text)
        if out_path:
            open(out_path, "w").write(text + "\n")
    if world > 1:
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)     # captured graphs hold the communicator (see bench.py)


if __name__ == "__main__":
    main()
