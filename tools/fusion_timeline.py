"""Device timeline of the fusion block alone (12-layer fusion VideoMAEEncoder, fwd+bwd, one CUDA-graph replay) from
CUPTI kernel records (torch.profiler): per kernel family busy time and the idle gap that follows each launch, so the
share of the block that is math, memory passes and inter-kernel gaps is measured, not guessed.

Usage: python tools/fusion_timeline.py [B=16] [S=323] [out.txt]     (no nsys in this image: kineto gives the same
per-kernel start/duration records)"""
import collections
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from transformers import VideoMAEConfig  # noqa: E402

from multi_modal_emotion_b200 import synthetic as syn  # noqa: E402
from multi_modal_emotion_b200.tavformer import VideoMAEEncoder  # noqa: E402


def short(name):
    name = name.replace("tavk::", "")
    for cut in ("(", "<unnamed>::"):
        pass
    if name.startswith("void "):
        name = name[5:]
    return name[:70]


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 323
    out_path = sys.argv[3] if len(sys.argv) > 3 else None
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"]
    enc = VideoMAEEncoder(VideoMAEConfig(), 12)
    enc.load_state_dict(syn.synth_state_dict(enc, seed=3))
    enc = enc.cuda()
    x = torch.randn(B, S, 768, device="cuda", requires_grad=True)
    T, K = 70, 104
    Ta = S - T - K
    mask = syn.reference_masks(B, T, Ta, K, torch.full((B,), T), torch.full((B,), Ta)).cuda()
    go = torch.full((B, S, 768), 1.0 / (B * S * 768), device="cuda")

    def fstep():
        y = enc(x, mask)
        y.backward(go)

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fstep()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fstep()
    for _ in range(3):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms_plain = e0.elapsed_time(e1) / 10
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        g.replay()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
    ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in evs), key=lambda t: t[0])
    lines = []
    if not ks:
        print("no CUDA kernel records captured")
        return
    span = ks[-1][1] - ks[0][0]
    busy = collections.defaultdict(float)
    gap = collections.defaultdict(float)
    cnt = collections.Counter()
    cover_end = ks[0][0]
    union = 0.0
    for i, (s, e, n) in enumerate(ks):
        n = short(n)
        busy[n] += e - s
        cnt[n] += 1
        if i + 1 < len(ks):
            gap[n] += max(0.0, ks[i + 1][0] - max(e, cover_end))
        union += max(0.0, e - max(s, cover_end))
        cover_end = max(cover_end, e)
    flops = 3 * 12 * (14155776 * S + 3072 * S * S) * B
    lines.append("# fusion block alone: B=%d S=%d, one CUDA-graph replay; plain (unprofiled) %.3f ms/replay = %.1f TFLOP/s = %.1f%% of "
                 "measured bf16 burst peak %.0f" % (B, S, ms_plain, flops / ms_plain / 1e9, 100 * flops / ms_plain / 1e9 / peak, peak))
    lines.append("# under the profiler: %d device records, span %.1f us, GPU busy (union) %.1f us = %.1f%%, idle gaps %.1f us" % (
        len(ks), span, union, 100 * union / span, span - union))
    lines.append("%-72s %6s %10s %7s %10s %9s" % ("kernel", "count", "busy_us", "share", "gap_after", "avg_us"))
    for n in sorted(busy, key=lambda k: -busy[k]):
        lines.append("%-72s %6d %10.1f %6.1f%% %10.1f %9.2f" % (n, cnt[n], busy[n], 100 * busy[n] / span, gap[n], busy[n] / cnt[n]))
    # one forward layer and one backward layer in launch order (layer 6 of 12)
    per_layer_f = None
    names = [short(n) for _, _, n in ks]
    ln_idx = [i for i, n in enumerate(names) if n.startswith("layernorm_fwd")]
    if len(ln_idx) >= 14:
        a, b = ln_idx[12], ln_idx[14]
        lines.append("# forward layer 6 in launch order (start offset us, duration us, gap after us)")
        for i in range(a, b):
            s, e, _ = ks[i]
            lines.append("  %8.1f %8.2f %7.2f  %s" % (s - ks[a][0], e - s, ks[i + 1][0] - e, names[i]))
    lb_idx = [i for i, n in enumerate(names) if n.startswith("layernorm_bwd")]
    if len(lb_idx) >= 14:
        a, b = lb_idx[11] + 1, lb_idx[13] + 1
        lines.append("# backward layer (6th from the top) in launch order")
        for i in range(a, min(b, len(ks) - 1)):
            s, e, _ = ks[i]
            lines.append("  %8.1f %8.2f %7.2f  %s" % (s - ks[a][0], e - s, ks[i + 1][0] - e, names[i]))
    text = "\n".join(lines)
    print(text)
    if out_path:
        with open(out_path, "w") as f:
            f.write(text + "\n")


if __name__ == "__main__":
    main()
