"""Stand-alone fwd+bwd time of each independent sub-graph of the TAV step (CUDA-graph replay, B200): the three encoders
of TAVForMAE, the fusion encoder, PreFormer, and the optimiser step.  Their sum against the step time shows how much
the branch streams (tav.branch_streams) can hide.   Usage: python tools/branch_times.py [B=16] [cfg=C2]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multi_modal_emotion_b200 import engine, hf_adapters as hf, synthetic as syn, tav  # noqa: E402
from multi_modal_emotion_b200.optim import FusedAdamW  # noqa: E402


def timed(fn, reps=5):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    cfg = sys.argv[2] if len(sys.argv) > 2 else "C2"
    tav.set_encoder_variant("baseline")
    tav.branch_streams = False
    torch.manual_seed(0)
    C = syn.CONFIGS[cfg]["C"]
    model = tav.TAVForMAE({"output_dim": C, "dropout": 0.4, "learn_PosEmbeddings": True, "num_layers": 12}).cuda().train()
    pre = tav.PreFormer().cuda().train()
    inputs, labels = syn.make_batch(cfg, seed=1, B=B)
    inputs = [{k: v.cuda() for k, v in d.items()} for d in inputs]
    ids, tmask = inputs[0]["input_ids"], inputs[0]["attention_mask"]
    wav, amask = inputs[1]["audio_features"], inputs[1]["attention_mask"]
    vid, vmask = inputs[2]["visual_embeds"], inputs[2]["attention_mask"]
    keep = int(vmask[0].sum().item())
    S = syn.fused_len(cfg)
    out = {}

    def bw(y):
        y.backward(torch.ones_like(y) / y.numel())

    out["videomae (TAVForMAE, S=1464)"] = timed(lambda: bw(hf.run_videomae(model.videomae, vid, vmask, 1568 - keep)))
    out["wav2vec2 (conv front-end + pos-conv + 12 layers)"] = timed(lambda: bw(hf.run_wav2vec2(model.wav2vec2, wav)))
    out["roberta (embeddings + 12 layers)"] = timed(lambda: bw(hf.run_roberta(model.bert, ids, tmask)[1]))
    x = torch.randn(B, S, 768, device="cuda", requires_grad=True)
    Ta = syn.conv_frames(syn.CONFIGS[cfg]["L"])
    mask = syn.reference_masks(B, syn.CONFIGS[cfg]["T"], Ta, syn.CONFIGS[cfg]["K"], torch.full((B,), syn.CONFIGS[cfg]["T"]),
                               torch.full((B,), Ta)).cuda()
    out["fusion encoder (12 layers, S=%d)" % S] = timed(lambda: bw(model.random_mae_encoder(x, mask)))
    pre.static_keep_count = keep

    def pre_fn():
        t, _, _ = pre(input_ids=ids, audio_features=wav, video_embeds=vid, text_mask=tmask, audio_mask=amask,
                      visual_mask=vmask, device="cuda", train=True)
        bw(t)

    out["PreFormer (embeddings, conv front-end, patch embed)"] = timed(pre_fn)
    params = [p for p in list(model.parameters()) + list(pre.parameters()) if p.requires_grad and p.grad is not None]
    opt = FusedAdamW(params, lr=1e-5, weight_decay=1e-4)
    opt.step(max_grad_norm=1.0)
    out["clip + AdamW (%.0f M parameters)" % (opt.flat.numel / 1e6)] = timed(lambda: opt.step(max_grad_norm=1.0))
    tot = sum(out.values())
    print("# stand-alone fwd+bwd per sub-graph, B=%d %s (CUDA-graph replays)" % (B, cfg))
    for k, v in out.items():
        print("%-60s %8.3f ms" % (k, v))
    print("%-60s %8.3f ms" % ("sum", tot))


if __name__ == "__main__":
    main()
