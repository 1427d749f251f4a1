"""Diagnostic probe (not a test): runs the tcgen05 GEMM and the attention kernels over a grid of configurations and
prints error statistics instead of asserting, so a single GPU call shows which operand-major / epilogue / shape
combinations are wrong.  Usage on the GPU box: python tools/probe_kernels.py > gpurun_out/probe.log"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_modal_emotion_b200 import _lib as L  # noqa: E402


def stats(name, got, ref):
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs()
    rel = err.norm() / ref.norm().clamp_min(1e-30)
    print("%-58s max_abs %.3e rel_l2 %.3e ref_absmax %.3e nan %d" % (name, err.max().item(), rel.item(),
                                                                    ref.abs().max().item(), int(torch.isnan(got).sum())),
          flush=True)
    return rel.item()


def probe_gemm():
    dev = "cuda"
    g = torch.Generator(device="cpu").manual_seed(0)
    for (M, N, K) in [(128, 128, 64), (128, 256, 128), (256, 256, 256), (300, 768, 768), (5168, 2304, 768), (1000, 3072, 776)]:
        for a_mn in (False, True):
            for b_mn in (False, True):
                for bn in (128, 256):
                    A = (torch.randn(M, K, generator=g) * 0.5).to(dev).bfloat16()
                    B = (torch.randn(N, K, generator=g) * 0.5).to(dev).bfloat16()
                    ref = A.float() @ B.float().t()
                    Ain = A.t().contiguous() if a_mn else A
                    Bin = B.t().contiguous() if b_mn else B
                    if (a_mn and M % 8) or (b_mn and N % 8) or (not a_mn and K % 8) or (not b_mn and K % 8):
                        continue
                    out = torch.full((M, N), float("nan"), device=dev, dtype=torch.float32)
                    try:
                        L.gemm(Ain, Bin, out, M=M, N=N, K=K, a_mn=a_mn, b_mn=b_mn, block_n=bn)
                        torch.cuda.synchronize()
                        stats("gemm M%d N%d K%d a_mn%d b_mn%d bn%d" % (M, N, K, a_mn, b_mn, bn), out, ref)
                    except Exception as e:  # noqa: BLE001
                        print("gemm M%d N%d K%d a_mn%d b_mn%d bn%d EXC %s" % (M, N, K, a_mn, b_mn, bn, e), flush=True)
                        return False
    # split-K accumulate (wgrad shape)
    M, N, K = 768, 768, 5168
    A = (torch.randn(M, K, generator=g) * 0.5).to(dev).bfloat16()
    B = (torch.randn(N, K, generator=g) * 0.5).to(dev).bfloat16()
    ref = A.float() @ B.float().t()
    out = torch.zeros((M, N), device=dev, dtype=torch.float32)
    L.gemm(A.t().contiguous(), B.t().contiguous(), out, M=M, N=N, K=K, a_mn=True, b_mn=True, accumulate=True, k_splits=8)
    torch.cuda.synchronize()
    stats("gemm wgrad split-K 8", out, ref)
    # timing of the forward shapes
    for (M, N, K) in [(5168, 2304, 768), (5168, 3072, 768), (5168, 768, 3072), (23424, 3072, 768), (8192, 8192, 8192)]:
        A = torch.randn(M, K, device=dev).bfloat16()
        B = torch.randn(N, K, device=dev).bfloat16()
        out = torch.empty((M, N), device=dev, dtype=torch.bfloat16)
        for bn in (128, 256):
            for _ in range(3):
                L.gemm(A, B, out, M=M, N=N, K=K, block_n=bn)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                L.gemm(A, B, out, M=M, N=N, K=K, block_n=bn)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            print("gemm time M%d N%d K%d bn%d: %.3f ms  %.1f TFLOP/s" % (M, N, K, bn, ms, 2.0 * M * N * K / ms / 1e9), flush=True)
        for _ in range(3):
            torch.matmul(A, B.t())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            torch.matmul(A, B.t())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print("cublas time M%d N%d K%d: %.3f ms  %.1f TFLOP/s" % (M, N, K, ms, 2.0 * M * N * K / ms / 1e9), flush=True)
    return True


def probe_attn():
    dev = "cuda"
    g = torch.Generator(device="cpu").manual_seed(1)
    for (B, S, nh) in [(2, 64, 2), (2, 185, 12), (3, 323, 12), (1, 1464, 12)]:
        for use_bias in (False, True):
            H = nh * 64
            qkv = (torch.randn(B, S, 3 * H, generator=g)).to(dev).bfloat16()
            bias = None
            if use_bias:
                bias = torch.zeros(B, S)
                bias[:, S - S // 3:] = -65504.0
                bias = bias.to(dev)
            q, k, v = qkv[..., :H], qkv[..., H:2 * H], qkv[..., 2 * H:]
            o = torch.empty(B, S, H, device=dev, dtype=torch.bfloat16)
            lse = torch.empty(B, nh, S, device=dev, dtype=torch.float32)
            L.attn_fwd(q, k, v, o, lse, B=B, S=S, nh=nh, ld_qkv=3 * H, ld_o=H, key_bias=bias)
            torch.cuda.synchronize()
            qf = q.float().view(B, S, nh, 64).transpose(1, 2).detach().requires_grad_(True)
            kf = k.float().view(B, S, nh, 64).transpose(1, 2).detach().requires_grad_(True)
            vf = v.float().view(B, S, nh, 64).transpose(1, 2).detach().requires_grad_(True)
            sc = qf @ kf.transpose(-1, -2) * 0.125
            if bias is not None:
                sc = sc + bias[:, None, None, :]
            P = torch.softmax(sc, dim=-1)
            ref = (P @ vf).transpose(1, 2).reshape(B, S, H)
            tag = "attn B%d S%d nh%d bias%d" % (B, S, nh, use_bias)
            stats(tag + " fwd O", o, ref)
            stats(tag + " fwd lse", lse, torch.logsumexp(sc, dim=-1))
            do = torch.randn(B, S, H, generator=g).to(dev).bfloat16()
            ref.backward(do.float())
            dqkv = torch.full((B, S, 3 * H), float("nan"), device=dev, dtype=torch.bfloat16)
            delta = torch.empty(B, nh, S, device=dev, dtype=torch.float32)
            L.attn_bwd(q, k, v, o, do, lse, delta, dqkv[..., :H], dqkv[..., H:2 * H], dqkv[..., 2 * H:], B=B, S=S, nh=nh,
                       ld_qkv=3 * H, ld_o=H, ld_dqkv=3 * H, key_bias=bias)
            torch.cuda.synchronize()
            stats(tag + " bwd dQ", dqkv[..., :H], qf.grad.transpose(1, 2).reshape(B, S, H))
            stats(tag + " bwd dK", dqkv[..., H:2 * H], kf.grad.transpose(1, 2).reshape(B, S, H))
            stats(tag + " bwd dV", dqkv[..., 2 * H:], vf.grad.transpose(1, 2).reshape(B, S, H))
    # timing at the MELD fusion shape and the VideoMAE shape
    for (B, S, nh) in [(16, 323, 12), (16, 1464, 12)]:
        H = nh * 64
        qkv = torch.randn(B, S, 3 * H, device=dev).bfloat16()
        q, k, v = qkv[..., :H], qkv[..., H:2 * H], qkv[..., 2 * H:]
        o = torch.empty(B, S, H, device=dev, dtype=torch.bfloat16)
        do = torch.randn(B, S, H, device=dev).bfloat16()
        lse = torch.empty(B, nh, S, device=dev, dtype=torch.float32)
        delta = torch.empty_like(lse)
        dqkv = torch.empty_like(qkv)
        fl = 4.0 * B * nh * S * S * 64
        for name, fn, mult in (
            ("fwd", lambda: L.attn_fwd(q, k, v, o, lse, B=B, S=S, nh=nh, ld_qkv=3 * H, ld_o=H), 1.0),
            ("bwd", lambda: L.attn_bwd(q, k, v, o, do, lse, delta, dqkv[..., :H], dqkv[..., H:2 * H], dqkv[..., 2 * H:], B=B,
                                       S=S, nh=nh, ld_qkv=3 * H, ld_o=H, ld_dqkv=3 * H), 2.5)):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            print("attn %s B%d S%d: %.3f ms  %.1f TFLOP/s (algorithmic)" % (name, B, S, ms, fl * mult / ms / 1e9), flush=True)


if __name__ == "__main__":
    L.require_device()
    print("device", torch.cuda.get_device_name(0), "sms", L.lib().tavk_sm_count(), flush=True)
    t0 = time.time()
    which = sys.argv[1:] or ["gemm", "attn"]
    if "attn" in which:
        probe_attn()
    if "gemm" in which:
        probe_gemm()
    print("probe done in %.1fs" % (time.time() - t0))
