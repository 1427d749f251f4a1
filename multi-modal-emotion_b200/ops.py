"""``torch.ops.tavk.*`` — the kernel library as torch custom operators ("Python side registers
torch.library custom ops + register_autograd, so loss.backward() drives the backward kernels").

Each operator is a thin shim over one or two C-ABI entry points of libtavk.so (include/tavk.h): it allocates the outputs
as torch tensors, passes raw pointers and torch's current stream, and registers a fake (meta) implementation and an
autograd formula whose backward is itself a custom operator.  They are the per-kernel public surface (usable from any
PyTorch code, traceable, ``torch.library.opcheck``-clean); the drop-in modules reach the same entry points through
``engine.py``, which additionally fuses a whole encoder stack into ONE autograd node and accumulates parameter gradients
in place — something a per-op graph cannot express.

  tavk::layer_norm(x, weight, bias, eps)         nn.LayerNorm over the last dim   (reference models/tav.py:486-490)
  tavk::mean_pool(x)                             x.mean(dim=1), x [B,S,H]         (models/tav.py:478,481,488)
  tavk::small_linear(x, w, b)                    fp32 F.linear, launch-bound sizes (models/tav.py:499)
  tavk::softmax_ce(logits, target, weight)       (sum_i w_yi l_i, sum_i w_yi)     (utils/global_functions.py:63-83)
  tavk::attention(qkv, heads, key_bias)          softmax(QK^T/sqrt(d) + bias) V from a packed [B,S,3H] bf16 tensor
                                                 (utils/TAVFormer.py:357-387, :60-86), head_dim 64
  tavk::linear(x, w, b)                          F.linear on the tcgen05 GEMM: bf16 operands, fp32 accumulate/out; backward
                                                 = the dgrad / wgrad forms of the same entry point (utils/TAVFormer.py:348-350,
                                                 :404,:422,:434; models/tav.py:363,478)
  tavk::embed_add(x, idx, table)                 x + table[idx]                   (models/tav.py:474)
  tavk::dropout(x, p, seed, counter)             nn.Dropout(p) training mode, counter-based generator (models/tav.py:497-498)
  tavk::adamw_step(p, m, v, g, ...)              clip_grad_norm_ + AdamW.step over flat buffers, device-resident clock
                                                 (train_model/tav_train.py:61-63,148-149); mutates its arguments

There is no CPU implementation: calling one with CPU tensors raises (the fake implementations only propagate shapes)."""
from typing import Optional, Tuple

import torch

from . import _lib as L

_F32, _BF16 = torch.float32, torch.bfloat16


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("tavk operators run on the sm_100a kernel library only (got a %s tensor; there is no CPU "
                               "fallback)" % t.device.type)


# ------------------------------------------------------------------------------------------------ layer_norm
@torch.library.custom_op("tavk::layer_norm_fwd", mutates_args=())
def layer_norm_fwd(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    _need_cuda(x, weight, bias)
    x2 = x.contiguous().view(-1, x.shape[-1]).float()
    _, y, mean, rstd = L.layernorm_fwd(x2, weight.float().contiguous(), bias.float().contiguous(), eps, want_bf16=False, want_f32=True)
    return y.view(x.shape), mean, rstd


@layer_norm_fwd.register_fake
def _(x, weight, bias, eps):
    rows = x.numel() // x.shape[-1]
    return x.new_empty(x.shape, dtype=_F32), x.new_empty((rows,), dtype=_F32), x.new_empty((rows,), dtype=_F32)


@torch.library.custom_op("tavk::layer_norm_bwd", mutates_args=())
def layer_norm_bwd(dy: torch.Tensor, x: torch.Tensor, mean: torch.Tensor, rstd: torch.Tensor, weight: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    _need_cuda(dy, x)
    H = x.shape[-1]
    x2 = x.contiguous().view(-1, H).float()
    dw = torch.zeros((H,), dtype=_F32, device=x.device)
    db = torch.zeros((H,), dtype=_F32, device=x.device)
    dx, _ = L.layernorm_bwd(dy.contiguous().view(-1, H).float(), x2, mean, rstd, weight.float().contiguous(), dw, db)
    return dx.view(x.shape), dw, db


@layer_norm_bwd.register_fake
def _(dy, x, mean, rstd, weight):
    H = x.shape[-1]
    return x.new_empty(x.shape, dtype=_F32), x.new_empty((H,), dtype=_F32), x.new_empty((H,), dtype=_F32)


def _ln_setup(ctx, inputs, output):
    x, weight, _, _ = inputs
    _, mean, rstd = output
    ctx.save_for_backward(x, weight, mean, rstd)


def _ln_backward(ctx, dy, _dmean, _drstd):
    x, weight, mean, rstd = ctx.saved_tensors
    dx, dw, db = layer_norm_bwd(dy, x, mean, rstd, weight)
    return dx, dw, db, None


layer_norm_fwd.register_autograd(_ln_backward, setup_context=_ln_setup)


def layer_norm(x, weight, bias, eps=1e-5):
    return layer_norm_fwd(x, weight, bias, eps)[0]


# ------------------------------------------------------------------------------------------------ mean_pool
@torch.library.custom_op("tavk::mean_pool", mutates_args=())
def mean_pool(x: torch.Tensor) -> torch.Tensor:
    _need_cuda(x)
    B, S, H = x.shape
    x = x.contiguous().float()
    y = torch.empty((B, H), dtype=_F32, device=x.device)
    L.call("tavk_mean_pool_fwd", x.data_ptr(), y.data_ptr(), B, S, H)
    return y


@mean_pool.register_fake
def _(x):
    return x.new_empty((x.shape[0], x.shape[2]), dtype=_F32)


@torch.library.custom_op("tavk::mean_pool_bwd", mutates_args=())
def mean_pool_bwd(dy: torch.Tensor, S: int) -> torch.Tensor:
    _need_cuda(dy)
    B, H = dy.shape
    dy = dy.contiguous().float()
    dx = torch.empty((B, S, H), dtype=_F32, device=dy.device)
    L.call("tavk_mean_pool_bwd", dy.data_ptr(), dx.data_ptr(), None, B, S, H)
    return dx


@mean_pool_bwd.register_fake
def _(dy, S):
    return dy.new_empty((dy.shape[0], S, dy.shape[1]), dtype=_F32)


mean_pool.register_autograd(lambda ctx, dy: mean_pool_bwd(dy, ctx.S),
                            setup_context=lambda ctx, inputs, output: setattr(ctx, "S", inputs[0].shape[1]))


# ------------------------------------------------------------------------------------------------ small_linear
@torch.library.custom_op("tavk::small_linear", mutates_args=())
def small_linear(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor]) -> torch.Tensor:
    _need_cuda(x, w, b)
    x = x.contiguous().float()
    w = w.contiguous().float()
    M, K = x.shape
    N = w.shape[0]
    y = torch.empty((M, N), dtype=_F32, device=x.device)
    L.call("tavk_small_linear_fwd", x.data_ptr(), w.data_ptr(), L._ptr(None if b is None else b.contiguous().float()),
           y.data_ptr(), M, N, K)
    return y


@small_linear.register_fake
def _(x, w, b):
    return x.new_empty((x.shape[0], w.shape[0]), dtype=_F32)


@torch.library.custom_op("tavk::small_linear_bwd", mutates_args=())
def small_linear_bwd(dy: torch.Tensor, x: torch.Tensor, w: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    _need_cuda(dy, x, w)
    x, w, dy = x.contiguous().float(), w.contiguous().float(), dy.contiguous().float()
    M, K = x.shape
    N = w.shape[0]
    dx = torch.empty_like(x)
    L.call("tavk_small_linear_bwd_x", dy.data_ptr(), w.data_ptr(), dx.data_ptr(), M, N, K, 0)
    dw = torch.zeros_like(w)
    db = torch.zeros((N,), dtype=_F32, device=x.device)
    L.call("tavk_small_linear_bwd_w", dy.data_ptr(), x.data_ptr(), dw.data_ptr(), db.data_ptr(), M, N, K)
    return dx, dw, db


@small_linear_bwd.register_fake
def _(dy, x, w):
    return x.new_empty(x.shape, dtype=_F32), w.new_empty(w.shape, dtype=_F32), w.new_empty((w.shape[0],), dtype=_F32)


def _sl_setup(ctx, inputs, output):
    x, w, b = inputs
    ctx.save_for_backward(x, w)
    ctx.has_b = b is not None


def _sl_backward(ctx, dy):
    x, w = ctx.saved_tensors
    dx, dw, db = small_linear_bwd(dy, x, w)
    return dx, dw, (db if ctx.has_b else None)


small_linear.register_autograd(_sl_backward, setup_context=_sl_setup)


# ------------------------------------------------------------------------------------------------ softmax_ce
@torch.library.custom_op("tavk::softmax_ce", mutates_args=())
def softmax_ce(logits: torch.Tensor, target: torch.Tensor, weight: Optional[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(numerator, denominator, probabilities): weighted-mean CE = numerator / denominator."""
    _need_cuda(logits, target, weight)
    B, C = logits.shape
    logits = logits.contiguous().float()
    target = target.contiguous().long()
    w = None if weight is None else weight.contiguous().float()
    probs = torch.empty((B, C), dtype=_F32, device=logits.device)
    nd = torch.empty((2,), dtype=_F32, device=logits.device)
    L.call("tavk_softmax_ce_fwd", logits.data_ptr(), target.data_ptr(), L._ptr(w), probs.data_ptr(), nd.data_ptr(),
           nd.data_ptr() + 4, B, C)
    return nd[0].clone(), nd[1].clone(), probs


@softmax_ce.register_fake
def _(logits, target, weight):
    return logits.new_empty((), dtype=_F32), logits.new_empty((), dtype=_F32), logits.new_empty(logits.shape, dtype=_F32)


@torch.library.custom_op("tavk::softmax_ce_bwd", mutates_args=())
def softmax_ce_bwd(dnum: torch.Tensor, probs: torch.Tensor, target: torch.Tensor, weight: Optional[torch.Tensor]) -> torch.Tensor:
    _need_cuda(dnum, probs, target, weight)
    B, C = probs.shape
    w = None if weight is None else weight.contiguous().float()
    dl = torch.empty_like(probs)
    g = dnum.contiguous().float().view(1)
    L.call("tavk_softmax_ce_bwd", probs.data_ptr(), target.contiguous().long().data_ptr(), L._ptr(w), g.data_ptr(),
           dl.data_ptr(), B, C)
    return dl


@softmax_ce_bwd.register_fake
def _(dnum, probs, target, weight):
    return probs.new_empty(probs.shape)


def _ce_setup(ctx, inputs, output):
    _, target, weight = inputs
    ctx.save_for_backward(output[2], target, *([] if weight is None else [weight]))


def _ce_backward(ctx, dnum, _dden, _dprobs):
    probs, target, *w = ctx.saved_tensors
    return softmax_ce_bwd(dnum, probs, target, w[0] if w else None), None, None


softmax_ce.register_autograd(_ce_backward, setup_context=_ce_setup)


def cross_entropy(logits, target, weight=None):
    """nn.CrossEntropyLoss(weight)(logits, target), mean reduction."""
    num, den, _ = softmax_ce(logits, target, weight)
    return num / den


# ------------------------------------------------------------------------------------------------ attention
@torch.library.custom_op("tavk::attention", mutates_args=())
def attention(qkv: torch.Tensor, heads: int, key_bias: Optional[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
    """qkv bf16 [B,S,3H] = q | k | v (head-major inside each), head_dim 64; key_bias f32 [B,S] or None -> (o bf16 [B,S,H],
    lse f32 [B,heads,S])."""
    _need_cuda(qkv, key_bias)
    B, S, H3 = qkv.shape
    H = H3 // 3
    if H // heads != 64 or qkv.dtype != _BF16:
        raise ValueError("tavk::attention needs a bf16 packed qkv with head_dim 64")
    qkv = qkv.contiguous()
    o = torch.empty((B, S, H), dtype=_BF16, device=qkv.device)
    lse = torch.empty((B, heads, S), dtype=_F32, device=qkv.device)
    L.attn_fwd(qkv[..., :H], qkv[..., H:2 * H], qkv[..., 2 * H:], o, lse, B=B, S=S, nh=heads, ld_qkv=3 * H, ld_o=H,
               key_bias=None if key_bias is None else key_bias.contiguous().float(), scale=0.125)
    return o, lse


@attention.register_fake
def _(qkv, heads, key_bias):
    B, S, H3 = qkv.shape
    return qkv.new_empty((B, S, H3 // 3)), qkv.new_empty((B, heads, S), dtype=_F32)


@torch.library.custom_op("tavk::attention_bwd", mutates_args=())
def attention_bwd(d_o: torch.Tensor, qkv: torch.Tensor, o: torch.Tensor, lse: torch.Tensor, heads: int, key_bias: Optional[torch.Tensor]) -> torch.Tensor:
    _need_cuda(d_o, qkv, o, lse, key_bias)
    B, S, H3 = qkv.shape
    H = H3 // 3
    qkv = qkv.contiguous()
    d_o = d_o.contiguous().to(_BF16)
    dqkv = torch.empty_like(qkv)
    delta = torch.empty((B, heads, S), dtype=_F32, device=qkv.device)
    L.attn_bwd(qkv[..., :H], qkv[..., H:2 * H], qkv[..., 2 * H:], o, d_o, lse, delta, dqkv[..., :H], dqkv[..., H:2 * H],
               dqkv[..., 2 * H:], B=B, S=S, nh=heads, ld_qkv=3 * H, ld_o=H, ld_dqkv=3 * H,
               key_bias=None if key_bias is None else key_bias.contiguous().float(), scale=0.125)
    return dqkv


@attention_bwd.register_fake
def _(d_o, qkv, o, lse, heads, key_bias):
    return qkv.new_empty(qkv.shape)


def _attn_setup(ctx, inputs, output):
    qkv, heads, key_bias = inputs
    ctx.save_for_backward(qkv, output[0], output[1], *([] if key_bias is None else [key_bias]))
    ctx.heads = heads


def _attn_backward(ctx, d_o, _dlse):
    qkv, o, lse, *kb = ctx.saved_tensors
    return attention_bwd(d_o, qkv, o, lse, ctx.heads, kb[0] if kb else None), None, None


attention.register_autograd(_attn_backward, setup_context=_attn_setup)


# ------------------------------------------------------------------------------------------------ linear (tcgen05 GEMM)
@torch.library.custom_op("tavk::linear", mutates_args=())
def linear(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor]) -> torch.Tensor:
    _need_cuda(x, w, b)
    K, N = x.shape[-1], w.shape[0]
    x2 = x.contiguous().view(-1, K)
    M = x2.shape[0]
    x_bf = x2 if x2.dtype == _BF16 else L.cast_bf16(x2.float())
    w_bf = w if w.dtype == _BF16 else L.cast_bf16(w.detach().float())
    y = torch.empty((M, N), dtype=_F32, device=x.device)
    L.gemm(x_bf, w_bf.contiguous(), y, M=M, N=N, K=K, bias=None if b is None else b.float().contiguous())
    return y.view(*x.shape[:-1], N)


@linear.register_fake
def _(x, w, b):
    return x.new_empty((*x.shape[:-1], w.shape[0]), dtype=_F32)


@torch.library.custom_op("tavk::linear_bwd", mutates_args=())
def linear_bwd(dy: torch.Tensor, x: torch.Tensor, w: torch.Tensor, has_bias: bool) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(dx, dw, db): dgrad (A = dY K-major, B = W MN-major), wgrad (both MN-major, split-K, accumulate) and the bias column sum."""
    _need_cuda(dy, x, w)
    K, N = x.shape[-1], w.shape[0]
    x2 = x.contiguous().view(-1, K)
    M = x2.shape[0]
    x_bf = x2 if x2.dtype == _BF16 else L.cast_bf16(x2.float())
    w_bf = (w if w.dtype == _BF16 else L.cast_bf16(w.detach().float())).contiguous()
    dy2 = dy.contiguous().view(M, N).float()
    dy_bf = L.cast_bf16(dy2)
    dx = torch.empty((M, K), dtype=_F32, device=x.device)
    L.gemm(dy_bf, w_bf, dx, M=M, N=K, K=N, b_mn=True)
    dw = torch.zeros((N, K), dtype=_F32, device=x.device)
    tiles = ((N + 127) // 128) * ((K + 255) // 256)
    L.gemm(dy_bf, x_bf, dw, M=N, N=K, K=M, a_mn=True, b_mn=True, accumulate=True,
           k_splits=max(1, min(148 // max(tiles, 1), (M + 511) // 512)))
    db = torch.zeros((N if has_bias else 0,), dtype=_F32, device=x.device)
    if has_bias:
        L.colsum(dy2, db, M=M, N=N)
    return dx.view(x.shape), dw, db


@linear_bwd.register_fake
def _(dy, x, w, has_bias):
    return x.new_empty(x.shape, dtype=_F32), w.new_empty(w.shape, dtype=_F32), w.new_empty((w.shape[0] if has_bias else 0,), dtype=_F32)


def _linear_setup(ctx, inputs, output):
    x, w, b = inputs
    ctx.save_for_backward(x, w)
    ctx.has_b = b is not None


def _linear_backward(ctx, dy):
    x, w = ctx.saved_tensors
    dx, dw, db = linear_bwd(dy, x, w, ctx.has_b)
    return dx.to(x.dtype), dw.to(w.dtype), (db if ctx.has_b else None)


linear.register_autograd(_linear_backward, setup_context=_linear_setup)


# ------------------------------------------------------------------------------------------------ embed_add
@torch.library.custom_op("tavk::embed_add", mutates_args=())
def embed_add(x: torch.Tensor, idx: torch.Tensor, table: torch.Tensor) -> torch.Tensor:
    _need_cuda(x, idx, table)
    B, S, H = x.shape
    x = x.contiguous().float()
    y = torch.empty_like(x)
    L.call("tavk_embed_add_fwd", x.data_ptr(), idx.contiguous().long().data_ptr(), table.float().contiguous().data_ptr(),
           y.data_ptr(), B * S, H, table.shape[0])
    return y


@embed_add.register_fake
def _(x, idx, table):
    return x.new_empty(x.shape, dtype=_F32)


@torch.library.custom_op("tavk::embed_add_bwd", mutates_args=())
def embed_add_bwd(dy: torch.Tensor, idx: torch.Tensor, n_embed: int) -> torch.Tensor:
    _need_cuda(dy, idx)
    B, S, H = dy.shape
    dy = dy.contiguous().float()
    dt = torch.zeros((n_embed, H), dtype=_F32, device=dy.device)
    L.call("tavk_embed_add_bwd", dy.data_ptr(), idx.contiguous().long().data_ptr(), dt.data_ptr(), B * S, H, n_embed)
    return dt


@embed_add_bwd.register_fake
def _(dy, idx, n_embed):
    return dy.new_empty((n_embed, dy.shape[-1]), dtype=_F32)


def _embed_setup(ctx, inputs, output):
    _, idx, table = inputs
    ctx.save_for_backward(idx)
    ctx.n = table.shape[0]


def _embed_backward(ctx, dy):
    (idx,) = ctx.saved_tensors
    return dy, None, embed_add_bwd(dy, idx, ctx.n)


embed_add.register_autograd(_embed_backward, setup_context=_embed_setup)


# ------------------------------------------------------------------------------------------------ dropout
@torch.library.custom_op("tavk::dropout", mutates_args=())
def dropout(x: torch.Tensor, p: float, seed: int, counter: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(y, keep mask uint8).  ``counter``: device int64 scalar added to the stream position, so a captured CUDA graph draws
    a fresh mask per replay when the caller increments it (engine.dropout does)."""
    _need_cuda(x, counter)
    x = x.contiguous().float()
    y = torch.empty_like(x)
    keep = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    L.call("tavk_dropout", x.data_ptr(), y.data_ptr(), keep.data_ptr(), x.numel(), float(p), int(seed) & 0x7FFFFFFFFFFFFFFF, 0,
           counter.data_ptr())
    return y, keep


@dropout.register_fake
def _(x, p, seed, counter):
    return x.new_empty(x.shape, dtype=_F32), x.new_empty(x.shape, dtype=torch.uint8)


@torch.library.custom_op("tavk::dropout_bwd", mutates_args=())
def dropout_bwd(dy: torch.Tensor, keep: torch.Tensor, p: float) -> torch.Tensor:
    _need_cuda(dy, keep)
    dy = dy.contiguous().float()
    dx = torch.empty_like(dy)
    L.call("tavk_dropout_bwd", dy.data_ptr(), keep.data_ptr(), dx.data_ptr(), dy.numel(), float(p))
    return dx


@dropout_bwd.register_fake
def _(dy, keep, p):
    return dy.new_empty(dy.shape, dtype=_F32)


def _drop_setup(ctx, inputs, output):
    ctx.save_for_backward(output[1])
    ctx.p = inputs[1]


def _drop_backward(ctx, dy, _dkeep):
    (keep,) = ctx.saved_tensors
    return dropout_bwd(dy, keep, ctx.p), None, None, None


dropout.register_autograd(_drop_backward, setup_context=_drop_setup)


# ------------------------------------------------------------------------------------------------ optimiser step
@torch.library.custom_op("tavk::adamw_step", mutates_args=("p", "m", "v", "g", "p_bf16", "step", "hyper", "sqnorm"))
def adamw_step(p: torch.Tensor, m: torch.Tensor, v: torch.Tensor, g: torch.Tensor, p_bf16: torch.Tensor, step: torch.Tensor,
               hyper: torch.Tensor, sqnorm: torch.Tensor, beta1: float, beta2: float, eps: float, weight_decay: float,
               max_norm: float, grad_prescale: float) -> None:
    """clip_grad_norm_(max_norm) + AdamW over flat fp32 buffers, in place; writes the bf16 mirror of the parameters and zeroes
    the gradient.  ``step`` (int32[1]) and ``hyper`` (f32[4], hyper[0] = learning rate) are the device-resident optimiser
    clock: the op advances them, so it is CUDA-graph safe (include/tavk.h tavk_adamw_prep / tavk_adamw_dev)."""
    _need_cuda(p, m, v, g, p_bf16, step, hyper, sqnorm)
    n = p.numel()
    use_clip = max_norm > 0
    L.call("tavk_adamw_prep", step.data_ptr(), hyper.data_ptr(), sqnorm.data_ptr() if use_clip else None, beta1, beta2)
    if use_clip:
        L.call("tavk_grad_sqnorm", g.data_ptr(), n, sqnorm.data_ptr())
    L.call("tavk_adamw_dev", p.data_ptr(), m.data_ptr(), v.data_ptr(), g.data_ptr(), p_bf16.data_ptr(), n, hyper.data_ptr(),
           beta1, beta2, eps, weight_decay, sqnorm.data_ptr() if use_clip else None, max_norm if use_clip else 0.0,
           grad_prescale, 1)


OPS = ("layer_norm_fwd", "layer_norm_bwd", "mean_pool", "mean_pool_bwd", "small_linear", "small_linear_bwd", "softmax_ce",
       "softmax_ce_bwd", "attention", "attention_bwd", "linear", "linear_bwd", "embed_add", "embed_add_bwd", "dropout",
       "dropout_bwd", "adamw_step")
