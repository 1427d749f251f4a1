"""Transformer-layer engine: forward and backward of a stack of encoder layers executed entirely by libtavk.so kernels.

One ``torch.autograd.Function`` (``EncoderStackFn``) covers every layer family on the TAV path:

  * fusion ``VideoMAELayer`` of the reference (utils/TAVFormer.py:230-271): pre-LN, q/v bias, POST-softmax mask add,
    handled as unmasked attention plus a rank-1 fp32 term routed through the out-projection;
  * HF ``VideoMAELayer`` and ``Wav2Vec2EncoderLayerStableLayerNorm``: pre-LN, no mask;
  * HF ``RobertaLayer`` / ``Wav2Vec2EncoderLayer`` and the reference's custom ``TransformerBlock``
    (utils/TAVFormer.py:93-142): post-LN, optional pre-softmax key bias, optional scrambled head concat.

Numerics: the residual stream, LayerNorm statistics, biases, the rank-1 term and all parameter gradients are fp32;
GEMM / attention operands are bf16 with fp32 accumulation (tcgen05 / mma.sync).  Weights stay fp32 ``nn.Parameter``s
(state_dict compatible); bf16 operand copies ("shadows", QKV packed to [3H,H]) are cached per layer and refreshed when
a parameter's version counter changes or ``invalidate_shadows()`` is called by the fused optimiser."""
import os
from dataclasses import dataclass

import torch

from . import _lib as L

_generation = 0


def invalidate_shadows():
    """Called after an out-of-band parameter update (fused AdamW through raw pointers)."""
    global _generation
    _generation += 1


@dataclass(frozen=True)
class LayerSpec:
    hidden: int = 768
    heads: int = 12
    inter: int = 3072
    pre_ln: bool = True
    eps: float = 1e-12
    mask_mode: str = "none"      # "none" | "key_bias" (pre-softmax additive) | "rank1" (post-softmax additive, Q1)
    scrambled_concat: bool = False  # reference MultiHeadAttention concat quirk (Q5)
    dropout: float = 0.0         # post-LN blocks only: the reference TransformerBlock's three nn.Dropout(p) in training mode
    dropout_salt: int = 0        # distinguishes the random streams of different encoders


# flat parameter order per layer (None allowed for absent biases)
PARAM_SLOTS = ("ln1_w", "ln1_b", "wq", "wk", "wv", "bq", "bk", "bv", "wo", "bo", "ln2_w", "ln2_b", "w1", "b1", "w2", "b2")
N_SLOTS = len(PARAM_SLOTS)


def _flat_shadow(t):
    """(bf16 view of parameter `t` inside its optimiser's flat shadow buffer, offset) or (None, None)."""
    ref = getattr(t, "_tavk_flat", None)
    if ref is None:
        return None, None
    flat, off = ref
    if flat.shadow is None or t.data_ptr() != flat.flat.data_ptr() + 4 * off:
        return None, None          # the parameter was re-pointed elsewhere since FlatParams was built
    return flat.shadow_view(t, off), off


class LayerShadow:
    """bf16 operand copies of one layer's matrices + packed fp32 QKV bias.

    Two regimes.  Before the fused optimiser exists (or for parameters it does not own) the copies are private buffers,
    re-cast whenever a parameter's version counter or the optimiser generation changes.  Once optim.FlatParams owns
    the parameters, the AdamW kernel writes a bf16 mirror of the whole flat buffer in its update pass and the copies
    are views of that mirror (Q, K, V as one [3H, H] view when their slices are adjacent): nothing is cast per step;
    only a direct write to a parameter (version counter) triggers a re-cast of its slice."""

    def __init__(self):
        self.key = None
        self.wqkv = self.wo = self.w1 = self.w2 = self.bqkv = None
        self.own = {}

    def _own(self, name, shape, dev):
        t = self.own.get(name)
        if t is None or t.shape != shape:
            t = self.own[name] = torch.empty(shape, dtype=torch.bfloat16, device=dev)
        return t

    def refresh(self, p, spec):
        H = spec.hidden
        d = dict(zip(PARAM_SLOTS, p))
        dev = d["wq"].device
        mats = ("wq", "wk", "wv", "wo", "w1", "w2")
        views = {n: _flat_shadow(d[n]) for n in mats}
        flat_all = all(v[0] is not None for v in views.values())
        ver = tuple((t.data_ptr(), t._version) if t is not None else None for t in p)
        key = (flat_all, ver) if flat_all else (flat_all, _generation, ver)
        bias_key = (_generation, ver)
        if key != self.key:
            with torch.no_grad():
                if flat_all:
                    # the mirror is kept current by the optimiser (raw-pointer updates do not touch version counters);
                    # re-cast only a slice whose parameter was written directly since the mirror last matched it
                    for n in mats:
                        if d[n]._version != getattr(d[n], "_tavk_mirror_version", None):
                            L.cast_bf16(d[n].detach(), views[n][0])
                            d[n]._tavk_mirror_version = d[n]._version
                    oq, ok, ov = views["wq"][1], views["wk"][1], views["wv"][1]
                    if ok - oq == H * H and ov - ok == H * H:
                        self.wqkv = d["wq"]._tavk_flat[0].shadow[oq:oq + 3 * H * H].view(3 * H, H)
                        self._qkv_views = None
                    else:
                        self.wqkv = self._own("wqkv", (3 * H, H), dev)
                        self._qkv_views = [views[n][0] for n in ("wq", "wk", "wv")]
                    self.wo, self.w1, self.w2 = views["wo"][0], views["w1"][0], views["w2"][0]
                else:
                    self._qkv_views = None
                    self.wqkv = self._own("wqkv", (3 * H, H), dev)
                    self.wo = self._own("wo", (H, H), dev)
                    self.w1 = self._own("w1", (spec.inter, H), dev)
                    self.w2 = self._own("w2", (H, spec.inter), dev)
                    for i, n in enumerate(("wq", "wk", "wv")):
                        L.cast_bf16(d[n].detach(), self.wqkv[i * H:(i + 1) * H])
                    L.cast_bf16(d["wo"].detach(), self.wo)
                    L.cast_bf16(d["w1"].detach(), self.w1)
                    L.cast_bf16(d["w2"].detach(), self.w2)
            self.key = key
        if bias_key != getattr(self, "bias_key", None):
            with torch.no_grad():
                if getattr(self, "_qkv_views", None) is not None:
                    # Q, K, V live apart in the flat buffer (biases in between): gather the three bf16 slices
                    for i, v in enumerate(self._qkv_views):
                        self.wqkv[i * H:(i + 1) * H].copy_(v)
                if self.bqkv is None:
                    self.bqkv = torch.zeros((3 * H,), dtype=torch.float32, device=dev)
                for i, n in enumerate(("bq", "bk", "bv")):
                    if d[n] is not None:
                        self.bqkv[i * H:(i + 1) * H].copy_(d[n].detach())
                    else:
                        self.bqkv[i * H:(i + 1) * H].zero_()
            self.bias_key = bias_key
        return self


def _wgrad_splits(rows_out, cols_out, k_tokens):
    tiles = ((rows_out + 127) // 128) * ((cols_out + 255) // 256)
    sms = 148
    s = max(1, min(sms // max(tiles, 1), (k_tokens + 511) // 512))
    return s


def _wgrad(dy_bf, x_bf, rows_out, cols_out, tokens, out=None, lda=None):
    """dW[rows_out, cols_out] (+)= dY^T · X over `tokens` rows; both operands MN-major views of row-major activations.
    ``out``: an existing fp32 gradient buffer to accumulate into (the parameter's .grad slice); else a new tensor."""
    if out is None:
        out = torch.zeros((rows_out, cols_out), dtype=torch.float32, device=dy_bf.device)
    ks = _wgrad_splits(rows_out, cols_out, tokens)
    L.gemm(dy_bf, x_bf, out, M=rows_out, N=cols_out, K=tokens, a_mn=True, b_mn=True, accumulate=True, k_splits=ks, lda=lda)
    return out


# Gradient sink: when a parameter already owns a contiguous fp32 ``.grad`` (optim.FlatParams points every .grad at a
# slice of one flat buffer, zeroed by the fused AdamW kernel), the stack's backward accumulates straight into it — the
# wgrad GEMMs' red.global.add epilogue, the LayerNorm / column-sum atomics — and returns None for that input, instead
# of materialising a temporary and letting autograd launch one add_ per parameter.  ``grad_written_hook(params)`` is
# called after each layer's kernels are enqueued (dp.GradBuckets uses it to launch bucket all-reduces during backward,
# layer by layer, instead of when autograd gets to the stack's parameters after the whole stack has run; autograd's own
# post-accumulate-grad hooks fire for these parameters too, with the None gradient — GradBuckets counts each once).
grad_sink_enabled = True
grad_written_hook = None
# Row-sparse gradients: an embedding backward offers (table, token ids, gradient rows, the table's gradient, padding index)
# after its own scatter-add; the data-parallel layer answers True when it adds the other ranks' rows itself
# (dp.GradBuckets.exchange_rows) instead of all-reducing the dense table.
row_sparse_hook = None


def _sink(t):
    if not grad_sink_enabled or t is None:
        return None
    g = t.grad
    if g is None or g.dtype != torch.float32 or not g.is_contiguous() or g.shape != t.shape or not g.is_cuda:
        return None
    return g


def _f32(shape, dev):
    return torch.empty(shape, dtype=torch.float32, device=dev)


def _bf16(shape, dev):
    return torch.empty(shape, dtype=torch.bfloat16, device=dev)


class _Saved:
    __slots__ = ("x", "x_bf", "h1", "mean1", "rstd1", "qkv", "o", "o_used", "lse", "c", "x1", "h2", "mean2", "rstd2",
                 "pre", "act", "f", "drop")

    def __init__(self):
        for s in self.__slots__:
            setattr(self, s, None)


def _attention_fwd(spec, qkv, B, S, key_bias):
    H, nh = spec.hidden, spec.heads
    dev = qkv.device
    o = _bf16((B * S, H), dev)
    lse = _f32((B, nh, S), dev)
    L.attn_fwd(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], o, lse, B=B, S=S, nh=nh, ld_qkv=3 * H, ld_o=H,
               key_bias=key_bias if spec.mask_mode == "key_bias" else None, scale=(H // nh) ** -0.5)
    return o, lse


# The rank-1 path of the fusion attention — weighted column sums and three tiny fp32 linears over [B, H] — is
# latency, not work: ~6 short kernels per layer that depend only on V (forward) or dx1 (backward) and are needed again at
# the out-projection / the dK/dV kernel.  They run on a side stream next to the attention forward (forward) and next to
# the out-projection wgrad + dgrad GEMMs (backward) instead of between them; a CUDA-graph capture records the fork and
# join as dependencies.  Buffers are allocated on the calling stream before the fork and every use of them on the side
# stream is joined back before the calling stream touches or frees them, so the caching allocator needs no record_stream.
rank1_side_stream = os.environ.get("TAVK_RANK1_SIDE", "1") != "0"
_RANK1_STREAMS = {}
# (Running the weight-gradient GEMMs on a second stream beside the dgrad chain was tried the same way and lost: two
# persistent one-CTA-per-SM GEMMs cannot share an SM, so they only interleave at CTA granularity — fusion block 5.67 ms
# against 5.64 ms, the full step 396 against 407 samples/s.)


class _SideWork:
    """with _SideWork(dev): kernels enqueued inside run on the device's rank-1 side stream, after everything enqueued so
    far on the calling stream; join() makes the calling stream wait for them."""

    def __init__(self, dev):
        self.main = torch.cuda.current_stream(dev)
        self.side = None
        if rank1_side_stream:
            self.side = _RANK1_STREAMS.get(dev)
            if self.side is None:
                self.side = _RANK1_STREAMS[dev] = torch.cuda.Stream(device=dev, priority=-1)
        self._ctx = None
        self.used = False

    def __enter__(self):
        self.used = True
        if self.side is not None:
            self.side.wait_stream(self.main)
            self._ctx = torch.cuda.stream(self.side)
            self._ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self._ctx is not None:
            self._ctx.__exit__(*exc)
            self._ctx = None
        return False

    def join(self):
        if self.side is not None and self.used:
            self.main.wait_stream(self.side)
            self.used = False


debug_dropout_masks = None   # tests set this to a list: every keep-mask drawn by the layer engine is appended (uint8)


def _drop_fwd(x, p, li, site, spec):
    """nn.Dropout(p) in training mode on an fp32 [M,H] buffer: (y, keep_mask uint8).  Counter-based generator keyed on
    (torch.initial_seed(), encoder salt, layer, site) plus the device-side step counter (fresh masks on graph replays)."""
    c = _dropout_counter.get(x.device)
    if c is None:
        c = _dropout_counter[x.device] = torch.zeros(1, dtype=torch.int64, device=x.device)
    y = torch.empty_like(x)
    keep = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    stream_id = ((spec.dropout_salt & 0xFFFF) << 16) | ((li * 4 + site) & 0xFFFF)
    seed = (torch.initial_seed() + 0x9E3779B97F4A7C15 * (stream_id + 1)) & 0x7FFFFFFFFFFFFFFF
    L.call("tavk_dropout", x.data_ptr(), y.data_ptr(), keep.data_ptr(), x.numel(), float(p), seed, 0, c.data_ptr())
    if debug_dropout_masks is not None:
        debug_dropout_masks.append(keep)
    return y, keep


def _drop_bwd(dy, keep, p, resid=None):
    dx = torch.empty_like(dy)
    if resid is None:
        L.call("tavk_dropout_bwd", dy.data_ptr(), keep.data_ptr(), dx.data_ptr(), dy.numel(), float(p))
    else:
        L.call("tavk_dropout_bwd_add", dy.data_ptr(), keep.data_ptr(), resid.data_ptr(), dx.data_ptr(), dy.numel(), float(p))
    return dx


def _layer_fwd(spec, p, sh, x, x_bf, B, S, mask2d, keep, li=0, same_as_training=False):
    """x: f32 [M,H].  Returns (y_f32, y_bf16 or None, saved).  ``same_as_training``: use the training forward's epilogues
    even though nothing is kept (selective recompute: the first pass and the pass re-run in backward must agree bit for bit)."""
    train_epi = keep or same_as_training
    H, I = spec.hidden, spec.inter
    M = B * S
    dev = x.device
    d = dict(zip(PARAM_SLOTS, p))
    sv = _Saved() if keep else None
    if spec.pre_ln:
        h1, _, mean1, rstd1 = L.layernorm_fwd(x, d["ln1_w"], d["ln1_b"], spec.eps)
        qkv = _bf16((M, 3 * H), dev)
        L.gemm(h1, sh.wqkv, qkv, M=M, N=3 * H, K=H, bias=sh.bqkv)
        rb = c = None
        if spec.mask_mode == "rank1":
            # P + m  =>  ctx += sum_k m[b,k] V[b,k,:]  (same vector for every query); keep it fp32 end to end
            c, rb = _f32((B, H), dev), _f32((B, H), dev)
            with _SideWork(dev) as sw:      # next to the attention forward
                L.masked_colsum(qkv[:, 2 * H:], mask2d, c, B=B, S=S, N=H, ld=3 * H)
                L.call("tavk_small_linear_fwd", c.data_ptr(), d["wo"].data_ptr(), None, rb.data_ptr(), B, H, H)
            o, lse = _attention_fwd(spec, qkv, B, S, mask2d)
            sw.join()
        else:
            o, lse = _attention_fwd(spec, qkv, B, S, mask2d)
        x1 = _f32((M, H), dev)
        L.gemm(o, sh.wo, x1, M=M, N=H, K=H, bias=d["bo"], resid=x, rowbias=rb, rows_per_group=S)
        h2, _, mean2, rstd2 = L.layernorm_fwd(x1, d["ln2_w"], d["ln2_b"], spec.eps)
        # `pre` receives gelu'(pre-activation) when a backward will follow (EPI_GELU_GRAD): the backward's dgrad GEMM then
        # only multiplies by it (EPI_MUL) instead of evaluating the GELU derivative in its (issue-bound) epilogue
        pre, act = _bf16((M, I), dev), _bf16((M, I), dev)
        L.gemm(h2, sh.w1, pre, M=M, N=I, K=H, bias=d["b1"], out2=act, epilogue=L.EPI_GELU_GRAD if train_epi else L.EPI_GELU)
        y = _f32((M, H), dev)
        L.gemm(act, sh.w2, y, M=M, N=H, K=I, bias=d["b2"], resid=x1)
        if keep:
            sv.x, sv.h1, sv.mean1, sv.rstd1, sv.qkv, sv.o, sv.lse, sv.c = x, h1, mean1, rstd1, qkv, o, lse, c
            sv.x1, sv.h2, sv.mean2, sv.rstd2, sv.pre, sv.act = x1, h2, mean2, rstd2, pre, act
        return y, None, sv
    # post-LN
    if x_bf is None:
        x_bf = L.cast_bf16(x)
    qkv = _bf16((M, 3 * H), dev)
    L.gemm(x_bf, sh.wqkv, qkv, M=M, N=3 * H, K=H, bias=sh.bqkv)
    o, lse = _attention_fwd(spec, qkv, B, S, mask2d)
    o_used = o
    if spec.scrambled_concat:
        o_used = _bf16((M, H), dev)
        L.call("tavk_permute_bshd_bhds", o.data_ptr(), o_used.data_ptr(), B, S, spec.heads, H // spec.heads, 0)
    a = _f32((M, H), dev)
    L.gemm(o_used, sh.wo, a, M=M, N=H, K=H, bias=d["bo"], resid=x)
    pd = spec.dropout
    drop = None
    if pd > 0.0:
        # reference TransformerBlock in training mode (utils/TAVFormer.py:130-141): dropout1 on the residual sum before
        # norm1, the Dropout that opens feed_forward (the FFN residual keeps the un-dropped norm1 output), dropout2
        # before norm2
        a, m1 = _drop_fwd(a, pd, li, 0, spec)
    y_bf, y_f32, mean1, rstd1 = L.layernorm_fwd(a, d["ln1_w"], d["ln1_b"], spec.eps, want_bf16=True, want_f32=True)
    ff_in = y_bf
    if pd > 0.0:
        yd, m2 = _drop_fwd(y_f32, pd, li, 1, spec)
        ff_in = L.cast_bf16(yd)
    pre, act = _bf16((M, I), dev), _bf16((M, I), dev)
    L.gemm(ff_in, sh.w1, pre, M=M, N=I, K=H, bias=d["b1"], out2=act, epilogue=L.EPI_GELU_GRAD if train_epi else L.EPI_GELU)
    f = _f32((M, H), dev)
    L.gemm(act, sh.w2, f, M=M, N=H, K=I, bias=d["b2"], resid=y_f32)
    if pd > 0.0:
        f, m3 = _drop_fwd(f, pd, li, 2, spec)
        drop = (m1, m2, m3)
    z_bf, z_f32, mean2, rstd2 = L.layernorm_fwd(f, d["ln2_w"], d["ln2_b"], spec.eps, want_bf16=True, want_f32=True)
    if keep:
        sv.x_bf, sv.qkv, sv.o, sv.o_used, sv.lse, sv.x1, sv.mean1, sv.rstd1 = x_bf, qkv, o, o_used, lse, a, mean1, rstd1
        sv.h2, sv.pre, sv.act, sv.f, sv.mean2, sv.rstd2, sv.drop = ff_in, pre, act, f, mean2, rstd2, drop
    return z_f32, z_bf, sv


def _attention_bwd(spec, sv, do, B, S, mask2d, dc, go=None):
    """dqkv [M, 3H] bf16; when ``go`` is given the q/k/v projection bias gradients (column sums of dq/dk/dv) are
    accumulated into its targets by the backward kernels themselves."""
    H, nh = spec.hidden, spec.heads
    dev = do.device
    dqkv = _bf16((B * S, 3 * H), dev)
    delta = _f32((B, nh, S), dev)
    qkv = sv.qkv
    db = {n: (go.target(n) if go is not None else None) for n in ("bq", "bk", "bv")}
    L.attn_bwd(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], sv.o, do, sv.lse, delta, dqkv[:, :H], dqkv[:, H:2 * H],
               dqkv[:, 2 * H:], B=B, S=S, nh=nh, ld_qkv=3 * H, ld_o=H, ld_dqkv=3 * H,
               key_bias=mask2d if spec.mask_mode == "key_bias" else None,
               dv_rowscale=mask2d if dc is not None else None, dv_rank1=dc, scale=(H // nh) ** -0.5,
               dbq=db["bq"], dbk=db["bk"], dbv=db["bv"])
    return dqkv


class _GradOut:
    """Per-layer gradient targets: the parameter's own .grad (sink) or a fresh zero buffer returned to autograd."""

    def __init__(self, p, dev):
        self.d = dict(zip(PARAM_SLOTS, p))
        self.sink = {n: _sink(t) for n, t in self.d.items()}
        self.tmp = {}
        self.dev = dev

    def target(self, name):
        """fp32 buffer to ACCUMULATE the gradient of slot `name` into (None when the layer has no such parameter)."""
        t = self.d[name]
        if t is None:
            return None
        if self.sink[name] is not None:
            return self.sink[name]
        if name not in self.tmp:
            self.tmp[name] = torch.zeros(t.shape, dtype=torch.float32, device=self.dev)
        return self.tmp[name]

    def results(self):
        return [self.tmp.get(n) if (self.d[n] is not None and self.sink[n] is None) else None for n in PARAM_SLOTS]

    def written(self):
        return [self.d[n] for n in PARAM_SLOTS if self.d[n] is not None and self.sink[n] is not None]


def _qkv_grads(go, dqkv, x_bf, H, M):
    """dW_q/k/v from the packed dqkv [M, 3H]: one [3H, H] wgrad GEMM when the three weight-gradient targets are
    adjacent in memory (flat-buffer order q, k, v), else one GEMM per matrix.  (The bias gradients come out of the
    attention backward, see _attention_bwd.)"""
    tq, tk, tv = go.target("wq"), go.target("wk"), go.target("wv")
    step = H * H * 4
    if tk.data_ptr() - tq.data_ptr() == step and tv.data_ptr() - tk.data_ptr() == step:
        _wgrad(dqkv, x_bf, 3 * H, H, M, out=tq)     # tq is the base of the adjacent [3H, H] block
    else:
        for i, t in enumerate((tq, tk, tv)):
            _wgrad(dqkv[:, i * H:(i + 1) * H], x_bf, H, H, M, out=t, lda=3 * H)


def _layer_bwd(spec, p, sh, sv, dy, dy_bf, B, S, mask2d, need_dx_bf, b2_done=False, dx_colsum=None):
    """Returns (dx_f32, dx_bf16|None, grads list aligned with PARAM_SLOTS — None where the gradient went to the sink —,
    parameters whose .grad was written directly).
    ``b2_done``: the column sums of dy (this layer's b2 gradient) were already accumulated by the layer above's
    LayerNorm backward; ``dx_colsum``: where to accumulate the column sums of dx (the b2 gradient of the layer below)."""
    H, I = spec.hidden, spec.inter
    M = B * S
    dev = dy.device
    d = dict(zip(PARAM_SLOTS, p))
    go = _GradOut(p, dev)
    if spec.pre_ln:
        if dy_bf is None:
            dy_bf = L.cast_bf16(dy)
        # FFN down
        if not b2_done and d["b2"] is not None:
            L.colsum(dy, go.target("b2"), M=M, N=H, accumulate=True)
        _wgrad(dy_bf, sv.act, H, I, M, out=go.target("w2"))
        dpre = _bf16((M, I), dev)
        # the GELU' dgrad GEMM also accumulates the column sums of dpre (= b1 gradient) in its epilogue
        L.gemm(dy_bf, sh.w2, dpre, M=M, N=I, K=H, b_mn=True, aux=sv.pre, epilogue=L.EPI_MUL, colsum=go.target("b1"))
        # FFN up
        _wgrad(dpre, sv.h2, I, H, M, out=go.target("w1"))
        dh2 = _f32((M, H), dev)
        L.gemm(dpre, sh.w1, dh2, M=M, N=H, K=I, b_mn=True)
        # LayerNorm backward also emits the column sums of dx1 (= out-projection bias gradient)
        dx1, dx1_bf = L.layernorm_bwd(dh2, sv.x1, sv.mean2, sv.rstd2, d["ln2_w"], go.target("ln2_w"), go.target("ln2_b"),
                                      resid=dy, want_f32=True, want_bf16=True, dx_colsum=go.target("bo"))
        dc = sw = None
        wo_t = go.target("wo")
        if spec.mask_mode == "rank1":
            drb, dc = _f32((B, H), dev), _f32((B, H), dev)
            with _SideWork(dev) as sw:      # next to the out-projection wgrad / dgrad GEMMs (dW_o: atomic adds on both sides)
                L.masked_colsum(dx1, None, drb, B=B, S=S, N=H, ld=H)
                L.call("tavk_small_linear_bwd_x", drb.data_ptr(), d["wo"].data_ptr(), dc.data_ptr(), B, H, H, 0)
                L.call("tavk_small_linear_bwd_w", drb.data_ptr(), sv.c.data_ptr(), wo_t.data_ptr(), None, B, H, H)
        _wgrad(dx1_bf, sv.o, H, H, M, out=wo_t)
        do = _bf16((M, H), dev)
        L.gemm(dx1_bf, sh.wo, do, M=M, N=H, K=H, b_mn=True)
        if sw is not None:
            sw.join()
        dqkv = _attention_bwd(spec, sv, do, B, S, mask2d, dc, go)
        _qkv_grads(go, dqkv, sv.h1, H, M)
        dh1 = _f32((M, H), dev)
        L.gemm(dqkv, sh.wqkv, dh1, M=M, N=H, K=3 * H, b_mn=True)
        dx, dx_bf = L.layernorm_bwd(dh1, sv.x, sv.mean1, sv.rstd1, d["ln1_w"], go.target("ln1_w"), go.target("ln1_b"),
                                    resid=dx1, want_f32=True, want_bf16=need_dx_bf, dx_colsum=dx_colsum)
    else:
        pd = spec.dropout if sv.drop is not None else 0.0
        if pd > 0.0:
            # gradients of the three dropouts (see _layer_fwd): the LayerNorm backward can no longer emit the bias column
            # sums or the bf16 copies itself, because both are taken AFTER the dropout mask is applied
            df, _ = L.layernorm_bwd(dy, sv.f, sv.mean2, sv.rstd2, d["ln2_w"], go.target("ln2_w"), go.target("ln2_b"),
                                    want_f32=True, want_bf16=False)
            df = _drop_bwd(df, sv.drop[2], pd)
            df_bf = L.cast_bf16(df)
            if d["b2"] is not None:
                L.colsum(df, go.target("b2"), M=M, N=H, accumulate=True)
        else:
            df, df_bf = L.layernorm_bwd(dy, sv.f, sv.mean2, sv.rstd2, d["ln2_w"], go.target("ln2_w"), go.target("ln2_b"),
                                        want_f32=True, want_bf16=True, dx_colsum=go.target("b2"))
        _wgrad(df_bf, sv.act, H, I, M, out=go.target("w2"))
        dpre = _bf16((M, I), dev)
        L.gemm(df_bf, sh.w2, dpre, M=M, N=I, K=H, b_mn=True, aux=sv.pre, epilogue=L.EPI_MUL, colsum=go.target("b1"))
        _wgrad(dpre, sv.h2, I, H, M, out=go.target("w1"))
        dyl = _f32((M, H), dev)
        if pd > 0.0:
            L.gemm(dpre, sh.w1, dyl, M=M, N=H, K=I, b_mn=True)
            dyl = _drop_bwd(dyl, sv.drop[1], pd, resid=df)
            da, _ = L.layernorm_bwd(dyl, sv.x1, sv.mean1, sv.rstd1, d["ln1_w"], go.target("ln1_w"), go.target("ln1_b"),
                                    want_f32=True, want_bf16=False)
            da = _drop_bwd(da, sv.drop[0], pd)
            da_bf = L.cast_bf16(da)
            if d["bo"] is not None:
                L.colsum(da, go.target("bo"), M=M, N=H, accumulate=True)
        else:
            L.gemm(dpre, sh.w1, dyl, M=M, N=H, K=I, b_mn=True, resid=df)
            da, da_bf = L.layernorm_bwd(dyl, sv.x1, sv.mean1, sv.rstd1, d["ln1_w"], go.target("ln1_w"), go.target("ln1_b"),
                                        want_f32=True, want_bf16=True, dx_colsum=go.target("bo"))
        _wgrad(da_bf, sv.o_used, H, H, M, out=go.target("wo"))
        do = _bf16((M, H), dev)
        L.gemm(da_bf, sh.wo, do, M=M, N=H, K=H, b_mn=True)
        if spec.scrambled_concat:
            do2 = _bf16((M, H), dev)
            L.call("tavk_permute_bshd_bhds", do.data_ptr(), do2.data_ptr(), B, S, spec.heads, H // spec.heads, 1)
            do = do2
        dqkv = _attention_bwd(spec, sv, do, B, S, mask2d, None, go)
        _qkv_grads(go, dqkv, sv.x_bf, H, M)
        dx = _f32((M, H), dev)
        L.gemm(dqkv, sh.wqkv, dx, M=M, N=H, K=3 * H, b_mn=True, resid=da)
        dx_bf = None
    return dx, dx_bf, go.results(), go.written()


# Selective recompute (BASELINE configs[4] sweeps the per-GPU batch to 256): when set, a stack keeps only each
# layer's fp32 INPUT (4·H bytes per token instead of ~58·H: LN outputs, packed QKV, attention output, FFN activations ...)
# and re-runs the layer's forward inside the backward.  One extra forward per layer (+1/3 of the stack's FLOPs) buys a
# ~12x smaller activation footprint: at 256 samples per GPU the VideoMAE stack alone would save 166 GB, against 14 GB.
recompute_layers = False


class EncoderStackFn(torch.autograd.Function):
    """y = layers_n(...layers_1(x)); x f32 [B,S,H]; mask2d f32 [B,S] or None; params = N_SLOTS entries per layer."""

    @staticmethod
    def forward(ctx, spec, shadows, x, mask2d, *params):
        L.require_device()
        B, S, H = x.shape
        assert H == spec.hidden and x.dtype == torch.float32
        n_layers = len(params) // N_SLOTS
        keep = any(ctx.needs_input_grad)  # inference: nothing is saved
        recompute = keep and recompute_layers and spec.dropout == 0.0   # (dropout masks are not regenerated)
        cur = x.contiguous().view(B * S, H)
        if mask2d is not None:
            mask2d = mask2d.contiguous().float()
        cur_bf = None
        saved = []
        for li in range(n_layers):
            p = params[li * N_SLOTS:(li + 1) * N_SLOTS]
            sh = shadows[li].refresh(p, spec)
            if recompute:
                x_in = cur
                cur, cur_bf, _ = _layer_fwd(spec, p, sh, cur, cur_bf, B, S, mask2d, False, li, same_as_training=True)
                saved.append(x_in)
            else:
                cur, cur_bf, sv = _layer_fwd(spec, p, sh, cur, cur_bf, B, S, mask2d, keep, li)
                saved.append(sv)
        if spec.dropout > 0.0 and not spec.pre_ln:
            _dropout_counter[x.device].add_(1)   # device-side: the next call (or graph replay) draws fresh masks
        ctx.spec, ctx.shadows, ctx.saved, ctx.mask2d, ctx.params, ctx.dims = spec, shadows, saved, mask2d, params, (B, S, H)
        ctx.recompute = recompute
        return cur.view(B, S, H)

    @staticmethod
    def backward(ctx, dy):
        spec, B, S, H = ctx.spec, *ctx.dims
        n_layers = len(ctx.params) // N_SLOTS
        dy = dy.contiguous().view(B * S, H)
        if dy.dtype != torch.float32:
            dy = dy.float()
        dy_bf = None
        grads = [None] * len(ctx.params)
        b2_done = False
        i_b2 = PARAM_SLOTS.index("b2")
        for li in range(n_layers - 1, -1, -1):
            p = ctx.params[li * N_SLOTS:(li + 1) * N_SLOTS]
            # pre-LN stacks: this layer's last LayerNorm backward also accumulates the column sums of its dx, which
            # are the b2 gradient of the layer below (into that parameter's .grad, or a buffer handed back to autograd)
            below_b2 = ctx.params[(li - 1) * N_SLOTS + i_b2] if (li > 0 and spec.pre_ln) else None
            cs_t, cs_tmp = None, None
            if below_b2 is not None:
                cs_t = _sink(below_b2)
                if cs_t is None:
                    cs_t = cs_tmp = torch.zeros(below_b2.shape, dtype=torch.float32, device=dy.device)
            sv = ctx.saved[li]
            if ctx.recompute:      # sv is the layer's input: rebuild what the backward needs
                _, _, sv = _layer_fwd(spec, p, ctx.shadows[li], sv, None, B, S, ctx.mask2d, True, li)
            dy, dy_bf, g, written = _layer_bwd(spec, p, ctx.shadows[li], sv, dy, dy_bf, B, S, ctx.mask2d,
                                               need_dx_bf=(li > 0 and spec.pre_ln), b2_done=b2_done, dx_colsum=cs_t)
            ctx.saved[li] = None  # free activations as we go
            for k, gk in enumerate(g):
                if gk is not None:
                    grads[li * N_SLOTS + k] = gk if grads[li * N_SLOTS + k] is None else grads[li * N_SLOTS + k] + gk
            if cs_tmp is not None:
                grads[(li - 1) * N_SLOTS + i_b2] = cs_tmp
            b2_done = below_b2 is not None
            if grad_written_hook is not None and written:
                grad_written_hook(written)
        return (None, None, dy.view(B, S, H), None, *grads)


def run_stack(spec, shadows, x, mask2d, layer_params):
    """layer_params: list (per layer) of N_SLOTS-long lists of tensors/None."""
    flat = [t for p in layer_params for t in p]
    return EncoderStackFn.apply(spec, shadows, x, mask2d, *flat)


# ------------------------------------------------------------------------------------------------ small fused ops
class _LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, eps):
        shp = x.shape
        x2 = x.contiguous().view(-1, shp[-1]).float()
        _, y, mean, rstd = L.layernorm_fwd(x2, w, b, eps, want_bf16=False, want_f32=True)
        ctx.save_for_backward(x2, w, mean, rstd)
        ctx.shp = shp
        return y.view(shp)

    @staticmethod
    def backward(ctx, dy):
        x2, w, mean, rstd = ctx.saved_tensors
        H = x2.shape[1]
        dw = torch.zeros((H,), dtype=torch.float32, device=x2.device)
        db = torch.zeros((H,), dtype=torch.float32, device=x2.device)
        dx, _ = L.layernorm_bwd(dy.contiguous().view(-1, H).float(), x2, mean, rstd, w, dw, db)
        return dx.view(ctx.shp), dw, db, None


# The small operators below are the torch custom ops of ops.py (torch.ops.tavk.*: schema + fake implementation +
# register_autograd over the same C-ABI entry points) — the modules reach the kernels through that layer, as the boundary
# asks.  TAVK_ENGINE_DIRECT=1 keeps the in-file autograd Functions instead (identical kernels).
_VIA_OPS = os.environ.get("TAVK_ENGINE_DIRECT", "0") != "1"


def layer_norm(x, w, b, eps=1e-5):
    """nn.LayerNorm over the last dim through tavk_layernorm_{fwd,bwd} (reference models/tav.py:486,488-490)."""
    if _VIA_OPS:
        from . import ops

        return ops.layer_norm(x, w, b, eps)
    return _LayerNormFn.apply(x, w, b, eps)


class _MeanPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, lengths):
        B, S, H = x.shape
        x = x.contiguous().float()
        y = torch.empty((B, H), dtype=torch.float32, device=x.device)
        if lengths is not None:
            lengths = lengths.to(device=x.device, dtype=torch.int32).contiguous()
        L.call("tavk_masked_mean_pool_fwd", x.data_ptr(), L._ptr(lengths), y.data_ptr(), B, S, H)
        ctx.dims = (B, S, H)
        ctx.lengths = lengths
        return y

    @staticmethod
    def backward(ctx, dy):
        B, S, H = ctx.dims
        dy = dy.contiguous().float()
        dx = torch.empty((B, S, H), dtype=torch.float32, device=dy.device)
        L.call("tavk_masked_mean_pool_bwd", dy.data_ptr(), L._ptr(ctx.lengths), dx.data_ptr(), None, B, S, H)
        return dx, None


def mean_pool(x, lengths=None):
    """torch.mean(x, dim=1) (reference models/tav.py:478,481,488; unmasked — SURVEY Q3).  ``lengths`` (int [B]) selects the
    masked mean over the first lengths[b] tokens of each sample instead (the boundary's optional argument; the
    reference itself never passes one)."""
    if _VIA_OPS and lengths is None:
        from . import ops

        return ops.mean_pool(x)
    return _MeanPoolFn.apply(x, lengths)


class _EmbedAddFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, idx, table):
        B, S, H = x.shape
        x = x.contiguous().float()
        idx = idx.contiguous().long()
        y = torch.empty_like(x)
        L.call("tavk_embed_add_fwd", x.data_ptr(), idx.data_ptr(), table.data_ptr(), y.data_ptr(), B * S, H, table.shape[0])
        ctx.save_for_backward(idx)
        ctx.dims = (B, S, H, table.shape[0])
        return y

    @staticmethod
    def backward(ctx, dy):
        (idx,) = ctx.saved_tensors
        B, S, H, n = ctx.dims
        dt = None
        if ctx.needs_input_grad[2]:
            dy = dy.contiguous()
            dt = torch.zeros((n, H), dtype=torch.float32, device=dy.device)
            L.call("tavk_embed_add_bwd", dy.data_ptr(), idx.data_ptr(), dt.data_ptr(), B * S, H, n)
        return dy, None, dt


def embed_add(x, idx, table):
    """x + table[idx] (reference models/tav.py:474)."""
    if _VIA_OPS:
        from . import ops  # noqa: F401  (registers torch.ops.tavk)

        return torch.ops.tavk.embed_add(x, idx, table)
    return _EmbedAddFn.apply(x, idx, table)


class _RobertaEmbedFn(torch.autograd.Function):
    """word[ids] + type[0] + pos[position ids] (HF RobertaEmbeddings before its LayerNorm) as one gather kernel; the
    backward scatters rows straight into the tables' gradients — into the parameter's own ``.grad`` slice of the flat
    buffer when the fused optimiser owns it (gradient sink), so the 154 MB word-embedding gradient is never zero-filled,
    written and added as three dense passes."""

    @staticmethod
    def forward(ctx, ids, word, pos, typ, pad_id, pads):
        B, T = ids.shape
        V, H = word.shape
        ids = ids.contiguous().long()
        y = torch.empty((B, T, H), dtype=torch.float32, device=word.device)
        pos_ids = torch.empty((B, T), dtype=torch.int64, device=word.device)
        L.call("tavk_roberta_embed_fwd", ids.data_ptr(), word.data_ptr(), pos.data_ptr(), typ.data_ptr(), y.data_ptr(),
               pos_ids.data_ptr(), B, T, H, V, pos.shape[0], int(pad_id))
        word._tavk_row_sparse = True      # its gradient has at most B*T non-zero rows (see row_sparse_hook)
        ctx.save_for_backward(ids, pos_ids)
        ctx.tables = (word, pos, typ)
        ctx.pads = pads
        return y

    @staticmethod
    def backward(ctx, dy):
        ids, pos_ids = ctx.saved_tensors
        word, pos, typ = ctx.tables
        B, T = ids.shape
        H = word.shape[1]
        dy = dy.contiguous().float()
        outs, written = [], []
        for table, idx, need, skip in ((word, ids, ctx.needs_input_grad[1], ctx.pads[0]),
                                       (pos, pos_ids, ctx.needs_input_grad[2], ctx.pads[1])):
            if not need:
                outs.append(None)
                continue
            g = _sink(table)
            if g is None:
                g = torch.zeros_like(table)
                outs.append(g)
            else:
                outs.append(None)
                written.append(table)
            L.call("tavk_embedding_scatter_add", dy.data_ptr(), idx.data_ptr(), g.data_ptr(), B * T, H, table.shape[0],
                   -1 if skip is None else int(skip))
            if table is word and row_sparse_hook is not None:
                row_sparse_hook(table, idx, dy, g, skip)
        dtyp = None
        if ctx.needs_input_grad[3]:
            g = _sink(typ)
            if g is None:
                dtyp = torch.zeros_like(typ)
                g = dtyp
            else:
                written.append(typ)
            colsum_target = g[0]          # every token uses type row 0
            L.colsum(dy.view(B * T, H), colsum_target, M=B * T, N=H, accumulate=True)
        if grad_written_hook is not None and written:
            grad_written_hook(written)
        return None, outs[0], outs[1], dtyp, None, None


def roberta_embeddings(emb, input_ids):
    """HF RobertaEmbeddings.forward(input_ids=...) (reference models/tav.py:349 and inside bert(...) at :485): gather-sum
    kernel, then the LayerNorm kernel (dropout: the HF sub-models stay in eval mode)."""
    if emb.training and emb.dropout.p > 0:
        raise NotImplementedError("RobertaEmbeddings dropout in training mode (the reference keeps the HF models in eval)")
    # nn.Embedding(padding_idx=...) rows receive no gradient: word and position tables both carry one in HF RoBERTa
    x = _RobertaEmbedFn.apply(input_ids, emb.word_embeddings.weight, emb.position_embeddings.weight,
                              emb.token_type_embeddings.weight, emb.padding_idx,
                              (emb.word_embeddings.padding_idx, emb.position_embeddings.padding_idx))
    return layer_norm(x, emb.LayerNorm.weight, emb.LayerNorm.bias, emb.LayerNorm.eps)


class _SmallLinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        x = x.contiguous().float()
        M, K = x.shape
        N = w.shape[0]
        y = torch.empty((M, N), dtype=torch.float32, device=x.device)
        L.call("tavk_small_linear_fwd", x.data_ptr(), w.data_ptr(), L._ptr(b), y.data_ptr(), M, N, K)
        ctx.save_for_backward(x, w)
        ctx.has_b = b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        M, K = x.shape
        N = w.shape[0]
        dy = dy.contiguous().float()
        dx = torch.empty_like(x)
        L.call("tavk_small_linear_bwd_x", dy.data_ptr(), w.data_ptr(), dx.data_ptr(), M, N, K, 0)
        dw = torch.zeros_like(w)
        db = torch.zeros((N,), dtype=torch.float32, device=x.device) if ctx.has_b else None
        L.call("tavk_small_linear_bwd_w", dy.data_ptr(), x.data_ptr(), dw.data_ptr(), L._ptr(db), M, N, K)
        return dx, dw, db


def small_linear(x, w, b=None):
    """fp32 linear for launch-bound shapes: the classifier head Linear(3072, C) (reference models/tav.py:499)."""
    if _VIA_OPS:
        from . import ops

        return ops.small_linear(x, w, b)
    return _SmallLinearFn.apply(x, w, b)


class _LinearBf16Fn(torch.autograd.Function):
    """y = x W^T + b with the tcgen05 GEMM (bf16 operands, fp32 accumulate/output): the 1024->768 audio projections
    (reference models/tav.py:363,478)."""

    @staticmethod
    def forward(ctx, x, w, b):
        shp = x.shape
        K, N = shp[-1], w.shape[0]
        x2 = x.contiguous().view(-1, K)
        M = x2.shape[0]
        x_bf = L.cast_bf16(x2.float()) if x2.dtype != torch.bfloat16 else x2
        w_bf = L.cast_bf16(w.detach())
        y = torch.empty((M, N), dtype=torch.float32, device=x.device)
        L.gemm(x_bf, w_bf, y, M=M, N=N, K=K, bias=b)
        ctx.save_for_backward(x_bf, w_bf)
        ctx.meta = (shp, M, N, K, b is not None)
        return y.view(*shp[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        x_bf, w_bf = ctx.saved_tensors
        shp, M, N, K, has_b = ctx.meta
        dy2 = dy.contiguous().view(M, N).float()
        dy_bf = L.cast_bf16(dy2)
        dx = torch.empty((M, K), dtype=torch.float32, device=dy.device)
        L.gemm(dy_bf, w_bf, dx, M=M, N=K, K=N, b_mn=True)
        dw = _wgrad(dy_bf, x_bf, N, K, M)
        db = None
        if has_b:
            db = torch.empty((N,), dtype=torch.float32, device=dy.device)
            L.colsum(dy2, db, M=M, N=N)
        return dx.view(shp), dw, db


def linear_bf16(x, w, b=None):
    if _VIA_OPS:
        from . import ops  # noqa: F401

        return torch.ops.tavk.linear(x, w, b)
    return _LinearBf16Fn.apply(x, w, b)


_dropout_counter = {}


class _DropoutFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p, seed, counter):
        x = x.contiguous().float()
        y = torch.empty_like(x)
        keep = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
        L.call("tavk_dropout", x.data_ptr(), y.data_ptr(), keep.data_ptr(), x.numel(), float(p), int(seed), 0,
               counter.data_ptr())
        counter.add_(1)  # device-side: also advances on every CUDA-graph replay
        ctx.save_for_backward(keep)
        ctx.p = float(p)
        return y

    @staticmethod
    def backward(ctx, dy):
        (keep,) = ctx.saved_tensors
        dy = dy.contiguous().float()
        dx = torch.empty_like(dy)
        L.call("tavk_dropout_bwd", dy.data_ptr(), keep.data_ptr(), dx.data_ptr(), dy.numel(), ctx.p)
        return dx, None, None, None


def dropout(x, p, seed=None):
    """nn.Dropout(p) in training mode (the classifier-head dropout, reference models/tav.py:497-498) with a
    counter-based generator keyed on (torch.initial_seed(), device-side call counter): reproducible after
    torch.manual_seed + reset_dropout_counter(), but not torch's own random stream."""
    if p <= 0.0:
        return x
    c = _dropout_counter.get(x.device)
    if c is None:
        c = _dropout_counter[x.device] = torch.zeros(1, dtype=torch.int64, device=x.device)
    seed = (torch.initial_seed() if seed is None else seed) & 0x7FFFFFFFFFFFFFFF
    if _VIA_OPS:
        from . import ops  # noqa: F401

        y, _ = torch.ops.tavk.dropout(x, float(p), seed, c)
        c.add_(1)  # device-side: also advances on every CUDA-graph replay
        return y
    return _DropoutFn.apply(x, p, seed, c)


def reset_dropout_counter():
    for c in _dropout_counter.values():
        c.zero_()


class _SoftmaxCEFn(torch.autograd.Function):
    """Returns (numerator, denominator) of the weighted mean CE so data-parallel ranks can all-reduce the denominator."""

    @staticmethod
    def forward(ctx, logits, target, weight):
        B, C = logits.shape
        logits = logits.contiguous().float()
        target = target.contiguous().long()
        probs = torch.empty((B, C), dtype=torch.float32, device=logits.device)
        nd = torch.empty((2,), dtype=torch.float32, device=logits.device)
        L.call("tavk_softmax_ce_fwd", logits.data_ptr(), target.data_ptr(), L._ptr(weight), probs.data_ptr(),
               nd.data_ptr(), nd.data_ptr() + 4, B, C)
        ctx.save_for_backward(probs, target)
        ctx.weight = weight
        ctx.mark_non_differentiable(nd[1:])
        return nd[0], nd[1]

    @staticmethod
    def backward(ctx, dnum, dden):
        probs, target = ctx.saved_tensors
        B, C = probs.shape
        dl = torch.empty_like(probs)
        gscale = dnum.contiguous().float().view(1)
        L.call("tavk_softmax_ce_bwd", probs.data_ptr(), target.data_ptr(), L._ptr(ctx.weight), gscale.data_ptr(),
               dl.data_ptr(), B, C)
        return dl, None, None


def softmax_ce_parts(logits, target, weight=None):
    return _SoftmaxCEFn.apply(logits, target, weight)


def cross_entropy(logits, target, weight=None):
    """nn.CrossEntropyLoss(weight=w)(logits, target) with mean reduction (reference utils/global_functions.py:63-64)."""
    num, den = softmax_ce_parts(logits, target, weight)
    return num / den
