"""Front-end contractions of the TAV path moved from cuDNN onto the tcgen05 GEMM:

* VideoMAE patch embedding — ``Conv3d(3, 768, kernel=stride=(2,16,16))`` (HF ``VideoMAEPatchEmbeddings``, used at
  reference models/tav.py:368 and :480) is exactly ``[B·1568, 1536] x [1536, 768]``.  The im2col is a single
  cast+permute of the video (shared between PreFormer and TAVForMAE, which read the same clip), the kept-token gather
   happens BEFORE the GEMM (PreFormer keeps 104 of 1568 tokens), and the sinusoid position rows are added
  in the GEMM epilogue.
* Wav2Vec2 positional convolution — ``Conv1d(H, H, k=128, pad=64, groups=16)`` with weight-norm (HF
  ``Wav2Vec2PositionalConvEmbedding``, used at reference models/tav.py:360 and inside ``wav2vec2(...)`` :476):
  16 grouped GEMMs ``[B·T, 48·128] x [48·128, 48]`` over a bf16 sliding-window matrix; backward is the same routine
  with flipped/transposed weights (dgrad) and a split-K wgrad.  cuDNN ran this as 32 tiny TF32 kernels per call
  (≈20 ms of a 90 ms step at B=16).
"""
import weakref

import torch
import torch.nn.functional as F

from . import _lib as L
from .engine import _wgrad


def _fe_rows(L, kernels, strides):
    """Valid frames T_l per layer and padded rows-per-sample R_l with R_{l-1} = stride_l * R_l (and s_0 * R_0 >= L):
    with that padding the (b, t) rows of every layer form ONE matrix whose row pitch is stride*C."""
    T, t = [], L
    for k, st in zip(kernels, strides):
        t = (t - k) // st + 1
        T.append(t)
    n = len(T)
    suffix = [1] * n                      # product of strides of the layers above l
    for l in range(n - 2, -1, -1):
        suffix[l] = suffix[l + 1] * strides[l + 1]
    r_last = max(-(-T[l] // suffix[l]) for l in range(n))
    r_last = max(r_last, -(-L // (suffix[0] * strides[0])))
    return T, [r_last * suffix[l] for l in range(n)]


class _ConvFeatureEncoderFn(torch.autograd.Function):
    """HF Wav2Vec2FeatureEncoder (feat_extract_norm="group", no conv bias) in channels-last layout: [B, L] f32 ->
    [B, T, C] f32.  Layer 0 (C_in = 1) + GroupNorm + GELU are streaming kernels (csrc/conv_frontend.cu); every later
    Conv1d(C, C, k, stride s) + GELU is one tcgen05 GEMM over the whole batch: A = previous activations with row pitch
    s*C and row length k*C (overlapping TMA rows), B = the weight reordered to [C_out, (tap, c_in)], GELU epilogue.
    Backward per layer: wgrad GEMM (both operands MN-major, split-K) and a dgrad GEMM that writes s input rows per
    output row at once — A = [dpre(t-D+1) .. dpre(t)] (row pitch C, row length D*C, D = ceil(k/s)), B = the taps
    regrouped by (input phase r, delay d) — with the previous layer's GELU' fused in its epilogue."""

    @staticmethod
    def forward(ctx, wav, gn_w, gn_b, eps, kernels, strides, *weights):
        L.require_device()
        B, Ls = wav.shape
        C = weights[0].shape[0]
        n = len(weights)
        dev = wav.device
        T, R = _fe_rows(Ls, kernels, strides)
        wav = wav.contiguous().float()
        Lp = Ls                            # reads past the clip (padding rows only) are predicated to zero
        pad = 8                            # zeroed tail rows: the last padded rows of the last sample read past B*R
        u0 = torch.empty((B * R[0], C), dtype=torch.bfloat16, device=dev)
        L.call("tavk_conv0_fwd", wav.data_ptr(), weights[0].data_ptr(), None, u0.data_ptr(), B, Lp, R[0], T[0], C,
               kernels[0], strides[0])
        acts = [torch.zeros((B * R[l] + pad, C), dtype=torch.bfloat16, device=dev) for l in range(n)]
        pres = [torch.empty((B * R[l], C), dtype=torch.bfloat16, device=dev) for l in range(n)]
        mean = torch.empty((B, C), dtype=torch.float32, device=dev)
        rstd = torch.empty((B, C), dtype=torch.float32, device=dev)
        sums = torch.empty((B, 2, C), dtype=torch.float32, device=dev)
        L.call("tavk_groupnorm_gelu_fwd", u0.data_ptr(), gn_w.data_ptr(), gn_b.data_ptr(), pres[0].data_ptr(),
               acts[0].data_ptr(), mean.data_ptr(), rstd.data_ptr(), sums.data_ptr(), B, R[0], T[0], C, float(eps))
        for l in range(1, n):
            k, st = kernels[l], strides[l]
            wk = L.cast_bf16(weights[l].detach().permute(0, 2, 1).contiguous().view(C, k * C))
            L.gemm(acts[l - 1], wk, pres[l], M=B * R[l], N=C, K=k * C, lda=st * C, out2=acts[l][:B * R[l]],
                   epilogue=L.EPI_GELU)
        ctx.save_for_backward(wav, u0, mean, rstd, gn_w, *weights, *acts[:-1], *pres)
        ctx.meta = (B, Lp, C, n, T, R, tuple(kernels), tuple(strides))
        return acts[-1][:B * R[-1]].view(B, R[-1], C)[:, :T[-1]].float()

    @staticmethod
    def backward(ctx, dy):
        B, Lp, C, n, T, R, kernels, strides = ctx.meta
        sv = ctx.saved_tensors
        wav, u0, mean, rstd, gn_w = sv[:5]
        weights, acts, pres = sv[5:5 + n], sv[5 + n:5 + 2 * n - 1], sv[5 + 2 * n - 1:]
        dev = dy.device
        # gradient w.r.t. the last pre-activation, in the padded row layout (padding rows stay exactly zero)
        pre_top = pres[-1].view(B, R[-1], C)[:, :T[-1]].float()
        dtop = torch.ops.aten.gelu_backward(dy.contiguous().float(), pre_top)
        D = [-(-kernels[l] // strides[l]) for l in range(n)]
        lead = D[-1] - 1
        dpre = torch.zeros((lead + B * R[-1], C), dtype=torch.bfloat16, device=dev)
        dpre[lead:].view(B, R[-1], C)[:, :T[-1]].copy_(dtop)
        grads = [None] * n
        for l in range(n - 1, 0, -1):
            k, st, d = kernels[l], strides[l], D[l]
            M = B * R[l]
            cur = dpre[d - 1:]                           # [M, C] rows of this layer's dpre
            # wgrad: dWk[n, (tap, c)] = sum_m dpre[m, n] * a_{l-1}[m*st*C + (tap, c)]
            dwk = torch.zeros((C, k * C), dtype=torch.float32, device=dev)
            tiles = ((C + 127) // 128) * ((k * C + 255) // 256)
            ks = max(1, min(148 // tiles, (M + 511) // 512))
            L.gemm(cur, acts[l - 1], dwk, M=C, N=k * C, K=M, lda=C, ldb=st * C, a_mn=True, b_mn=True, accumulate=True,
                   k_splits=ks)
            grads[l] = dwk.view(C, k, C).permute(0, 2, 1)
            # dgrad (+ GELU' of layer l-1): out rows = s input rows each
            w = weights[l].detach()
            b2 = torch.zeros((st, C, d, C), dtype=torch.bfloat16, device=dev)
            for r in range(st):
                for dd in range(d):
                    tap = r + st * dd
                    if tap < k:
                        b2[r, :, d - 1 - dd, :].copy_(w[:, :, tap].t())
            lead_prev = D[l - 1] - 1 if l - 1 >= 1 else 0
            nxt = torch.empty((lead_prev + B * R[l - 1], C), dtype=torch.bfloat16, device=dev)
            if lead_prev:
                nxt[:lead_prev].zero_()
            L.gemm(dpre, b2.view(st * C, d * C), nxt[lead_prev:].view(M, st * C), M=M, N=st * C, K=d * C, lda=C,
                   aux=pres[l - 1].view(M, st * C), epilogue=L.EPI_GELU_BWD)
            dpre = nxt
        # layer 0 (dpre now holds dz0): GroupNorm backward in place -> du0; affine gradients from its per-sample sums;
        # conv0 weight gradient = wgrad GEMM of du0 against the waveform's window matrix (C_in = 1: 16 bf16 per row)
        k0, s0 = kernels[0], strides[0]
        sums = torch.empty((B, 2, C), dtype=torch.float32, device=dev)
        L.call("tavk_groupnorm_bwd", dpre.data_ptr(), u0.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gn_w.data_ptr(),
               dpre.data_ptr(), sums.data_ptr(), B, R[0], T[0], C)
        dgb, dgw = sums[:, 0].sum(dim=0), sums[:, 1].sum(dim=0)
        win = torch.empty((B * R[0], 16), dtype=torch.bfloat16, device=dev)
        L.call("tavk_wave_windows", wav.data_ptr(), win.data_ptr(), B, Lp, R[0], T[0], k0, s0)
        dw0 = torch.zeros((C, 16), dtype=torch.float32, device=dev)
        M0 = B * R[0]
        L.gemm(dpre, win, dw0, M=C, N=16, K=M0, lda=C, ldb=16, a_mn=True, b_mn=True, accumulate=True,
               k_splits=max(1, min(148 // ((C + 127) // 128), (M0 + 2047) // 2048)))
        grads[0] = dw0[:, :k0].reshape(weights[0].shape)
        return (None, dgw, dgb, None, None, None, *grads)


def _regroup_taps_bf16(w, st, d, k):
    """conv weight [C_out, C_in, k] -> bf16 [st*C_in, d*C_out]: B operand of the dgrad GEMM that writes `st` input rows per
    output row (taps regrouped by input phase r and delay dd; see _ConvFeatureEncoderFn)."""
    C = w.shape[0]
    b2 = torch.zeros((st, C, d, C), dtype=torch.bfloat16, device=w.device)
    for r in range(st):
        for dd in range(d):
            tap = r + st * dd
            if tap < k:
                b2[r, :, d - 1 - dd, :].copy_(w[:, :, tap].t())
    return b2.view(st * C, d * C)


class _ConvFeatureEncoderLNFn(torch.autograd.Function):
    """HF Wav2Vec2FeatureEncoder with feat_extract_norm="layer" (conv bias, LayerNorm over channels, GELU per layer — the
    wav2vec2-large family the reference loads, models/tav.py:257,455) in the same channels-last / padded-row layout as
    _ConvFeatureEncoderFn: layer 0 is the streaming conv kernel, every later Conv1d + bias is one tcgen05 GEMM with
    overlapping TMA rows (bf16 output through the TMA-store epilogue), and LayerNorm + GELU is one warp-per-row kernel
    (csrc/conv_frontend.cu chan_ln_gelu_*).  Backward per layer: LN'/GELU' kernel (z recomputed), conv-bias gradient as a
    column sum, wgrad GEMM, dgrad GEMM writing `stride` input rows per output row."""

    @staticmethod
    def forward(ctx, wav, eps, kernels, strides, *params):
        L.require_device()
        n = len(kernels)
        ws, bs, gs, bes = params[0:n], params[n:2 * n], params[2 * n:3 * n], params[3 * n:4 * n]
        B, Ls = wav.shape
        C = ws[0].shape[0]
        dev = wav.device
        T, R = _fe_rows(Ls, kernels, strides)
        wav = wav.contiguous().float()
        pad = 8
        us = [torch.empty((B * R[l], C), dtype=torch.bfloat16, device=dev) for l in range(n)]
        acts = [torch.zeros((B * R[l] + pad, C), dtype=torch.bfloat16, device=dev) for l in range(n)]
        means = [torch.empty((B * R[l],), dtype=torch.float32, device=dev) for l in range(n)]
        rstds = [torch.empty((B * R[l],), dtype=torch.float32, device=dev) for l in range(n)]
        L.call("tavk_conv0_fwd", wav.data_ptr(), ws[0].data_ptr(), bs[0].data_ptr(), us[0].data_ptr(), B, Ls, R[0], T[0], C,
               kernels[0], strides[0])
        for l in range(n):
            if l >= 1:
                k, st = kernels[l], strides[l]
                wk = L.cast_bf16(ws[l].detach().permute(0, 2, 1).contiguous().view(C, k * C))
                L.gemm(acts[l - 1], wk, us[l], M=B * R[l], N=C, K=k * C, lda=st * C, bias=bs[l])
            L.call("tavk_chan_ln_gelu_fwd", us[l].data_ptr(), gs[l].data_ptr(), bes[l].data_ptr(), acts[l].data_ptr(),
                   means[l].data_ptr(), rstds[l].data_ptr(), B, R[l], T[l], C, float(eps))
        ctx.save_for_backward(wav, *params, *us, *acts[:-1], *means, *rstds)
        ctx.meta = (B, Ls, C, n, T, R, tuple(kernels), tuple(strides))
        return acts[-1][:B * R[-1]].view(B, R[-1], C)[:, :T[-1]].float()

    @staticmethod
    def backward(ctx, dy):
        B, Ls, C, n, T, R, kernels, strides = ctx.meta
        sv = ctx.saved_tensors
        wav = sv[0]
        ws, bs, gs, bes = sv[1:1 + n], sv[1 + n:1 + 2 * n], sv[1 + 2 * n:1 + 3 * n], sv[1 + 3 * n:1 + 4 * n]
        o = 1 + 4 * n
        us, acts = sv[o:o + n], sv[o + n:o + 2 * n - 1]
        means, rstds = sv[o + 2 * n - 1:o + 3 * n - 1], sv[o + 3 * n - 1:o + 4 * n - 1]
        dev = dy.device
        D = [-(-kernels[l] // strides[l]) for l in range(n)]
        da = torch.zeros((B * R[-1], C), dtype=torch.bfloat16, device=dev)
        da.view(B, R[-1], C)[:, :T[-1]].copy_(dy)
        dws, dbs, dgs, dbes = [None] * n, [None] * n, [None] * n, [None] * n
        for l in range(n - 1, -1, -1):
            M = B * R[l]
            lead = D[l] - 1 if l >= 1 else 0
            du = torch.zeros((lead + M, C), dtype=torch.bfloat16, device=dev)     # leading zero rows: dgrad row windows
            dgs[l] = torch.zeros((C,), dtype=torch.float32, device=dev)
            dbes[l] = torch.zeros((C,), dtype=torch.float32, device=dev)
            L.call("tavk_chan_ln_gelu_bwd", da.data_ptr(), us[l].data_ptr(), means[l].data_ptr(), rstds[l].data_ptr(),
                   gs[l].data_ptr(), bes[l].data_ptr(), du[lead:].data_ptr(), dgs[l].data_ptr(), dbes[l].data_ptr(), B,
                   R[l], T[l], C)
            dbs[l] = torch.empty((C,), dtype=torch.float32, device=dev)
            L.colsum(du[lead:], dbs[l], M=M, N=C)
            if l == 0:
                break
            k, st, d = kernels[l], strides[l], D[l]
            dwk = torch.zeros((C, k * C), dtype=torch.float32, device=dev)
            tiles = ((C + 127) // 128) * ((k * C + 255) // 256)
            ks = max(1, min(148 // tiles, (M + 511) // 512))
            L.gemm(du[lead:], acts[l - 1], dwk, M=C, N=k * C, K=M, lda=C, ldb=st * C, a_mn=True, b_mn=True, accumulate=True,
                   k_splits=ks)
            dws[l] = dwk.view(C, k, C).permute(0, 2, 1)
            da = torch.empty((B * R[l - 1], C), dtype=torch.bfloat16, device=dev)
            L.gemm(du, _regroup_taps_bf16(ws[l].detach(), st, d, k), da.view(M, st * C), M=M, N=st * C, K=d * C, lda=C)
        # layer 0: conv weight gradient = wgrad GEMM of du0 against the waveform's window matrix
        k0, s0 = kernels[0], strides[0]
        M0 = B * R[0]
        win = torch.empty((M0, 16), dtype=torch.bfloat16, device=dev)
        L.call("tavk_wave_windows", wav.data_ptr(), win.data_ptr(), B, Ls, R[0], T[0], k0, s0)
        dw0 = torch.zeros((C, 16), dtype=torch.float32, device=dev)
        L.gemm(du, win, dw0, M=C, N=16, K=M0, lda=C, ldb=16, a_mn=True, b_mn=True, accumulate=True,
               k_splits=max(1, min(148 // ((C + 127) // 128), (M0 + 2047) // 2048)))
        dws[0] = dw0[:, :k0].reshape(ws[0].shape)
        return (None, None, None, None, *dws, *dbs, *dgs, *dbes)


def feature_extractor_cl(w2v, wav):
    """HF Wav2Vec2FeatureEncoder.forward in channels-last form: [B, L] -> [B, frames, C] (fp32), both norm families on
    the kernel library (there is no cuDNN / autocast fallback: unsupported configurations raise)."""
    fe = w2v.feature_extractor
    c = w2v.config
    layers = fe.conv_layers
    if not wav.is_cuda:
        raise RuntimeError("the Wav2Vec2 feature encoder runs on the sm_100a kernel path only (no CPU fallback)")
    same_c = len(set(l.conv.weight.shape[0] for l in layers)) == 1
    C = layers[0].conv.weight.shape[0]
    if c.feat_extract_activation != "gelu" or max(c.conv_kernel) > 16 or not same_c:
        raise NotImplementedError("Wav2Vec2 feature encoder: GELU, kernels <= 16 and one channel count expected")
    if c.feat_extract_norm == "group" and not c.conv_bias and C % 16 == 0:
        gn = layers[0].layer_norm
        return _ConvFeatureEncoderFn.apply(wav, gn.weight, gn.bias, gn.eps, tuple(c.conv_kernel), tuple(c.conv_stride),
                                           *[l.conv.weight for l in layers])
    if c.feat_extract_norm == "layer" and c.conv_bias and C % 256 == 0 and C <= 1024:
        return _ConvFeatureEncoderLNFn.apply(wav, layers[0].layer_norm.eps, tuple(c.conv_kernel), tuple(c.conv_stride),
                                             *[l.conv.weight for l in layers], *[l.conv.bias for l in layers],
                                             *[l.layer_norm.weight for l in layers], *[l.layer_norm.bias for l in layers])
    raise NotImplementedError("Wav2Vec2 feature encoder: feat_extract_norm=%r with conv_bias=%r is not on the kernel path"
                              % (c.feat_extract_norm, c.conv_bias))


# ------------------------------------------------------------------------------------------------ patch embedding
_cols_cache = {"src": None, "key": None, "cols": None}


def video_patches_bf16(pixel_values, tubelet, patch):
    """[B, T, C, H, W] -> bf16 [B, (T/tub)(H/p)(W/p), C*tub*p*p] in Conv3d weight order (c, dt, dy, dx); cached for
    the second consumer of the SAME tensor within a step (PreFormer and TAVForMAE read the same clip).  The cache is
    keyed on the tensor object (weak reference) and its version counter, never on its address: the allocator hands the
    block of a freed clip to the next batch, which then has the same data_ptr, shape and version 0."""
    src = _cols_cache["src"]
    key = (pixel_values._version, tubelet, patch)
    if src is not None and src() is pixel_values and _cols_cache["key"] == key:
        return _cols_cache["cols"]
    B, T, C, H, W = pixel_values.shape
    tp, hp, wp = T // tubelet, H // patch, W // patch
    v = pixel_values.reshape(B, tp, tubelet, C, hp, patch, wp, patch).permute(0, 1, 4, 6, 3, 2, 5, 7)
    cols = torch.empty((B, tp, hp, wp, C, tubelet, patch, patch), dtype=torch.bfloat16, device=pixel_values.device)
    cols.copy_(v)  # one fused cast + permute pass over the clip
    cols = cols.view(B, tp * hp * wp, C * tubelet * patch * patch)
    _cols_cache["src"], _cols_cache["key"], _cols_cache["cols"] = weakref.ref(pixel_values), key, cols
    return cols


class _PatchProjFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cols, weight, bias, pos_rows):
        M, K = cols.shape
        N = weight.shape[0]
        w_bf = L.cast_bf16(weight.detach().reshape(N, K))
        y = torch.empty((M, N), dtype=torch.float32, device=cols.device)
        L.gemm(cols, w_bf, y, M=M, N=N, K=K, bias=bias, resid=pos_rows)
        ctx.save_for_backward(cols)
        ctx.wshape = weight.shape
        return y

    @staticmethod
    def backward(ctx, dy):
        (cols,) = ctx.saved_tensors
        M, K = cols.shape
        dy = dy.contiguous().float()
        N = dy.shape[1]
        dy_bf = L.cast_bf16(dy)
        dw = _wgrad(dy_bf, cols, N, K, M).view(ctx.wshape)
        db = torch.empty((N,), dtype=torch.float32, device=dy.device)
        L.colsum(dy, db, M=M, N=N)
        return None, dw, db, None


def video_embeddings(emb, pixel_values, bool_masked_pos, keep_count=None):
    """HF VideoMAEEmbeddings.forward(pixel_values, bool_masked_pos): patch projection + fixed sinusoid table, keeping
    the rows where ``~bool_masked_pos`` (every row keeps the same number of tokens: HF requirement)."""
    pe = emb.patch_embeddings
    B, T, C, H, W = pixel_values.shape
    if C != pe.num_channels:
        raise ValueError("Make sure that the channel dimension of the pixel values match with the one set in the configuration.")
    if H != pe.image_size[0] or W != pe.image_size[1]:
        raise ValueError(f"Input image size ({H}*{W}) doesn't match model ({pe.image_size[0]}*{pe.image_size[1]}).")
    if pe.patch_size[0] != pe.patch_size[1]:
        raise NotImplementedError("square patches only")
    cols = video_patches_bf16(pixel_values.float(), pe.tubelet_size, pe.patch_size[0])
    n_tok, K = cols.shape[1], cols.shape[2]
    pos = getattr(emb, "_tavk_pos", None)   # HF keeps the sinusoid table as a plain CPU tensor: cache a device copy
    if pos is None or pos.device != cols.device:
        pos = emb._tavk_pos = emb.position_embeddings.detach().to(device=cols.device, dtype=torch.float32).reshape(n_tok, -1)
    Hd = pos.shape[1]
    if bool_masked_pos is None:
        rows, pos_rows, keep_count = cols.reshape(B * n_tok, K), pos.repeat(B, 1), n_tok
    else:
        keep = ~bool_masked_pos.to(cols.device)
        if keep_count is None:
            keep_count = int(keep[0].sum().item())
        # stable descending sort of the keep flags lists kept positions first, in their original order
        idx = torch.sort(keep.to(torch.uint8), dim=1, descending=True, stable=True).indices[:, :keep_count]
        rows = torch.gather(cols, 1, idx[:, :, None].expand(-1, -1, K)).reshape(B * keep_count, K)
        pos_rows = pos[idx.reshape(-1)]
    y = _PatchProjFn.apply(rows, pe.projection.weight, pe.projection.bias, pos_rows)
    return y.view(B, keep_count, Hd)


# ------------------------------------------------------------------------------------------------ positional conv
def _pad_rows_bf16(x_bf, left, Tp, tail):
    """x_bf [B,T,H] bf16 -> zero-padded [B*Tp + tail, H]: sample b occupies rows b*Tp + left .. b*Tp + left + T - 1."""
    B, T, H = x_bf.shape
    buf = torch.zeros((B * Tp + tail, H), dtype=torch.bfloat16, device=x_bf.device)
    buf[:B * Tp].view(B, Tp, H)[:, left:left + T].copy_(x_bf)
    return buf


def _pack_taps_bf16(w, Cg, k):
    """conv weight [H, Cg, k] -> bf16 [H, k*64] with column (tap, ic) and zeros for ic >= Cg: one 64-wide k-block
    per tap, so a k-block of the GEMM is a TMA box of 64 channels of one (shifted) activation row."""
    H = w.shape[0]
    wp = torch.zeros((H, k, 64), dtype=torch.bfloat16, device=w.device)
    wp[:, :, :Cg].copy_(w.permute(0, 2, 1))
    return wp.view(H, k * 64)


def _sliding_conv(xpad, wp, bias, rows, H, G, Cg, k):
    """out[m, g*Cg + oc] = sum_tap sum_ic xpad[m + tap, g*Cg + ic] * wp[g*Cg + oc, tap*64 + ic]   (m < rows): an
    implicit GEMM — the k-block walk of the A operand steps one activation ROW per tap (a_kstep = row pitch), the 16
    groups are 16 problems of one launch; nothing like an im2col matrix is ever written."""
    out = torch.empty((rows, H), dtype=torch.float32, device=xpad.device)
    L.gemm(xpad, wp, out, M=rows, N=Cg, K=k * 64, lda=H, ldb=k * 64, bias=bias, groups=G, a_kstep=H, a_g_k=Cg,
           b_g_mn=Cg, out_g_col=Cg, a_rows=rows, a_cols=k * H, b_rows=H, b_cols=k * 64)
    return out


class _GroupedConv1dSameFn(torch.autograd.Function):
    """y[b,t,:] = sum_k conv_weight[:, :, k] · xpad[b, t+k, :] for t < T (padding k/2 each side, last frame dropped)."""

    @staticmethod
    def forward(ctx, x, weight, bias, groups):
        B, T, H = x.shape
        Cg, k = weight.shape[1], weight.shape[2]
        if Cg > 64 or Cg % 8 != 0:
            raise NotImplementedError("positional conv: at most 64 channels per group, multiple of 8 (got %d)" % Cg)
        Tp = T + k
        x_bf = L.cast_bf16(x.contiguous().float())
        xpad = _pad_rows_bf16(x_bf, k // 2, Tp, k)
        y = _sliding_conv(xpad, _pack_taps_bf16(weight.detach(), Cg, k), bias, B * Tp, H, groups, Cg, k)
        ctx.save_for_backward(xpad, weight)
        ctx.meta = (B, T, H, Cg, k, groups, bias is not None)
        return y.view(B, Tp, H)[:, :T].contiguous()

    @staticmethod
    def backward(ctx, dy):
        xpad, weight = ctx.saved_tensors
        B, T, H, Cg, k, G, has_bias = ctx.meta
        Tp = T + k
        dy = dy.contiguous().float()
        db = None
        if has_bias:
            db = torch.empty((H,), dtype=torch.float32, device=dy.device)
            L.colsum(dy.view(B * T, H), db, M=B * T, N=H)
        left = k // 2 - 1
        dypad = _pad_rows_bf16(L.cast_bf16(dy), left, Tp, k)
        # wgrad: dW[g*Cg+oc, ic, tap] = sum_m dY[m, g*Cg+oc] * xpad[m + tap, g*Cg+ic]  (m over the padded row grid, where
        # dY rows are zero outside the T valid frames): A = dY (MN-major), B = xpad (MN-major), one 64-wide box of N per
        # tap reading rows shifted by the tap (b_box_k_shift = 1); output [H, k*64] regrouped to [H, Cg, k] afterwards.
        dwt = torch.zeros((H, k * 64), dtype=torch.float32, device=dy.device)
        rows = B * Tp
        ks = max(1, min(4, 148 * 2 // (G * (k * 64 // 256))))
        L.gemm(dypad[left:], xpad, dwt, M=Cg, N=k * 64, K=rows, lda=H, ldb=H, a_mn=True, b_mn=True, accumulate=True,
               k_splits=ks, groups=G, a_g_mn=Cg, b_g_mn=Cg, b_box_k_shift=1, out_g_row=Cg, a_rows=rows, a_cols=H,
               b_rows=rows + k, b_cols=H)
        dw = dwt.view(H, k, 64)[:, :, :Cg].permute(0, 2, 1)
        # dgrad: dX[s] = sum_u dYpad[s+u] · W[.., k-1-u] with dYpad shifted by k/2-1 -> same routine, flipped taps and
        # (oc, ic) transposed inside every group
        wf = weight.detach().view(G, Cg, Cg, k).flip(-1).transpose(1, 2).reshape(H, Cg, k)
        dx = _sliding_conv(dypad, _pack_taps_bf16(wf, Cg, k), None, rows, H, G, Cg, k)
        return dx.view(B, Tp, H)[:, :T].contiguous(), dw, db, None


def _weight_normed(conv):
    """weight = g * v / ||v|| (norm over all dims but `dim`) from the parametrization's own tensors with plain torch
    ops (autograd differentiates them): avoids ATen's slow last-dim weight_norm kernels on the 768x48x128 weight."""
    par = getattr(conv, "parametrizations", None)
    if par is None or not hasattr(par, "weight"):
        return conv.weight
    g, v = par.weight.original0, par.weight.original1
    dims = [d for d in range(v.dim()) if g.shape[d] == 1]
    return v * (g / v.pow(2).sum(dim=dims, keepdim=True).sqrt())


def pos_conv_embed(pc, hidden):
    """HF Wav2Vec2PositionalConvEmbedding.forward: GELU(SamePad(Conv1d(hidden^T)))^T."""
    conv = pc.conv
    k = conv.kernel_size[0]
    if k % 2 != 0 or conv.padding[0] != k // 2 or conv.stride[0] != 1 or conv.dilation[0] != 1:
        raise NotImplementedError("positional conv: even kernel with padding k/2 expected")
    y = _GroupedConv1dSameFn.apply(hidden, _weight_normed(conv), conv.bias, conv.groups)
    return F.gelu(y)
