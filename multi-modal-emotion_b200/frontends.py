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
import torch
import torch.nn.functional as F

from . import _lib as L
from .engine import _wgrad

# The 7-layer Conv1d feature extractor still runs in cuDNN; under bf16 autocast it uses the
# bf16 tensor-core kernels (fp32 accumulate; GroupNorm / LayerNorm stay fp32) instead of TF32 + layout conversions.
FE_AUTOCAST = True


def feature_extractor(w2v, wav):
    """HF Wav2Vec2FeatureEncoder.forward: [B, L] -> [B, 512, frames] (fp32 out)."""
    if FE_AUTOCAST and wav.is_cuda:
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return w2v.feature_extractor(wav).float()
    return w2v.feature_extractor(wav)


# ------------------------------------------------------------------------------------------------ patch embedding
_cols_cache = {"key": None, "cols": None}


def video_patches_bf16(pixel_values, tubelet, patch):
    """[B, T, C, H, W] -> bf16 [B, (T/tub)(H/p)(W/p), C*tub*p*p] in Conv3d weight order (c, dt, dy, dx); cached for
    the second consumer of the same clip within a step."""
    key = (pixel_values.data_ptr(), pixel_values._version, tuple(pixel_values.shape), tubelet, patch)
    if _cols_cache["key"] == key:
        return _cols_cache["cols"]
    B, T, C, H, W = pixel_values.shape
    tp, hp, wp = T // tubelet, H // patch, W // patch
    v = pixel_values.view(B, tp, tubelet, C, hp, patch, wp, patch).permute(0, 1, 4, 6, 3, 2, 5, 7)
    cols = torch.empty((B, tp, hp, wp, C, tubelet, patch, patch), dtype=torch.bfloat16, device=pixel_values.device)
    cols.copy_(v)  # one fused cast + permute pass over the clip
    cols = cols.view(B, tp * hp * wp, C * tubelet * patch * patch)
    _cols_cache["key"], _cols_cache["cols"] = key, cols
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
def _windows_bf16(x_bf, k, left, right):
    """x_bf [B,T,H] bf16 -> sliding-window matrix [B*T, H*k] with column order (channel, tap): window t covers padded
    rows t .. t+k-1."""
    B, T, H = x_bf.shape
    xp = F.pad(x_bf, (0, 0, left, right))
    win = xp.unfold(1, k, 1)[:, :T]          # [B, T, H, k] view
    out = torch.empty((B, T, H, k), dtype=torch.bfloat16, device=x_bf.device)
    out.copy_(win)
    return out.view(B * T, H * k)


def _grouped_gemm(cols, w_bf, bias, M, H, G, k):
    """out[:, g*Cg:(g+1)*Cg] = cols[:, g*Cg*k:(g+1)*Cg*k] @ w_bf[g*Cg:(g+1)*Cg]^T (+ bias)."""
    Cg = H // G
    Kg = Cg * k
    out = torch.empty((M, H), dtype=torch.float32, device=cols.device)
    for g in range(G):
        L.gemm(cols[:, g * Kg:(g + 1) * Kg], w_bf[g * Cg:(g + 1) * Cg], out[:, g * Cg:(g + 1) * Cg], M=M, N=Cg, K=Kg,
               lda=H * k, ldb=Kg, bias=None if bias is None else bias[g * Cg:(g + 1) * Cg])
    return out


class _GroupedConv1dSameFn(torch.autograd.Function):
    """y[b,t,:] = sum_k conv_weight[:, :, k] · xpad[b, t+k, :] for t < T (padding k/2 each side, last frame dropped)."""

    @staticmethod
    def forward(ctx, x, weight, bias, groups):
        B, T, H = x.shape
        Cg, k = weight.shape[1], weight.shape[2]
        x_bf = L.cast_bf16(x.contiguous().float())
        cols = _windows_bf16(x_bf, k, k // 2, k // 2)
        w_bf = L.cast_bf16(weight.detach().reshape(H, Cg * k))
        y = _grouped_gemm(cols, w_bf, bias, B * T, H, groups, k)
        ctx.save_for_backward(x_bf, weight)
        ctx.meta = (B, T, H, Cg, k, groups, bias is not None)
        return y.view(B, T, H)

    @staticmethod
    def backward(ctx, dy):
        x_bf, weight = ctx.saved_tensors
        B, T, H, Cg, k, G, has_bias = ctx.meta
        M = B * T
        dy = dy.contiguous().float().view(M, H)
        dy_bf = L.cast_bf16(dy)
        db = None
        if has_bias:
            db = torch.empty((H,), dtype=torch.float32, device=dy.device)
            L.colsum(dy, db, M=M, N=H)
        # wgrad: dW_g[oc, (ic,k)] = sum_m dY[m, g*Cg+oc] * cols[m, g*Cg*k + (ic,k)]   (window matrix recomputed)
        cols = _windows_bf16(x_bf, k, k // 2, k // 2)
        dw = torch.zeros((H, Cg * k), dtype=torch.float32, device=dy.device)
        Kg = Cg * k
        ks = max(1, min(148 // ((Kg + 255) // 256), (M + 511) // 512))
        for g in range(G):
            L.gemm(dy_bf[:, g * Cg:(g + 1) * Cg], cols[:, g * Kg:(g + 1) * Kg], dw[g * Cg:(g + 1) * Cg], M=Cg, N=Kg, K=M,
                   lda=H, ldb=H * k, a_mn=True, b_mn=True, accumulate=True, k_splits=ks)
        del cols
        # dgrad: dX[s] = sum_u dYpad[s+u] · W[.., k-1-u] with dYpad shifted by k/2-1 -> same routine, flipped taps and
        # (oc, ic) transposed inside every group
        wf = weight.detach().view(G, Cg, Cg, k).flip(-1).transpose(1, 2).reshape(H, Cg * k)
        wf_bf = L.cast_bf16(wf.contiguous())
        dcols = _windows_bf16(dy_bf.view(B, T, H), k, k // 2 - 1, k // 2)
        dx = _grouped_gemm(dcols, wf_bf, None, M, H, G, k)
        return dx.view(B, T, H), dw.view(H, Cg, k), db, None


def pos_conv_embed(pc, hidden):
    """HF Wav2Vec2PositionalConvEmbedding.forward: GELU(SamePad(Conv1d(hidden^T)))^T."""
    conv = pc.conv
    k = conv.kernel_size[0]
    if k % 2 != 0 or conv.padding[0] != k // 2 or conv.stride[0] != 1 or conv.dilation[0] != 1:
        raise NotImplementedError("positional conv: even kernel with padding k/2 expected")
    y = _GroupedConv1dSameFn.apply(hidden, conv.weight, conv.bias, conv.groups)
    return F.gelu(y)
