"""GPU parity tests of every libtavk.so entry point against a plain PyTorch fp32 restatement of the same op.
All calls go through the C ABI (ctypes, multi_modal_emotion_b200._lib).  Tolerances are stated per test: bf16
tensor-core ops compare against an fp32 computation on the same bf16-rounded inputs."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    from multi_modal_emotion_b200 import _lib

    _lib.require_device()
    return _lib


def rel_l2(a, b):
    """Relative L2 error; a reference that is (numerically) zero is compared on an absolute 1e-6 floor."""
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-6 * max(1.0, b.numel() ** 0.5))).item()


def gen(seed=0):
    return torch.Generator(device="cpu").manual_seed(seed)


# ------------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("shape", [(128, 128, 64), (300, 768, 768), (5168, 2304, 768), (1000, 3072, 776), (16, 768, 1024)])
@pytest.mark.parametrize("block_n", [128, 256])
@pytest.mark.parametrize("cta_pair", [1, 2])       # single-CTA tiles | CTA pairs (cta_group::2) wherever M > 128
def test_gemm_operand_majors(L, a_mn, b_mn, shape, block_n, cta_pair):
    M, N, K = shape
    if (a_mn and M % 8) or (b_mn and N % 8):
        pytest.skip("MN-major operand needs a row count that is a multiple of 8")
    g = gen(1)
    A = (torch.randn(M, K, generator=g) * 0.5).cuda().bfloat16()
    B = (torch.randn(N, K, generator=g) * 0.5).cuda().bfloat16()
    ref = A.float() @ B.float().t()
    out = torch.full((M, N), float("nan"), device="cuda")
    L.gemm(A.t().contiguous() if a_mn else A, B.t().contiguous() if b_mn else B, out, M=M, N=N, K=K, a_mn=a_mn,
           b_mn=b_mn, block_n=block_n, cta_pair=cta_pair)
    # fp32 accumulation of exact bf16 products: only summation-order noise is allowed
    assert rel_l2(out, ref) < 1e-5
    assert (out - ref).abs().max().item() < 1e-3 * math.sqrt(K)


@pytest.mark.parametrize("cta_pair", [1, 2])
def test_gemm_epilogues(L, cta_pair, monkeypatch):
    _gemm = L.gemm
    monkeypatch.setattr(L, "gemm", lambda *a, **k: _gemm(*a, cta_pair=cta_pair, **k))
    g = gen(2)
    M, N, K, S = 646, 768, 768, 323
    A = (torch.randn(M, K, generator=g) * 0.5).cuda().bfloat16()
    W = (torch.randn(N, K, generator=g) * 0.05).cuda().bfloat16()
    bias = torch.randn(N, generator=g).cuda()
    resid = torch.randn(M, N, generator=g).cuda()
    rowbias = torch.randn(M // S, N, generator=g).cuda()
    acc = A.float() @ W.float().t()
    # bias + per-sample row-bias + residual -> f32
    out = torch.empty(M, N, device="cuda")
    L.gemm(A, W, out, M=M, N=N, K=K, bias=bias, resid=resid, rowbias=rowbias, rows_per_group=S)
    ref = acc + bias + resid + rowbias.repeat_interleave(S, dim=0)
    assert (out - ref).abs().max().item() < 2e-4
    # bf16 output with alpha
    outb = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    L.gemm(A, W, outb, M=M, N=N, K=K, bias=bias, alpha=0.5)
    assert rel_l2(outb, 0.5 * acc + bias) < 4e-3
    # GELU epilogue: pre-activation and activation, both bf16
    pre = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    act = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    L.gemm(A, W, pre, M=M, N=N, K=K, bias=bias, out2=act, epilogue=L.EPI_GELU)
    assert rel_l2(pre, acc + bias) < 4e-3
    assert rel_l2(act, torch.nn.functional.gelu(acc + bias)) < 4e-3
    # GELU backward epilogue: acc * gelu'(aux)
    x = pre.float().requires_grad_(True)
    torch.nn.functional.gelu(x).backward(torch.ones_like(x))
    outg = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    L.gemm(A, W, outg, M=M, N=N, K=K, aux=pre, epilogue=L.EPI_GELU_BWD)
    assert rel_l2(outg, acc * x.grad) < 4e-3
    # GELU_GRAD epilogue: the forward stores gelu'(pre) (bf16) next to gelu(pre); MUL epilogue: acc * aux (+ column sums)
    dg = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    act2 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    L.gemm(A, W, dg, M=M, N=N, K=K, bias=bias, out2=act2, epilogue=L.EPI_GELU_GRAD)
    xe = (acc + bias).requires_grad_(True)
    torch.nn.functional.gelu(xe).backward(torch.ones_like(xe))
    assert rel_l2(act2, torch.nn.functional.gelu(acc + bias)) < 4e-3
    assert rel_l2(dg, xe.grad) < 4e-3
    assert (dg.float() - xe.grad).abs().max().item() < 8e-3          # bf16 rounding of values in [-0.13, 1.13]
    outm = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    cs = torch.zeros((N,), device="cuda")
    L.gemm(A, W, outm, M=M, N=N, K=K, aux=dg, epilogue=L.EPI_MUL, colsum=cs)
    assert rel_l2(outm, acc * dg.float()) < 4e-3
    assert rel_l2(cs, (acc * dg.float()).sum(dim=0)) < 2e-3
    # fused column sums of the stored result (the bias gradient of the producing Linear), accumulated into colsum
    cs = torch.full((N,), 2.0, device="cuda")
    L.gemm(A, W, outg, M=M, N=N, K=K, aux=pre, epilogue=L.EPI_GELU_BWD, colsum=cs)
    assert rel_l2(cs, 2.0 + (acc * x.grad).sum(dim=0)) < 2e-3
    cs = torch.zeros((N,), device="cuda")
    out32 = torch.empty(M, N, device="cuda")
    L.gemm(A, W, out32, M=M, N=N, K=K, bias=bias, colsum=cs)
    assert rel_l2(cs, (acc + bias).sum(dim=0)) < 2e-3
    # accumulate + split-K (wgrad form)
    base = torch.randn(N, K, generator=g).cuda()
    out = base.clone()
    dY = (torch.randn(M, N, generator=g) * 0.5).cuda().bfloat16()
    L.gemm(dY, A, out, M=N, N=K, K=M, a_mn=True, b_mn=True, accumulate=True, k_splits=4)
    ref = base + dY.float().t() @ A.float()
    assert rel_l2(out, ref) < 1e-5


def test_gemm_argument_errors(L):
    A = torch.zeros(16, 60, device="cuda", dtype=torch.bfloat16)
    out = torch.zeros(16, 16, device="cuda")
    with pytest.raises(L.TavkError):
        L.gemm(A, A, out, M=16, N=16, K=60, lda=60, ldb=60)  # lda not a multiple of 8
    with pytest.raises(L.TavkError):
        L.gemm(A, A, out, M=16, N=12, K=64, lda=64, ldb=64)  # N % 8


# ------------------------------------------------------------------------------------------------ attention
def _attn_ref(q, k, v, bias, B, S, nh):
    qf = q.float().reshape(B, S, nh, 64).transpose(1, 2).detach().requires_grad_(True)
    kf = k.float().reshape(B, S, nh, 64).transpose(1, 2).detach().requires_grad_(True)
    vf = v.float().reshape(B, S, nh, 64).transpose(1, 2).detach().requires_grad_(True)
    sc = qf @ kf.transpose(-1, -2) * 0.125
    if bias is not None:
        sc = sc + bias[:, None, None, :]
    o = (torch.softmax(sc, dim=-1) @ vf).transpose(1, 2).reshape(B, S, nh * 64)
    return o, torch.logsumexp(sc, dim=-1), (qf, kf, vf)


@pytest.mark.parametrize("B,S,nh", [(2, 64, 2), (2, 185, 12), (2, 323, 12), (1, 1464, 12), (3, 1, 1), (2, 70, 16)])
@pytest.mark.parametrize("use_bias", [False, True])
def test_attention_fwd_bwd(L, B, S, nh, use_bias):
    g = gen(3)
    H = nh * 64
    qkv = torch.randn(B, S, 3 * H, generator=g).cuda().bfloat16()
    bias = None
    if use_bias:
        bias = torch.zeros(B, S)
        bias[:, S - S // 3:] = -65504.0  # HF-style additive key-padding mask
        bias = bias.cuda()
    q, k, v = qkv[..., :H], qkv[..., H:2 * H], qkv[..., 2 * H:]
    o = torch.empty(B, S, H, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B, nh, S, device="cuda")
    L.attn_fwd(q, k, v, o, lse, B=B, S=S, nh=nh, ld_qkv=3 * H, ld_o=H, key_bias=bias)
    ref, ref_lse, (qf, kf, vf) = _attn_ref(q, k, v, bias, B, S, nh)
    assert rel_l2(o, ref) < 6e-3  # bf16 P and bf16 output rounding
    assert (lse - ref_lse).abs().max().item() < 2e-3
    do = torch.randn(B, S, H, generator=g).cuda().bfloat16()
    ref.backward(do.float())
    dqkv = torch.full((B, S, 3 * H), float("nan"), device="cuda", dtype=torch.bfloat16)
    delta = torch.empty(B, nh, S, device="cuda")
    # rank-1 dV term of the post-softmax mask quirk
    rowscale = torch.randn(B, S, generator=g).cuda()
    rank1 = torch.randn(B, H, generator=g).cuda()
    db = [torch.full((H,), 0.25, device="cuda") for _ in range(3)]   # bias-gradient accumulators (+= semantics)
    L.attn_bwd(q, k, v, o, do, lse, delta, dqkv[..., :H], dqkv[..., H:2 * H], dqkv[..., 2 * H:], B=B, S=S, nh=nh,
               ld_qkv=3 * H, ld_o=H, ld_dqkv=3 * H, key_bias=bias, dv_rowscale=rowscale, dv_rank1=rank1,
               dbq=db[0], dbk=db[1], dbv=db[2])
    for i in range(3):   # column sums of dq / dk / dv over all B*S rows (fp32 sums of the values before bf16 rounding)
        want = dqkv[..., i * H:(i + 1) * H].float().sum(dim=(0, 1))
        assert (db[i] - 0.25 - want).abs().max().item() < 2e-2 * max(1.0, want.abs().max().item())
    unpack = lambda t: t.transpose(1, 2).reshape(B, S, H)  # noqa: E731
    if S == 1:  # softmax over one key is constant: dQ = dK = 0 exactly; only rounding noise may remain
        assert dqkv[..., :2 * H].float().abs().max().item() < 1e-6
    else:
        assert rel_l2(dqkv[..., :H], unpack(qf.grad)) < 1.5e-2
        assert rel_l2(dqkv[..., H:2 * H], unpack(kf.grad)) < 1.5e-2
    assert rel_l2(dqkv[..., 2 * H:], unpack(vf.grad) + rowscale[:, :, None] * rank1[:, None, :]) < 1.5e-2


@pytest.mark.parametrize("B,S", [(16, 1464), (16, 323)])
def test_attention_full_size_deterministic(L, B, S):
    """BASELINE configs[1] shapes (VideoMAE S=1464, fusion S=323; 192 (b,h) pairs -> every SM holds two CTAs):
    the tensor-core attention has no atomics, so repeated launches must agree bit for bit, forward and backward."""
    nh, H = 12, 768
    g = gen(11)
    qkv = (torch.randn(B, S, 3 * H, generator=g) * 0.7).cuda().bfloat16()
    do = torch.randn(B, S, H, generator=g).cuda().bfloat16()
    q, k, v = qkv[..., :H], qkv[..., H:2 * H], qkv[..., 2 * H:]
    outs = []
    for _ in range(4):
        o = torch.full((B, S, H), float("nan"), device="cuda", dtype=torch.bfloat16)
        lse = torch.empty(B, nh, S, device="cuda")
        L.attn_fwd(q, k, v, o, lse, B=B, S=S, nh=nh, ld_qkv=3 * H, ld_o=H)
        dqkv = torch.full((B, S, 3 * H), float("nan"), device="cuda", dtype=torch.bfloat16)
        delta = torch.empty(B, nh, S, device="cuda")
        L.attn_bwd(q, k, v, o, do, lse, delta, dqkv[..., :H], dqkv[..., H:2 * H], dqkv[..., 2 * H:], B=B, S=S, nh=nh,
                   ld_qkv=3 * H, ld_o=H, ld_dqkv=3 * H)
        outs.append((o, lse, dqkv))
    for o, lse, dqkv in outs[1:]:
        assert torch.equal(o, outs[0][0]) and torch.equal(lse, outs[0][1]) and torch.equal(dqkv, outs[0][2])
    # one (batch, head) slice against fp32 torch
    b, h = B - 1, nh - 1
    sl = slice(h * 64, (h + 1) * 64)
    ref = torch.softmax((q[b, :, sl].float() @ k[b, :, sl].float().t()) * 0.125, dim=-1) @ v[b, :, sl].float()
    assert rel_l2(outs[0][0][b, :, sl], ref) < 6e-3
    assert torch.isfinite(outs[0][2].float()).all()


# ------------------------------------------------------------------------------------------------ LayerNorm
@pytest.mark.parametrize("M,H,eps", [(646, 768, 1e-12), (5168, 768, 1e-5), (37, 1024, 1e-5), (2, 768, 1e-5)])
@pytest.mark.parametrize("big", [False, True])
def test_layernorm_fwd_bwd(L, M, H, eps, big):
    g = gen(4)
    x = torch.randn(M, H, generator=g)
    if big:  # residual-stream magnitudes produced by the reference's post-softmax mask
        x = x + 3.0e6 * torch.randn(1, H, generator=g)
    x = x.cuda()
    gamma = (1.0 + 0.1 * torch.randn(H, generator=g)).cuda()
    beta = (0.1 * torch.randn(H, generator=g)).cuda()
    yb, yf, mean, rstd = L.layernorm_fwd(x, gamma, beta, eps, want_bf16=True, want_f32=True)
    xd = x.double().requires_grad_(True)
    gd, bd = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xd, (H,), gd, bd, eps)
    tol = 2e-5 if not big else 2e-4
    assert (yf.double() - ref).abs().max().item() < tol
    assert rel_l2(yb, ref) < 4e-3
    dy = torch.randn(M, H, generator=g).cuda()
    resid = torch.randn(M, H, generator=g).cuda()
    ref.backward(dy.double())
    dgamma = torch.zeros(H, device="cuda")
    dbeta = torch.zeros(H, device="cuda")
    dcs = torch.full((H,), 1.5, device="cuda")
    dxf, dxb = L.layernorm_bwd(dy, x, mean, rstd, gamma, dgamma, dbeta, resid=resid, want_f32=True, want_bf16=True,
                               dx_colsum=dcs)
    ref_dx = xd.grad + resid.double()
    assert rel_l2(dcs, 1.5 + ref_dx.sum(dim=0)) < (1e-5 if not big else 1e-3)   # fused bias-gradient column sums
    assert rel_l2(dxf, ref_dx) < (1e-5 if not big else 1e-3)
    assert rel_l2(dxb, ref_dx) < 5e-3
    assert rel_l2(dgamma, gd.grad) < (1e-5 if not big else 1e-3)
    assert rel_l2(dbeta, bd.grad) < 1e-5


# ------------------------------------------------------------------------------------------------ pointwise
def test_embed_add_and_pool(L):
    g = gen(5)
    B, S, H = 4, 323, 768
    x = torch.randn(B, S, H, generator=g).cuda()
    idx = torch.cat([torch.zeros(B, 70), torch.ones(B, 149), 2 * torch.ones(B, 104)], dim=1).long().cuda()
    table = torch.randn(3, H, generator=g).cuda()
    y = torch.empty_like(x)
    L.call("tavk_embed_add_fwd", x.data_ptr(), idx.data_ptr(), table.data_ptr(), y.data_ptr(), B * S, H, 3)
    assert torch.equal(y, x + table[idx])
    dy = torch.randn(B, S, H, generator=g).cuda()
    dt = torch.zeros(3, H, device="cuda")
    L.call("tavk_embed_add_bwd", dy.data_ptr(), idx.data_ptr(), dt.data_ptr(), B * S, H, 3)
    ref = torch.zeros(3, H, device="cuda", dtype=torch.float64).index_add_(0, idx.view(-1), dy.view(-1, H).double())
    assert rel_l2(dt, ref) < 1e-5
    p = torch.empty(B, H, device="cuda")
    L.call("tavk_mean_pool_fwd", x.data_ptr(), p.data_ptr(), B, S, H)
    assert (p - x.mean(dim=1)).abs().max().item() < 1e-5
    dp = torch.randn(B, H, generator=g).cuda()
    dx = torch.empty_like(x)
    dxb = torch.empty(B, S, H, device="cuda", dtype=torch.bfloat16)
    L.call("tavk_mean_pool_bwd", dp.data_ptr(), dx.data_ptr(), dxb.data_ptr(), B, S, H)
    ref = (dp / S)[:, None, :].expand(B, S, H)
    assert (dx - ref).abs().max().item() < 1e-7
    assert rel_l2(dxb, ref) < 4e-3


def test_colsums_and_small_linear(L):
    g = gen(6)
    B, S, N = 3, 323, 768
    xf = torch.randn(B * S, N, generator=g).cuda()
    xb = xf.bfloat16()
    out = torch.empty(N, device="cuda")
    L.colsum(xf, out, M=B * S, N=N)
    assert rel_l2(out, xf.double().sum(0)) < 1e-5
    L.colsum(xb, out, M=B * S, N=N, accumulate=True)
    assert rel_l2(out, xf.double().sum(0) + xb.double().sum(0)) < 1e-5
    w = torch.randn(B, S, generator=g).cuda() * 65505.0
    mo = torch.empty(B, N, device="cuda")
    L.masked_colsum(xb, w, mo, B=B, S=S, N=N, ld=N)
    ref = torch.einsum("bs,bsn->bn", w.double(), xb.double().view(B, S, N))
    assert rel_l2(mo, ref) < 1e-5
    # strided view: V slice of a packed QKV buffer
    qkv = torch.randn(B * S, 3 * N, generator=g).cuda().bfloat16()
    L.masked_colsum(qkv[:, 2 * N:], w, mo, B=B, S=S, N=N, ld=3 * N)
    ref = torch.einsum("bs,bsn->bn", w.double(), qkv[:, 2 * N:].double().reshape(B, S, N))
    assert rel_l2(mo, ref) < 1e-5
    # small linears (classifier head shape and the rank-1 GEMV shape)
    for (M, Nn, K) in [(16, 7, 3072), (16, 768, 768), (2, 2, 3072)]:
        x = torch.randn(M, K, generator=g).cuda()
        wt = torch.randn(Nn, K, generator=g).cuda() * 0.05
        b = torch.randn(Nn, generator=g).cuda()
        y = torch.empty(M, Nn, device="cuda")
        L.call("tavk_small_linear_fwd", x.data_ptr(), wt.data_ptr(), b.data_ptr(), y.data_ptr(), M, Nn, K)
        assert rel_l2(y, x.double() @ wt.double().t() + b.double()) < 1e-5
        dy = torch.randn(M, Nn, generator=g).cuda()
        dx = torch.empty(M, K, device="cuda")
        L.call("tavk_small_linear_bwd_x", dy.data_ptr(), wt.data_ptr(), dx.data_ptr(), M, Nn, K, 0)
        assert rel_l2(dx, dy.double() @ wt.double()) < 1e-5
        dw = torch.zeros(Nn, K, device="cuda")
        db = torch.zeros(Nn, device="cuda")
        L.call("tavk_small_linear_bwd_w", dy.data_ptr(), x.data_ptr(), dw.data_ptr(), db.data_ptr(), M, Nn, K)
        assert rel_l2(dw, dy.double().t() @ x.double()) < 1e-5
        assert rel_l2(db, dy.double().sum(0)) < 1e-5


def test_cast_scale_dropout_permute(L):
    g = gen(7)
    x = torch.randn(100003, generator=g).cuda()
    assert torch.equal(L.cast_bf16(x), x.bfloat16())
    y = torch.empty_like(x)
    L.call("tavk_scale_f32", x.data_ptr(), y.data_ptr(), 0.25, x.numel())
    assert torch.equal(y, x * 0.25)
    keep = torch.empty(x.numel(), device="cuda", dtype=torch.uint8)
    L.call("tavk_dropout", x.data_ptr(), y.data_ptr(), keep.data_ptr(), x.numel(), 0.4, 1234, 0, None)
    frac = keep.float().mean().item()
    assert abs(frac - 0.6) < 0.01
    assert torch.allclose(y, torch.where(keep.bool(), x / 0.6, torch.zeros_like(x)), rtol=1e-6, atol=0)
    keep2 = torch.empty_like(keep)
    L.call("tavk_dropout", x.data_ptr(), y.data_ptr(), keep2.data_ptr(), x.numel(), 0.4, 1234, 0, None)
    assert torch.equal(keep, keep2)  # counter-based: reproducible for (seed, offset)
    ctr = torch.ones(1, dtype=torch.int64, device="cuda")
    L.call("tavk_dropout", x.data_ptr(), y.data_ptr(), keep2.data_ptr(), x.numel(), 0.4, 1234, 0, ctr.data_ptr())
    assert not torch.equal(keep, keep2)  # device-side step counter changes the stream (CUDA-graph replays)
    dx = torch.empty_like(x)
    L.call("tavk_dropout_bwd", y.data_ptr(), keep.data_ptr(), dx.data_ptr(), x.numel(), 0.4)
    assert torch.allclose(dx, torch.where(keep.bool(), y / 0.6, torch.zeros_like(y)), rtol=1e-6, atol=0)
    B, S, nh, d = 2, 185, 12, 64
    t = torch.randn(B, S, nh, d, generator=g).cuda().bfloat16()
    o = torch.empty(B, nh, d, S, device="cuda", dtype=torch.bfloat16)
    L.call("tavk_permute_bshd_bhds", t.data_ptr(), o.data_ptr(), B, S, nh, d, 0)
    assert torch.equal(o, t.permute(0, 2, 3, 1).contiguous())
    back = torch.empty_like(t)
    L.call("tavk_permute_bshd_bhds", o.data_ptr(), back.data_ptr(), B, S, nh, d, 1)
    assert torch.equal(back, t)


# ------------------------------------------------------------------------------------------------ loss / optimiser
KAT_LOGITS = None  # the notebook KAT lives in tests/test_oracle_cpu.py (oracle side); here: kernel vs torch


@pytest.mark.parametrize("B,C,weighted", [(16, 7, False), (16, 7, True), (32, 2, True), (1, 7, True)])
def test_softmax_ce(L, B, C, weighted):
    g = gen(8)
    logits = (torch.randn(B, C, generator=g) * 3).cuda()
    target = torch.randint(0, C, (B,), generator=g).cuda()
    w = (torch.rand(C, generator=g) + 0.5).cuda() if weighted else None
    probs = torch.empty(B, C, device="cuda")
    num = torch.empty(1, device="cuda")
    den = torch.empty(1, device="cuda")
    L.call("tavk_softmax_ce_fwd", logits.data_ptr(), target.data_ptr(), L._ptr(w), probs.data_ptr(), num.data_ptr(),
           den.data_ptr(), B, C)
    ld = logits.double().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(ld, target, weight=None if w is None else w.double())
    assert abs((num / den).item() - ref.item()) < 1e-5 * max(1.0, abs(ref.item()))
    ref.backward()
    gscale = (1.0 / den).contiguous()
    dl = torch.empty(B, C, device="cuda")
    L.call("tavk_softmax_ce_bwd", probs.data_ptr(), target.data_ptr(), L._ptr(w), gscale.data_ptr(), dl.data_ptr(), B, C)
    assert rel_l2(dl, ld.grad) < 1e-5


def test_adamw_matches_torch(L):
    g = gen(9)
    n = 1_000_003
    p0 = torch.randn(n, generator=g).cuda()
    grads = [torch.randn(n, generator=g).cuda() * s for s in (1.0, 0.1, 5.0)]
    lr, wd, clip = 1e-3, 1e-2, 1.0
    ref_p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref_p], lr=lr, weight_decay=wd)
    pad = (-n) % 4
    p = torch.cat([p0, torch.zeros(pad, device="cuda")])
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    shadow = torch.empty(p.numel(), device="cuda", dtype=torch.bfloat16)
    sq = torch.zeros(1, device="cuda")
    for step, gr in enumerate(grads, start=1):
        ref_p.grad = gr.clone()
        torch.nn.utils.clip_grad_norm_([ref_p], clip)
        opt.step()
        gbuf = torch.cat([gr, torch.zeros(pad, device="cuda")])
        sq.zero_()
        L.call("tavk_grad_sqnorm", gbuf.data_ptr(), gbuf.numel(), sq.data_ptr())
        assert abs(sq.item() - gr.double().pow(2).sum().item()) / gr.double().pow(2).sum().item() < 1e-5
        L.call("tavk_adamw", p.data_ptr(), m.data_ptr(), v.data_ptr(), gbuf.data_ptr(), shadow.data_ptr(), p.numel(),
               lr, 0.9, 0.999, 1e-8, wd, step, sq.data_ptr(), clip, 1.0, 1)
        assert gbuf.abs().max().item() == 0.0  # zero_grad fused
    assert (p[:n] - ref_p.data).abs().max().item() < 2e-6
    assert torch.equal(shadow[:n], p[:n].bfloat16())


def test_masked_mean_pool_lengths(L):
    """optional lengths on the mean-pool entry points (NULL = the reference's unmasked mean)."""
    from multi_modal_emotion_b200 import engine

    g = gen(21)
    B, S, H = 5, 77, 768
    x = torch.randn(B, S, H, generator=g).cuda().requires_grad_(True)
    lengths = torch.tensor([77, 1, 40, 0, 200])
    y = engine.mean_pool(x, lengths)
    w = torch.randn(B, H, generator=g).cuda()
    (y * w).sum().backward()
    xr = x.detach().clone().requires_grad_(True)
    rows = []
    for b in range(B):
        n = int(min(max(int(lengths[b]), 0), S))
        rows.append(xr[b, :n].mean(dim=0) if n > 0 else xr[b].sum(dim=0) * 0.0)
    yr = torch.stack(rows)
    (yr * w).sum().backward()
    assert (y - yr).abs().max().item() < 1e-5
    assert (x.grad - xr.grad).abs().max().item() < 1e-6
    # no lengths: the reference's plain mean
    assert (engine.mean_pool(x.detach()) - x.detach().mean(dim=1)).abs().max().item() < 1e-5


@pytest.mark.parametrize("cta_pair", [1, 2])
@pytest.mark.parametrize("M,N,K", [(300, 136, 200), (129, 72, 64), (5168, 768, 768)])
def test_gemm_tma_store_epilogue_edges(L, M, N, K, cta_pair, monkeypatch):
    """bf16 outputs leave through TMA stores (csrc/gemm_tcgen05.cu epilogue_loop_tma): ragged M / N edges are clipped by
    the TMA unit, bias arrives as broadcast loads, the GELU' multiplier as TMA-loaded boxes; checked against torch and
    against the coalesced-store epilogue (TAVK_GEMM_TMA_EPI=0 is read once per process, so the reference here is torch).
    cta_pair = 2: the same through CTA pairs, whose lower CTA may own rows past M."""
    _gemm = L.gemm
    monkeypatch.setattr(L, "gemm", lambda *a, **k: _gemm(*a, cta_pair=cta_pair, **k))
    g = gen(31)
    A = (torch.randn(M, K, generator=g) * 0.5).cuda().bfloat16()
    W = (torch.randn(N, K, generator=g) * 0.1).cuda().bfloat16()
    bias = torch.randn(N, generator=g).cuda()
    acc = A.float() @ W.float().t()
    guard = 7.0     # rows / columns past the edges must stay untouched
    out = torch.full((M + 3, N + 8), guard, device="cuda", dtype=torch.bfloat16)
    L.gemm(A, W, out, M=M, N=N, K=K, bias=bias, alpha=0.5, ldo=N + 8)
    assert rel_l2(out[:M, :N], 0.5 * acc + bias) < 4e-3
    assert (out[M:] == guard).all() and (out[:, N:] == guard).all()
    dg = torch.full((M + 3, N + 8), guard, device="cuda", dtype=torch.bfloat16)
    act = torch.full((M + 3, N + 8), guard, device="cuda", dtype=torch.bfloat16)
    L.gemm(A, W, dg, M=M, N=N, K=K, bias=bias, out2=act, epilogue=L.EPI_GELU_GRAD, ldo=N + 8)
    xe = (acc + bias).requires_grad_(True)
    torch.nn.functional.gelu(xe).backward(torch.ones_like(xe))
    assert rel_l2(act[:M, :N], torch.nn.functional.gelu(acc + bias)) < 4e-3
    assert rel_l2(dg[:M, :N], xe.grad) < 4e-3
    assert (dg[M:] == guard).all() and (act[:, N:] == guard).all()
    aux = dg[:M, :N].contiguous()
    outm = torch.full((M + 3, N), guard, device="cuda", dtype=torch.bfloat16)
    cs = torch.full((N,), 3.0, device="cuda")
    L.gemm(A, W, outm, M=M, N=N, K=K, aux=aux, epilogue=L.EPI_MUL, colsum=cs)
    assert rel_l2(outm[:M], acc * aux.float()) < 4e-3
    assert rel_l2(cs, 3.0 + (acc * aux.float()).sum(dim=0)) < 3e-3
    assert (outm[M:] == guard).all()
    pre = (acc + bias).bfloat16()
    outg = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
    L.gemm(A, W, outg, M=M, N=N, K=K, aux=pre, epilogue=L.EPI_GELU_BWD)
    xp = pre.float().requires_grad_(True)
    torch.nn.functional.gelu(xp).backward(torch.ones_like(xp))
    assert rel_l2(outg, acc * xp.grad) < 4e-3
