"""torch.ops.tavk.* on a B200: forward and backward of every operator against plain PyTorch (fp32 ops: 1e-4 / 1e-3;
attention: the bf16 tolerances of tests/test_kernels_gpu.py), plus torch.library.opcheck (schema, fake tensors, autograd
registration)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def test_fp32_operators_match_torch_and_pass_opcheck():
    from multi_modal_emotion_b200 import _lib, ops

    _lib.require_device()
    g = torch.Generator().manual_seed(0)
    x = torch.randn(3, 37, 768, generator=g).cuda().requires_grad_(True)
    w = (1 + 0.1 * torch.randn(768, generator=g)).cuda().requires_grad_(True)
    b = (0.1 * torch.randn(768, generator=g)).cuda().requires_grad_(True)
    lw = (0.05 * torch.randn(7, 768, generator=g)).cuda().requires_grad_(True)
    lb = (0.1 * torch.randn(7, generator=g)).cuda().requires_grad_(True)
    cw = torch.tensor([0.5, 0.9, 1.0, 0.9, 0.8, 1.0, 0.9]).cuda()
    target = torch.tensor([0, 3, 6]).cuda()

    def run(ln, pool, lin, ce):
        for t in (x, w, b, lw, lb):
            t.grad = None
        loss = ce(lin(pool(ln(x, w, b, 1e-5)), lw, lb), target, cw)
        loss.backward()
        return loss.detach(), [t.grad.clone() for t in (x, w, b, lw, lb)]

    ours = run(ops.layer_norm, ops.mean_pool, ops.small_linear, ops.cross_entropy)
    ref = run(lambda x_, w_, b_, e: F.layer_norm(x_, (768,), w_, b_, e), lambda t: t.mean(dim=1), F.linear,
              lambda lg, tg, cw_: F.cross_entropy(lg, tg, weight=cw_))
    assert abs(ours[0].item() - ref[0].item()) < 1e-4
    for a, r in zip(ours[1], ref[1]):
        assert rel(a, r) < 1e-3
    checks = ("test_schema", "test_faketensor", "test_autograd_registration")
    xd = x.detach()
    torch.library.opcheck(torch.ops.tavk.layer_norm_fwd, (xd.clone().requires_grad_(True), w.detach().clone().requires_grad_(True),
                                                          b.detach().clone().requires_grad_(True), 1e-5), test_utils=checks)
    torch.library.opcheck(torch.ops.tavk.mean_pool, (xd.clone().requires_grad_(True),), test_utils=checks)
    torch.library.opcheck(torch.ops.tavk.small_linear, (xd[:, 0].clone().requires_grad_(True), lw.detach().clone().requires_grad_(True),
                                                        lb.detach().clone().requires_grad_(True)), test_utils=checks)
    torch.library.opcheck(torch.ops.tavk.softmax_ce, (torch.randn(3, 7, device="cuda", requires_grad=True), target, cw),
                          test_utils=checks)


@pytest.mark.parametrize("use_bias", [False, True])
def test_attention_operator_matches_torch(use_bias):
    from multi_modal_emotion_b200 import _lib

    _lib.require_device()
    from multi_modal_emotion_b200 import ops  # noqa: F401

    B, S, nh = 2, 323, 12
    H = nh * 64
    g = torch.Generator().manual_seed(1)
    qkv = torch.randn(B, S, 3 * H, generator=g).cuda().bfloat16().requires_grad_(True)
    bias = None
    if use_bias:
        bias = torch.zeros(B, S)
        bias[:, S - 40:] = -65504.0
        bias = bias.cuda()
    o, lse = torch.ops.tavk.attention(qkv, nh, bias)
    do = torch.randn(B, S, H, generator=g).cuda().bfloat16()
    o.backward(do)
    q, k, v = (t.float().reshape(B, S, nh, 64).transpose(1, 2).detach().requires_grad_(True) for t in qkv.detach().split(H, dim=-1))
    sc = q @ k.transpose(-1, -2) * 0.125
    if bias is not None:
        sc = sc + bias[:, None, None, :]
    ref = (torch.softmax(sc, dim=-1) @ v).transpose(1, 2).reshape(B, S, H)
    ref.backward(do.float())
    assert rel(o, ref) < 6e-3 and (lse - torch.logsumexp(sc, dim=-1)).abs().max().item() < 2e-3
    want = torch.cat([t.grad.transpose(1, 2).reshape(B, S, H) for t in (q, k, v)], dim=-1)
    assert rel(qkv.grad, want) < 1.5e-2
    torch.library.opcheck(torch.ops.tavk.attention, (qkv.detach().clone().requires_grad_(True), nh, bias),
                          test_utils=("test_schema", "test_faketensor", "test_autograd_registration"))


def test_linear_embed_add_dropout_adamw_operators():
    """tavk::linear (tcgen05 GEMM forward / dgrad / wgrad), embed_add, dropout and the fused optimiser step as torch custom
    ops: against torch (bf16-operand tolerance for linear, fp32 for the rest) and through torch.library.opcheck."""
    from multi_modal_emotion_b200 import _lib, ops  # noqa: F401

    _lib.require_device()
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 149, 1024, generator=g).cuda().requires_grad_(True)
    w = (torch.randn(768, 1024, generator=g) * 0.03).cuda().requires_grad_(True)
    b = (torch.randn(768, generator=g) * 0.1).cuda().requires_grad_(True)
    table = torch.randn(3, 768, generator=g).cuda().requires_grad_(True)
    idx = torch.randint(0, 3, (2, 149), generator=g).cuda()
    probe = torch.randn(2, 149, 768, generator=g).cuda()

    def run(lin, emb):
        for t in (x, w, b, table):
            t.grad = None
        y = emb(lin(x, w, b), idx, table)
        (y * probe).sum().backward()
        return y.detach(), [t.grad.clone() for t in (x, w, b, table)]

    ours = run(torch.ops.tavk.linear, torch.ops.tavk.embed_add)
    ref = run(F.linear, lambda h, i, t: h + t[i])
    assert rel(ours[0], ref[0]) < 6e-3
    for a, r, tol in zip(ours[1], ref[1], (1e-2, 1e-2, 1e-4, 1e-4)):
        assert rel(a, r) < tol
    counter = torch.zeros(1, dtype=torch.int64, device="cuda")
    xd = torch.randn(64, 768, generator=g).cuda().requires_grad_(True)
    y, keep = torch.ops.tavk.dropout(xd, 0.4, 1234, counter)
    assert abs(keep.float().mean().item() - 0.6) < 2e-2
    assert torch.equal(y, torch.where(keep.bool(), xd.detach() / 0.6, torch.zeros_like(y)))
    y.sum().backward()
    assert torch.equal(xd.grad, keep.float() / 0.6)
    counter.add_(1)
    assert not torch.equal(torch.ops.tavk.dropout(xd, 0.4, 1234, counter)[1], keep)
    # fused optimiser step vs torch.optim.AdamW + clip_grad_norm_ (three steps: the device clock advances inside the op)
    n = 4096 + 64
    p0 = torch.randn(n, generator=g).cuda()
    refp = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([refp], lr=3e-3, weight_decay=1e-2)
    p, m, v = p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    pb = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    step = torch.zeros(1, dtype=torch.int32, device="cuda")
    hyper = torch.tensor([3e-3, 0, 0, 0], device="cuda")
    sq = torch.zeros(1, device="cuda")
    for i in range(3):
        gr = torch.randn(n, generator=g).cuda() * 3
        refp.grad = gr.clone()
        torch.nn.utils.clip_grad_norm_([refp], 1.0)
        opt.step()
        gg = gr.clone()
        torch.ops.tavk.adamw_step(p, m, v, gg, pb, step, hyper, sq, 0.9, 0.999, 1e-8, 1e-2, 1.0, 1.0)
        assert gg.abs().max().item() == 0.0                       # the update zeroes the gradient
    assert int(step.item()) == 3
    assert rel(p, refp.detach()) < 1e-6 and torch.equal(pb.float(), p.bfloat16().float())
    checks = ("test_schema", "test_faketensor", "test_autograd_registration")
    torch.library.opcheck(torch.ops.tavk.linear, (x.detach().clone().requires_grad_(True), w.detach().clone().requires_grad_(True),
                                                  b.detach().clone().requires_grad_(True)), test_utils=checks)
    torch.library.opcheck(torch.ops.tavk.embed_add, (ours[0].clone().requires_grad_(True), idx, table.detach().clone().requires_grad_(True)),
                          test_utils=checks)
    torch.library.opcheck(torch.ops.tavk.dropout, (xd.detach().clone().requires_grad_(True), 0.4, 1234, counter), test_utils=checks)
    torch.library.opcheck(torch.ops.tavk.adamw_step, (p.clone(), m.clone(), v.clone(), torch.randn(n, device="cuda"), pb.clone(),
                                                      step.clone(), hyper.clone(), sq.clone(), 0.9, 0.999, 1e-8, 1e-2, 1.0, 1.0),
                          test_utils=("test_schema", "test_faketensor"))
