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
