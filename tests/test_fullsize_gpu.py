"""BASELINE.json configs[1] sizes (B=16 MELD: VideoMAE M=23424 rows, fused S=323, 3 s of audio): the kernels against
plain PyTorch on the GPU and size-independent properties, complementing the small-size oracle / golden parity tests."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def L():
    from multi_modal_emotion_b200 import _lib

    _lib.require_device()
    return _lib


def test_gemm_full_size_epilogues_and_wgrad(L):
    g = torch.Generator().manual_seed(0)
    M, H, I = 23424, 768, 3072
    x = (torch.randn(M, H, generator=g) * 0.5).cuda().bfloat16()
    w1 = (torch.randn(I, H, generator=g) * 0.03).cuda().bfloat16()
    b1 = torch.randn(I, generator=g).cuda() * 0.1
    pre = torch.empty(M, I, device="cuda", dtype=torch.bfloat16)
    act = torch.empty_like(pre)
    L.gemm(x, w1, pre, M=M, N=I, K=H, bias=b1, out2=act, epilogue=L.EPI_GELU)
    ref_pre = x.float() @ w1.float().t() + b1
    assert rel(pre, ref_pre) < 4e-3
    assert rel(act, torch.nn.functional.gelu(ref_pre)) < 4e-3
    # dgrad with GELU' and fused column sums; fp32 residual epilogue; split-K wgrad accumulating into an existing buffer
    dy = (torch.randn(M, H, generator=g) * 0.1).cuda().bfloat16()
    w2 = (torch.randn(H, I, generator=g) * 0.03).cuda().bfloat16()
    dpre = torch.empty(M, I, device="cuda", dtype=torch.bfloat16)
    cs = torch.zeros(I, device="cuda")
    L.gemm(dy, w2, dpre, M=M, N=I, K=H, b_mn=True, aux=pre, epilogue=L.EPI_GELU_BWD, colsum=cs)
    xp = pre.float().requires_grad_(True)
    torch.nn.functional.gelu(xp).backward(dy.float() @ w2.float())
    assert rel(dpre, xp.grad) < 5e-3
    assert rel(cs, xp.grad.sum(dim=0)) < 2e-3
    resid = torch.randn(M, H, generator=g).cuda()
    y = torch.empty(M, H, device="cuda")
    L.gemm(act, w2, y, M=M, N=H, K=I, resid=resid)
    assert rel(y, act.float() @ w2.float().t() + resid) < 1e-4
    dw = torch.ones(H, I, device="cuda")
    L.gemm(dy, act, dw, M=H, N=I, K=M, a_mn=True, b_mn=True, accumulate=True, k_splits=2)
    assert rel(dw, 1.0 + dy.float().t() @ act.float()) < 1e-4


def test_layernorm_full_size(L):
    g = torch.Generator().manual_seed(1)
    M, H = 23424, 768
    x = (torch.randn(M, H, generator=g) * 3 + 1).cuda()
    gamma, beta = (1 + 0.1 * torch.randn(H, generator=g)).cuda(), (0.1 * torch.randn(H, generator=g)).cuda()
    yb, yf, mean, rstd = L.layernorm_fwd(x, gamma, beta, 1e-12, want_bf16=True, want_f32=True)
    xr = x.clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr, (H,), gamma, beta, 1e-12)
    assert (yf - ref).abs().max().item() < 5e-5
    dy = torch.randn(M, H, generator=g).cuda()
    ref.backward(dy)
    dg, db, dc = torch.zeros(H, device="cuda"), torch.zeros(H, device="cuda"), torch.zeros(H, device="cuda")
    dx, _ = L.layernorm_bwd(dy, x, mean, rstd, gamma, dg, db, want_f32=True, dx_colsum=dc)
    assert rel(dx, xr.grad) < 1e-5 and rel(dc, xr.grad.sum(dim=0)) < 1e-3
    assert rel(db, dy.sum(dim=0)) < 1e-5


def test_conv_feature_encoder_full_size_vs_hf():
    """B=16, 3 s of 16 kHz audio through the group-norm conv stack: output and parameter gradients vs the HF module in
    fp32 on the GPU; a sample's frames do not depend on its batch neighbours."""
    from transformers import Wav2Vec2Config, Wav2Vec2Model

    from multi_modal_emotion_b200 import frontends, synthetic as syn

    m = Wav2Vec2Model(Wav2Vec2Config(num_hidden_layers=1)).cuda().eval()
    m.load_state_dict({k: v.cuda() for k, v in syn.synth_state_dict(m, seed=9).items()})
    g = torch.Generator().manual_seed(5)
    wav = (0.1 * torch.randn(16, 48000, generator=g)).cuda()
    ref = m.feature_extractor(wav).transpose(1, 2)
    probe = torch.randn(ref.shape, generator=g).cuda()
    (ref * probe).sum().backward()
    ref_grads = {k: p.grad.clone() for k, p in m.feature_extractor.named_parameters()}
    m.zero_grad()
    out = frontends.feature_extractor_cl(m, wav)
    assert out.shape == (16, 149, 512) and rel(out, ref) < 2e-2
    (out * probe).sum().backward()
    for k, p in m.feature_extractor.named_parameters():
        assert rel(p.grad, ref_grads[k]) < 3e-2, k
    with torch.no_grad():
        one = frontends.feature_extractor_cl(m, wav[7:8])
    # not bit-equal: the GroupNorm sums are accumulated with atomics, their last-bit noise flips a few bf16 roundings of
    # the normalised activations, and six more bf16 layers carry that forward (same size as the run-to-run noise)
    assert rel(one, out[7:8].detach()) < 1e-2


def test_full_training_step_b16_properties():
    """One MELD-shaped step at B=16 through the public runner: finite loss and gradients, every trainable parameter
    that the reference trains gets a gradient, and a second step on the same batch lowers the loss."""
    from multi_modal_emotion_b200 import dp, synthetic as syn, tav
    from multi_modal_emotion_b200.losses import NewCrossEntropyLoss
    from multi_modal_emotion_b200.optim import FusedAdamW

    tav.set_encoder_variant("baseline")
    torch.manual_seed(0)
    model = tav.TAVForMAE({"output_dim": 7, "dropout": 0.4, "learn_PosEmbeddings": True, "num_layers": 12}).cuda().train()
    pre = tav.PreFormer().cuda().train()
    crit = NewCrossEntropyLoss(torch.tensor(syn.MELD_CLASS_WEIGHTS), epoch_switch=2)
    params = [p for p in model.parameters() if p.requires_grad] + [p for p in pre.parameters() if p.requires_grad]
    opt = FusedAdamW(params, lr=1e-5, weight_decay=1e-4)      # the benchmark's (and the reference sweeps') step size
    runner = dp.DataParallelTAV(model, pre, crit, opt, clip=1.0)
    inputs, labels = syn.make_batch("C2")
    inputs = [{k: v.cuda() for k, v in d.items()} for d in inputs]
    labels = labels.cuda()
    pre.static_keep_count, model.static_keep_count = 104, 1568 - 104
    l0 = runner.train_step(inputs, labels, 1, "val").item()
    assert torch.isfinite(opt.flat.flat).all() and torch.isfinite(opt.grad_norm()).all()
    trained = {id(p) for p in opt.flat.params}
    for name in ("linear1.weight", "embedding.weight", "wav_2_768_2.weight", "random_mae_encoder.layer.0.intermediate.dense.weight",
                 "videomae.encoder.layer.11.output.dense.weight", "wav2vec2.feature_extractor.conv_layers.0.conv.weight",
                 "bert.encoder.layer.0.attention.self.query.weight"):
        assert id(dict(model.named_parameters())[name]) in trained, name
    ls = [l0] + [runner.train_step(inputs, labels, 1, "val").item() for _ in range(4)]
    assert all(map(lambda v: v == v and abs(v) < 1e4, ls)), ls
    assert min(ls[1:]) < l0, ls
