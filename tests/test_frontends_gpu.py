"""GPU parity of the GEMM-based front-ends (multi_modal_emotion_b200.frontends) against the HF / torch modules they
replace: VideoMAE patch embedding (Conv3d) with token drop, and the Wav2Vec2 grouped positional convolution with
weight-norm, forward and backward.  Tolerance: bf16 operands, fp32 accumulation -> 1e-2 relative-L2."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("masked", [True, False])
def test_patch_embedding_matches_hf(masked):
    from transformers import VideoMAEConfig, VideoMAEModel

    from multi_modal_emotion_b200 import frontends, synthetic as syn

    m = VideoMAEModel(VideoMAEConfig(num_hidden_layers=1)).cuda()
    m.load_state_dict({k: v.cuda() for k, v in syn.synth_state_dict(m, seed=4).items()})
    g = torch.Generator().manual_seed(0)
    B, K = 3, 104
    video = torch.randn(B, 16, 3, 224, 224, generator=g).cuda()
    mask = None
    if masked:
        mask = torch.zeros(B, 1568, dtype=torch.bool)
        for b in range(B):
            mask[b, torch.randperm(1568, generator=g)[:1568 - K]] = True   # masked positions are dropped
        mask = mask.cuda()
    ref = m.embeddings(video, mask)
    probe = torch.randn(ref.shape, generator=torch.Generator().manual_seed(1)).cuda()
    (ref * probe).sum().backward()
    gw, gb = m.embeddings.patch_embeddings.projection.weight.grad.clone(), m.embeddings.patch_embeddings.projection.bias.grad.clone()
    m.zero_grad()
    out = frontends.video_embeddings(m.embeddings, video, mask)
    assert out.shape == ref.shape
    assert rel(out, ref) < 1e-2
    (out * probe).sum().backward()
    assert rel(m.embeddings.patch_embeddings.projection.weight.grad, gw) < 1e-2
    assert rel(m.embeddings.patch_embeddings.projection.bias.grad, gb) < 1e-3


@pytest.mark.parametrize("hidden,T", [(768, 149), (1024, 49), (768, 1)])
def test_positional_conv_matches_hf(hidden, T):
    from transformers import Wav2Vec2Config
    from transformers.models.wav2vec2.modeling_wav2vec2 import Wav2Vec2PositionalConvEmbedding

    from multi_modal_emotion_b200 import frontends

    torch.manual_seed(0)
    pc = Wav2Vec2PositionalConvEmbedding(Wav2Vec2Config(hidden_size=hidden)).cuda()
    with torch.no_grad():
        pc.conv.bias.normal_(0, 0.1)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(4, T, hidden, generator=g).cuda()
    probe = torch.randn(4, T, hidden, generator=g).cuda()
    xr = x.clone().requires_grad_(True)
    ref = pc(xr)
    (ref * probe).sum().backward()
    ref_grads = {k: p.grad.clone() for k, p in pc.named_parameters()}
    pc.zero_grad()
    xo = x.clone().requires_grad_(True)
    out = frontends.pos_conv_embed(pc, xo)
    assert rel(out, ref) < 1e-2
    (out * probe).sum().backward()
    assert rel(xo.grad, xr.grad) < 1e-2
    for k, p in pc.named_parameters():
        assert rel(p.grad, ref_grads[k]) < 1.5e-2, k


@pytest.mark.parametrize("B,L", [(2, 16000), (3, 48000), (1, 4000)])
def test_conv_feature_encoder_matches_hf(B, L):
    """wav2vec2-base feature encoder (group-norm family) on the kernel path vs the HF module in fp32: output and every
    parameter gradient.  bf16 activations/operands with fp32 accumulation through 7 layers -> 2e-2 relative-L2."""
    from transformers import Wav2Vec2Config, Wav2Vec2Model

    from multi_modal_emotion_b200 import frontends, synthetic as syn

    torch.manual_seed(0)
    m = Wav2Vec2Model(Wav2Vec2Config(num_hidden_layers=1)).cuda().eval()
    m.load_state_dict({k: v.cuda() for k, v in syn.synth_state_dict(m, seed=9).items()})
    g = torch.Generator().manual_seed(3)
    wav = torch.randn(B, L, generator=g).cuda()
    ref = m.feature_extractor(wav).transpose(1, 2)            # [B, T, C]
    probe = torch.randn(ref.shape, generator=g).cuda()
    (ref * probe).sum().backward()
    ref_grads = {k: p.grad.clone() for k, p in m.feature_extractor.named_parameters()}
    m.zero_grad()
    T, R = frontends._fe_rows(L, m.config.conv_kernel, m.config.conv_stride)
    assert T[-1] == ref.shape[1] and all(r >= t for r, t in zip(R, T))
    assert all(R[i - 1] == m.config.conv_stride[i] * R[i] for i in range(1, len(R)))
    out = frontends.feature_extractor_cl(m, wav)
    assert out.shape == ref.shape and out.dtype == torch.float32
    assert rel(out, ref) < 2e-2
    (out * probe).sum().backward()
    for k, p in m.feature_extractor.named_parameters():
        assert p.grad is not None, k
        assert rel(p.grad, ref_grads[k]) < 3e-2, (k, rel(p.grad, ref_grads[k]))


def test_roberta_embeddings_and_feature_projection_vs_hf_modules():
    """RobertaEmbeddings (gather-sum kernel + LayerNorm kernel, row-scatter backward) and Wav2Vec2FeatureProjection
    (LayerNorm kernel + tcgen05 GEMM) against the HF modules they replace (reference models/tav.py:349,356,476,485).
    The embedding path is fp32 end to end (1e-5); the projection has bf16 GEMM operands (2e-2 relative-L2)."""
    from transformers import RobertaModel, Wav2Vec2Model

    from multi_modal_emotion_b200 import engine, hf_adapters as hf, synthetic as syn, tav

    cfg = tav.encoder_configs("tiny_base")
    bert = RobertaModel(cfg["text"]).eval()
    bert.load_state_dict(syn.synth_state_dict(bert, seed=31))
    bert = bert.cuda()
    g = torch.Generator().manual_seed(5)
    B, T = 4, 70
    ids = torch.randint(3, 50265, (B, T), generator=g)
    lens = torch.tensor([70, 33, 1, 12])
    ids = torch.where(torch.arange(T)[None, :] < lens[:, None], ids, torch.ones_like(ids)).cuda()    # pad id 1
    probe = torch.randn(B, T, 768, generator=g).cuda()
    emb = bert.embeddings
    ref = emb(input_ids=ids)
    (ref * probe).sum().backward()
    want = {k: p.grad.clone() for k, p in emb.named_parameters() if p.grad is not None}
    emb.zero_grad(set_to_none=True)
    ours = engine.roberta_embeddings(emb, ids)
    (ours * probe).sum().backward()
    assert (ours - ref).abs().max().item() < 1e-5
    got = {k: p.grad for k, p in emb.named_parameters() if p.grad is not None}
    assert set(got) == set(want)
    for k in want:
        assert rel(got[k], want[k]) < 1e-5, k
    # feature projection
    w2v = Wav2Vec2Model(cfg["audio"]).eval()
    w2v.load_state_dict(syn.synth_state_dict(w2v, seed=32))
    w2v = w2v.cuda()
    feats = torch.randn(3, 149, 512, generator=g).cuda().requires_grad_(True)
    probe2 = torch.randn(3, 149, 768, generator=g).cuda()
    ref2, _ = w2v.feature_projection(feats)
    (ref2 * probe2).sum().backward()
    want2 = {k: p.grad.clone() for k, p in w2v.feature_projection.named_parameters()}
    dfe = feats.grad.clone()
    w2v.zero_grad(set_to_none=True)
    feats.grad = None
    ours2 = hf.feature_projection(w2v, feats)
    (ours2 * probe2).sum().backward()
    assert rel(ours2, ref2) < 2e-2 and rel(feats.grad, dfe) < 2e-2
    for k, p in w2v.feature_projection.named_parameters():
        assert rel(p.grad, want2[k]) < 2e-2, k


@pytest.mark.parametrize("B,L", [(2, 16000), (3, 48000), (1, 4000)])
def test_conv_feature_encoder_layer_norm_family_matches_hf(B, L):
    """wav2vec2-large family (feat_extract_norm="layer", conv bias — the checkpoints the reference names,
    models/tav.py:257,455) on the kernel path vs the HF module in fp32: output and every parameter gradient
    (conv weights and biases, LayerNorm weights and biases of all 7 layers).  Same tolerance as the group-norm family."""
    from transformers import Wav2Vec2Config, Wav2Vec2Model

    from multi_modal_emotion_b200 import frontends, synthetic as syn

    torch.manual_seed(0)
    cfg = Wav2Vec2Config(num_hidden_layers=1, feat_extract_norm="layer", conv_bias=True, do_stable_layer_norm=True,
                         hidden_size=1024, num_attention_heads=16, intermediate_size=4096)
    m = Wav2Vec2Model(cfg).cuda().eval()
    m.load_state_dict({k: v.cuda() for k, v in syn.synth_state_dict(m, seed=10).items()})
    g = torch.Generator().manual_seed(4)
    wav = torch.randn(B, L, generator=g).cuda()
    ref = m.feature_extractor(wav).transpose(1, 2)            # [B, T, C]
    probe = torch.randn(ref.shape, generator=g).cuda()
    (ref * probe).sum().backward()
    ref_grads = {k: p.grad.clone() for k, p in m.feature_extractor.named_parameters()}
    assert len(ref_grads) == 28
    m.zero_grad()
    out = frontends.feature_extractor_cl(m, wav)
    assert out.shape == ref.shape and out.dtype == torch.float32
    e = rel(out, ref)
    (out * probe).sum().backward()
    worst = ("", 0.0)
    for k, p in m.feature_extractor.named_parameters():
        assert p.grad is not None, k
        ek = rel(p.grad, ref_grads[k])
        if ek > worst[1]:
            worst = (k, ek)
    print("layer-norm conv family B=%d L=%d: out rel-L2 %.2e, worst parameter gradient %.2e (%s)" % (B, L, e, worst[1], worst[0]))
    assert e < 2e-2
    assert worst[1] < 3e-2, worst
