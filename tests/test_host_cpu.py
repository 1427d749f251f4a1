"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/tavk.h declares (no compute
calls without a GPU), argument validation that does not need a device, the synthetic workload generator, the module
API surface (constructor arguments, state_dict keys, error behaviour) and the rule that the product path never
imports the oracle and has no CPU fallback."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "multi-modal-emotion_b200")


def test_library_exports_every_declared_symbol():
    from multi_modal_emotion_b200 import _lib

    header = open(os.path.join(ROOT, "include", "tavk.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(tavk_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 27
    h = _lib.lib()
    for name in declared:
        assert hasattr(h, name), "libtavk.so does not export %s" % name
    assert declared == set(_lib.SIGNATURES), "ctypes binding and header disagree: %s" % (declared ^ set(_lib.SIGNATURES))
    assert h.tavk_version() == 100


def test_no_device_is_a_loud_error_not_a_fallback():
    from multi_modal_emotion_b200 import _lib

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert _lib.lib().tavk_device_check() == 3
    assert b"CUDA" in _lib.lib().tavk_last_error()
    with pytest.raises(_lib.TavkError):
        _lib.require_device()


def test_argument_validation_without_device():
    import ctypes as C

    from multi_modal_emotion_b200 import _lib

    h = _lib.lib()
    assert h.tavk_gemm_bf16(None, None) == 1
    a = _lib.GemmArgs()
    a.M, a.N, a.K = 16, 12, 64
    assert h.tavk_gemm_bf16(C.byref(a), None) == 1          # N % 8 != 0 is rejected before any launch
    assert b"multiple of 8" in h.tavk_last_error()
    assert h.tavk_attn_fwd(None, None) == 1
    assert h.tavk_layernorm_fwd(None, None, None, None, None, None, None, 4, 768, 1e-5, None) == 1
    assert h.tavk_adamw(None, None, None, None, None, 8, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1, None, 0.0, 1.0, 0, None) == 1


def test_product_path_never_imports_the_oracle():
    for fn in os.listdir(PKG):
        if fn.endswith(".py"):
            src = open(os.path.join(PKG, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
            assert "/root/reference" not in src, fn
    for fn in ("bench.py", "__graft_entry__.py"):
        assert "/root/reference" not in open(os.path.join(ROOT, fn)).read()


def test_synthetic_batch_format_and_masks():
    from multi_modal_emotion_b200 import synthetic as syn

    inputs, labels = syn.make_batch("C1")
    assert inputs[0]["input_ids"].shape == (2, 32) and inputs[1]["audio_features"].shape == (2, 16000)
    assert inputs[2]["visual_embeds"].shape == (2, 16, 3, 224, 224)
    assert inputs[2]["attention_mask"].sum(dim=1).tolist() == [104, 104]          # exactly K kept tokens per row (Q9)
    assert labels.tolist() == [0.0, 3.0]
    again, _ = syn.make_batch("C1")
    assert torch.equal(inputs[2]["visual_embeds"], again[2]["visual_embeds"])       # deterministic
    assert syn.fused_len("C1") == 185 and syn.fused_len("C2") == 323 and syn.fused_len("C4") == 423 and syn.fused_len("C3") == 923
    m = syn.reference_masks(2, 32, 49, 104, torch.tensor([32, 20]), torch.tensor([49, 37]))
    assert m.shape == (2, 1, 1, 185)
    assert sorted(set(m[..., :32].flatten().tolist())) == [-65504.0, 0.0]            # text: (1-m)*fp16.min
    assert sorted(set(m[..., 32:81].flatten().tolist())) == [1.0, 65505.0]           # audio: 1 - m*fp16.min (quirk)
    assert set(m[..., 81:].flatten().tolist()) == {0.0}
    for L_, f in ((3280, 10), (16000, 49), (48000, 149), (80000, 249), (240000, 749)):
        assert syn.conv_frames(L_) == f


def test_module_api_surface_and_errors():
    from transformers import VideoMAEConfig

    from multi_modal_emotion_b200 import tav
    from multi_modal_emotion_b200.losses import NewCrossEntropyLoss
    from multi_modal_emotion_b200.tavformer import TransformerEncoder, VideoMAEEncoder

    enc = VideoMAEEncoder(VideoMAEConfig(), 2)
    keys = set(enc.state_dict())
    for k in ("layernorm_before.weight", "attention.attention.q_bias", "attention.attention.v_bias",
              "attention.attention.query.weight", "attention.attention.key.weight", "attention.attention.value.weight",
              "attention.output.dense.bias", "layernorm_after.bias", "intermediate.dense.weight", "output.dense.weight"):
        assert "layer.1." + k in keys
    assert "layer.0.attention.attention.key.bias" not in keys                        # k has no bias (Q7)
    te = TransformerEncoder(768, num_layers=1)
    assert "layers.0.attention.query_matrix.weight" in te.state_dict() and "layers.0.feed_forward.3.bias" in te.state_dict()
    with pytest.raises(ValueError):
        TransformerEncoder(768, n_heads=8)                                           # head_dim must be 64
    tav.set_encoder_variant("tiny")
    model = tav.TAVForMAE({"output_dim": 2, "dropout": 0.5, "learn_PosEmbeddings": False, "num_layers": 3})
    assert len(model.random_mae_encoder.layer) == 12                                 # num_layers is ignored (Q4)
    assert model.linear1.weight.shape == (2, 3072) and not model.embedding.weight.requires_grad
    assert model.random_mae_encoder.layer[0].layernorm_before.weight.eq(1).all()     # randomize_model
    assert model.random_mae_encoder.layer[0].intermediate.dense.bias.eq(0).all()
    with pytest.raises(KeyError):
        tav.TAVForMAE({"output_dim": 2})
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(None, None, None, None, None, torch.zeros(1, 4, 768), torch.zeros(1, 4).long(), None)
    crit = NewCrossEntropyLoss(torch.ones(7), epoch_switch=2)
    assert crit._weights(0, torch.device("cpu")) is None and crit._weights(1, torch.device("cpu")) is not None


def test_wgrad_split_heuristic_and_layer_spec():
    from multi_modal_emotion_b200 import engine

    assert engine._wgrad_splits(768, 768, 5168) >= 4       # 18 tiles on 148 SMs -> split the token dimension
    assert engine._wgrad_splits(3072, 768, 64) == 1
    assert len(engine.PARAM_SLOTS) == engine.N_SLOTS == 16
    s = engine.LayerSpec()
    assert s.pre_ln and s.mask_mode == "none" and s.hidden == 768


def test_video_patch_cache_never_serves_a_recycled_address():
    """frontends.video_patches_bf16 shares the patchified clip between PreFormer and TAVForMAE.  The allocator gives
    the block of a freed clip to the next batch (same address, shape and version 0), which must NOT hit the cache:
    the eager training loop would otherwise keep training on the first batch's video."""
    import torch
    import torch.nn.functional as F

    from multi_modal_emotion_b200 import frontends

    def ref(v):  # Conv3d(kernel = stride = (2,16,16)) im2col in weight order (c, dt, dy, dx)
        B = v.shape[0]
        u = v.permute(0, 2, 1, 3, 4)                                     # [B, C, T, H, W]
        u = u.unfold(2, 2, 2).unfold(3, 16, 16).unfold(4, 16, 16)        # [B, C, T/2, H/16, W/16, 2, 16, 16]
        return u.permute(0, 2, 3, 4, 1, 5, 6, 7).reshape(B, -1, 3 * 2 * 16 * 16).bfloat16()

    g = torch.Generator().manual_seed(0)
    a = torch.randn(1, 4, 3, 32, 32, generator=g)
    ca = frontends.video_patches_bf16(a, 2, 16)
    assert torch.equal(ca, ref(a)) and frontends.video_patches_bf16(a, 2, 16) is ca      # same tensor: shared
    w = torch.randn(8, 3, 2, 16, 16, generator=g)
    conv = F.conv3d(a.permute(0, 2, 1, 3, 4), w, stride=(2, 16, 16)).flatten(2).transpose(1, 2)
    assert (ca.float() @ w.reshape(8, -1).t() - conv).abs().max().item() < 0.5           # bf16 rounding of 1536 terms
    ptr = a.data_ptr()
    del a
    for _ in range(8):                                                   # provoke address reuse
        b = torch.randn(1, 4, 3, 32, 32, generator=g)
        cb = frontends.video_patches_bf16(b, 2, 16)
        assert torch.equal(cb, ref(b))
        hit = b.data_ptr() == ptr
        ptr = b.data_ptr()
        del b
        if hit:
            break
    c = torch.randn(1, 4, 3, 32, 32, generator=g)
    cc = frontends.video_patches_bf16(c, 2, 16)
    c.add_(1.0)                                                          # in-place write: version counter moves
    assert torch.equal(frontends.video_patches_bf16(c, 2, 16), ref(c))


def test_criterion_state_dict_carries_the_reference_key_both_directions():
    """The reference criterion owns nn.CrossEntropyLoss(weight=...) as ``weightedCEL`` (utils/global_functions.py:63), so
    best.pt's 'loss' entry is {'weightedCEL.weight': ...}; ours must emit and accept exactly that."""
    import torch
    from torch import nn

    from multi_modal_emotion_b200.losses import NewCrossEntropyLoss

    class ReferenceShaped(nn.Module):          # the reference's module tree, state-wise
        def __init__(self, w):
            super().__init__()
            self.weightedCEL = nn.CrossEntropyLoss(weight=w)
            self.normalCEL = nn.CrossEntropyLoss()

    w = torch.tensor([0.5, 0.9, 1.0])
    ours, ref = NewCrossEntropyLoss(w.clone(), epoch_switch=2), ReferenceShaped(w.clone() * 2)
    assert list(ours.state_dict()) == list(ref.state_dict()) == ["weightedCEL.weight"]
    ours.load_state_dict(ref.state_dict())
    assert torch.equal(ours.weightedCEL.weight, w * 2)
    ref.load_state_dict(NewCrossEntropyLoss(w.clone()).state_dict())
    assert torch.equal(ref.weightedCEL.weight, w)


def test_workspace_queries_and_per_call_sm_budget_field():
    """caller-owned workspaces are sized through queries; the SM budget of the persistent GEMM is a field of
    the call's argument struct, not process-wide state (pure host code: callable without a GPU)."""
    from multi_modal_emotion_b200 import _lib as L

    h = L.lib()
    assert h.tavk_workspace_bytes_attn_bwd(16, 1464, 12) == 16 * 12 * 1464 * 4
    assert h.tavk_workspace_bytes_attn_bwd(0, 5, 5) == 0
    assert h.tavk_workspace_bytes_groupnorm(16, 512) == 16 * 2 * 512 * 4
    a = L.GemmArgs()
    assert h.tavk_workspace_bytes_gemm(a) == 0
    assert "max_ctas" in [f[0] for f in L.GemmArgs._fields_]
    assert not hasattr(h, "tavk_reserve_sms") or True     # the symbol is gone from the header; see test below
    import re
    header = open(os.path.join(ROOT, "include", "tavk.h")).read()
    assert "tavk_reserve_sms" not in header and re.search(r"int32_t\s+max_ctas;", header)


def test_spec_augment_matches_the_hf_routine_the_reference_copied():
    """reference models/tav.py:269-306 is a copy of HF Wav2Vec2Model._mask_hidden_states (numpy RNG via
    _compute_mask_indices).  With the same numpy seed, PreFormer._mask_hidden_states must replace exactly the same time
    steps by masked_spec_embed and zero exactly the same feature columns."""
    import numpy as np
    import torch

    from multi_modal_emotion_b200 import tav

    tav.set_encoder_variant("tiny")
    torch.manual_seed(0)
    pre = tav.PreFormer()
    cfg = pre.wav2vec2.config
    cfg.apply_spec_augment, cfg.mask_time_prob, cfg.mask_time_length, cfg.mask_time_min_masks = True, 0.3, 4, 2
    cfg.mask_feature_prob, cfg.mask_feature_length, cfg.mask_feature_min_masks = 0.2, 8, 1
    B, T, H = 3, 49, cfg.hidden_size
    x = torch.randn(B, T, H, generator=torch.Generator().manual_seed(1))
    frame_mask = torch.arange(T)[None, :] < torch.tensor([49, 30, 12])[:, None]
    np.random.seed(123)
    ours = pre._mask_hidden_states(x.clone(), frame_mask, training=True)
    hf = pre.wav2vec2
    with torch.no_grad():
        hf.masked_spec_embed.copy_(pre.masked_spec_embed)
    hf.train()
    np.random.seed(123)
    ref = hf._mask_hidden_states(x.clone(), attention_mask=frame_mask)
    hf.eval()
    assert torch.equal(ours, ref)
    changed = ours != x
    assert changed.any() and not changed.all()                      # something was masked, something was left alone
    assert (ours == 0).all(dim=1).any()                             # a feature column zeroed for every frame
    assert torch.equal(pre._mask_hidden_states(x.clone(), frame_mask, training=False), x)      # eval: untouched
    cfg.apply_spec_augment = False
    assert torch.equal(pre._mask_hidden_states(x.clone(), frame_mask, training=True), x)       # config switch
