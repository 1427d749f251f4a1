"""Deterministic synthetic weights and inputs for the TAV path (there is no network: no checkpoints, no datasets).

Weights: ``synth_state_dict`` fills a module's state_dict key by key from a per-key seeded CPU generator, so the
reference modules, the CPU oracle and the CUDA modules can be given bit-identical parameters without relying on
construction order.  Inputs: ``make_batch`` builds the dict-of-tensors batch format the reference's
``collate_batch`` emits (reference models/tav.py:235-246) for the BASELINE.json configurations."""
import math
import zlib

import torch

MELD_CLASS_WEIGHTS = [0.5285, 0.8794, 0.9732, 0.9316, 0.8255, 0.9729, 0.8890]

# name -> (batch, text_len, wav_len, kept_video_tokens, classes)
CONFIGS = {
    "C1": dict(B=2, T=32, L=16000, K=104, C=7),     # BASELINE.json configs[0]: CPU-runnable case
    "C2": dict(B=16, T=70, L=48000, K=104, C=7),    # configs[1]: MELD 7-class, batch 16 (the benched workload)
    "C3": dict(B=8, T=70, L=240000, K=104, C=7),    # configs[2]: IEMOCAP-shape 15 s audio (full TAV, fused S=923)
    "C4": dict(B=32, T=70, L=80000, K=104, C=2),    # configs[3]: MUStARD++ shape
}


def conv_frames(L):
    """Wav2Vec2 feature-extractor output length (reference models/tav.py:308-324): floor((L-400)/320)+1."""
    for k, s in zip((10, 3, 3, 3, 3, 2, 2), (5, 2, 2, 2, 2, 2, 2)):
        L = (L - k) // s + 1
    return L


def _fill(key, t, seed):
    g = torch.Generator(device="cpu").manual_seed((zlib.crc32(key.encode()) + 7919 * seed) & 0x7FFFFFFF)
    low = key.lower()
    if t.dim() >= 2:
        rf = 1
        for d in t.shape[2:]:
            rf *= d
        fan_in, fan_out = t.shape[1] * rf, t.shape[0] * rf
        a = math.sqrt(6.0 / (fan_in + fan_out))
        return (torch.rand(t.shape, generator=g) * 2 - 1) * a
    if low.endswith("weight") and ("norm" in low):
        return 1.0 + 0.05 * torch.randn(t.shape, generator=g)
    return 0.02 * torch.randn(t.shape, generator=g)


def synth_state_dict(module, seed=0, dtype=torch.float32):
    """Deterministic values for every floating-point entry of ``module.state_dict()`` (integer buffers untouched)."""
    out = {}
    for key, t in module.state_dict().items():
        out[key] = _fill(key, t, seed).to(dtype) if t.is_floating_point() else t.clone()
    return out


def make_batch(cfg="C1", seed=1234, B=None):
    """Returns (inputs, labels) in the reference batch format: inputs = [text, audio, video] dicts."""
    c = dict(CONFIGS[cfg]) if isinstance(cfg, str) else dict(cfg)
    if B is not None:
        c["B"] = B
    B_, T, L, K, C = c["B"], c["T"], c["L"], c["K"], c["C"]
    g = torch.Generator(device="cpu").manual_seed(seed)
    lo_t, lo_a = max(4, T * 12 // 70), max(1024, L // 3)
    tl = torch.linspace(T, lo_t, B_).round().long()
    al = torch.linspace(L, lo_a, B_).round().long()
    ids = torch.randint(3, 50265, (B_, T), generator=g)
    tmask = (torch.arange(T)[None, :] < tl[:, None]).long()
    ids = torch.where(tmask.bool(), ids, torch.ones_like(ids))  # pad id 1
    wav = 0.1 * torch.randn(B_, L, generator=g)
    amask = (torch.arange(L)[None, :] < al[:, None]).long()
    wav = wav * amask
    video = torch.randn(B_, 16, 3, 224, 224, generator=g)
    vmask = torch.zeros(B_, 1568, dtype=torch.bool)
    for b in range(B_):  # exactly K kept tokens per row
        vmask[b, torch.randperm(1568, generator=g)[:K]] = True
    labels = torch.tensor([(3 * i) % C for i in range(B_)], dtype=torch.float32)
    inputs = [
        {"input_ids": ids, "attention_mask": tmask},
        {"audio_features": wav, "attention_mask": amask},
        {"visual_embeds": video, "attention_mask": vmask},
    ]
    return inputs, labels


def fused_len(cfg):
    c = CONFIGS[cfg] if isinstance(cfg, str) else cfg
    return c["T"] + conv_frames(c["L"]) + c["K"]


def reference_masks(B, T, Ta, K, text_len, audio_frames, dtype=torch.float32):
    """The additive mask exactly as the reference PreFormer builds it (models/tav.py:383,390,397,409):
    text (1-m)*fp16.min, audio 1 - m*fp16.min (precedence quirk -> +65505 valid / +1 pad), video 0."""
    fmin = torch.finfo(torch.float16).min
    tm = (torch.arange(T)[None, :] < text_len[:, None]).to(dtype)
    am = (torch.arange(Ta)[None, :] < audio_frames[:, None])
    text = (1.0 - tm[:, None, None, :]) * fmin
    audio = 1.0 - am[:, None, None, :] * fmin
    video = torch.zeros((B, 1, 1, K), dtype=dtype)
    return torch.cat((text, audio.to(dtype), video), dim=-1)
