"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.pt by importing and running the UNMODIFIED reference modules
from /root/reference (authoring container only) on deterministic synthetic weights/inputs, and prints how far the
restatement in oracle/tav_oracle.py is from them.  Re-run:  python -m oracle.make_golden [fusion custom tav ce]

Fixtures are kept small: outputs are stored sub-sampled plus norms; weights are never stored (they are regenerated
from per-key seeds by multi_modal_emotion_b200.synthetic.synth_state_dict on both sides)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from multi_modal_emotion_b200 import synthetic as syn  # noqa: E402
from oracle import ref_loader, tav_oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def versions():
    import transformers

    return {"torch": str(torch.__version__), "transformers": str(transformers.__version__)}


def sub(t):
    """Deterministic sub-sample of a [B,S,H] tensor that keeps fixtures small."""
    return t[:, ::8, ::16].contiguous()


def fusion_inputs(B=2, T=32, Ta=49, K=104, seed=5):
    g = torch.Generator().manual_seed(seed)
    S = T + Ta + K
    x = torch.randn(B, S, 768, generator=g)
    probe = torch.randn(B, S, 768, generator=g) / S  # loss = <y, probe>
    text_len = torch.tensor([32, 20][:B])
    frames = torch.tensor([49, 37][:B])
    masks = {
        "R": syn.reference_masks(B, T, Ta, K, text_len, frames),          # exactly what PreFormer emits (Q1+Q2)
        "zero": torch.zeros(B, 1, 1, S),
        "none": None,
    }
    return x, probe, masks


def gold_fusion():
    ns = ref_loader.load_reference("tiny")
    from transformers import VideoMAEConfig

    out = {"versions": versions(), "cases": {}}
    enc = ns.VideoMAEEncoder(VideoMAEConfig(), 2).eval()
    sd = syn.synth_state_dict(enc, seed=11)
    enc.load_state_dict(sd)
    x, probe, masks = fusion_inputs()
    for name, mask in masks.items():
        case = {}
        for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            e = enc.to(dt)
            e.zero_grad(set_to_none=True)
            xi = x.detach().to(dt).clone().requires_grad_(True)
            y = e(xi, None if mask is None else mask.to(dt))
            (y * probe.to(dt)).sum().backward()
            case["y_" + tag] = sub(y.detach()).float() if dt == torch.float32 else sub(y.detach())
            case["y_norm_" + tag] = y.detach().norm().item()
            case["dx_" + tag] = sub(xi.grad)
            case["dx_norm_" + tag] = xi.grad.norm().item()
            case["grad_norms_" + tag] = {k: p.grad.norm().item() for k, p in e.named_parameters()}
            case["grad_slices_" + tag] = {k: p.grad.flatten()[:: max(1, p.numel() // 64)][:64].clone()
                                          for k, p in e.named_parameters() if k.startswith("layer.0.")}
            # restatement check
            xo = x.detach().to(dt).clone().requires_grad_(True)
            yo = O.fusion_encoder(xo, None if mask is None else mask.to(dt), {k: v.to(dt) for k, v in sd.items()})
            print("fusion[%s,%s] |ref-oracle| max %.3e  (|y| max %.3e)" % (name, tag, (yo - y).abs().max().item(),
                                                                        y.abs().max().item()))
        out["cases"][name] = case
    torch.save(out, os.path.join(GOLD, "fusion_encoder.pt"))


def gold_custom():
    ns = ref_loader.load_reference("tiny")
    out = {"versions": versions(), "cases": {}}
    x, probe, _ = fusion_inputs()
    B, S = x.shape[:2]
    key_pad = torch.zeros(B, 1, 1, S)
    key_pad[1, :, :, 150:] = -65504.0
    for early in (False, True):
        enc = ns.TransformerEncoder(768, num_layers=2, dropout=0.0, early_div=early).eval()
        sd = syn.synth_state_dict(enc, seed=12)
        enc.load_state_dict(sd)
        for mname, mask in (("pad", key_pad), ("none", None)):
            enc.zero_grad(set_to_none=True)
            xi = x.clone().requires_grad_(True)
            y = enc(xi, mask)
            (y * probe).sum().backward()
            case = {"y_f32": sub(y.detach()), "y_norm_f32": y.detach().norm().item(), "dx_f32": sub(xi.grad),
                    "dx_norm_f32": xi.grad.norm().item(),
                    "grad_norms_f32": {k: p.grad.norm().item() for k, p in enc.named_parameters()}}
            yo = O.custom_encoder(x, mask, sd, 2, early_div=early)
            print("custom[early=%s,%s] |ref-oracle| max %.3e" % (early, mname, (yo - y).abs().max().item()))
            out["cases"]["early%d_%s" % (early, mname)] = case
    torch.save(out, os.path.join(GOLD, "custom_encoder.pt"))


def gold_ce():
    ns = ref_loader.load_reference("tiny")
    g = torch.Generator().manual_seed(21)
    logits = torch.randn(16, 7, generator=g) * 2
    target = torch.randint(0, 7, (16,), generator=g)
    w = torch.tensor(syn.MELD_CLASS_WEIGHTS)
    crit = ns.NewCrossEntropyLoss(class_weights=w, epoch_switch=2)
    out = {"versions": versions(), "logits": logits, "target": target, "weights": w, "epoch_switch": 2, "loss": {}, "dlogits": {}}
    for epoch in range(4):
        lg = logits.clone().requires_grad_(True)
        loss = crit(lg, target, epoch=epoch)
        loss.backward()
        out["loss"][epoch] = loss.item()
        out["dlogits"][epoch] = lg.grad.clone()
        print("CE epoch %d: ref %.6f oracle %.6f" % (epoch, loss.item(), O.new_cross_entropy(logits, target, epoch, w, 2).item()))
    torch.save(out, os.path.join(GOLD, "new_ce.pt"))


GRAD_KEYS = [
    "TAVForMAE/linear1.weight", "TAVForMAE/random_mae_encoder.layer.11.attention.attention.query.weight",
    "TAVForMAE/random_mae_encoder.layer.0.attention.attention.query.weight",
    "TAVForMAE/random_mae_encoder.layer.0.intermediate.dense.weight",
    "TAVForMAE/random_mae_encoder.layer.0.layernorm_before.weight", "TAVForMAE/wav_2_768_2.weight",
    "TAVForMAE/videomae.encoder.layer.0.attention.attention.query.weight",
    "TAVForMAE/videomae.encoder.layer.0.intermediate.dense.weight", "TAVForMAE/bert.encoder.layer.0.attention.self.query.weight",
    "TAVForMAE/wav2vec2.encoder.layers.0.attention.q_proj.weight", "PreFormer/wav_2_768.weight",
    "PreFormer/videomae.embeddings.patch_embeddings.projection.weight", "PreFormer/bert.embeddings.word_embeddings.weight",
]


def gold_tav(variant="tiny", cfg="C1"):
    ns = ref_loader.load_reference(variant)
    t0 = time.time()
    torch.manual_seed(0)
    model = ns.TAVForMAE({"output_dim": syn.CONFIGS[cfg]["C"], "dropout": 0.4, "learn_PosEmbeddings": True, "num_layers": 12})
    pre = ns.PreFormer()
    pre_sd, tav_sd = syn.synth_state_dict(pre, seed=1), syn.synth_state_dict(model, seed=2)
    pre.load_state_dict(pre_sd)
    model.load_state_dict(tav_sd)
    inputs, labels = syn.make_batch(cfg)
    ids, tm = inputs[0]["input_ids"], inputs[0]["attention_mask"]
    wav, am = inputs[1]["audio_features"], inputs[1]["attention_mask"]
    video, vm = inputs[2]["visual_embeds"], inputs[2]["attention_mask"]
    tav, pos, mask = pre(input_ids=ids, audio_features=wav, video_embeds=video, text_mask=tm, audio_mask=am,
                         visual_mask=vm, device="cpu", train=False)
    logits = model(input_ids=ids, text_attention_mask=tm, audio_features=wav, video_embeds=video, visual_mask=vm,
                   hidden_states=tav, pos_embed=pos, attention_mask=mask, batch_size=len(labels), check="val")
    w = torch.tensor(syn.MELD_CLASS_WEIGHTS) if logits.shape[1] == 7 else torch.tensor([0.5, 0.5])
    crit = ns.NewCrossEntropyLoss(class_weights=w, epoch_switch=2)
    loss = crit(logits, labels.long(), epoch=1)  # weighted branch
    loss.backward()
    print("reference %s/%s fwd+bwd %.1fs logits\n%s loss %.6f" % (variant, cfg, time.time() - t0, logits.detach(), loss.item()))
    grads = {}
    for tag, m in (("TAVForMAE", model), ("PreFormer", pre)):
        for k, p in m.named_parameters():
            if p.grad is not None:
                grads["%s/%s" % (tag, k)] = p.grad
    out = {"versions": versions(), "variant": variant, "cfg": cfg, "logits": logits.detach().clone(), "loss": loss.item(),
           "tav_sub": sub(tav.detach()), "tav_norm": tav.detach().norm().item(), "pos": pos.clone(), "mask": mask.clone(),
           "grad_norms": {k: grads[k].norm().item() for k in grads},
           "grad_slices": {k: grads[k].flatten()[:: max(1, grads[k].numel() // 64)][:64].clone() for k in GRAD_KEYS if k in grads}}
    # restatement check
    orc = O.OracleTAV(ns.configs).load(pre_sd, tav_sd)
    lo = orc.forward(inputs)
    lo_loss = O.new_cross_entropy(lo, labels.long(), 1, w, 2)
    lo_loss.backward()
    og = orc.named_grads()
    print("oracle  logits max |diff| %.3e  loss diff %.3e" % ((lo - logits).abs().max().item(), abs(lo_loss.item() - loss.item())))
    rels = {k: ((og[k] - grads[k]).norm() / grads[k].norm().clamp_min(1e-30)).item() for k in GRAD_KEYS if k in grads and k in og}
    for k, v in rels.items():
        print("   %-80s rel %.3e  |g| %.3e" % (k, v, grads[k].norm().item()))
    worst = max(rels.values())
    print("oracle  worst rel-L2 grad diff over probe keys %.3e; grads present ref %d oracle %d" % (worst, len(grads), len(og)))
    torch.save(out, os.path.join(GOLD, "tav_%s_%s.pt" % (variant, cfg)))


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    which = sys.argv[1:] or ["ce", "fusion", "custom", "tav"]
    if "ce" in which:
        gold_ce()
    if "fusion" in which:
        gold_fusion()
    if "custom" in which:
        gold_custom()
    if "tav" in which:
        gold_tav("tiny", "C1")
    if "tav_baseline" in which:
        gold_tav("baseline", "C1")
