"""Parity cases the round-1 review asked for on the GPU path:

  * integer / mask outputs of PreFormer (reference models/tav.py:308-342 frame lengths, :381-411 modality ids and additive
    masks) bit-exact against the golden vectors of the UNMODIFIED reference, on ragged text/audio lengths;
  * NewCrossEntropyLoss on the GPU against the reference's own losses / logit gradients for epochs 0-3 (new_ce.pt);
  * five input seeds per configuration with argmax equality on every row and the top-1/top-2 margin table printed;
  * the BENCHMARK configuration itself — full-depth `baseline` encoders, MELD shape, B=16 — against the CPU oracle
    (logits, loss, whole-model flat gradient), the oracle running the batch in chunks of 2 samples (the model is
    per-sample; only the weighted-CE normaliser couples the samples).

Stated tolerance (bf16 tensor-core operands, fp32 elsewhere): logits 3e-2 absolute, loss 2e-2, flat gradient
2e-2 relative-L2; integer and mask tensors bit-exact; the loss kernel 1e-5 / 1e-6 (pure fp32)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _tiny(variant, C, seeds=(1, 2)):
    from multi_modal_emotion_b200 import synthetic as syn, tav

    tav.set_encoder_variant(variant)
    torch.manual_seed(0)
    model = tav.TAVForMAE({"output_dim": C, "dropout": 0.4, "learn_PosEmbeddings": True, "num_layers": 12})
    pre = tav.PreFormer()
    pre_sd, tav_sd = syn.synth_state_dict(pre, seed=seeds[0]), syn.synth_state_dict(model, seed=seeds[1])
    pre.load_state_dict(pre_sd)
    model.load_state_dict(tav_sd)
    return model.cuda(), pre.cuda(), pre_sd, tav_sd


def test_preformer_integer_and_mask_outputs_bit_exact_vs_reference_golden():
    from multi_modal_emotion_b200 import synthetic as syn

    gold = torch.load(os.path.join(GOLD, "tav_tiny_C1.pt"))
    _, pre, _, _ = _tiny("tiny", 7)
    inputs, _ = syn.make_batch("C1")            # ragged: text lengths 32 -> 5, audio 16000 -> 5333 samples
    ids, tm = inputs[0]["input_ids"], inputs[0]["attention_mask"]
    wav, am = inputs[1]["audio_features"], inputs[1]["attention_mask"]
    vid, vm = inputs[2]["visual_embeds"], inputs[2]["attention_mask"]
    assert len(set(tm.sum(1).tolist())) > 1 and len(set(am.sum(1).tolist())) > 1
    with torch.no_grad():
        tav_h, tav_embed, attention_mask = pre(input_ids=ids.cuda(), audio_features=wav.cuda(), video_embeds=vid.cuda(),
                                               text_mask=tm.cuda(), audio_mask=am.cuda(), visual_mask=vm, device="cuda",
                                               train=False)
    assert tav_embed.dtype == torch.long and torch.equal(tav_embed.cpu(), gold["pos"])
    assert attention_mask.dtype == torch.float32 and torch.equal(attention_mask.cpu(), gold["mask"])
    assert sorted(set(attention_mask.flatten().tolist())) == [-65504.0, 0.0, 1.0, 65505.0]
    # frame lengths / frame mask on the device: the notebook KAT f(3280) = 10 and the per-sample valid-frame counts
    lens = pre._get_feat_extract_output_lengths(torch.tensor([3280, 16000, 400, 399 + 320], device="cuda"))
    assert lens.tolist() == [10, 49, 1, 1]
    fm = pre._get_feature_vector_attention_mask(49, am.cuda())
    want = torch.arange(49)[None, :] < torch.tensor([syn.conv_frames(int(n)) for n in am.sum(1)])[:, None]
    assert fm.dtype == torch.bool and torch.equal(fm.cpu(), want)
    T, Ta = ids.shape[1], 49
    assert torch.equal(attention_mask[:, 0, 0, T:T + Ta].cpu(), 1.0 - want.float() * torch.finfo(torch.float16).min)


def test_new_cross_entropy_gpu_vs_reference_golden_epochs_0_to_3():
    from multi_modal_emotion_b200.losses import NewCrossEntropyLoss

    gold = torch.load(os.path.join(GOLD, "new_ce.pt"))
    crit = NewCrossEntropyLoss(gold["weights"].clone(), epoch_switch=gold["epoch_switch"])
    for epoch in range(4):
        logits = gold["logits"].clone().cuda().requires_grad_(True)
        loss = crit(logits, gold["target"].cuda(), epoch)
        loss.backward()
        assert abs(loss.item() - gold["loss"][epoch]) < 1e-5, (epoch, loss.item(), gold["loss"][epoch])
        assert (logits.grad.cpu() - gold["dlogits"][epoch]).abs().max().item() < 1e-6, epoch
    assert gold["loss"][0] != gold["loss"][1]           # the switch between unweighted and weighted CE is exercised


@pytest.mark.parametrize("variant,cfg,B", [("tiny", "C1", 2), ("tiny_base", "C2", 2), ("tiny_base", "C4", 2)])
def test_five_seeds_argmax_equal_everywhere_with_margin_table(variant, cfg, B):
    from multi_modal_emotion_b200 import synthetic as syn, tav
    from oracle import tav_oracle as O

    C = syn.CONFIGS[cfg]["C"]
    model, pre, pre_sd, tav_sd = _tiny(variant, C, seeds=(5, 6))
    orc = O.OracleTAV(tav.encoder_configs(variant)).load(pre_sd, tav_sd)
    rows = []
    for seed in (101, 202, 303, 404, 505):
        inputs, _ = syn.make_batch(cfg, seed=seed, B=B)
        with torch.no_grad():
            lo = orc.forward(inputs)
            dev = [{k: v.cuda() for k, v in d.items()} for d in inputs]
            t, pos, mask = pre(input_ids=dev[0]["input_ids"], audio_features=dev[1]["audio_features"],
                               video_embeds=dev[2]["visual_embeds"], text_mask=dev[0]["attention_mask"],
                               audio_mask=dev[1]["attention_mask"], visual_mask=inputs[2]["attention_mask"], device="cuda")
            lg = model(dev[0]["input_ids"], dev[0]["attention_mask"], dev[1]["audio_features"], dev[2]["visual_embeds"],
                       inputs[2]["attention_mask"], t, pos, mask, batch_size=B, check="val").cpu()
        top2 = lo.topk(2, dim=1).values
        for b in range(B):
            rows.append((seed, b, int(lo[b].argmax()), int(lg[b].argmax()), float(top2[b, 0] - top2[b, 1]),
                         float((lg[b] - lo[b]).abs().max())))
    print("%s %s: seed row oracle_argmax gpu_argmax top1-top2_margin max|dlogit|" % (variant, cfg))
    for r in rows:
        print("   %4d %3d %6d %6d %10.4f %10.2e" % r)
    assert all(r[5] < 3e-2 for r in rows)
    assert all(r[2] == r[3] for r in rows), [r for r in rows if r[2] != r[3]]


def test_benchmark_configuration_full_depth_b16_vs_cpu_oracle():
    """The configuration bench.py times (C2: MELD shape, B=16, 12-layer RoBERTa-base / Wav2Vec2-base / VideoMAE-base,
    12 fusion layers) checked end to end: logits, loss and the whole-model flat gradient against the CPU oracle."""
    import gc

    from multi_modal_emotion_b200 import synthetic as syn, tav
    from multi_modal_emotion_b200.losses import NewCrossEntropyLoss
    from multi_modal_emotion_b200.tav_train import get_statistics
    from oracle import tav_oracle as O

    B, chunk = 16, 2
    model, pre, pre_sd, tav_sd = _tiny("baseline", 7, seeds=(7, 8))
    inputs, labels = syn.make_batch("C2", seed=77, B=B)
    w = torch.tensor(syn.MELD_CLASS_WEIGHTS)
    crit = NewCrossEntropyLoss(class_weights=w, epoch_switch=2)
    cap = {}
    h = model.register_forward_hook(lambda m, i, o: cap.__setitem__("logits", o.detach().clone()))
    loss = get_statistics(inputs, labels, model, pre, crit, None, check="val", epoch=1)
    h.remove()
    loss.backward()
    grads = {}
    for tag, m in (("TAVForMAE", model), ("PreFormer", pre)):
        for k, p in m.named_parameters():
            if p.grad is not None:
                grads["%s/%s" % (tag, k)] = p.grad.detach().cpu()
    logits, loss_v = cap["logits"].cpu(), loss.item()
    del model, pre, loss, cap
    gc.collect()
    torch.cuda.empty_cache()
    # oracle: chunks of 2 samples, each back-propagating  sum_i w[y_i] l_i / sum_{all 16} w[y_i]
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    orc = O.OracleTAV(tav.encoder_configs("baseline")).load(pre_sd, tav_sd)
    y = labels.long()
    den = w[y].sum()
    lo_all, num = [], 0.0
    for c0 in range(0, B, chunk):
        sub = [{k: v[c0:c0 + chunk] for k, v in d.items()} for d in inputs]
        lo = orc.forward(sub)
        part = torch.nn.functional.cross_entropy(lo, y[c0:c0 + chunk], weight=w, reduction="sum") / den
        part.backward()
        num += part.item()
        lo_all.append(lo.detach())
    lo = torch.cat(lo_all)
    og = orc.named_grads()
    err = (logits - lo).abs().max().item()
    top2 = lo.topk(2, dim=1).values
    margin = top2[:, 0] - top2[:, 1]
    assert set(og) == set(grads)
    d2 = sum((grads[k] - g).norm().item() ** 2 for k, g in og.items())
    n2 = sum(g.norm().item() ** 2 for g in og.values())
    print("benchmark configuration (baseline, C2, B=16): logits max abs err %.3e (min margin %.3f), loss %.6f vs %.6f, "
          "whole-model flat gradient rel-L2 %.3e over %d tensors" % (err, margin.min().item(), loss_v, num, (d2 / n2) ** 0.5, len(og)))
    assert err < 3e-2
    assert abs(loss_v - num) < 2e-2
    sure = margin > 3e-2
    assert torch.equal(logits.argmax(dim=1)[sure], lo.argmax(dim=1)[sure])
    assert (d2 / n2) ** 0.5 < 2e-2
