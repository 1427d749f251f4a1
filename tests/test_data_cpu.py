"""Host-side batch format and checkpoint layout — no GPU needed."""
import os

import torch

from multi_modal_emotion_b200 import data, synthetic as syn
from multi_modal_emotion_b200.checkpoint import CheckpointIO


def test_collate_batch_matches_reference_layout():
    ds = data.SyntheticTAVDataset(5, cfg="C1", seed=7)
    items = [ds[i] for i in range(4)]
    g = torch.Generator().manual_seed(0)
    (text, audio, video), labels = data.collate_batch(items, "train", generator=g)
    c = syn.CONFIGS["C1"]
    assert text["input_ids"].shape == (4, c["T"]) and text["input_ids"].dtype == torch.long
    assert text["attention_mask"].shape == (4, c["T"]) and text["attention_mask"].dtype == torch.float32
    lens = [len(it[0][1]) for it in items]
    assert audio["audio_features"].shape == (4, max(lens))
    assert audio["attention_mask"].sum(dim=1).tolist() == [float(n) for n in lens]
    for b, n in enumerate(lens):                       # zero padding beyond each clip, data intact before it
        assert torch.equal(audio["audio_features"][b, :n], items[b][0][1])
        assert audio["audio_features"][b, n:].abs().sum() == 0
    assert video["visual_embeds"].shape == (4, 16, 3, 224, 224)
    assert torch.equal(video["visual_embeds"][2, 5, 1], items[2][0][2][1, 5])   # [3,16,H,W] -> [16,3,H,W] per sample
    m = video["attention_mask"]
    assert m.shape == (4, 1568) and m.dtype == torch.bool and set(m.sum(dim=1).tolist()) == {104}
    assert labels.dtype == torch.float32 and labels.shape == (4,)
    # same keys / nesting as synthetic.make_batch, which the GPU parity tests feed to get_statistics
    ref_inputs, _ = syn.make_batch("C1")
    assert [sorted(d) for d in (text, audio, video)] == [sorted(d) for d in ref_inputs]


def test_reference_style_video_mask_quirk():
    g = torch.Generator().manual_seed(3)
    for B in (1, 3, 8):
        m = data.video_token_mask(B, g, equal_rows=False)
        assert m.shape == (B, 1568)
        assert (1568 * B - int(m.sum())) % B == 0      # the reference's divisibility fix-up (models/tav.py:212-218)
        frac = m.float().mean().item()
        assert 0.04 < frac < 0.10                      # keep probability 1/15


def test_grad_accum_bookkeeping_matches_reference_dataset():
    ds = data.SyntheticTAVDataset(6, cfg="C1", dialog_lengths=[2, 3, 1])
    got = [ds.retGradAccum(i) for i in range(6)]
    assert got == [(2, 2), (2, 2), (3, 5), (3, 5), (3, 5), (1, 6)]
    assert ds.ctr == 0                                  # wrapped around after the last dialogue


def test_checkpoint_roundtrip_reference_keys(tmp_path):
    torch.manual_seed(0)
    model, pre = torch.nn.Linear(4, 3), torch.nn.Linear(2, 2)
    params = list(model.parameters()) + list(pre.parameters())
    opt = torch.optim.AdamW(params, lr=1e-3, weight_decay=1e-2)
    sched = torch.optim.lr_scheduler.CosineAnnealingWarmRestarts(opt, T_0=5)
    crit = torch.nn.CrossEntropyLoss(weight=torch.tensor([1.0, 2.0, 3.0]))
    (model(torch.randn(5, 4)).sum() + pre(torch.randn(5, 2)).sum()).backward()
    opt.step()
    sched.step(0.5)
    io = CheckpointIO(str(tmp_path / "proj" / "sweep" / "run"))
    io.save(model, pre, opt, crit, sched, epoch=2, step=17)
    ck = torch.load(io.path)
    assert set(ck) == {"epoch", "step", "model_state_dict", "optimizer_state_dict", "loss", "scheduler", "PREFormer"}
    model2, pre2 = torch.nn.Linear(4, 3), torch.nn.Linear(2, 2)
    opt2 = torch.optim.AdamW(list(model2.parameters()) + list(pre2.parameters()), lr=1e-3, weight_decay=1e-2)
    crit2 = torch.nn.CrossEntropyLoss(weight=torch.ones(3))
    io.load(model2, pre2, opt2, crit2)
    assert torch.equal(model2.weight, model.weight) and torch.equal(pre2.bias, pre.bias)
    assert torch.equal(crit2.weight, crit.weight)
    s1, s2 = opt.state_dict()["state"], opt2.state_dict()["state"]
    assert all(torch.equal(s1[k]["exp_avg_sq"], s2[k]["exp_avg_sq"]) for k in s1)
    assert io.last["epoch"] == 2 and io.last["step"] == 17 and not os.path.exists(io.path + ".tmp")
