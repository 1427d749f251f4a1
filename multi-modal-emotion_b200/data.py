"""Batch format of the TAV path: the dict-of-tensors wire format that ``get_statistics`` consumes
(reference ``collate_batch``, models/tav.py:174-246, fed by ``TextAudioVideoDataset``, utils/data_loaders.py:12-65).

Decoding mp4 / wav files and tokenising text are out of scope (no datasets, no network): ``SyntheticTAVDataset`` yields
items with the same structure the reference's collate sees AFTER decoding — a tokenised text dict, a waveform, a video
clip ``[3, 16, 224, 224]`` — and ``collate_batch`` performs the reference's batching steps on them: stack the token
rows, zero-pad the waveforms to the longest clip with a sample-level mask, stack the clips to ``[B, 16, 3, 224, 224]``
and draw the random video-token mask."""
import torch
from torch.nn.utils.rnn import pad_sequence
from torch.utils.data import Dataset

from . import synthetic as syn

VIDEO_TOKENS = 1568   # 8*14*14, hard-coded in the reference (models/tav.py:207)


def video_token_mask(batch_size, generator=None, equal_rows=True, n_tokens=VIDEO_TOKENS):
    """The reference draws ``randint(-13, 2) > 0`` (keep probability 1/15, ~104.5 tokens per clip) and then flips zeros
    to ones until the number of zeros is divisible by the batch size (models/tav.py:207-218).  HF's VideoMAE
    nevertheless needs the SAME number of masked tokens in every row (``embeddings[~bool_masked_pos].reshape(B, -1, C)``;
    SURVEY Q9), which the reference only meets by chance for B > 1.  ``equal_rows=True`` (default) keeps exactly
    floor(n_tokens/15) = 104 tokens per row at random positions; ``equal_rows=False`` reproduces the reference draw."""
    if equal_rows:
        k = n_tokens // 15
        mask = torch.zeros(batch_size, n_tokens, dtype=torch.bool)
        for b in range(batch_size):
            mask[b, torch.randperm(n_tokens, generator=generator)[:k]] = True
        return mask
    mask = torch.randint(-13, 2, (batch_size, n_tokens), generator=generator)
    mask[mask < 0] = 0
    mask = mask.bool()
    rem = (n_tokens * batch_size - int(mask.sum())) % batch_size
    if rem != 0:
        zeros = torch.where(mask.view(-1) == 0)[0]
        pick = zeros[torch.randperm(len(zeros), generator=generator)[:rem]]
        mask.view(-1)[pick] = True
    return mask


def collate_batch(batch, check="train", generator=None, equal_rows=True):
    """[(item, label)] -> ([text, audio_features, visual_embeds], labels) exactly in the layout of the reference's
    ``collate_batch`` (models/tav.py:235-246): item = [text_dict, waveform, clip]; ``check`` is accepted for signature
    compatibility (the reference only forwards it to the video decoder)."""
    ids, tmask, waves, clips, labels = [], [], [], [], []
    for (text, wav, clip), label in batch:
        ids.append(torch.as_tensor(text["input_ids"]).reshape(-1).long())
        tmask.append(torch.as_tensor(text["attention_mask"]).reshape(-1).float())
        waves.append(torch.as_tensor(wav).reshape(-1).float())
        clips.append(torch.as_tensor(clip).float())
        labels.append(float(label))
    B = len(labels)
    lens = torch.tensor([len(w) for w in waves])
    audio = pad_sequence(waves, batch_first=True, padding_value=0.0)                       # models/tav.py:228
    audio_mask = (torch.arange(audio.shape[1])[None, :] < lens[:, None]).float()            # PROC(padding=True) mask, :225
    video = torch.stack(clips).permute(0, 2, 1, 3, 4).contiguous()                          # [B,3,16,H,W] -> [B,16,3,H,W], :243
    text = {"input_ids": torch.stack(ids), "attention_mask": torch.stack(tmask)}
    audio_features = {"audio_features": audio, "attention_mask": audio_mask}
    visual_embeds = {"visual_embeds": video, "attention_mask": video_token_mask(B, generator, equal_rows)}
    return [text, audio_features, visual_embeds], torch.tensor(labels)


class SyntheticTAVDataset(Dataset):
    """Same item schema as the reference's ``TextAudioVideoDataset.__getitem__`` after decoding, generated from
    per-index seeds (shapes of a synthetic.CONFIGS entry; variable text and audio lengths)."""

    def __init__(self, n, cfg="C2", seed=1234, dialog_lengths=None):
        self.n, self.cfg, self.seed = n, dict(syn.CONFIGS[cfg]) if isinstance(cfg, str) else dict(cfg), seed
        # gradient-accumulation bookkeeping of the reference (utils/data_loaders.py:23-25,46-57): utterances per dialogue
        self.grad = list(dialog_lengths) if dialog_lengths else [n]
        self.grad_sum = [sum(self.grad[:i + 1]) for i in range(len(self.grad))]
        self.ctr = 0

    def retGradAccum(self, i):
        g, gs = self.grad[self.ctr], self.grad_sum[self.ctr]
        if i + 1 == self.grad_sum[self.ctr]:
            self.ctr += 1
        if self.ctr == len(self.grad):
            self.resetCtr()
        return g, gs

    def resetCtr(self):
        self.ctr = 0

    def __len__(self):
        return self.n

    def __getitem__(self, idx):
        c = self.cfg
        g = torch.Generator().manual_seed(self.seed * 1000003 + idx)
        T, L = c["T"], c["L"]
        tl = int(torch.randint(max(4, T // 6), T + 1, (1,), generator=g))
        al = int(torch.randint(max(1024, L // 3), L + 1, (1,), generator=g))
        ids = torch.ones(1, T, dtype=torch.long)                       # pad id 1, as the RoBERTa tokenizer pads
        ids[0, :tl] = torch.randint(3, 50265, (tl,), generator=g)
        mask = torch.zeros(1, T, dtype=torch.long)
        mask[0, :tl] = 1
        wav = 0.1 * torch.randn(al, generator=g)
        clip = torch.randn(3, 16, 224, 224, generator=g)
        label = int(torch.randint(0, c["C"], (1,), generator=g))
        return [{"input_ids": ids, "attention_mask": mask}, wav, clip], label
