"""``NewCrossEntropyLoss`` drop-in (reference utils/global_functions.py:51-83) on the fused softmax-CE kernel."""
import torch
from torch import nn

from . import engine


class NewCrossEntropyLoss(nn.Module):
    """Alternates between unweighted CE (``epoch % epoch_switch == 0``) and class-weighted CE, both with torch's
    mean reduction (weighted: sum_i w[y_i] l_i / sum_i w[y_i]).  Same constructor / call signature as the reference.

    Data-parallel note: ``parts()`` returns the numerator and denominator separately so that
    ``dp.DataParallelTAV`` can all-reduce the denominator and reproduce the single-process loss on the global batch."""

    def __init__(self, class_weights, epoch_switch=2):
        super().__init__()
        self.class_weights = class_weights
        self.epoch_switch = epoch_switch
        # Same child modules as the reference (utils/global_functions.py:63-64): they are state holders here (the
        # fused kernel does the arithmetic), so ``state_dict()`` carries the reference's only key,
        # ``weightedCEL.weight``, and a ``best.pt`` 'loss' entry loads in either direction.
        self.weightedCEL = nn.CrossEntropyLoss(weight=class_weights)
        self.normalCEL = nn.CrossEntropyLoss()
        self.iter1 = self.iter2 = self.iter3 = 1

    def _weights(self, epoch, device):
        w = self.weightedCEL.weight
        if epoch % self.epoch_switch == 0 or w is None:
            return None
        if w.device != device or w.dtype != torch.float32:
            w = w.to(device=device, dtype=torch.float32)
            self.weightedCEL.weight = w            # re-registers the buffer on the compute device
        return w

    def parts(self, logits, target, epoch):
        return engine.softmax_ce_parts(logits, target, self._weights(epoch, logits.device))

    def forward(self, logits, target, epoch):
        num, den = self.parts(logits, target, epoch)
        return num / den


class CrossEntropyLoss(nn.Module):
    """torch.nn.CrossEntropyLoss(weight=...) with mean reduction on the fused kernel (reference tav_nn.py:83-89 uses
    the stock module when no class weights are configured).  Accepts and ignores ``epoch`` like the training loop
    passes it (train_model/tav_train.py:47)."""

    def __init__(self, weight=None):
        super().__init__()
        self.weight = weight

    def parts(self, logits, target, epoch=None):
        w = self.weight
        if w is not None and (w.device != logits.device or w.dtype != torch.float32):
            w = self.weight = w.to(device=logits.device, dtype=torch.float32)
        return engine.softmax_ce_parts(logits, target, w)

    def forward(self, logits, target, epoch=None):
        num, den = self.parts(logits, target, epoch)
        return num / den
