"""Checkpoint round trip in the reference's ``best.pt`` layout (reference ``save_model`` /
``load_model``, utils/global_functions.py:199-258): keys ``epoch, step, model_state_dict, optimizer_state_dict, loss,
scheduler, PREFormer``.  Parameter names are the reference's, and ``optim.FusedAdamW.state_dict()`` is emitted in
``torch.optim.AdamW`` layout, so files move between the reference and this path in both directions.  Plug an instance
into ``tav_train.checkpoint_io`` to get the reference's save-on-improvement / reload-after-epoch behaviour."""
import os

import torch


class CheckpointIO:
    def __init__(self, directory, filename="best.pt"):
        self.path = os.path.join(directory, filename)

    def save(self, model, PREFormer, optimizer, criterion, scheduler, epoch, step):
        os.makedirs(os.path.dirname(self.path) or ".", exist_ok=True)
        payload = {
            "epoch": epoch,
            "step": step,
            "model_state_dict": model.state_dict(),
            "optimizer_state_dict": optimizer.state_dict(),
            "loss": criterion.state_dict(),
            "scheduler": scheduler.state_dict(),
        }
        if PREFormer is not None:
            payload["PREFormer"] = PREFormer.state_dict()
        tmp = self.path + ".tmp"
        torch.save(payload, tmp)
        os.replace(tmp, self.path)          # never leave a half-written best.pt behind

    def load(self, model, PREFormer, optimizer, criterion, map_location=None):
        """Restores everything in place and returns ``(model, PREFormer, optimizer, criterion)`` like the reference.
        (The reference rebuilds a default-hyper-parameter AdamW before loading the state, utils/global_functions.py:253;
        loading into the live optimiser keeps lr / weight decay and, for FusedAdamW, the flat buffers.)"""
        ckpt = torch.load(self.path, map_location=map_location)
        model.load_state_dict(ckpt["model_state_dict"])
        if PREFormer is not None and "PREFormer" in ckpt:
            PREFormer.load_state_dict(ckpt["PREFormer"])
        optimizer.load_state_dict(ckpt["optimizer_state_dict"])
        criterion.load_state_dict(ckpt["loss"])
        self.last = {"epoch": ckpt["epoch"], "step": ckpt["step"], "scheduler": ckpt.get("scheduler")}
        return model, PREFormer, optimizer, criterion
