"""Training / evaluation loop of the TAV model with the reference's public functions and argument order
(reference train_model/tav_train.py): ``get_statistics``, ``not_grad_accum``, ``grad_accum``, ``validate``,
``one_epoch``, ``train_tav_network``, ``evaluate_tav``.

What changed underneath: inputs are moved to the device once; the B==1 assert on the video tensor
(tav_train.py:32) is relaxed to a shape check so batched runs work; ``clip_grad_norm_`` + ``AdamW.step`` +
``zero_grad`` are one fused step over flat buffers (optim.FusedAdamW); when ``torch.distributed`` is initialised with
more than one rank, ``train_tav_network`` drives every training iteration through ``dp.DataParallelTAV`` (global-loss
normalisation, bucketed gradient all-reduce overlapped with backward), so N ranks reproduce the single-process
step on the concatenated batch instead of training N unsynchronised replicas.  Quirks kept on purpose: ``grad_accum`` steps
the optimiser every iteration and once more at dialogue boundaries (Q10); the scheduler is stepped with the
fractional epoch.  wandb logging and checkpoint files are host-side bookkeeping outside the hot path: logging goes
through ``log_fn`` (defaults to wandb when it is importable and active, else a no-op) and checkpoints through
``checkpoint_io`` when the caller provides one."""
import torch
from torch.optim.lr_scheduler import CosineAnnealingWarmRestarts

from .optim import FusedAdamW

PATIENCE_ITER = 0
LOG_VAL = 2400  # reference: log / validate every 2400 iterations (tav_train.py:136)


def _default_log(payload):
    try:
        import wandb

        if wandb.run is not None:
            wandb.log(payload)
    except Exception:  # noqa: BLE001 — logging must never break a step
        pass


log_fn = _default_log
_dp_runner = None     # dp.DataParallelTAV while train_tav_network runs under torch.distributed (world_size > 1)
checkpoint_io = None  # optional object with save(model, PREFormer, optimizer, criterion, scheduler, epoch, step) / load(...)


def get_statistics(input, label, model, PREFormer, criterion, Metric, check="train", epoch=None):
    """One forward pass + loss (reference tav_train.py:15-48).  ``input`` = [text, audio, video] dicts as emitted by
    the reference's collate function (models/tav.py:235-246)."""
    device = next(model.parameters()).device
    batch_size = len(label)
    text, audio, video = input[0], input[1], input[2]
    ids, text_mask = text["input_ids"], text["attention_mask"]
    wav, audio_mask = audio["audio_features"], audio["attention_mask"]
    vid, vid_mask = video["visual_embeds"], video["attention_mask"]
    if tuple(vid.shape[1:]) != (16, 3, 224, 224):
        raise AssertionError(f"Shape of video is {vid.shape}")
    keep = int(vid_mask[0].sum()) if not vid_mask.is_cuda else None
    nb = dict(non_blocking=True)
    ids, text_mask, wav, audio_mask = ids.to(device, **nb), text_mask.to(device, **nb), wav.to(device, **nb), audio_mask.to(device, **nb)
    vid, vid_mask_d = vid.to(device, **nb), vid_mask.to(device, **nb)
    tav, tav_embed, attention_mask = PREFormer(input_ids=ids, audio_features=wav, video_embeds=vid, text_mask=text_mask,
                                               audio_mask=audio_mask, visual_mask=vid_mask if keep is not None else vid_mask_d,
                                               device=device, train=(check == "train"))
    output = model(input_ids=ids, text_attention_mask=text_mask, audio_features=wav, video_embeds=vid,
                   visual_mask=vid_mask if keep is not None else vid_mask_d, hidden_states=tav, pos_embed=tav_embed,
                   attention_mask=attention_mask, batch_size=batch_size, check=check)
    label = label.to(device, **nb).long()
    if label.dim() > 1:
        label = label.view(-1)
    if Metric is not None:
        Metric.update_metrics(torch.argmax(output, dim=1), label)
    batch_loss = None
    if criterion is not None:
        batch_loss = criterion(output, label, epoch=epoch if epoch is not None else 1)
    return batch_loss


def _optimizer_step(model, PREFormer, optimizer, scheduler, clip, t):
    if isinstance(optimizer, FusedAdamW):
        optimizer.step(max_grad_norm=clip)   # sqnorm + clip + AdamW + zero_grad, fused
        scheduler.step(t)
    else:  # stock torch optimiser: the reference sequence verbatim (tav_train.py:61-65)
        params = [p for p in model.parameters() if p.requires_grad] + [p for p in PREFormer.parameters() if p.requires_grad]
        torch.nn.utils.clip_grad_norm_(params, clip)
        optimizer.step()
        scheduler.step(t)
        model.zero_grad()
        PREFormer.zero_grad()


def _dp_step(train_input, train_label, Metric, epoch, loss_scale, clip, scheduler, t):
    """One data-parallel iteration: forward on the local shard, GLOBAL loss, backward overlapped with the bucketed
    all-reduce, fused clip + AdamW on identical gradients, scheduler step.  Returns the global loss as a float."""
    loss = _dp_runner._eager_step(train_input, train_label, epoch, "train", loss_scale=loss_scale, clip=clip,
                                  Metric=Metric)     # Metric accumulates this rank's shard
    scheduler.step(t)
    return loss.item()


def _validate_and_track(epoch, batch_idx, val_dataloader, model, PREFormer, criterion, optimizer, scheduler, Metric,
                        prev_val_loss, total_loss_train, iters, patience):
    global PATIENCE_ITER
    log(Metric, total_loss_train / iters, "train")
    val_loss = validate(val_dataloader, model, PREFormer, criterion, Metric, name="val")
    if val_loss < prev_val_loss:
        PATIENCE_ITER = 0
        prev_val_loss = val_loss
        if checkpoint_io is not None:
            checkpoint_io.save(model, PREFormer, optimizer, criterion, scheduler, epoch, batch_idx)
    else:
        PATIENCE_ITER += 1
    return prev_val_loss, PATIENCE_ITER == patience


def not_grad_accum(epoch, train_dataloader, val_dataloader, model, PREFormer, criterion, optimizer, scheduler, clip,
                   patience, Metric, prev_val_loss, total_loss_train, iters, log_val, path):
    for batch_idx, (train_input, train_label) in enumerate(train_dataloader):
        if _dp_runner is not None:
            total_loss_train += _dp_step(train_input, train_label, Metric, epoch, 1.0, clip, scheduler, epoch + batch_idx / iters)
        else:
            loss = get_statistics(train_input, train_label, model, PREFormer, criterion, Metric, check="train", epoch=epoch)
            total_loss_train += loss.item()
            loss.backward()
            _optimizer_step(model, PREFormer, optimizer, scheduler, clip, epoch + batch_idx / iters)
        if ((batch_idx + 1) % log_val == 0) or (batch_idx + 1 == iters):
            prev_val_loss, stop = _validate_and_track(epoch, batch_idx, val_dataloader, model, PREFormer, criterion,
                                                      optimizer, scheduler, Metric, prev_val_loss, total_loss_train,
                                                      iters, patience)
            if stop:
                break
    return prev_val_loss


def grad_accum(epoch, train_dataloader, val_dataloader, model, PREFormer, criterion, optimizer, scheduler, clip,
               patience, Metric, prev_val_loss, total_loss_train, iters, log_val, path):
    for batch_idx, (train_input, train_label) in enumerate(train_dataloader):
        accum_iter, accum_sum = train_dataloader.dataset.retGradAccum(i=batch_idx)
        if _dp_runner is not None:
            total_loss_train += _dp_step(train_input, train_label, Metric, epoch, 1.0 / accum_iter, clip, scheduler,
                                         epoch + batch_idx / iters)
        else:
            loss = get_statistics(train_input, train_label, model, PREFormer, criterion, Metric, check="train", epoch=epoch) / accum_iter
            total_loss_train += loss.item()
            loss.backward()
            _optimizer_step(model, PREFormer, optimizer, scheduler, clip, epoch + batch_idx / iters)
        if ((batch_idx + 1) % accum_sum == 0) or (batch_idx + 1 == iters):   # second step at dialogue boundaries (Q10)
            _optimizer_step(model, PREFormer, optimizer, scheduler, None, epoch + batch_idx / iters)
        if ((batch_idx + 1) % log_val == 0) or (batch_idx + 1 == iters):
            prev_val_loss, stop = _validate_and_track(epoch, batch_idx, val_dataloader, model, PREFormer, criterion,
                                                      optimizer, scheduler, Metric, prev_val_loss, total_loss_train,
                                                      iters, patience)
            if stop:
                break
    return prev_val_loss


def validate(val_dataloader, model, PREFormer, criterion, Metric, name="val"):
    total = 0
    with torch.no_grad():
        for val_input, val_label in val_dataloader:
            loss = get_statistics(val_input, val_label, model, PREFormer, criterion, Metric, name, epoch=None)
            if criterion is not None:
                total += loss.item()
        log(Metric, total / len(val_dataloader) if criterion is not None else 0, name)
    return total / len(val_dataloader)


def one_epoch(epoch, train_dataloader, val_dataloader, model, PREFormer, criterion, optimizer, scheduler, clip,
              epoch_switch, patience, Metric, prev_val_loss):
    iters = len(train_dataloader)
    body = not_grad_accum if epoch % epoch_switch == 0 else grad_accum
    prev_val_loss = body(epoch, train_dataloader, val_dataloader, model, PREFormer, criterion, optimizer, scheduler,
                         clip, patience, Metric, prev_val_loss, 0, iters, LOG_VAL, None)
    if checkpoint_io is not None:   # the reference reloads the best checkpoint after every epoch (tav_train.py:143)
        model, PREFormer, optimizer, criterion = checkpoint_io.load(model, PREFormer, optimizer, criterion)
    return model, PREFormer, optimizer, criterion, scheduler, prev_val_loss


def train_tav_network(model, PREFormer, train_dataloader, val_dataloader, criterion, learning_rate, epochs,
                      weight_decay, T_max, Metric, patience, clip, epoch_switch, checkpoint=None):
    global PATIENCE_ITER, _dp_runner
    from . import dp

    PATIENCE_ITER = 0
    params = [p for p in model.parameters() if p.requires_grad] + [p for p in PREFormer.parameters() if p.requires_grad]
    optimizer = FusedAdamW(params, lr=learning_rate, weight_decay=weight_decay)
    scheduler = CosineAnnealingWarmRestarts(optimizer, T_0=T_max)
    prev_val_loss = 100
    if checkpoint is not None:
        optimizer.load_state_dict(checkpoint["optimizer_state_dict"])
        scheduler.load_state_dict(checkpoint["scheduler_state_dict"])
    # under torchrun every rank holds a shard of each batch: couple the replicas
    _dp_runner = dp.DataParallelTAV(model, PREFormer, criterion, optimizer, clip=clip) if dp.is_distributed() else None
    try:
        for epoch_num in range(epochs):
            log_fn({"epoch": epoch_num, "learning_rate": scheduler.get_last_lr()[0]})
            optimizer.zero_grad()
            model, PREFormer, optimizer, criterion, scheduler, prev_val_loss = one_epoch(
                epoch_num, train_dataloader, val_dataloader, model, PREFormer, criterion, optimizer, scheduler, clip,
                epoch_switch, patience, Metric, prev_val_loss)
            if PATIENCE_ITER == patience:
                return model, PREFormer
    finally:
        _dp_runner = None
    return model, PREFormer


def evaluate_tav(model, PREFormer, test_dataloader, Metric):
    validate(test_dataloader, model, PREFormer, None, Metric, name="test")


def log(Metric, loss, check="train"):
    if Metric is None:
        log_fn({f"{check}/loss": loss})
        return
    multiAcc, multiF1, multiRec, multiPrec, Acc, F1Macro, F1Weighted, Rec, Prec, _ = Metric.compute_scores(f"{check}")
    log_fn({f"{check}/loss": loss, f"{check}/acc": Acc, f"{check}/precision": Prec, f"{check}/recall": Rec,
            f"{check}/weighted-f1-score": F1Weighted, f"{check}/macro-f1-score": F1Macro,
            **multiF1, **multiRec, **multiPrec, **multiAcc})
    Metric.reset_metrics()
