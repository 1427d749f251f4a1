"""Fused optimiser step over flat fp32 buffers: ``clip_grad_norm_`` + ``torch.optim.AdamW.step`` of the reference
training loop (train_model/tav_train.py:61-62,148) as two kernel launches (tavk_grad_sqnorm, tavk_adamw).

``FusedAdamW`` is a ``torch.optim.Optimizer`` (so ``CosineAnnealingWarmRestarts`` can drive ``param_groups[0]['lr']``
exactly like the reference does, tav_train.py:149,63).  At construction every parameter is re-pointed at a slice of
one flat fp32 buffer and its ``.grad`` at the matching slice of a flat gradient buffer; gradient buckets for the
data-parallel all-reduce (dp.py) are contiguous ranges of that same buffer."""
import torch

from . import _lib as L, engine

_ALIGN = 64  # elements; keeps every parameter slice 256-byte aligned (float4 kernels need 16 bytes)


class FlatParams:
    """Owns flat fp32 parameter / gradient buffers; parameters become views (state_dict keys are unaffected)."""

    def __init__(self, params):
        params = [p for p in params if p.requires_grad]
        seen, uniq = set(), []
        for p in params:
            if id(p) not in seen:
                seen.add(id(p))
                uniq.append(p)
        self.params = uniq
        if not uniq:
            raise ValueError("no trainable parameters")
        dev = uniq[0].device
        if dev.type != "cuda":
            raise RuntimeError("FusedAdamW needs CUDA parameters (there is no CPU fallback)")
        self.offsets = []
        off = 0
        for p in uniq:
            self.offsets.append(off)
            off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.numel = off
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(off, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, o in zip(uniq, self.offsets):
                view = self.flat[o:o + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
                p.grad = self.grad[o:o + p.numel()].view(p.shape)
                p._tavk_flat = (self, o)     # lets engine.LayerShadow find the bf16 shadow of this parameter
        # bf16 mirror of the whole parameter buffer: written by the AdamW kernel in the same pass as the fp32 update, so the
        # GEMM operand copies ("shadows") of the layer engine are plain views of it — no per-step cast kernels
        self.shadow = torch.empty(off, dtype=torch.bfloat16, device=dev)
        L.call("tavk_cast_f32_bf16", self.flat.data_ptr(), self.shadow.data_ptr(), off)
        for p in uniq:
            p._tavk_mirror_version = p._version

    def shadow_view(self, p, offset):
        return self.shadow[offset:offset + p.numel()].view(p.shape)

    def attach_grads(self):
        """(Re-)point .grad at the flat buffer (after someone set grads to None)."""
        for p, o in zip(self.params, self.offsets):
            g = self.grad[o:o + p.numel()].view(p.shape)
            if p.grad is None:
                p.grad = g
            elif p.grad.data_ptr() != g.data_ptr():
                g.copy_(p.grad)
                p.grad = g


class FusedAdamW(torch.optim.Optimizer):
    """``clip_grad_norm_`` + ``torch.optim.AdamW.step`` + ``zero_grad`` as three launches over flat buffers.

    The optimiser clock is DEVICE-resident (``step_dev`` int32, ``hyper`` = {lr, 1-b1^t, sqrt(1-b2^t)}): a small prep
    kernel advances it, so a captured CUDA graph (dp.DataParallelTAV) applies the bias correction of the true step on
    every replay, and the learning rate is a device scalar the host refreshes (``upload_lr``) from
    ``param_groups[0]['lr']`` — i.e. from the reference's ``CosineAnnealingWarmRestarts.step(epoch + i/iters)``
    (train_model/tav_train.py:63,149) — before each step or replay.  ``state_dict()`` / ``load_state_dict()`` use
    ``torch.optim.AdamW``'s layout indexed over ALL parameters the optimiser was given (the reference builds AdamW over
    every ``requires_grad`` parameter, utils/global_functions.py:253), with state only for parameters that have
    received a gradient — exactly what stock AdamW holds."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_grad_norm=None):
        params = list(params)
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        if len(self.param_groups) != 1:
            raise ValueError("FusedAdamW supports a single parameter group (as the reference loop uses)")
        self.flat = None
        self.step_count = 0          # host mirror of step_dev (kept in step with it; never read back in the hot path)
        self.max_grad_norm = max_grad_norm
        self.grad_prescale = 1.0  # e.g. 1/world_size after a summing all-reduce
        self.on_materialize = None  # dp.py hooks bucket construction here
        self._pending_state = None   # a state dict loaded before the flat buffers exist (resume-before-first-step)
        self._extra_state = {}       # loaded state of parameters that have no gradient here: kept for the round trip
        self._index = None

    # ------------------------------------------------------------------ flat buffers
    def _all_params(self):
        return self.param_groups[0]["params"]

    def materialize(self):
        """Flatten at the first step: like torch.optim.AdamW, parameters that never receive a gradient (e.g. the
        unused halves of PreFormer's encoders, ``masked_spec_embed``) are left untouched — not even decayed."""
        if self.flat is not None:
            return
        allp = self._all_params()
        live = [p for p in allp if p.requires_grad and p.grad is not None]
        grads = [p.grad for p in live]
        self.flat = FlatParams(live)
        with torch.no_grad():
            for p, g0 in zip(self.flat.params, grads):
                p.grad.copy_(g0)
        pos = {id(p): i for i, p in enumerate(allp)}
        self._index = [pos[id(p)] for p in self.flat.params]        # flat slot -> index in the full parameter list
        live_ids = {id(p) for p in self.flat.params}
        self._dead = [p for p in allp if p.requires_grad and id(p) not in live_ids]
        n = self.flat.numel
        dev = self.flat.flat.device
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.sqnorm = torch.zeros(1, dtype=torch.float32, device=dev)
        self.step_dev = torch.full((1,), self.step_count, dtype=torch.int32, device=dev)
        self.hyper = torch.zeros(4, dtype=torch.float32, device=dev)
        if self._pending_state is not None:
            sd, self._pending_state = self._pending_state, None
            self._apply_state(sd)
        if self.on_materialize is not None:
            self.on_materialize(self.flat)

    def upload_lr(self):
        """Write param_groups[0]['lr'] (what the scheduler last set) into the device scalar the update kernel reads.
        ``step()`` does it itself except while a CUDA graph is being captured: a replayer calls this before each
        replay (a fill kernel with the value as its argument: no host-to-device copy, no sync)."""
        self.hyper[0:1].fill_(float(self.param_groups[0]["lr"]))

    def note_replayed(self, n=1):
        """A captured step was replayed n times: keep the host mirror of the device step counter current."""
        self.step_count += n

    @torch.no_grad()
    def step(self, closure=None, max_grad_norm=None):
        """One fused update.  ``max_grad_norm`` (or the constructor's) applies clip_grad_norm_ semantics:
        g *= min(1, max_norm / (||g||_2 + 1e-6)) over ALL parameters jointly, computed on device (no host sync)."""
        if closure is not None:
            raise NotImplementedError("closures are not supported")
        self.materialize()
        capturing = torch.cuda.is_current_stream_capturing()
        if not capturing:
            for p in self._dead:
                if p.grad is not None:
                    raise RuntimeError(
                        "FusedAdamW: a parameter of shape %s received its first gradient after the flat buffers were "
                        "built; it would never be optimised.  Build the optimiser after a representative backward pass "
                        "or freeze the parameter." % (tuple(p.shape),))
        self.flat.attach_grads()
        g = self.param_groups[0]
        clip = max_grad_norm if max_grad_norm is not None else self.max_grad_norm
        use_clip = clip is not None and clip > 0
        if not capturing:
            self.upload_lr()
            self.step_count += 1
        b1, b2 = float(g["betas"][0]), float(g["betas"][1])
        L.call("tavk_adamw_prep", self.step_dev.data_ptr(), self.hyper.data_ptr(),
               self.sqnorm.data_ptr() if use_clip else None, b1, b2)
        sq_ptr = None
        if use_clip:
            L.call("tavk_grad_sqnorm", self.flat.grad.data_ptr(), self.flat.numel, self.sqnorm.data_ptr())
            sq_ptr = self.sqnorm.data_ptr()
        L.call("tavk_adamw_dev", self.flat.flat.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
               self.flat.grad.data_ptr(), self.flat.shadow.data_ptr(), self.flat.numel, self.hyper.data_ptr(), b1, b2,
               float(g["eps"]), float(g["weight_decay"]), sq_ptr, float(clip) if use_clip else 0.0,
               float(self.grad_prescale), 1)
        engine.invalidate_shadows()  # parameters changed behind torch's version counters

    def zero_grad(self, set_to_none=False):
        if self.flat is None:
            return super().zero_grad(set_to_none=True)
        # the update kernel already zeroed the flat gradient; keep the views attached (set_to_none would detach them)
        self.flat.grad.zero_()
        self.flat.attach_grads()

    def grad_norm(self):
        """||g||_2 of the last step (device scalar; reading it synchronises)."""
        return self.sqnorm.sqrt() * abs(self.grad_prescale)

    # ------------------------------------------------------------------ snapshot / restore (side-effect-free warm-up)
    def snapshot(self):
        """Everything a step mutates besides the parameters themselves (dp.py snapshots those)."""
        if self.flat is None:
            return {"fresh": True, "step_count": self.step_count}
        return {"fresh": False, "step_count": self.step_count, "exp_avg": self.exp_avg.clone(),
                "exp_avg_sq": self.exp_avg_sq.clone(), "step_dev": self.step_dev.clone(), "hyper": self.hyper.clone()}

    @torch.no_grad()
    def restore(self, snap):
        """Undo the optimiser-state side of warm-up steps; call after the parameters were restored (re-casts the bf16
        mirror of the flat parameter buffer and clears the gradient)."""
        self.step_count = snap["step_count"]
        if self.flat is None:
            return
        if snap["fresh"]:
            self.exp_avg.zero_()
            self.exp_avg_sq.zero_()
            self.step_dev.fill_(self.step_count)
            self.hyper.zero_()
        else:
            self.exp_avg.copy_(snap["exp_avg"])
            self.exp_avg_sq.copy_(snap["exp_avg_sq"])
            self.step_dev.copy_(snap["step_dev"])
            self.hyper.copy_(snap["hyper"])
        self.flat.grad.zero_()
        L.call("tavk_cast_f32_bf16", self.flat.flat.data_ptr(), self.flat.shadow.data_ptr(), self.flat.numel)
        engine.invalidate_shadows()

    # ------------------------------------------------------------------ torch.optim.AdamW-compatible state
    def state_dict(self):
        """torch.optim.AdamW layout over the FULL parameter list: ``param_groups[0]['params']`` =
        range(len(all parameters)); ``state[i]`` = {step, exp_avg, exp_avg_sq} for every parameter that has been
        updated.  Before the first backward (nothing flattened yet) a loaded-but-not-yet-applied state is returned
        unchanged."""
        g = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        g["params"] = list(range(len(self._all_params())))
        if self.flat is None:
            state = dict(self._pending_state["state"]) if self._pending_state is not None else {}
            return {"state": state, "param_groups": [g]}
        state = dict(self._extra_state)
        if self.step_count > 0:
            for i, p, o in zip(self._index, self.flat.params, self.flat.offsets):
                n = p.numel()
                state[i] = {"step": torch.tensor(float(self.step_count)),
                            "exp_avg": self.exp_avg[o:o + n].view(p.shape).clone(),
                            "exp_avg_sq": self.exp_avg_sq[o:o + n].view(p.shape).clone()}
        return {"state": dict(sorted(state.items())), "param_groups": [g]}

    def load_state_dict(self, sd):
        """Accepts a ``torch.optim.AdamW`` (or own) state dict over the same parameter list.  May be called before the
        first backward — the reference resumes exactly like that (train_model/tav_train.py:162) — in which case the
        moments are applied when the flat buffers are built."""
        groups = sd["param_groups"]
        if len(groups) != 1:
            raise ValueError("FusedAdamW.load_state_dict: expected one parameter group, got %d" % len(groups))
        n_all = len(self._all_params())
        if len(groups[0]["params"]) != n_all:
            raise ValueError("loaded state dict contains a parameter group that doesn't match the size of optimizer's "
                             "group (%d vs %d parameters)" % (len(groups[0]["params"]), n_all))
        for k, v in groups[0].items():
            if k != "params":
                self.param_groups[0][k] = v
        remap = {saved: i for i, saved in enumerate(groups[0]["params"])}
        state = {remap[k]: v for k, v in sd["state"].items()}
        sd = {"state": state}
        if self.flat is None:
            self._pending_state = sd
            steps = [int(float(st["step"])) for st in state.values()]
            self.step_count = max(steps) if steps else 0
            return
        self._apply_state(sd)

    @torch.no_grad()
    def _apply_state(self, sd):
        state = sd["state"]
        slot = {i: k for k, i in enumerate(self._index)}
        self._extra_state = {}
        steps = []
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        for i, st in state.items():
            k = slot.get(i)
            if k is None:
                self._extra_state[i] = st      # a parameter without a gradient in this run: carried, not applied
                continue
            p, o = self.flat.params[k], self.flat.offsets[k]
            n = p.numel()
            if st["exp_avg"].numel() != n:
                raise ValueError("FusedAdamW.load_state_dict: state %d has %d elements, parameter has %d"
                                 % (i, st["exp_avg"].numel(), n))
            self.exp_avg[o:o + n].copy_(st["exp_avg"].reshape(-1))
            self.exp_avg_sq[o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
            steps.append(int(float(st["step"])))
        # torch keeps one step per parameter; they advance together here (every live parameter is updated every step)
        self.step_count = max(steps) if steps else 0
        self.step_dev.fill_(self.step_count)
