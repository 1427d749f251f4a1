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
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_grad_norm=None):
        params = list(params)
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        if len(self.param_groups) != 1:
            raise ValueError("FusedAdamW supports a single parameter group (as the reference loop uses)")
        self.flat = None
        self.step_count = 0
        self.max_grad_norm = max_grad_norm
        self.grad_prescale = 1.0  # e.g. 1/world_size after a summing all-reduce
        self.on_materialize = None  # dp.py hooks bucket construction here

    def materialize(self):
        """Flatten at the first step: like torch.optim.AdamW, parameters that never receive a gradient (e.g. the
        unused halves of PreFormer's encoders, ``masked_spec_embed``) are left untouched — not even decayed."""
        if self.flat is not None:
            return
        live = [p for p in self.param_groups[0]["params"] if p.requires_grad and p.grad is not None]
        grads = [p.grad for p in live]
        self.flat = FlatParams(live)
        with torch.no_grad():
            for p, g0 in zip(self.flat.params, grads):
                p.grad.copy_(g0)
        n = self.flat.numel
        dev = self.flat.flat.device
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.sqnorm = torch.zeros(1, dtype=torch.float32, device=dev)
        if self.on_materialize is not None:
            self.on_materialize(self.flat)

    @torch.no_grad()
    def step(self, closure=None, max_grad_norm=None):
        """One fused update.  ``max_grad_norm`` (or the constructor's) applies clip_grad_norm_ semantics:
        g *= min(1, max_norm / (||g||_2 + 1e-6)) over ALL parameters jointly, computed on device (no host sync)."""
        if closure is not None:
            raise NotImplementedError("closures are not supported")
        self.materialize()
        self.flat.attach_grads()
        g = self.param_groups[0]
        clip = max_grad_norm if max_grad_norm is not None else self.max_grad_norm
        self.step_count += 1
        sq_ptr = None
        if clip is not None and clip > 0:
            self.sqnorm.zero_()
            L.call("tavk_grad_sqnorm", self.flat.grad.data_ptr(), self.flat.numel, self.sqnorm.data_ptr())
            sq_ptr = self.sqnorm.data_ptr()
        L.call("tavk_adamw", self.flat.flat.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
               self.flat.grad.data_ptr(), self.flat.shadow.data_ptr(), self.flat.numel, float(g["lr"]), float(g["betas"][0]),
               float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]), self.step_count, sq_ptr,
               float(clip) if clip else 0.0, float(self.grad_prescale), 1)
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

    def state_dict(self):
        """torch.optim.AdamW-compatible layout: per-parameter exp_avg / exp_avg_sq / step."""
        self.materialize()
        state = {}
        for i, (p, o) in enumerate(zip(self.flat.params, self.flat.offsets)):
            n = p.numel()
            state[i] = {"step": torch.tensor(float(self.step_count)), "exp_avg": self.exp_avg[o:o + n].view(p.shape).clone(),
                        "exp_avg_sq": self.exp_avg_sq[o:o + n].view(p.shape).clone()}
        g = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        g["params"] = list(range(len(self.flat.params)))
        return {"state": state, "param_groups": [g]}

    def load_state_dict(self, sd):
        self.materialize()
        for i, (p, o) in enumerate(zip(self.flat.params, self.flat.offsets)):
            st = sd["state"].get(i)
            if st is None:
                continue
            n = p.numel()
            self.exp_avg[o:o + n].copy_(st["exp_avg"].reshape(-1))
            self.exp_avg_sq[o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
            self.step_count = int(float(st["step"]))
        for k, v in sd["param_groups"][0].items():
            if k != "params":
                self.param_groups[0][k] = v
