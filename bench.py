#!/usr/bin/env python
"""bench.py — TAV training-step throughput on N B200s of one node (contract: see DESIGN.md §Measurement).

  python bench.py --gpus 1 --steps 20 --warmup 5            # our arm (sm_100a kernels, CUDA-graphed step)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
         bench.py --gpus N --steps K --warmup W             # one rank per GPU, NCCL
  python bench.py --impl reference --steps 2 --warmup 1      # the reference's CPU path (oracle port) on host cores

A "step" = one pass of the hot path over one batch: PreFormer + TAVForMAE forward, weighted CE, backward, bucketed
gradient all-reduce (N>1), clip_grad_norm + AdamW (the `not_grad_accum` body, reference train_model/tav_train.py:56-65)
on the workload BASELINE.json's metric is quoted on (configs[1]: MELD 7-class, batch 16 per GPU, RoBERTa-base +
Wav2Vec2-base + VideoMAE-base, random init, synthetic inputs).  Prints ONE JSON line on rank 0."""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "TAV fwd+bwd samples/s"
UNIT = "samples/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cfg", default="C2")
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (default: the config's, 16 for C2)")
    ap.add_argument("--variant", default="baseline")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of the captured CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--recompute", action="store_true",
                    help="selective recompute: keep only layer inputs, re-run each layer's forward in backward (large batches)")
    ap.add_argument("--no-fusion-128", action="store_true", help="skip the extra 128-sample fusion-block measurement")
    ap.add_argument("--no-roofline", action="store_true", help="skip the per-signature GEMM roofline and fusion-block legs (sweeps)")
    ap.add_argument("--bucket-mb", type=int, default=128, help="gradient all-reduce bucket size (measured at 2 GPUs: 128 MB 43.6 ms/step, 32 MB 44.8)")
    ap.add_argument("--grad-reduce", default="fp32", choices=["fp32", "bf16"],
                    help="dtype of the gradient all-reduce (bf16 = opt-in compression; default exact fp32 sums)")
    ap.add_argument("--cpu-batch", type=int, default=2, help="samples per CPU-baseline step (bounded sample)")
    return ap.parse_args()


def workload_name(cfg, B, variant):
    from multi_modal_emotion_b200 import synthetic as syn

    c = syn.CONFIGS[cfg]
    shape = {"C1": "MELD", "C2": "MELD", "C3": "IEMOCAP-shape (15 s audio)", "C4": "MUStARD++-shape"}.get(cfg, cfg)
    return ("TAV %s %d-class train step (PreFormer+TAVForMAE fwd, weighted CE, bwd, clip+AdamW), batch %d/GPU, "
            "T=%d, wav=%d samples, 16x3x224x224 video, fused S=%d, encoders=%s (random init)" % (
                shape, c["C"], B, c["T"], c["L"], syn.fused_len(cfg), variant))


# ------------------------------------------------------------------------------------------------ reference arm (CPU)
def _unmodified_reference_run(args, steps, warmup, cores):
    """The UNMODIFIED reference modules (PreFormer, TAVForMAE, NewCrossEntropyLoss) imported from where the reference tree
    lies — baseline/_ref when someone dropped it there, the read-only checkout of the authoring container — on the same
    synthetic weights and inputs.  Returns None when the tree is absent (the GPU box) or cannot be imported."""
    import torch

    from multi_modal_emotion_b200 import synthetic as syn
    from oracle import ref_loader

    if not ref_loader.available() or args.variant not in ("baseline", "reference", "tiny"):
        return None
    try:
        ns = ref_loader.load_reference(args.variant)
        B = args.cpu_batch
        t0 = time.time()
        C = syn.CONFIGS[args.cfg]["C"]
        torch.manual_seed(0)
        model = ns.TAVForMAE({"output_dim": C, "dropout": 0.4, "learn_PosEmbeddings": True, "num_layers": 12})
        pre = ns.PreFormer()
        pre.load_state_dict(syn.synth_state_dict(pre, seed=1))
        model.load_state_dict(syn.synth_state_dict(model, seed=2))
        inputs, labels = syn.make_batch(args.cfg, B=B)
        w = torch.tensor(syn.MELD_CLASS_WEIGHTS if C == 7 else [0.5, 0.5])
        crit = ns.NewCrossEntropyLoss(class_weights=w, epoch_switch=2)
        params = [p for p in model.parameters() if p.requires_grad] + [p for p in pre.parameters() if p.requires_grad]
        opt = torch.optim.AdamW(params, lr=1e-5, weight_decay=1e-4)
        ids, tm = inputs[0]["input_ids"], inputs[0]["attention_mask"]
        wav, am = inputs[1]["audio_features"], inputs[1]["attention_mask"]
        video, vm = inputs[2]["visual_embeds"], inputs[2]["attention_mask"]
        build_s = time.time() - t0

        def step():     # the body of train_model/tav_train.py:15-65 (get_statistics, backward, clip, AdamW)
            t, pos, mask = pre(input_ids=ids, audio_features=wav, video_embeds=video, text_mask=tm, audio_mask=am,
                               visual_mask=vm, device="cpu", train=False)
            logits = model(input_ids=ids, text_attention_mask=tm, audio_features=wav, video_embeds=video, visual_mask=vm,
                           hidden_states=t, pos_embed=pos, attention_mask=mask, batch_size=B, check="val")
            loss = crit(logits, labels.long(), epoch=1)
            loss.backward()
            torch.nn.utils.clip_grad_norm_([p for p in params if p.grad is not None], 1.0)
            opt.step()
            opt.zero_grad(set_to_none=True)
            return loss.item()

        import contextlib

        with contextlib.redirect_stdout(sys.stderr):      # the reference prints progress lines; stdout carries the JSON line
            for _ in range(warmup):
                step()
            t0 = time.time()
            for _ in range(steps):
                step()
            dt = (time.time() - t0) / max(steps, 1)
    except Exception as e:  # noqa: BLE001  (an importable-but-broken drop must not take the bench down: fall back to the port)
        sys.stderr.write("reference arm: unmodified reference at %s not usable (%s); using the oracle port\n" % (ref_loader.REF_ROOT, e))
        return None
    return {"value": B / dt, "unit": UNIT, "cores": cores, "kind": "reference",
            "sample": "UNMODIFIED reference modules from %s: %d-sample batch of the same workload shape, %d timed step(s) after %d "
                      "warm-up, fp32, torch %s CPU ops, %d threads (model build %.0fs not timed)" % (
                          ref_loader.REF_ROOT, B, steps, warmup, torch.__version__, cores, build_s),
            "ms_per_step": dt * 1e3}


def cpu_reference_run(args, steps, warmup, quiet=False):
    """The reference's own CPU implementation of the path: the unmodified reference when its tree is present (see
    _unmodified_reference_run), else the oracle port (the reference is Python and its checkout does not travel to
    the GPU box); fp32, all host threads, on a bounded sample of the workload (args.cpu_batch samples per step — the
    reference at the bench's 16 samples needs ~45 GB of host memory for autograd and minutes per step)."""
    import torch

    from multi_modal_emotion_b200 import synthetic as syn, tav
    from oracle import tav_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    r = _unmodified_reference_run(args, steps, warmup, cores)
    if r is not None:
        return r
    tav.set_encoder_variant(args.variant)
    B = args.cpu_batch
    t0 = time.time()
    model = tav.TAVForMAE({"output_dim": syn.CONFIGS[args.cfg]["C"], "dropout": 0.4, "learn_PosEmbeddings": True, "num_layers": 12})
    pre = tav.PreFormer()
    pre_sd, tav_sd = syn.synth_state_dict(pre, seed=1), syn.synth_state_dict(model, seed=2)
    del model, pre
    orc = O.OracleTAV(tav.encoder_configs(args.variant)).load(pre_sd, tav_sd)
    del pre_sd, tav_sd
    params = [p for m in list(orc.pre.values()) + list(orc.tav.values()) for p in m.parameters()] + \
        list(orc.pre_w.values()) + list(orc.tav_w.values())
    opt = torch.optim.AdamW(params, lr=1e-5, weight_decay=1e-4)
    inputs, labels = syn.make_batch(args.cfg, B=B)
    w = torch.tensor(syn.MELD_CLASS_WEIGHTS if syn.CONFIGS[args.cfg]["C"] == 7 else [0.5, 0.5])
    build_s = time.time() - t0

    def step():
        logits = orc.forward(inputs)
        loss = O.new_cross_entropy(logits, labels.long(), 1, w, 2)
        loss.backward()
        torch.nn.utils.clip_grad_norm_([p for p in params if p.grad is not None], 1.0)
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss.item()

    for _ in range(warmup):
        step()
    t0 = time.time()
    for _ in range(steps):
        step()
    dt = (time.time() - t0) / max(steps, 1)
    return {"value": B / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d-sample batch of the same workload shape, %d timed step(s) after %d warm-up, fp32, torch %s CPU ops, "
                      "%d threads (model build %.0fs not timed)" % (B, steps, warmup, torch.__version__, cores, build_s),
            "ms_per_step": dt * 1e3}


def torch_gpu_baseline(args, B):
    """Informational (SURVEY 0 / 8c "beat eager PyTorch on the same B200"): the plain-PyTorch restatement of the reference
    (oracle/) moved to the same GPU — eager ATen / cuBLAS / SDPA kernels — in fp32 and under bf16 autocast, one
    fwd + weighted CE + bwd + clip + AdamW step at the bench's own batch.  Not the product path, never the thing shipped."""
    import torch

    from multi_modal_emotion_b200 import synthetic as syn, tav
    from oracle import tav_oracle as O

    out = {}
    try:
        tav.set_encoder_variant(args.variant)
        C = syn.CONFIGS[args.cfg]["C"]
        model = tav.TAVForMAE({"output_dim": C, "dropout": 0.4, "learn_PosEmbeddings": True, "num_layers": 12})
        pre = tav.PreFormer()
        pre_sd, tav_sd = syn.synth_state_dict(pre, seed=1), syn.synth_state_dict(model, seed=2)
        del model, pre
        orc = O.OracleTAV(tav.encoder_configs(args.variant)).load(pre_sd, tav_sd).to("cuda")
        params = [p for m in list(orc.pre.values()) + list(orc.tav.values()) for p in m.parameters()] + \
            list(orc.pre_w.values()) + list(orc.tav_w.values())
        opt = torch.optim.AdamW(params, lr=1e-5, weight_decay=1e-4)
        inputs, labels = syn.make_batch(args.cfg, B=B)
        inputs = [{k: v.cuda() for k, v in d.items()} for d in inputs]
        labels = labels.cuda().long()
        w = torch.tensor(syn.MELD_CLASS_WEIGHTS if C == 7 else [0.5, 0.5]).cuda()
        for name, ctx in (("bf16_autocast", lambda: torch.autocast("cuda", dtype=torch.bfloat16)), ("fp32", None)):
            def step():
                if ctx is None:
                    logits = orc.forward(inputs)
                else:
                    with ctx():
                        logits = orc.forward(inputs)
                loss = O.new_cross_entropy(logits.float(), labels, 1, w, 2)
                loss.backward()
                torch.nn.utils.clip_grad_norm_([p for p in params if p.grad is not None], 1.0)
                opt.step()
                opt.zero_grad(set_to_none=True)

            for _ in range(3):
                step()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            n = 5
            for _ in range(n):
                step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            out[name] = {"ms_per_step": ms, "samples_per_s": B / ms * 1e3}
        out["what"] = ("oracle restatement of the reference on the same GPU, eager PyTorch %s (ATen/cuBLAS/SDPA), batch %d, "
                       "fwd + weighted CE + bwd + clip + AdamW, CUDA events, 3 warm-up + 5 timed steps" % (torch.__version__, B))
        del orc, params, opt
        torch.cuda.empty_cache()
    except Exception as e:  # noqa: BLE001  (informational leg: never fails the bench)
        out["error"] = str(e)[:300]
    return out


def reference_main(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from multi_modal_emotion_b200 import synthetic as syn  # noqa: F401

    steps, warmup = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
    r = cpu_reference_run(args, steps, warmup)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.cfg, args.cpu_batch, args.variant), "device": "host CPU",
                       "requested_steps": args.steps, "requested_warmup": args.warmup},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(power) if power else None}


# ------------------------------------------------------------------------------------------------ our arm
def main():
    args = parse()
    if args.impl == "reference":
        return reference_main(args)

    import torch
    import torch.distributed as dist

    from multi_modal_emotion_b200 import _lib as L, dp, synthetic as syn, tav
    from multi_modal_emotion_b200.losses import NewCrossEntropyLoss
    from multi_modal_emotion_b200.optim import FusedAdamW

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # one rank per GPU, each on its own slice of the host cores: the ranks' Python threads (graph launches, the
        # per-step loss read-back) otherwise share one affinity mask and migrate across each other
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = max(1, len(cores) // world)
            os.sched_setaffinity(0, set(cores[local * per:(local + 1) * per]) or set(cores))
        except (AttributeError, OSError):
            pass
        # Optional (TAVK_COMM_SMS=n): give NCCL a fixed CTA budget and keep as many SMs out of the persistent GEMM's grid,
        # either for every GEMM (TAVK_COMM_SMS_SCOPE=all: measured at 2 GPUs it costs more than it hides, 649 vs 676
        # samples/s with n=8) or only while backward runs and gradient all-reduces are in flight (default scope).
        comm_sms = int(os.environ.get("TAVK_COMM_SMS", "0"))
        comm_scope = os.environ.get("TAVK_COMM_SMS_SCOPE", "backward")
        if comm_sms > 0:
            os.environ.setdefault("NCCL_MAX_CTAS", str(comm_sms))
            os.environ.setdefault("NCCL_MIN_CTAS", str(min(4, comm_sms)))
        dist.init_process_group("nccl", device_id=dev)
        if comm_sms > 0 and comm_scope == "all":
            L.reserve_sms(comm_sms)
    L.require_device()   # fails loudly when the CUDA extension / an sm_100 device is missing: no fallback
    if args.no_fusion_128:
        os.environ["TAVK_BENCH_NO_FUSION128"] = "1"
    if args.recompute:
        from multi_modal_emotion_b200 import engine

        engine.recompute_layers = True

    peaks = {"bf16_tflops": 1590.0, "hbm_gbs": 6650.0, "source": "fallback"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            pj = json.load(f)
        peaks = {"bf16_tflops": pj["bf16_tflops"], "bf16_tflops_sustained": pj.get("bf16_tflops_sustained"),
                 "hbm_gbs": pj["hbm_gbs"], "source": "measured"}
    except Exception:  # noqa: BLE001
        pass

    cfg = args.cfg
    B = args.batch or syn.CONFIGS[cfg]["B"]
    C = syn.CONFIGS[cfg]["C"]
    tav.set_encoder_variant(args.variant)
    torch.manual_seed(0)
    model = tav.TAVForMAE({"output_dim": C, "dropout": 0.4, "learn_PosEmbeddings": True, "num_layers": 12})
    pre = tav.PreFormer()
    pre.load_state_dict(syn.synth_state_dict(pre, seed=1))
    model.load_state_dict(syn.synth_state_dict(model, seed=2))
    model, pre = model.to(dev).train(), pre.to(dev).train()
    crit = NewCrossEntropyLoss(torch.tensor(syn.MELD_CLASS_WEIGHTS if C == 7 else [0.5, 0.5]), epoch_switch=2)
    params = [p for p in model.parameters() if p.requires_grad] + [p for p in pre.parameters() if p.requires_grad]
    opt = FusedAdamW(params, lr=1e-5, weight_decay=1e-4)
    runner = dp.DataParallelTAV(model, pre, crit, opt, clip=1.0, bucket_mb=args.bucket_mb, use_cuda_graph=not args.no_graph,
                                grad_reduce_dtype=torch.bfloat16 if args.grad_reduce == "bf16" else None,
                                row_sparse=os.environ.get("TAVK_ROW_SPARSE", "1") != "0",
                                comm_sms=int(os.environ.get("TAVK_COMM_SMS", "0")) if (world > 1 and os.environ.get("TAVK_COMM_SMS_SCOPE", "backward") == "backward") else 0)

    host_inputs, host_labels = syn.make_batch(cfg, seed=1234 + rank, B=B)
    host_inputs = [{k: v.pin_memory() for k, v in d.items()} for d in host_inputs]
    host_labels = host_labels.pin_memory()
    h2d = sum(v.numel() * v.element_size() for d in host_inputs for v in d.values()) + host_labels.numel() * host_labels.element_size()

    # one eager step first: materialises the flat optimiser buffers / gradient buckets and counts our launches per step
    k0, c0 = L.kernel_count, L.launch_count
    L.gemm_log.clear()
    L.record_gemms = True
    runner._eager_step(host_inputs, host_labels, 1, "train")
    L.record_gemms = False
    torch.cuda.synchronize()
    kernels_per_step, calls_per_step = L.kernel_count - k0, L.launch_count - c0

    # ---- kernel-only throughput: inputs already resident in HBM
    dev_inputs = [{k: v.to(dev) for k, v in d.items()} for d in host_inputs]
    dev_labels = host_labels.to(dev)
    runner.train_step(dev_inputs, dev_labels, 1, "train")   # builds + captures the graph (its own warm-up inside)
    if runner.static_inputs() is not None:
        dev_inputs, dev_labels = runner.static_inputs()      # write-in-place buffers: no copies in the timed region
    for _ in range(max(args.warmup, 3)):
        runner.train_step(dev_inputs, dev_labels, 1, "train")
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        loss = runner.train_step(dev_inputs, dev_labels, 1, "train")
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = t.item()
    value = B * world / ms_max * 1e3
    last_loss = float(loss.item())

    # ---- end to end through the public call with HOST buffers: every step's inputs cross PCIe from pinned host memory
    # (the copy of batch i+1 is started by the call for batch i and overlaps its compute — what a DataLoader with
    # pin_memory + non_blocking copies gives) and the step's loss is read back to the host (a sync) every step.
    # Two host batches alternate so that consecutive steps really move different buffers.
    host_b = (host_inputs, host_labels)
    alt_in, alt_lab = syn.make_batch(cfg, seed=4321 + rank, B=B)
    host_c = ([{k: v.pin_memory() for k, v in d.items()} for d in alt_in], alt_lab.pin_memory())
    seq = [host_b, host_c]
    for i in range(2):
        runner.train_step(seq[i % 2][0], seq[i % 2][1], 1, "train", next_batch=seq[(i + 1) % 2]).item()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0.record()
    for i in range(args.steps):
        cur, nxt = seq[i % 2], seq[(i + 1) % 2]
        runner.train_step(cur[0], cur[1], 1, "train", next_batch=nxt).item()
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1) / args.steps
    t = torch.tensor([ms_e2e], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e2e = t.item()

    line = None
    if rank == 0:
        roof, fusion = (None, None) if args.no_roofline else roofline_and_fusion(torch, L, syn, model, B, cfg, ms_max, peaks)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_max, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16 (tensor-core operands; fp32 accumulate, residual stream, LayerNorm/softmax statistics, optimiser)",
            "data": "synthetic",
            "config": {"workload": workload_name(cfg, B, args.variant), "global_batch": B * world,
                       "parallelism": "dp%d" % world, "cuda_graph": not args.no_graph,
                       "l2_policy": "inputs (%.0f MB/step) and saved activations exceed the 126 MB L2; no explicit flush" % (h2d / 1e6),
                       "step": "fwd + loss + bwd + grad all-reduce + clip + AdamW",
                       "activation_recompute": bool(args.recompute), "grad_all_reduce": "%s, %d MB buckets" % (args.grad_reduce, args.bucket_mb)},
            "e2e": {"value": B * world / ms_e2e * 1e3, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e,
                    "how": "train_step(host batch, next_batch=...): pinned-host H2D of every batch (copy of batch i+1 "
                           "overlaps step i on a copy stream), graph replay, loss.item() every step"},
            "gpu_launches": kernels_per_step * args.steps,
            "library_calls_per_step": calls_per_step,
            "clocks": clocks, "roofline": roof, "fusion_block": fusion, "loss": last_loss,
            "peaks": peaks,
        }
        if world == 1 and not args.no_cpu_baseline:
            del runner, model, pre, opt
            torch.cuda.empty_cache()
            line["torch_gpu"] = torch_gpu_baseline(args, B)
            r = cpu_reference_run(args, 1, 1)
            line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    if world > 1:
        # Tear down without ncclCommDestroy: the communicator is referenced by the captured step graphs (the bucketed
        # all-reduces are graph nodes) and NCCL defers / blocks communicator destruction while such graphs are alive,
        # which hung the 2-GPU run at exit.  Everything is flushed, every rank is past the last collective: leave.
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    return 0


def roofline_and_fusion(torch, L, syn, model, B, cfg, step_ms, peaks):
    """Dominant kernel = the tcgen05 GEMM.  Every GEMM launch of one training step was recorded (shape, majors,
    epilogue) during the eager step; each distinct signature is re-launched back to back on the same stream and timed
    with CUDA events -> average launch duration per signature; achieved = sum(2MNK) / sum(count x duration)."""
    import collections

    sigs = collections.Counter(L.gemm_log)
    total_flops, total_ms, table = 0.0, 0.0, []
    for sig, count in sigs.items():
        M, N, K, a_mn, b_mn, epi, out_bf16, has_bias, has_resid, has_rb, acc, ks, groups = sig
        if groups > 1:
            continue   # the grouped implicit-GEMM convolutions (6 launches per step) are not re-created here
        A = torch.randn((K, M) if a_mn else (M, K), device="cuda").bfloat16()
        Bm = torch.randn((K, N) if b_mn else (N, K), device="cuda").bfloat16()
        out = torch.zeros((M, N), device="cuda", dtype=torch.bfloat16 if out_bf16 else torch.float32)
        kw = dict(M=M, N=N, K=K, a_mn=bool(a_mn), b_mn=bool(b_mn), epilogue=epi, accumulate=bool(acc), k_splits=ks)
        if has_bias:
            kw["bias"] = torch.zeros(N, device="cuda")
        if has_resid:
            kw["resid"] = torch.zeros((M, N), device="cuda")
        if has_rb:
            kw["rowbias"], kw["rows_per_group"] = torch.zeros((B, N), device="cuda"), max(1, M // B)
        if epi in (L.EPI_GELU, L.EPI_GELU_GRAD):
            kw["out2"] = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
        if epi in (L.EPI_GELU_BWD, L.EPI_MUL):
            kw["aux"] = torch.zeros((M, N), device="cuda", dtype=torch.bfloat16)
        L.record_gemms = False
        # device-side duration: `reps` launches captured in a CUDA graph (eager re-launches of the small shapes are paced
        # by the host's submission rate, ~15 us per call, not by the kernel), timed with CUDA events around replays
        reps = 10
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                L.gemm(A, Bm, out, **kw)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                L.gemm(A, Bm, out, **kw)
        g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / (2 * reps)
        del g
        total_ms += count * ms
        total_flops += count * 2.0 * M * N * K
        table.append((count * ms, count, ms, 2.0 * M * N * K / ms / 1e9, sig))
    if os.environ.get("TAVK_GEMM_TABLE"):
        with open(os.environ["TAVK_GEMM_TABLE"], "w") as f:
            f.write("total_ms count ms_each TFLOP/s (M,N,K,a_mn,b_mn,epi,out_bf16,bias,resid,rowbias,acc,k_splits,groups)\n")
            for row in sorted(table, reverse=True):
                f.write("%8.3f %4d %8.4f %7.1f %s\n" % row)
    achieved = total_flops / (total_ms * 1e-3) / 1e12 if total_ms > 0 else 0.0
    peak = peaks["bf16_tflops"]
    traffic, traffic_of = None, None
    try:   # DRAM bytes per launch of the heaviest signature, from the committed `ncu --set full` capture (profiles/)
        tpath = os.path.join(ROOT, "profiles", "r2_gemm_ncu_traffic.json")
        if not os.path.exists(tpath):
            tpath = os.path.join(ROOT, "profiles", "r1_gemm_ncu_traffic.json")
        with open(tpath) as f:
            tj = json.load(f)
        traffic, traffic_of = tj["traffic_bytes"], {k: tj[k] for k in ("signature", "algorithmic_bytes", "source")}
    except Exception:  # noqa: BLE001
        pass
    roof = {"bound": "tensor", "kernel": "gemm_bf16_tcgen05_kernel", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
            "frac": achieved / peak, "traffic": traffic, "traffic_of": traffic_of, "peak_source": peaks["source"] + " cuBLAS bf16 burst (MEASURED_PEAKS.json)",
            "launches_per_step": sum(sigs.values()), "distinct_shapes": len(sigs), "flops_per_step": total_flops,
            "avg_launch_ms": total_ms / max(1, sum(sigs.values())), "share_of_step": total_ms / step_ms,
            "method": "each distinct GEMM signature of the step: 10 launches captured in a CUDA graph, replayed twice between CUDA events on the launching stream (same operands every launch: L2-warm for the small shapes)"}
    # fusion block alone: 12-layer VideoMAEEncoder fwd+bwd at the workload's fused length, reference-faithful masks; at
    # the workload's batch and (BASELINE configs[4], the batch sweep) at 128 samples per GPU, where M = B*S fills the SMs
    S = syn.fused_len(cfg)
    c = syn.CONFIGS[cfg]
    Ta = syn.conv_frames(c["L"])
    enc = model.random_mae_encoder

    def fusion_at(Bf):
        x = torch.randn(Bf, S, 768, device="cuda", requires_grad=True)
        mask = syn.reference_masks(Bf, c["T"], Ta, c["K"], torch.full((Bf,), c["T"]), torch.full((Bf,), Ta)).cuda()
        go = torch.full((Bf, S, 768), 1.0 / (Bf * S * 768), device="cuda")

        def fstep():
            y = enc(x, mask)
            y.backward(go)

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                fstep()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fstep()
        for _ in range(3):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        fms = e0.elapsed_time(e1) / 10
        fl = 3 * 12 * (14155776 * S + 3072 * S * S) * Bf
        del g
        return {"ms_fwd_bwd": fms, "tflops": fl / fms / 1e9, "frac_of_bf16_peak": fl / fms / 1e9 / peak, "B": Bf, "S": S,
                "flops": fl, "mask_regime": "R (reference PreFormer masks)"}

    fusion = fusion_at(B)
    if os.environ.get("TAVK_BENCH_NO_FUSION128"):
        return roof, fusion
    try:
        fusion["batch_128"] = fusion_at(128)
    except Exception as e:  # noqa: BLE001  (e.g. out of memory on a smaller part): the B-sized figure stands
        fusion["batch_128"] = {"error": str(e)[:200]}
    return roof, fusion


if __name__ == "__main__":
    sys.exit(main())
