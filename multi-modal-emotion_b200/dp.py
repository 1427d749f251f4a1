"""Data-parallel execution of the TAV step: one process per GPU, NCCL over NVLink 5 / NVSwitch.

The reference has no distributed code at all — the batch shards trivially because every op on the
path is per-sample (§8e).  Ranks couple through exactly three quantities, each handled so that N ranks reproduce the
single-process step on the concatenated batch:
  1. parameter gradients  -> bucketed all-reduce(SUM) of contiguous ranges of the flat gradient buffer
     (optim.FlatParams), launched from post-accumulate-grad hooks on a side stream while backward is still running;
  2. the (weighted) CE normaliser sum_i w[y_i] -> one scalar all-reduce in the forward pass; every rank then
     back-propagates  num_local / den_global  so the summed gradients are the gradients of the global loss;
  3. clip_grad_norm_ -> computed after the reduction, when every rank holds the same gradient: no extra collective.
The bucket plumbing is device-agnostic (it is exercised on CPU with gloo, world_size 2, in tests/test_dp_cpu.py)."""
import torch
import torch.distributed as dist


def is_distributed():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def global_loss(num, den, group=None):
    """(loss to back-propagate on this rank, global loss value).  Gradients must then be SUM-reduced."""
    if not is_distributed():
        return num / den, (num / den).detach()
    den_g = den.detach().clone()
    dist.all_reduce(den_g, op=dist.ReduceOp.SUM, group=group)
    num_g = num.detach().clone()
    dist.all_reduce(num_g, op=dist.ReduceOp.SUM, group=group)
    return num / den_g, num_g / den_g


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def _scatter_rows(grad, idx, dy, skip_idx):
    """grad[idx[r], :] += dy[r, :] (rows with idx == skip_idx or outside the table are dropped): the embedding-backward
    kernel on the GPU; index_add_ for the CPU (gloo) tests of the plumbing."""
    if grad.is_cuda:
        from . import _lib as L

        L.call("tavk_embedding_scatter_add", dy.data_ptr(), idx.data_ptr(), grad.data_ptr(), idx.numel(), dy.shape[-1],
               grad.shape[0], -1 if skip_idx is None else int(skip_idx))
        return
    keep = (idx >= 0) & (idx < grad.shape[0])
    if skip_idx is not None and skip_idx >= 0:
        keep &= idx != skip_idx
    grad.index_add_(0, idx[keep], dy[keep])


class GradBuckets:
    """Splits a flat gradient buffer into parameter-aligned buckets and all-reduces each one as soon as every
    parameter in it has accumulated its gradient."""

    def __init__(self, flat, bucket_bytes=32 << 20, group=None, reduce_dtype=None, row_sparse=()):
        """``reduce_dtype=torch.bfloat16``: opt-in gradient compression — each bucket is cast to bf16 on the communication
        stream, summed by NCCL in bf16 and cast back (half the NVLink bytes and half the time the collective kernels hold
        SMs; the summed gradient carries bf16 rounding, 2^-9 relative).  Default: exact fp32 sums.

        ``row_sparse``: embedding tables whose gradient has at most (tokens per rank) non-zero rows — RoBERTa's two
        50265 x 768 word-embedding tables are 2 x 154 MB of the 1.58 GB gradient, produced last in backward, with 1120
        non-zero rows each.  They are left out of the dense buckets; their producer hands (token ids, gradient rows) to
        ``exchange_rows`` and every rank scatter-adds the other ranks' rows into its own table: 3.4 MB per rank on the
        wire instead of a 154 MB all-reduce at the very end of backward, and the same sum."""
        self.flat, self.group = flat, group
        self.reduce_dtype = reduce_dtype
        self.buckets = []   # (start, end, n_params)
        self.param_bucket = {}
        self.row_sparse = {id(p) for p in row_sparse}
        cap = max(1, bucket_bytes // 4)
        # backward produces gradients roughly in reverse registration order: build buckets from the tail; a bucket is a
        # contiguous range of the flat buffer, so a row-sparse table closes the bucket on either side of it
        i = len(flat.params) - 1
        while i >= 0:
            if id(flat.params[i]) in self.row_sparse:
                i -= 1
                continue
            end = flat.offsets[i] + (flat.params[i].numel() + 63) // 64 * 64
            if i == len(flat.params) - 1:
                end = flat.numel
            j = i
            while j > 0 and id(flat.params[j - 1]) not in self.row_sparse and end - flat.offsets[j - 1] <= cap:
                j -= 1
            start = flat.offsets[j]
            self.buckets.append((start, end, i - j + 1))
            for k in range(j, i + 1):
                self.param_bucket[id(flat.params[k])] = len(self.buckets) - 1
            i = j - 1
        self.pending = [0] * len(self.buckets)
        self.reported = set()
        self.streams = [set() for _ in self.buckets]
        self.works = []
        self.enabled = False
        self.cuda = flat.grad.is_cuda
        self.comm_stream = torch.cuda.Stream() if self.cuda else None
        self.handles = [p.register_post_accumulate_grad_hook(self._hook) for p in flat.params]
        self.launched = 0

    def start_backward(self):
        """Arm the buckets for one backward pass."""
        from . import engine

        self.pending = [n for (_, _, n) in self.buckets]
        self.reported = set()                           # parameters already counted in this backward pass (see _hook)
        self.streams = [set() for _ in self.buckets]    # streams that wrote gradients of each bucket (tav.branch_streams)
        self.works = []
        self.enabled = True
        self.launched = 0
        # parameters whose gradient the layer engine accumulates straight into the flat buffer are reported by the engine
        # layer by layer (autograd only sees a None gradient for them, after the whole stack's backward has returned)
        engine.grad_written_hook = self._written
        engine.row_sparse_hook = self.exchange_rows if self.row_sparse else None

    def exchange_rows(self, table, idx, dy, grad, skip_idx):
        """Called by an embedding backward after it has scatter-added this rank's rows ``dy`` [rows, H] (token ids ``idx``)
        into ``grad`` (the table's gradient): adds every other rank's rows too.  Returns False when ``table`` is not one
        of the row-sparse tables (the caller's gradient then travels in a dense bucket)."""
        if not self.enabled or id(table) not in self.row_sparse:
            return False
        world = dist.get_world_size(self.group)
        rank = dist.get_rank(self.group)
        rows, H = idx.numel(), dy.shape[-1]
        idx, dy = idx.reshape(rows).contiguous(), dy.reshape(rows, H).contiguous()
        cur = torch.cuda.current_stream() if self.cuda else None
        if self.cuda:
            self.comm_stream.wait_stream(cur)
            idx.record_stream(self.comm_stream)
            dy.record_stream(self.comm_stream)
        ctx = torch.cuda.stream(self.comm_stream) if self.cuda else _NullCtx()
        with ctx:
            ids_all = torch.empty((world, rows), dtype=idx.dtype, device=idx.device)
            dy_all = torch.empty((world, rows, H), dtype=dy.dtype, device=dy.device)
            dist.all_gather([ids_all[r] for r in range(world)], idx, group=self.group)     # (views of one buffer: no copies)
            dist.all_gather([dy_all[r] for r in range(world)], dy, group=self.group)
            for r in range(world):
                if r != rank:
                    _scatter_rows(grad, ids_all[r], dy_all[r], skip_idx)
        self.launched += 1
        return True

    def _written(self, params):
        for p in params:
            if id(p) in self.param_bucket:
                self._hook(p)

    def _launch(self, b):
        s, e, _ = self.buckets[b]
        view = self.flat.grad[s:e]
        if self.cuda:
            self.comm_stream.wait_stream(torch.cuda.current_stream())
            for st in self.streams[b]:
                self.comm_stream.wait_stream(st)     # a bucket may hold parameters of several concurrently running branches
            with torch.cuda.stream(self.comm_stream):
                if self.reduce_dtype is not None and self.reduce_dtype != view.dtype:
                    tmp = view.to(self.reduce_dtype)
                    dist.all_reduce(tmp, op=dist.ReduceOp.SUM, group=self.group)
                    view.copy_(tmp)
                    tmp.record_stream(self.comm_stream)
                else:
                    self.works.append(dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        else:
            self.works.append(dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        self.launched += 1

    def _hook(self, p):
        """A parameter's gradient for this backward pass is complete (its kernels are enqueued on the current stream).
        Counted ONCE per pass: a parameter handled by the engine's gradient sink is reported by the engine when its layer
        is done and then again by autograd — torch >= 2.x fires the post-accumulate-grad hook even for the None gradient
        the stack's backward returns for it (it did not when this was first written).  Counting both let a bucket that
        mixes such parameters with ones written later (another branch stream) reach zero early: it was all-reduced
        before those gradients existed and they stayed rank-local (tests/test_dp_nccl_gpu.py, tests/test_dp_cpu.py)."""
        if not self.enabled or id(p) in self.reported or id(p) not in self.param_bucket:
            return
        self.reported.add(id(p))
        b = self.param_bucket[id(p)]
        if self.cuda:
            self.streams[b].add(torch.cuda.current_stream())
        self.pending[b] -= 1
        if self.pending[b] == 0:
            self._launch(b)

    def finish(self):
        """Reduce whatever did not fire from a hook (parameters without a gradient this step) and wait."""
        from . import engine

        engine.grad_written_hook = None
        engine.row_sparse_hook = None
        if self.enabled:
            for b, n in enumerate(self.pending):
                if n > 0:
                    self.pending[b] = 0
                    self._launch(b)
        for w in self.works:
            w.wait()
        if self.cuda:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        self.works = []
        self.enabled = False

    def remove(self):
        for h in self.handles:
            h.remove()


class DataParallelTAV:
    """Drives one data-parallel training step of (PreFormer, TAVForMAE): forward on the local shard, global-loss
    normalisation, backward overlapped with bucketed gradient all-reduce, fused clip + AdamW on identical gradients."""

    def __init__(self, model, PREFormer, criterion, optimizer, clip=1.0, bucket_mb=128, group=None,
                 use_cuda_graph=False, graph_warmup=3, scheduler=None, grad_reduce_dtype=None, comm_sms=0, row_sparse=True):
        """``scheduler``: the reference's CosineAnnealingWarmRestarts (or any torch scheduler over ``optimizer``).
        ``train_step(..., sched_t=epoch + i/iters)`` steps it after the update exactly like the reference loop
        (train_model/tav_train.py:63); its learning rate reaches the captured graph through a device scalar."""
        self.model, self.pre, self.criterion, self.opt = model, PREFormer, criterion, optimizer
        self.scheduler = scheduler
        self.grad_reduce_dtype = grad_reduce_dtype
        # comm_sms > 0 (N > 1): while backward runs — the only time gradient all-reduces are in flight — every persistent
        # GEMM leaves that many SMs out of its grid (tavk_gemm_args.max_ctas), so its CTAs never queue behind the
        # collective kernels that hold those SMs; forward GEMMs keep the whole machine.  Pair it with NCCL_MAX_CTAS.
        self.comm_sms = int(comm_sms)
        self.row_sparse = bool(row_sparse)      # exchange large embedding-table gradients as (ids, rows), see GradBuckets
        self.clip, self.group, self.bucket_bytes = clip, group, bucket_mb << 20
        self.buckets = None
        self.use_cuda_graph, self.graph_warmup = use_cuda_graph, graph_warmup
        self._graph = None
        self.world = dist.get_world_size(group) if is_distributed() else 1
        if self.world > 1:
            # replicas must start identical: broadcast rank 0's parameters and buffers
            for m in (model, PREFormer):
                for t in list(m.parameters()) + list(m.buffers()):
                    dist.broadcast(t.data, src=0, group=group)
            optimizer.on_materialize = self._on_materialize

    def _on_materialize(self, flat):
        # embedding tables that announced themselves as row-sparse in the first (eager) step and are large enough to matter
        sparse = [p for p in flat.params if self.row_sparse and getattr(p, "_tavk_row_sparse", False) and p.numel() >= (1 << 22)]
        self.buckets = GradBuckets(flat, self.bucket_bytes, self.group, reduce_dtype=self.grad_reduce_dtype, row_sparse=sparse)

    # -- CUDA-graph path: the whole step (forward, loss, backward, bucketed all-reduce, clip + AdamW) is captured
    #    once per (shape, epoch-parity, check) and replayed; per step the host only enqueues the input copies and
    #    one graph launch.  Inputs may live on the host (pinned or pageable) or already on the device.
    def _graph_key(self, inputs, labels, epoch, check):
        shapes = tuple((k, tuple(v.shape), v.dtype) for d in inputs for k, v in d.items())
        weighted = getattr(self.criterion, "epoch_switch", None)
        return (shapes, tuple(labels.shape), check, (epoch % weighted) if weighted else 0)

    def _build_graph(self, inputs, labels, epoch, check):
        dev = next(self.model.parameters()).device
        st_in = [{k: torch.empty(v.shape, dtype=v.dtype, device=dev) for k, v in d.items()} for d in inputs]
        st_lab = torch.empty(labels.shape, dtype=labels.dtype, device=dev)
        # host-side facts the modules need without a device sync (token count kept by the video mask)
        self.pre.static_keep_count = int(inputs[2]["attention_mask"][0].sum())
        self.model.static_keep_count = int((~inputs[2]["attention_mask"][0]).sum())

        def load():
            for d, sd in zip(inputs, st_in):
                for k, v in d.items():
                    sd[k].copy_(v, non_blocking=True)
            st_lab.copy_(labels, non_blocking=True)

        load()
        # Warm-up steps (allocator / lazy-initialisation effects, the flat parameter buffers of the first step) must not
        # train: everything a step mutates is snapshotted here and put back after the capture, so the first replay is
        # the first optimisation step on this batch (capture itself executes nothing).
        snap = self._snapshot()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, self.graph_warmup)):
                self._eager_step(st_in, st_lab, epoch, check)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = self._eager_step(st_in, st_lab, epoch, check)
        self._restore(snap)
        torch.cuda.synchronize()
        return {"graph": g, "inputs": st_in, "labels": st_lab, "loss": out}

    def _stateful(self):
        seen, out = set(), []
        for m in (self.model, self.pre):
            for t in list(m.parameters()) + list(m.buffers()):
                if id(t) not in seen:
                    seen.add(id(t))
                    out.append(t)
        return out

    def _snapshot(self):
        from . import engine

        return {"tensors": [(t, t.detach().clone()) for t in self._stateful()], "opt": self.opt.snapshot(),
                "dropout": {d: c.clone() for d, c in engine._dropout_counter.items()}}

    @torch.no_grad()
    def _restore(self, snap):
        from . import engine

        for t, saved in snap["tensors"]:
            t.data.copy_(saved)          # .data: parameters are views of the flat buffer by now; keep them views
        self.opt.restore(snap["opt"])
        for d, c in engine._dropout_counter.items():
            if d in snap["dropout"]:
                c.copy_(snap["dropout"][d])
            else:
                c.zero_()

    def train_step(self, inputs, labels, epoch=1, check="train", next_batch=None, sched_t=None):
        """One optimisation step on this rank's shard.  Returns the GLOBAL loss as a device scalar.
        ``sched_t``: when a scheduler was given, ``scheduler.step(sched_t)`` runs after the update (the reference's
        ``scheduler.step(epoch + batch_idx / iters)``, train_model/tav_train.py:63).

        ``next_batch=(inputs, labels)`` (graph mode, host tensors): the batch the NEXT call will be given.  Its
        host-to-device copy is started on a copy stream into a staging set right after this step's graph launch, so it
        overlaps the step's compute (the reference's DataLoader hands batches over one at a time and copies them
        synchronously, tav_train.py:15-40).  The next call recognises the batch by identity and only moves it
        device-to-device into the graph's static input buffers."""
        if not self.use_cuda_graph:
            loss = self._eager_step(inputs, labels, epoch, check)
            self._sched(sched_t)
            return loss
        key = self._graph_key(inputs, labels, epoch, check)
        if self._graph is None or self._graph["key"] != key:
            self._graph = self._build_graph(inputs, labels, epoch, check)
            self._graph["key"] = key
            self._staged = None
        st = self._graph
        staged = getattr(self, "_staged", None)
        if staged is not None and self._same_batch(staged["batch"], inputs, labels):
            torch.cuda.current_stream().wait_event(staged["ready"])
            for sd, gd in zip(staged["inputs"], st["inputs"]):
                for k, v in sd.items():
                    gd[k].copy_(v, non_blocking=True)
            st["labels"].copy_(staged["labels"], non_blocking=True)
            self._staging_free = torch.cuda.Event()
            self._staging_free.record()
        else:
            for d, sd in zip(inputs, st["inputs"]):
                for k, v in d.items():
                    if sd[k].data_ptr() != v.data_ptr():
                        sd[k].copy_(v, non_blocking=True)
            if st["labels"].data_ptr() != labels.data_ptr():
                st["labels"].copy_(labels, non_blocking=True)
        self._staged = None
        self.opt.upload_lr()             # the scheduler's current learning rate -> the device scalar the graph reads
        st["graph"].replay()
        self.opt.note_replayed()
        self._sched(sched_t)
        if next_batch is not None:
            self._stage(next_batch[0], next_batch[1])
        return st["loss"]

    def _sched(self, t):
        if self.scheduler is not None:
            if t is None:
                self.scheduler.step()
            else:
                self.scheduler.step(t)

    @staticmethod
    def _same_batch(staged, inputs, labels):
        """The staged batch is recognised by OBJECT identity (the staging record keeps the tensors alive): a host
        address says nothing, pinned blocks of a dropped batch are handed to the next one."""
        st_in, st_lab = staged
        if st_lab is not labels or len(st_in) != len(inputs):
            return False
        return all(a.keys() == b.keys() and all(a[k] is b[k] for k in a) for a, b in zip(st_in, inputs))

    def _stage(self, inputs, labels):
        """Start the H2D copy of the next batch on the copy stream (pinned host memory makes it asynchronous)."""
        st = self._graph
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream()
            self._stage_bufs = ([{k: torch.empty_like(v) for k, v in d.items()} for d in st["inputs"]],
                                torch.empty_like(st["labels"]))
            self._staging_free = None
        if tuple((k, tuple(v.shape), v.dtype) for d in inputs for k, v in d.items()) != st["key"][0]:
            return                                   # different shapes: the next call rebuilds the graph anyway
        cs = self._copy_stream
        if self._staging_free is not None:
            cs.wait_event(self._staging_free)        # the previous staged batch has been moved out of the staging set
        with torch.cuda.stream(cs):
            for d, sd in zip(inputs, self._stage_bufs[0]):
                for k, v in d.items():
                    sd[k].copy_(v, non_blocking=True)
            self._stage_bufs[1].copy_(labels, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(cs)
        self._staged = {"batch": ([dict(d) for d in inputs], labels), "inputs": self._stage_bufs[0],
                        "labels": self._stage_bufs[1], "ready": ready}

    def static_inputs(self):
        """Device-resident input buffers of the captured step (write into them to skip the host copy)."""
        return (self._graph["inputs"], self._graph["labels"]) if self._graph else None

    def _eager_step(self, inputs, labels, epoch=1, check="train", loss_scale=1.0, clip="default", Metric=None):
        """forward + global loss + backward (+ bucketed all-reduce) + fused clip/AdamW.  ``loss_scale`` multiplies the
        back-propagated loss (the reference's gradient-accumulation branch divides by ``accum_iter``,
        train_model/tav_train.py:100); ``clip=None`` disables clipping for this step."""
        from .tav_train import get_statistics  # local import: tav_train imports optim, not dp

        crit = self.criterion
        holder = {}

        def parts_criterion(output, label, epoch=None):
            num, den = crit.parts(output, label, epoch)
            loss_bwd, loss_val = global_loss(num, den, self.group)
            holder["value"] = loss_val * loss_scale if loss_scale != 1.0 else loss_val
            return loss_bwd * loss_scale if loss_scale != 1.0 else loss_bwd

        if self.buckets is not None:
            self.buckets.start_backward()
        loss = get_statistics(inputs, labels, self.model, self.pre, parts_criterion, Metric, check=check, epoch=epoch)
        if self.world > 1 and self.comm_sms > 0 and self.buckets is not None:
            from . import _lib as L

            keep = L.gemm_reserved_sms
            L.gemm_reserved_sms = self.comm_sms
            try:
                loss.backward()
            finally:
                L.gemm_reserved_sms = keep
        else:
            loss.backward()
        if self.world > 1:
            if self.buckets is not None:
                self.buckets.finish()
            else:
                # first step: the flat buffer does not exist yet -> flatten now, reduce it in one piece
                self.opt.materialize()
                self.opt.flat.attach_grads()
                dist.all_reduce(self.opt.flat.grad, op=dist.ReduceOp.SUM, group=self.group)
        self.opt.step(max_grad_norm=self.clip if clip == "default" else clip)
        return holder["value"]
