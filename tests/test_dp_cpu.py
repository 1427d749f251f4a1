"""world_size-2 gloo test (CPU) of the data-parallel plumbing in multi_modal_emotion_b200.dp: parameter-aligned
gradient buckets fired from post-accumulate-grad hooks, the global (weighted) CE normaliser, and equivalence of the
2-rank result with a single process on the concatenated batch."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _toy():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(20, 64), torch.nn.GELU(), torch.nn.Linear(64, 64), torch.nn.GELU(),
                               torch.nn.Linear(64, 7))


def _parts(logits, target, w):
    logp = torch.log_softmax(logits, dim=-1)
    nll = -logp.gather(1, target[:, None]).squeeze(1)
    ww = w[target]
    return (ww * nll).sum(), ww.sum()


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from multi_modal_emotion_b200 import dp
    from multi_modal_emotion_b200.optim import FlatParams

    model = _toy()
    g = torch.Generator().manual_seed(7)
    x, y = torch.randn(8, 20, generator=g), torch.randint(0, 7, (8,), generator=g)
    w = torch.rand(7, generator=g) + 0.5
    xs, ys = x[rank * 4:(rank + 1) * 4], y[rank * 4:(rank + 1) * 4]
    flat = _cpu_flat(FlatParams, list(model.parameters()))
    buckets = dp.GradBuckets(flat, bucket_bytes=4096, group=None)     # tiny buckets -> several collectives
    assert len(buckets.buckets) >= 3
    buckets.start_backward()
    num, den = _parts(model(xs), ys, w)
    loss_bwd, loss_val = dp.global_loss(num, den)
    loss_bwd.backward()
    fired_from_hooks = buckets.launched
    buckets.finish()
    assert fired_from_hooks == len(buckets.buckets)                   # every bucket was launched during backward
    out[rank] = (flat.grad.clone(), loss_val.item())
    dist.destroy_process_group()


def _cpu_flat(FlatParams, params):
    f = FlatParams.__new__(FlatParams)
    f.params = params
    f.offsets, off = [], 0
    for p in params:
        f.offsets.append(off)
        off += (p.numel() + 63) // 64 * 64
    f.numel = off
    f.flat = torch.zeros(off)
    f.grad = torch.zeros(off)
    with torch.no_grad():
        for p, o in zip(params, f.offsets):
            v = f.flat[o:o + p.numel()].view(p.shape)
            v.copy_(p.data)
            p.data = v
            p.grad = f.grad[o:o + p.numel()].view(p.shape)
    return f


def test_two_rank_buckets_match_single_process():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    # single process on the concatenated batch
    model = _toy()
    g = torch.Generator().manual_seed(7)
    x, y = torch.randn(8, 20, generator=g), torch.randint(0, 7, (8,), generator=g)
    w = torch.rand(7, generator=g) + 0.5
    num, den = _parts(model(x), y, w)
    (num / den).backward()
    ref = torch.cat([torch.nn.functional.pad(p.grad.flatten(), (0, (-p.numel()) % 64)) for p in model.parameters()])
    for r in range(world):
        grad, loss = out[r]
        assert torch.allclose(grad, ref, rtol=1e-5, atol=1e-7)
        assert abs(loss - (num / den).item()) < 1e-6
    assert torch.equal(out[0][0], out[1][0])                          # ranks hold identical reduced gradients


def test_bucket_layout_covers_flat_buffer_once():
    from multi_modal_emotion_b200 import dp
    from multi_modal_emotion_b200.optim import FlatParams

    model = _toy()
    flat = _cpu_flat(FlatParams, list(model.parameters()))

    class _NoDist(dp.GradBuckets):
        pass

    b = _NoDist(flat, bucket_bytes=2048)
    spans = sorted((s, e) for s, e, _ in b.buckets)
    assert spans[0][0] == 0 and spans[-1][1] == flat.numel
    for (s0, e0), (s1, e1) in zip(spans, spans[1:]):
        assert e0 == s1
    assert sum(n for _, _, n in b.buckets) == len(flat.params)
    b.remove()


class _SinkLinear(torch.autograd.Function):
    """A layer that, like the engine's gradient sink, accumulates its weight gradient straight into ``w.grad`` and hands
    autograd None for it, then reports the parameter through ``engine.grad_written_hook``."""

    @staticmethod
    def forward(ctx, x, w):
        ctx.save_for_backward(x, w)
        return x @ w.t()

    @staticmethod
    def backward(ctx, dy):
        from multi_modal_emotion_b200 import engine

        x, w = ctx.saved_tensors
        w.grad.add_(dy.t() @ x)
        if engine.grad_written_hook is not None:
            engine.grad_written_hook([w])
        return dy @ w, None


def test_sink_parameter_is_counted_once_per_backward():
    """Regression (found by tests/test_dp_nccl_gpu.py on two B200s): torch fires a parameter's post-accumulate-grad hook
    even when the backward returned None for it, so a sink parameter was reported twice — by the engine and by autograd —
    and a bucket it shared with a parameter whose gradient arrives later was all-reduced before that gradient existed."""
    from multi_modal_emotion_b200 import dp
    from multi_modal_emotion_b200.optim import FlatParams

    torch.manual_seed(0)
    first = torch.nn.Parameter(torch.randn(6, 5))     # used first in forward -> its gradient arrives LAST in backward
    sink = torch.nn.Parameter(torch.randn(4, 6))      # gradient written early, by the sink
    flat = _cpu_flat(FlatParams, [first, sink])
    launches = []

    class Probe(dp.GradBuckets):
        def _launch(self, b):
            launches.append((b, first.grad.abs().sum().item() > 0, sink.grad.abs().sum().item() > 0))
            self.launched += 1

    buckets = Probe(flat, bucket_bytes=1 << 20, group=None)     # one bucket holds both
    assert len(buckets.buckets) == 1
    buckets.start_backward()
    x = torch.randn(3, 5)
    _SinkLinear.apply(torch.nn.functional.linear(x, first), sink).sum().backward()
    buckets.finish()
    assert launches == [(0, True, True)]      # launched once, after BOTH gradients had been written


class _SinkEmbedding(torch.autograd.Function):
    """Embedding lookup whose backward does what engine._RobertaEmbedFn does: scatter-add this rank's rows into the table's
    own .grad (gradient sink), offer (ids, rows) to ``engine.row_sparse_hook``, report the table as written."""

    @staticmethod
    def forward(ctx, ids, table):
        ctx.save_for_backward(ids)
        ctx.table = table
        return table[ids]

    @staticmethod
    def backward(ctx, dy):
        from multi_modal_emotion_b200 import engine

        (ids,) = ctx.saved_tensors
        t = ctx.table
        t.grad.index_add_(0, ids.reshape(-1), dy.reshape(-1, dy.shape[-1]))
        if engine.row_sparse_hook is not None:
            engine.row_sparse_hook(t, ids, dy, t.grad, None)
        if engine.grad_written_hook is not None:
            engine.grad_written_hook([t])
        return None, None


def _sparse_toy():
    torch.manual_seed(3)
    return torch.nn.Embedding(50, 16), torch.nn.Linear(16, 7)


def _sparse_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from multi_modal_emotion_b200 import dp
    from multi_modal_emotion_b200.optim import FlatParams

    emb, head = _sparse_toy()
    g = torch.Generator().manual_seed(11)
    ids, y = torch.randint(0, 50, (8, 5), generator=g), torch.randint(0, 7, (8,), generator=g)
    w = torch.rand(7, generator=g) + 0.5
    params = [emb.weight] + list(head.parameters())
    flat = _cpu_flat(FlatParams, params)
    buckets = dp.GradBuckets(flat, bucket_bytes=1 << 20, group=None, row_sparse=[emb.weight])
    assert id(emb.weight) not in buckets.param_bucket            # the table travels as rows, not in a dense bucket
    assert sum(n for _, _, n in buckets.buckets) == len(params) - 1
    buckets.start_backward()
    sl = slice(rank * 4, (rank + 1) * 4)
    num, den = _parts(head(_SinkEmbedding.apply(ids[sl], emb.weight).mean(1)), y[sl], w)
    loss_bwd, _ = dp.global_loss(num, den)
    loss_bwd.backward()
    buckets.finish()
    out[rank] = flat.grad.clone()
    dist.destroy_process_group()


def test_row_sparse_table_gradient_matches_dense_sum():
    """The embedding table's gradient exchanged as (token ids, rows) between two ranks equals the single-process gradient
    on the concatenated batch, like the dense buckets next to it."""
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_sparse_worker, args=(world, port, out), nprocs=world, join=True)
    emb, head = _sparse_toy()
    g = torch.Generator().manual_seed(11)
    ids, y = torch.randint(0, 50, (8, 5), generator=g), torch.randint(0, 7, (8,), generator=g)
    w = torch.rand(7, generator=g) + 0.5
    num, den = _parts(head(emb(ids).mean(1)), y, w)
    (num / den).backward()
    ref = torch.cat([torch.nn.functional.pad(p.grad.flatten(), (0, (-p.numel()) % 64)) for p in [emb.weight] + list(head.parameters())])
    for r in range(world):
        assert torch.allclose(out[r], ref, rtol=1e-5, atol=1e-7)
