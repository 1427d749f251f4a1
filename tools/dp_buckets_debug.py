"""One GPU: the bucketed data-parallel step with an identity all-reduce.  Prints the bucket layout around given parameter
names, the order in which buckets are launched, and compares step-2 gradients (gradient sink + buckets) with step-1
gradients (plain autograd) — with lr = 0 they must agree to bf16/atomics noise."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.distributed as dist
from multi_modal_emotion_b200 import dp, synthetic as syn, tav
from multi_modal_emotion_b200.losses import NewCrossEntropyLoss
from multi_modal_emotion_b200.optim import FusedAdamW

dev = torch.device("cuda", 0)
tav.set_encoder_variant("tiny_base")
torch.manual_seed(0)
model = tav.TAVForMAE({"output_dim": 7, "dropout": 0.0, "learn_PosEmbeddings": True, "num_layers": 12})
pre = tav.PreFormer()
pre.load_state_dict(syn.synth_state_dict(pre, seed=1))
model.load_state_dict(syn.synth_state_dict(model, seed=2))
crit = NewCrossEntropyLoss(class_weights=torch.tensor(syn.MELD_CLASS_WEIGHTS), epoch_switch=2)
model, pre = model.to(dev).train(), pre.to(dev).train()
names = {}
for tag, m in (("TAVForMAE", model), ("PreFormer", pre)):
    for k, p in m.named_parameters():
        names[id(p)] = "%s/%s" % (tag, k)
params = [p for p in model.parameters() if p.requires_grad] + [p for p in pre.parameters() if p.requires_grad]
opt = FusedAdamW(params, lr=0.0, weight_decay=0.0)
runner = dp.DataParallelTAV(model, pre, crit, opt, clip=1.0, bucket_mb=8)
runner.world = 2                      # pretend: take the bucketed path
opt.on_materialize = runner._on_materialize
launched = []


class _W:
    def wait(self):
        pass


def fake_all_reduce(t, op=None, group=None, async_op=False):
    return _W()


dist.all_reduce = fake_all_reduce
dp.is_distributed = lambda: False
inputs, labels = syn.make_batch("C2", seed=99, B=2)
seen = []
real_step = opt.step


def spy(*a, **kw):
    opt.materialize()
    seen.append({names[id(p)]: p.grad.detach().clone() for p in params if p.grad is not None})
    return real_step(*a, **kw)


opt.step = spy
for i in range(3):
    if runner.buckets is not None and i == 1:
        b = runner.buckets
        orig = b._launch

        def _launch(bi, orig=orig, b=b):
            launched.append((bi, b.launched))
            return orig(bi)
        b._launch = _launch
        orig_hook = b._hook
        hook_log = []

        def _hook(p, orig_hook=orig_hook, b=b):
            bi = b.param_bucket[id(p)]
            hook_log.append((len(hook_log), bi, b.pending[bi] - 1, names[id(p)], str(torch.cuda.current_stream().stream_id)))
            return orig_hook(p)
        b._hook = _hook
        # the post-accumulate hooks were registered with the bound method: re-register through the logging wrapper
        for h in b.handles:
            h.remove()
        b.handles = [p.register_post_accumulate_grad_hook(_hook) for p in b.flat.params]
        from multi_modal_emotion_b200 import engine as _e
        _orig_start = b.start_backward

        def _start(_orig_start=_orig_start, b=b):
            _orig_start()
            _e.grad_written_hook = lambda ps: [b._hook(p) for p in ps if id(p) in b.param_bucket]
        b.start_backward = _start
    runner._eager_step(inputs, labels, 1, "val")
torch.cuda.synchronize()
b = runner.buckets
flat = opt.flat
watch = sys.argv[1:] or ["wav2vec2.encoder.layers.1.final_layer_norm.bias", "TAVForMAE/bert_norm.bias"]
for bi, (s, e, n) in enumerate(b.buckets):
    members = [names[id(p)] for p, o in zip(flat.params, flat.offsets) if s <= o < e]
    if any(w in m for w in watch for m in members):
        print("bucket %d [%d, %d) %d params (%.1f MB):" % (bi, s, e, n, (e - s) * 4 / 2**20))
        for m in members:
            print("     ", m)
print("launch order (bucket, k-th):", launched[:12], "...")
n1 = len([h for h in hook_log])
print("hook calls:", n1)
for h in hook_log:
    if h[1] in (66, 76):
        print("   hook #%d bucket %d pending-> %d %s (stream %s)" % h)
for step in (1, 2):
    a, g = seen[0], seen[step]
    num = sum((g[k] - a[k]).norm().item() ** 2 for k in a)
    den = sum(a[k].norm().item() ** 2 for k in a)
    worst = sorted(((g[k] - a[k]).norm().item() / max(a[k].norm().item(), 1e-12), (g[k] - a[k]).norm().item(), a[k].norm().item(), k) for k in a)[-6:]
    print("step %d vs step 0: flat rel-L2 %.3e" % (step, (num / den) ** 0.5))
    for r, ae, gn, k in reversed(worst):
        print("    rel %.3e |err| %.3e |g| %.3e %s" % (r, ae, gn, k))
