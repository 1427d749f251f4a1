"""Device-side duration and HBM bandwidth of the fused AdamW step (+ the squared-norm pass) over a flat buffer of the
benchmark's size (395 M parameters): CUDA events around 5 calls after 2 warm-up calls.  python tools/adamw_probe.py [n_millions]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_modal_emotion_b200 import _lib as L  # noqa: E402

if os.environ.get("TAVK_PROBE_LIB"):      # A/B against another build of the library
    L.LIB_PATH = os.environ["TAVK_PROBE_LIB"]
L.require_device()
n = int(float(sys.argv[1]) * 1e6) if len(sys.argv) > 1 else 395_000_000
n -= n % 64
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
p, m, v = torch.randn(n, device="cuda") * 0.02, torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
g = torch.randn(n, device="cuda") * 1e-3
sh = torch.empty(n, device="cuda", dtype=torch.bfloat16)
step = torch.zeros(1, dtype=torch.int32, device="cuda")
hyper = torch.zeros(4, device="cuda")
hyper[0] = 1e-5
sq = torch.zeros(1, device="cuda")


def one():
    L.call("tavk_adamw_prep", step.data_ptr(), hyper.data_ptr(), sq.data_ptr(), 0.9, 0.999)
    L.call("tavk_grad_sqnorm", g.data_ptr(), n, sq.data_ptr())
    L.call("tavk_adamw_dev", p.data_ptr(), m.data_ptr(), v.data_ptr(), g.data_ptr(), sh.data_ptr(), n, hyper.data_ptr(), 0.9, 0.999,
           1e-8, 1e-4, sq.data_ptr(), 1.0, 1.0, 0)


for _ in range(2):
    one()
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e[0].record()
for _ in range(5):
    L.call("tavk_grad_sqnorm", g.data_ptr(), n, sq.data_ptr())
e[1].record()
for _ in range(5):
    L.call("tavk_adamw_dev", p.data_ptr(), m.data_ptr(), v.data_ptr(), g.data_ptr(), sh.data_ptr(), n, hyper.data_ptr(), 0.9, 0.999,
           1e-8, 1e-4, sq.data_ptr(), 1.0, 1.0, 1)
e[2].record()
torch.cuda.synchronize()
t_n, t_a = e[0].elapsed_time(e[1]) / 5, e[1].elapsed_time(e[2]) / 5
print("n = %.0f M parameters: grad_sqnorm %.3f ms (%.0f GB/s, %.2f of measured HBM); adamw %.3f ms (%.0f GB/s of 34 B/param, %.2f of measured HBM)" % (
    n / 1e6, t_n, 4 * n / t_n / 1e6, 4 * n / t_n / 1e6 / peak, t_a, 34 * n / t_a / 1e6, 34 * n / t_a / 1e6 / peak))
