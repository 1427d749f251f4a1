"""Device-side duration and achieved HBM bandwidth of the LayerNorm kernels at the step's row counts (20 launches in a
CUDA graph, CUDA events around 5 replays; operands larger than L2 rotate through 4 buffer sets).
python tools/ln_probe.py [M ...]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_modal_emotion_b200 import _lib as L  # noqa: E402

L.require_device()
H = 768
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]


def timed(fn, reps=20):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for i in range(4):
            fn(i)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * reps) * 1e3


for M in [int(a) for a in sys.argv[1:]] or [1120, 2384, 5168, 23424, 41344]:
    sets = []
    for _ in range(4):
        x = torch.randn(M, H, device="cuda")
        sets.append(dict(x=x, dy=torch.randn(M, H, device="cuda"), res=torch.randn(M, H, device="cuda"),
                         mean=x.mean(1), rstd=1.0 / (x.var(1, unbiased=False) + 1e-5).sqrt(),
                         dxf=torch.empty(M, H, device="cuda"), dxb=torch.empty(M, H, device="cuda", dtype=torch.bfloat16),
                         yb=torch.empty(M, H, device="cuda", dtype=torch.bfloat16), m2=torch.empty(M, device="cuda"), r2=torch.empty(M, device="cuda")))
    g, b = torch.ones(H, device="cuda"), torch.zeros(H, device="cuda")
    dg, db, dc = torch.zeros(H, device="cuda"), torch.zeros(H, device="cuda"), torch.zeros(H, device="cuda")

    def bwd(i):
        t = sets[i % 4]
        L.call("tavk_layernorm_bwd", t["dy"].data_ptr(), t["x"].data_ptr(), t["mean"].data_ptr(), t["rstd"].data_ptr(), g.data_ptr(),
               t["res"].data_ptr(), t["dxf"].data_ptr(), t["dxb"].data_ptr(), dg.data_ptr(), db.data_ptr(), dc.data_ptr(), M, H)

    def fwd(i):
        t = sets[i % 4]
        L.call("tavk_layernorm_fwd", t["x"].data_ptr(), g.data_ptr(), b.data_ptr(), t["yb"].data_ptr(), None, t["m2"].data_ptr(),
               t["r2"].data_ptr(), M, H, 1e-5)

    ub, uf = timed(bwd), timed(fwd)
    bb, bf = M * H * (12 + 6), M * H * (4 + 2)
    print("M=%6d  bwd %7.1f us  %6.0f GB/s (%.2f of measured HBM)   fwd %6.1f us  %6.0f GB/s (%.2f)" % (
        M, ub, bb / ub / 1e3, bb / ub / 1e3 / peak, uf, bf / uf / 1e3, bf / uf / 1e3 / peak), flush=True)
