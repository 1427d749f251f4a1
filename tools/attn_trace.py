"""Event timeline of the dK/dV attention-backward kernel (development tool).

Builds a PRIVATE copy of the library with -DTAVK_ATTN_TRACE (multi-modal-emotion_b200/build/libtavk_trace.so; libtavk.so
itself has the trace compiled out), runs one backward at B=16, S=1464 and dumps clock64() stamps of every hand-off of
every step for 8 CTAs to gpurun_out/attn_trace_<variant>.json.
  python tools/attn_trace.py build     # here (no GPU)
  python tools/attn_trace.py           # on the GPU box
  python tools/attn_trace.py summary gpurun_out/attn_trace_dkv.json
Slots per (cta, step): 0 producer passed qdo_empty, 1 producer arrived on qdo_full, 3 issuer about to issue S^T/dP^T,
4 issued, 5 issuer saw pds_full, 6 dV/dK issued, 7 elementwise warp 2 starts waiting for st_full, 8 got it, 9 released
S^T/dP^T (st_free), 10 first half computed, 11 passed pds_empty, 12 P^T/dS^T stored, 13 arrived on pds_full;
row 63 of a CTA = (smid, start clock)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multi_modal_emotion_b200 import build_ext  # noqa: E402

TRACE_LIB = os.path.join(build_ext.OBJ, "libtavk_trace.so")


def build():
    build_ext.build_library()
    nvcc = build_ext._nvcc()
    obj = os.path.join(build_ext.OBJ, "attention_tc_trace.o")
    subprocess.run([nvcc] + build_ext.NVCC_FLAGS + ["-DTAVK_ATTN_TRACE", "-c", os.path.join(build_ext.CSRC, "attention_tc.cu"),
                    "-o", obj], check=True)
    objs = [os.path.join(build_ext.OBJ, s.replace(".cu", ".o")) for s in build_ext.SOURCES if s != "attention_tc.cu"] + [obj]
    subprocess.run([nvcc, "-shared", "-cudart", "shared", "-o", TRACE_LIB] + objs, check=True)
    print(TRACE_LIB)


def run():
    import ctypes

    import torch

    from multi_modal_emotion_b200 import _lib as L

    L.LIB_PATH = TRACE_LIB
    h = L.lib()
    B, S, nh, H = 16, 1464, 12, 768
    qkv = torch.randn(B, S, 3 * H, device="cuda").bfloat16()
    q, k, v = qkv[..., :H], qkv[..., H:2 * H], qkv[..., 2 * H:]
    o = torch.empty(B, S, H, device="cuda", dtype=torch.bfloat16)
    do = torch.randn(B, S, H, device="cuda").bfloat16()
    lse = torch.empty(B, nh, S, device="cuda")
    delta = torch.empty_like(lse)
    dqkv = torch.empty_like(qkv)
    buf = torch.zeros(8 * 64 * 16, dtype=torch.int64, device="cuda")
    fn = h.tavk_debug_set_attn_trace
    fn.argtypes, fn.restype = [ctypes.c_void_p], ctypes.c_int
    assert fn(buf.data_ptr()) == 0
    L.attn_fwd(q, k, v, o, lse, B=B, S=S, nh=nh, ld_qkv=3 * H, ld_o=H)
    for it in range(3):
        buf.zero_()
        L.attn_bwd(q, k, v, o, do, lse, delta, dqkv[..., :H], dqkv[..., H:2 * H], dqkv[..., 2 * H:], B=B, S=S, nh=nh,
                   ld_qkv=3 * H, ld_o=H, ld_dqkv=3 * H)
        torch.cuda.synchronize()
    out = os.path.join(ROOT, "gpurun_out", "attn_trace_%s.json" % (sys.argv[1] if len(sys.argv) > 1 else "dkv"))
    os.makedirs(os.path.dirname(out), exist_ok=True)
    json.dump(buf.view(8, 64, 16).cpu().tolist(), open(out, "w"))
    print("wrote", out)


def summary(path, cta=2):
    """Average phase durations (cycles) over steps 4..19 of one traced CTA."""
    import statistics as st

    t = json.load(open(path))
    r = t[cta]
    steps = range(4, 20)
    avg = lambda f: st.mean(f(i) for i in steps)  # noqa: E731
    print("%s: CTA %d on SM %d" % (path, cta, r[63][0]))
    rows = [
        ("period per step (st_full(i) -> st_full(i+1))", lambda i: r[i + 1][8] - r[i][8]),
        ("elementwise: wait for S^T/dP^T (st_full)", lambda i: r[i][8] - r[i][7]),
        ("elementwise: busy (st_full -> arrive pds_full)", lambda i: r[i][13] - r[i][8]),
        ("  of which wait for pds_empty (dV/dK of step i-1 done)", lambda i: r[i][11] - r[i][10]),
        ("  st_full -> S^T/dP^T released (st_free)", lambda i: r[i][9] - r[i][8]),
        ("issuer: issue S^T/dP^T MMAs + commit", lambda i: r[i][4] - r[i][3]),
        ("issuer: issue dV/dK MMAs + commits", lambda i: r[i][6] - r[i][5]),
        ("issuer: P^T/dS^T staged -> noticed", lambda i: r[i][5] - r[i][13]),
        ("S^T/dP^T issued -> visible to the elementwise warps", lambda i: r[i][8] - r[i][4]),
        ("producer: stage free (qdo_empty) -> stage full (TMA issue + statistics + arrive)", lambda i: r[i][1] - r[i][0]),
        ("dV/dK(i) issued -> producer sees the stage of step i+2 free", lambda i: r[i + 2][0] - r[i][6]),
        ("S^T/dP^T released (i) -> S^T/dP^T of step i+1 being issued", lambda i: r[i + 1][3] - r[i][9]),
    ]
    for name, f in rows:
        print("  %-85s %7.0f" % (name, avg(f)))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "build":
        build()
    elif len(sys.argv) > 2 and sys.argv[1] == "summary":
        summary(sys.argv[2])
    else:
        run()
