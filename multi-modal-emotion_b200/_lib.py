"""ctypes binding of libtavk.so (C ABI declared in include/tavk.h).

This is the only place Python touches the kernel library.  There is no CPU fallback: ``lib()`` raises if the shared
object is missing and every wrapper raises ``TavkError`` on a non-zero return code.  Tensors are passed as raw device
pointers; outputs and workspaces are torch tensors owned by the caller (PyTorch caching allocator), and every kernel is
enqueued on torch's current CUDA stream."""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtavk.so")

F32, BF16 = 0, 1
EPI_LINEAR, EPI_GELU, EPI_GELU_BWD, EPI_GELU_GRAD, EPI_MUL = 0, 1, 2, 3, 4
ATTN_NONE, ATTN_KEY_BIAS = 0, 1


class TavkError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("lda", C.c_int64), ("a_mn_major", C.c_int32),
        ("B", C.c_void_p), ("ldb", C.c_int64), ("b_mn_major", C.c_int32),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("out", C.c_void_p), ("ldo", C.c_int64), ("out_dtype", C.c_int32),
        ("out2", C.c_void_p), ("ldo2", C.c_int64),
        ("bias", C.c_void_p),
        ("resid", C.c_void_p), ("ldr", C.c_int64),
        ("rowbias", C.c_void_p), ("rows_per_group", C.c_int32),
        ("aux", C.c_void_p), ("ldaux", C.c_int64),
        ("colsum", C.c_void_p),
        ("epilogue", C.c_int32), ("accumulate", C.c_int32), ("k_splits", C.c_int32), ("block_n", C.c_int32),
        ("alpha", C.c_float),
        ("groups", C.c_int32), ("a_kstep", C.c_int32),
        ("a_g_mn", C.c_int32), ("a_g_k", C.c_int32), ("b_g_mn", C.c_int32), ("b_g_k", C.c_int32),
        ("b_box_k_shift", C.c_int32), ("out_g_row", C.c_int32), ("out_g_col", C.c_int32),
        ("a_rows", C.c_int64), ("a_cols", C.c_int64), ("b_rows", C.c_int64), ("b_cols", C.c_int64),
        ("max_ctas", C.c_int32),
        ("sched_workspace", C.c_void_p),
        ("cta_pair", C.c_int32),
    ]


class AttnArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("ld_qkv", C.c_int64),
        ("o", C.c_void_p), ("ld_o", C.c_int64),
        ("lse", C.c_void_p), ("key_bias", C.c_void_p),
        ("B", C.c_int32), ("S", C.c_int32), ("nh", C.c_int32), ("mode", C.c_int32), ("scale", C.c_float),
    ]


class AttnBwdArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("ld_qkv", C.c_int64),
        ("o", C.c_void_p), ("d_o", C.c_void_p), ("ld_o", C.c_int64),
        ("lse", C.c_void_p), ("delta", C.c_void_p), ("key_bias", C.c_void_p),
        ("dq", C.c_void_p), ("dk", C.c_void_p), ("dv", C.c_void_p), ("ld_dqkv", C.c_int64),
        ("dv_rowscale", C.c_void_p), ("dv_rank1", C.c_void_p),
        ("B", C.c_int32), ("S", C.c_int32), ("nh", C.c_int32), ("mode", C.c_int32), ("scale", C.c_float),
        ("dbq", C.c_void_p), ("dbk", C.c_void_p), ("dbv", C.c_void_p),
    ]


_P, _I, _L, _F, _U64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_uint64

# name -> argtypes (restype is int unless listed in _RESTYPES); mirrors include/tavk.h one to one
SIGNATURES = {
    "tavk_last_error": [],
    "tavk_version": [],
    "tavk_device_check": [],
    "tavk_sm_count": [],
    "tavk_workspace_bytes_attn_bwd": [_I, _I, _I],
    "tavk_workspace_bytes_groupnorm": [_I, _I],
    "tavk_workspace_bytes_gemm": [C.POINTER(GemmArgs)],
    "tavk_gemm_bf16": [C.POINTER(GemmArgs), _P],
    "tavk_attn_fwd": [C.POINTER(AttnArgs), _P],
    "tavk_attn_bwd": [C.POINTER(AttnBwdArgs), _P],
    "tavk_layernorm_fwd": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _F, _P],
    "tavk_layernorm_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P],
    "tavk_embed_add_fwd": [_P, _P, _P, _P, _I, _I, _I, _P],
    "tavk_embed_add_bwd": [_P, _P, _P, _I, _I, _I, _P],
    "tavk_roberta_embed_fwd": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "tavk_embedding_scatter_add": [_P, _P, _P, _I, _I, _I, _I, _P],
    "tavk_mean_pool_fwd": [_P, _P, _I, _I, _I, _P],
    "tavk_mean_pool_bwd": [_P, _P, _P, _I, _I, _I, _P],
    "tavk_masked_mean_pool_fwd": [_P, _P, _P, _I, _I, _I, _P],
    "tavk_masked_mean_pool_bwd": [_P, _P, _P, _P, _I, _I, _I, _P],
    "tavk_colsum": [_P, _I, _L, _P, _I, _I, _I, _P],
    "tavk_masked_colsum": [_P, _I, _L, _P, _P, _I, _I, _I, _P],
    "tavk_small_linear_fwd": [_P, _P, _P, _P, _I, _I, _I, _P],
    "tavk_small_linear_bwd_x": [_P, _P, _P, _I, _I, _I, _I, _P],
    "tavk_small_linear_bwd_w": [_P, _P, _P, _P, _I, _I, _I, _P],
    "tavk_cast_f32_bf16": [_P, _P, _L, _P],
    "tavk_scale_f32": [_P, _P, _F, _L, _P],
    "tavk_dropout": [_P, _P, _P, _L, _F, _U64, _U64, _P, _P],
    "tavk_dropout_bwd": [_P, _P, _P, _L, _F, _P],
    "tavk_dropout_bwd_add": [_P, _P, _P, _P, _L, _F, _P],
    "tavk_permute_bshd_bhds": [_P, _P, _I, _I, _I, _I, _I, _P],
    "tavk_conv0_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "tavk_groupnorm_gelu_fwd": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P],
    "tavk_groupnorm_bwd": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "tavk_wave_windows": [_P, _P, _I, _I, _I, _I, _I, _I, _P],
    "tavk_chan_ln_gelu_fwd": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P],
    "tavk_chan_ln_gelu_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "tavk_softmax_ce_fwd": [_P, _P, _P, _P, _P, _P, _I, _I, _P],
    "tavk_softmax_ce_bwd": [_P, _P, _P, _P, _P, _I, _I, _P],
    "tavk_grad_sqnorm": [_P, _L, _P, _P],
    "tavk_adamw": [_P, _P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _I, _P, _F, _F, _I, _P],
    "tavk_adamw_prep": [_P, _P, _P, C.c_double, C.c_double, _P],
    "tavk_adamw_dev": [_P, _P, _P, _P, _P, _L, _P, C.c_double, C.c_double, _F, _F, _P, _F, _F, _I, _P],
}
_RESTYPES = {"tavk_last_error": C.c_char_p, "tavk_workspace_bytes_attn_bwd": C.c_int64,
             "tavk_workspace_bytes_groupnorm": C.c_int64, "tavk_workspace_bytes_gemm": C.c_int64}

_lib = None
launch_count = 0   # library entry points invoked
kernel_count = 0   # CUDA kernels those entry points launched (bench.py reports it as gpu_launches)
record_gemms = False
gemm_log = []      # (M, N, K, a_mn, b_mn, epilogue, out_bf16, bias, resid, rowbias, accumulate, k_splits, groups) per launch
# kernels launched per entry point (memset nodes are not counted)
_KERNELS = {"tavk_attn_bwd": 3, "tavk_groupnorm_gelu_fwd": 2, "tavk_groupnorm_bwd": 2}


def lib():
    """Load libtavk.so once.  Raises (never falls back) when the library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TavkError(
                "libtavk.so is missing (%s): build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
                "there is no CPU fallback for the TAV kernels" % LIB_PATH)
        h = C.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(h, name)
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, C.c_int)
        _lib = h
    return _lib


def _check(rc, name):
    if rc != 0:
        raise TavkError("%s failed (code %d): %s" % (name, rc, lib().tavk_last_error().decode()))


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def call(name, *args):
    """Invoke an entry point on torch's current stream (appended as the last argument)."""
    global launch_count, kernel_count
    launch_count += 1
    kernel_count += _KERNELS.get(name, 1)
    rc = getattr(lib(), name)(*args, torch.cuda.current_stream().cuda_stream)
    if rc != 0:
        _check(rc, name)


# Dynamic tile scheduling of the persistent GEMM (tavk_gemm_args.sched_workspace): every gemm() call takes the next
# {next tile, finished CTAs} pair of a per-device rotating pool, so any two launches fewer than _SCHED_SLOTS calls apart —
# in particular kernels of different branch streams or graph branches that may run concurrently — never share one; a
# captured graph node keeps the pair it was captured with (the kernel leaves it zeroed).  The pool is created by the first
# call, which is never inside a stream capture (graphs are captured after eager warm-up steps).
gemm_dynamic_tiles = os.environ.get("TAVK_GEMM_DYNAMIC", "1") != "0"
_SCHED_SLOTS = 8192
_sched_pools = {}


def _sched_slot():
    dev = torch.cuda.current_device()
    pool = _sched_pools.get(dev)
    if pool is None:
        if torch.cuda.is_current_stream_capturing():
            raise TavkError("the GEMM scheduler workspace must exist before a CUDA graph is captured: run one eager step first")
        pool = _sched_pools[dev] = [torch.zeros(2 * _SCHED_SLOTS, dtype=torch.int32, device=torch.device("cuda", dev)), 0]
        torch.cuda.synchronize(dev)
    i = pool[1]
    pool[1] = (i + 1) % _SCHED_SLOTS
    return pool[0].data_ptr() + 8 * i


gemm_reserved_sms = 0   # HOST-side policy: SMs every gemm() call leaves free (passed per call as tavk_gemm_args.max_ctas)


def reserve_sms(n):
    """Keep n SMs out of the persistent GEMM's grid (for concurrently running NCCL kernels).  The library itself holds no
    such state: the budget travels with every call."""
    global gemm_reserved_sms
    n = int(n)
    if n < 0 or n >= lib().tavk_sm_count():
        raise TavkError("reserve_sms: %d is outside [0, %d)" % (n, lib().tavk_sm_count()))
    gemm_reserved_sms = n


def require_device():
    _check(lib().tavk_device_check(), "tavk_device_check")


# ------------------------------------------------------------------------------------------------ typed wrappers
_GEMM_EXT = ("groups", "a_kstep", "a_g_mn", "a_g_k", "b_g_mn", "b_g_k", "b_box_k_shift", "out_g_row", "out_g_col",
             "a_rows", "a_cols", "b_rows", "b_cols")


def gemm(A, B, out, *, M, N, K, lda=None, ldb=None, a_mn=False, b_mn=False, out2=None, bias=None, resid=None,
         rowbias=None, rows_per_group=0, aux=None, colsum=None, epilogue=EPI_LINEAR, accumulate=False, k_splits=1,
         block_n=0, alpha=1.0, ldo=None, cta_pair=0, **ext):
    """out[M,N] = epi(alpha * A·B^T).  A/B bf16 2-D tensors (or views); K-major: [rows,K]; MN-major: [K,rows].
    ``ext``: the grouped / convolution-walk fields of tavk_gemm_args (groups, a_kstep, a_g_mn, ..., b_cols)."""
    a = GemmArgs()
    for k, v in ext.items():
        if k not in _GEMM_EXT:
            raise TypeError("gemm() got an unexpected keyword argument %r" % k)
        setattr(a, k, int(v))
    a.A, a.lda, a.a_mn_major = A.data_ptr(), (lda if lda is not None else A.stride(0)), int(a_mn)
    a.B, a.ldb, a.b_mn_major = B.data_ptr(), (ldb if ldb is not None else B.stride(0)), int(b_mn)
    a.M, a.N, a.K = M, N, K
    a.out, a.ldo = out.data_ptr(), (ldo if ldo is not None else out.stride(0))
    a.out_dtype = BF16 if out.dtype == torch.bfloat16 else F32
    a.out2, a.ldo2 = _ptr(out2), (out2.stride(0) if out2 is not None else 0)
    a.bias = _ptr(bias)
    a.resid, a.ldr = _ptr(resid), (resid.stride(0) if resid is not None else 0)
    a.rowbias, a.rows_per_group = _ptr(rowbias), rows_per_group
    a.aux, a.ldaux = _ptr(aux), (aux.stride(0) if aux is not None else 0)
    a.colsum = _ptr(colsum)
    a.epilogue, a.accumulate, a.k_splits, a.block_n, a.alpha = epilogue, int(accumulate), k_splits, block_n, alpha
    a.cta_pair = cta_pair      # 0 = library heuristic, 1 = single-CTA tiles, 2 = CTA pairs (cta_group::2) when possible
    if gemm_reserved_sms:
        a.max_ctas = lib().tavk_sm_count() - gemm_reserved_sms
    if gemm_dynamic_tiles:
        a.sched_workspace = _sched_slot()
    if record_gemms:
        gemm_log.append((M, N, K, int(a_mn), int(b_mn), epilogue, int(out.dtype == torch.bfloat16), int(bias is not None),
                         int(resid is not None), int(rowbias is not None), int(accumulate), k_splits,
                         int(ext.get("groups", 1) or 1)))
    call("tavk_gemm_bf16", C.byref(a))


def attn_fwd(q, k, v, o, lse, *, B, S, nh, ld_qkv, ld_o, key_bias=None, scale=0.125):
    a = AttnArgs()
    a.q, a.k, a.v, a.ld_qkv = q.data_ptr(), k.data_ptr(), v.data_ptr(), ld_qkv
    a.o, a.ld_o, a.lse, a.key_bias = o.data_ptr(), ld_o, _ptr(lse), _ptr(key_bias)
    a.B, a.S, a.nh = B, S, nh
    a.mode = ATTN_KEY_BIAS if key_bias is not None else ATTN_NONE
    a.scale = scale
    call("tavk_attn_fwd", C.byref(a))


def attn_bwd(q, k, v, o, d_o, lse, delta, dq, dk, dv, *, B, S, nh, ld_qkv, ld_o, ld_dqkv, key_bias=None,
             dv_rowscale=None, dv_rank1=None, scale=0.125, dbq=None, dbk=None, dbv=None):
    a = AttnBwdArgs()
    a.q, a.k, a.v, a.ld_qkv = q.data_ptr(), k.data_ptr(), v.data_ptr(), ld_qkv
    a.o, a.d_o, a.ld_o = o.data_ptr(), d_o.data_ptr(), ld_o
    a.lse, a.delta, a.key_bias = lse.data_ptr(), delta.data_ptr(), _ptr(key_bias)
    a.dq, a.dk, a.dv, a.ld_dqkv = dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), ld_dqkv
    a.dv_rowscale, a.dv_rank1 = _ptr(dv_rowscale), _ptr(dv_rank1)
    a.B, a.S, a.nh = B, S, nh
    a.mode = ATTN_KEY_BIAS if key_bias is not None else ATTN_NONE
    a.scale = scale
    a.dbq, a.dbk, a.dbv = _ptr(dbq), _ptr(dbk), _ptr(dbv)
    call("tavk_attn_bwd", C.byref(a))


def layernorm_fwd(x, gamma, beta, eps, *, want_bf16=True, want_f32=False):
    """x f32 [M,H] -> (y_bf16|None, y_f32|None, mean, rstd)."""
    M, H = x.shape
    yb = torch.empty((M, H), dtype=torch.bfloat16, device=x.device) if want_bf16 else None
    yf = torch.empty((M, H), dtype=torch.float32, device=x.device) if want_f32 else None
    mean = torch.empty((M,), dtype=torch.float32, device=x.device)
    rstd = torch.empty((M,), dtype=torch.float32, device=x.device)
    call("tavk_layernorm_fwd", x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), _ptr(yb), _ptr(yf), mean.data_ptr(),
         rstd.data_ptr(), M, H, float(eps))
    return yb, yf, mean, rstd


def layernorm_bwd(dy, x, mean, rstd, gamma, dgamma, dbeta, *, resid=None, want_f32=True, want_bf16=False,
                  dx_colsum=None):
    """Returns (dx_f32|None, dx_bf16|None); dgamma/dbeta (and dx_colsum, the column sums of dx) are accumulated into
    f32 [H] buffers."""
    M, H = x.shape
    dxf = torch.empty((M, H), dtype=torch.float32, device=x.device) if want_f32 else None
    dxb = torch.empty((M, H), dtype=torch.bfloat16, device=x.device) if want_bf16 else None
    call("tavk_layernorm_bwd", dy.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
         _ptr(resid), _ptr(dxf), _ptr(dxb), _ptr(dgamma), _ptr(dbeta), _ptr(dx_colsum), M, H)
    return dxf, dxb


def colsum(x, out, *, M, N, ld=None, accumulate=False):
    call("tavk_colsum", x.data_ptr(), BF16 if x.dtype == torch.bfloat16 else F32, ld if ld is not None else x.stride(0),
         out.data_ptr(), M, N, int(accumulate))


def masked_colsum(x, w, out, *, B, S, N, ld):
    call("tavk_masked_colsum", x.data_ptr(), BF16 if x.dtype == torch.bfloat16 else F32, ld, _ptr(w), out.data_ptr(), B,
         S, N)


def cast_bf16(x, out=None):
    x = x.contiguous()
    if out is None:
        out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    call("tavk_cast_f32_bf16", x.data_ptr(), out.data_ptr(), x.numel())
    return out
