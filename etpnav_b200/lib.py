"""ctypes binding of libetpnav_b200.so (C ABI in include/etpnav_b200.h).

The library is built in-tree by ``etpnav_b200.build`` and must be present: there is no CPU or
PyTorch fallback — every wrapper raises if the library is missing or a call fails.
Tensors are passed as raw device pointers on torch's current CUDA stream.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libetpnav_b200.so")

_lib = None


class EtpError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EtpError(f"{LIB_PATH} not built: run `python -m etpnav_b200.build` (no fallback path exists)")
        _lib = C.CDLL(LIB_PATH)
        _lib.etp_last_error.restype = C.c_char_p
        _declare(_lib)
    return _lib


def _check(rc, what):
    if rc != 0:
        raise EtpError(f"{what} failed ({rc}): {lib().etp_last_error().decode()}")


def require_device():
    _check(lib().etp_check_device(), "etp_check_device")


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    if t is None:
        return None
    assert t.is_cuda, "device tensor expected"
    return C.c_void_p(t.data_ptr())


p_void, p_f32, i32, f32 = C.c_void_p, C.c_void_p, C.c_int32, C.c_float


class GemmArgs(C.Structure):
    _fields_ = [("M", i32), ("N", i32), ("K", i32),
                ("A", p_void), ("lda", i32), ("a_mn", i32),
                ("B", p_void), ("ldb", i32), ("b_mn", i32),
                ("alpha", f32), ("bias", p_f32), ("act", i32), ("aux_mode", i32),
                ("aux", p_void), ("ld_aux", i32), ("resid", p_f32), ("ld_resid", i32),
                ("out_f32", p_f32), ("ld_f32", i32), ("atomic", i32),
                ("out_bf16", p_void), ("ld_bf16", i32), ("out_pre", p_void), ("ld_pre", i32),
                ("k_splits", i32), ("block_n", i32), ("colsum", p_f32), ("pre_mode", i32),
                ("drop_key", C.c_uint32), ("drop_thr", C.c_uint32), ("drop_scale", f32)]


class AttnArgs(C.Structure):
    _fields_ = [("B", i32), ("heads", i32), ("Sq", i32), ("Sk", i32),
                ("q", p_void), ("ldq", i32), ("k", p_void), ("ldk", i32), ("v", p_void), ("ldv", i32),
                ("scale", f32), ("key_valid", p_void), ("mask_value", f32),
                ("pair", p_f32), ("pair_w", f32), ("pair_b", f32), ("pair_w_dev", p_f32), ("pair_b_dev", p_f32),
                ("out", p_void), ("ldo", i32), ("lse", p_f32), ("impl", i32)]


class AttnBwdArgs(C.Structure):
    _fields_ = [("B", i32), ("heads", i32), ("Sq", i32), ("Sk", i32),
                ("q", p_void), ("k", p_void), ("v", p_void), ("ldq", i32), ("ldk", i32), ("ldv", i32),
                ("out", p_void), ("ldo", i32), ("dout", p_void), ("lddo", i32), ("lse", p_f32), ("dvec", p_f32),
                ("scale", f32), ("key_valid", p_void), ("mask_value", f32), ("pair", p_f32), ("pair_w", f32),
                ("pair_b", f32), ("dq", p_void), ("dk", p_void), ("dv", p_void), ("lddq", i32), ("lddk", i32),
                ("lddv", i32), ("dpair_w", p_f32), ("dpair_b", p_f32), ("impl", i32)]


class PanoPackArgs(C.Structure):
    _fields_ = [("rows", i32)] + [(n, p_void) for n in (
        "rgb_lin", "dep_lin", "loc_fts", "nav_types", "loc_w", "loc_b", "img_g", "img_b", "dep_g", "dep_b",
        "loc_g", "loc_bb", "out_g", "out_b", "nav_emb", "tok_emb1", "x_f32", "loc_lin", "sum_pre", "stats")]


class NodePackArgs(C.Structure):
    _fields_ = [("rows", i32)] + [(n, p_void) for n in (
        "img_fts", "step_ids", "pos_fts", "pos_w", "pos_b", "pos_g", "pos_bb", "step_emb",
        "x_f32", "x_bf16", "pos_lin", "stats")]


def _declare(L):
    L.etp_gemm.argtypes = [C.POINTER(GemmArgs), C.c_void_p]
    L.etp_attention_fwd.argtypes = [C.POINTER(AttnArgs), C.c_void_p]
    L.etp_attention_bwd.argtypes = [C.POINTER(AttnBwdArgs), C.c_void_p]
    L.etp_layernorm_fwd.argtypes = [p_void, p_void, p_void, f32, i32, i32, p_void, p_void, p_void, p_void, p_void]
    L.etp_layernorm_bwd.argtypes = [p_void, p_void, p_void, p_void, p_void, i32, i32, p_void, i32, p_void, p_void,
                                    p_void, p_void]
    L.etp_colsum_bf16.argtypes = [p_void, i32, i32, i32, p_void, p_void]
    L.etp_colsum_f32.argtypes = [p_void, i32, i32, i32, p_void, p_void]
    L.etp_cast_f32_to_bf16.argtypes = [p_void, p_void, C.c_int64, p_void]
    L.etp_pano_pack_fwd.argtypes = [C.POINTER(PanoPackArgs), C.c_void_p]
    L.etp_node_pack_fwd.argtypes = [C.POINTER(NodePackArgs), C.c_void_p]
    L.etp_sap_tail_fwd.argtypes = [p_void] * 7 + [i32, i32, p_void, p_void, p_void, p_void]


# ------------------------------------------------------------------------------------------------
# tensor-level wrappers (used by the parity tests; the planner module uses the same entry points)
# ------------------------------------------------------------------------------------------------
def gemm(A, B, *, a_mn=False, b_mn=False, alpha=1.0, bias=None, act=0, aux=None, aux_mode=0, resid=None,
         out_f32=None, out_bf16=None, out_pre=None, atomic=False, k_splits=1, block_n=0, M=None, N=None, K=None,
         colsum=None, pre_mode=0, drop=None):
    """D = epilogue(alpha * A @ B^T).  A: [M,K] (or [K,M] if a_mn), B: [N,K] (or [K,N] if b_mn), bf16, last dim contiguous."""
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16
    assert A.stride(-1) == 1 and B.stride(-1) == 1
    if M is None:
        M = A.shape[1] if a_mn else A.shape[0]
    if K is None:
        K = A.shape[0] if a_mn else A.shape[1]
    if N is None:
        N = B.shape[1] if b_mn else B.shape[0]
    g = GemmArgs()
    g.M, g.N, g.K = M, N, K
    g.A, g.lda, g.a_mn = ptr(A), A.stride(0), int(a_mn)
    g.B, g.ldb, g.b_mn = ptr(B), B.stride(0), int(b_mn)
    g.alpha = alpha
    g.bias = ptr(bias)
    g.act, g.aux_mode = act, aux_mode
    if aux is not None:
        g.aux, g.ld_aux = ptr(aux), aux.stride(0)
    if resid is not None:
        g.resid, g.ld_resid = ptr(resid), resid.stride(0)
    if out_f32 is not None:
        g.out_f32, g.ld_f32 = ptr(out_f32), out_f32.stride(0)
    if out_bf16 is not None:
        g.out_bf16, g.ld_bf16 = ptr(out_bf16), out_bf16.stride(0)
    if out_pre is not None:
        g.out_pre, g.ld_pre = ptr(out_pre), out_pre.stride(0)
    g.atomic, g.k_splits, g.block_n = int(atomic), k_splits, block_n
    g.colsum = ptr(colsum)
    g.pre_mode = pre_mode
    g.drop_scale = 1.0
    if drop is not None:  # (key, thr, scale) from dropout_site()
        g.drop_key, g.drop_thr, g.drop_scale = drop
    _check(lib().etp_gemm(C.byref(g), stream_ptr()), "etp_gemm")


def attention_fwd(q, k, v, out, *, B, heads, Sq, Sk, scale=0.125, key_valid=None, mask_value=-10000.0, pair=None,
                  pair_w=0.0, pair_b=0.0, lse=None, impl=0):
    """q/k/v/out: 2-D views [B*S, ld] (bf16) whose column h*64 starts head h."""
    a = AttnArgs()
    a.B, a.heads, a.Sq, a.Sk = B, heads, Sq, Sk
    a.q, a.ldq = ptr(q), q.stride(0)
    a.k, a.ldk = ptr(k), k.stride(0)
    a.v, a.ldv = ptr(v), v.stride(0)
    a.scale, a.key_valid, a.mask_value = scale, ptr(key_valid), mask_value
    a.pair, a.pair_w, a.pair_b = ptr(pair), pair_w, pair_b
    a.out, a.ldo, a.lse, a.impl = ptr(out), out.stride(0), ptr(lse), impl
    _check(lib().etp_attention_fwd(C.byref(a), stream_ptr()), "etp_attention_fwd")


def attention_bwd(q, k, v, out, dout, lse, dq, dk, dv, *, B, heads, Sq, Sk, scale=0.125, key_valid=None,
                  mask_value=-10000.0, pair=None, pair_w=0.0, pair_b=0.0, dpair_w=None, dpair_b=None, impl=0):
    a = AttnBwdArgs()
    a.B, a.heads, a.Sq, a.Sk = B, heads, Sq, Sk
    a.q, a.k, a.v, a.ldq, a.ldk, a.ldv = ptr(q), ptr(k), ptr(v), q.stride(0), k.stride(0), v.stride(0)
    a.out, a.ldo, a.dout, a.lddo, a.lse = ptr(out), out.stride(0), ptr(dout), dout.stride(0), ptr(lse)
    dvec = torch.empty(B * heads * Sq, device=q.device, dtype=torch.float32)
    a.dvec, a.scale, a.key_valid, a.mask_value = ptr(dvec), scale, ptr(key_valid), mask_value
    a.pair, a.pair_w, a.pair_b = ptr(pair), pair_w, pair_b
    a.dq, a.dk, a.dv, a.lddq, a.lddk, a.lddv = ptr(dq), ptr(dk), ptr(dv), dq.stride(0), dk.stride(0), dv.stride(0)
    a.dpair_w, a.dpair_b, a.impl = ptr(dpair_w), ptr(dpair_b), impl
    _check(lib().etp_attention_bwd(C.byref(a), stream_ptr()), "etp_attention_bwd")


def layernorm_fwd(x, gamma, beta, eps, y_f32=None, y_bf16=None, mean=None, rstd=None):
    rows, H = x.shape
    _check(lib().etp_layernorm_fwd(ptr(x), ptr(gamma), ptr(beta), eps, rows, H, ptr(y_f32), ptr(y_bf16), ptr(mean),
                                   ptr(rstd), stream_ptr()), "etp_layernorm_fwd")


def layernorm_bwd(dy, x, gamma, mean, rstd, dx_f32, accumulate_dx=False, dx_bf16=None, dgamma=None, dbeta=None):
    rows, H = x.shape
    _check(lib().etp_layernorm_bwd(ptr(dy), ptr(x), ptr(gamma), ptr(mean), ptr(rstd), rows, H, ptr(dx_f32),
                                   int(accumulate_dx), ptr(dx_bf16), ptr(dgamma), ptr(dbeta), stream_ptr()),
           "etp_layernorm_bwd")


def colsum(x, out):
    rows, cols = x.shape
    fn = lib().etp_colsum_bf16 if x.dtype == torch.bfloat16 else lib().etp_colsum_f32
    _check(fn(ptr(x), rows, cols, x.stride(0), ptr(out), stream_ptr()), "etp_colsum")


def cast_bf16(x, y):
    _check(lib().etp_cast_f32_to_bf16(ptr(x), ptr(y), x.numel(), stream_ptr()), "etp_cast_f32_to_bf16")


def pano_pack_fwd(**kw):
    a = PanoPackArgs()
    a.rows = kw.pop("rows")
    for k, v in kw.items():
        setattr(a, k, ptr(v))
    _check(lib().etp_pano_pack_fwd(C.byref(a), stream_ptr()), "etp_pano_pack_fwd")


def node_pack_fwd(**kw):
    a = NodePackArgs()
    a.rows = kw.pop("rows")
    for k, v in kw.items():
        setattr(a, k, ptr(v))
    _check(lib().etp_node_pack_fwd(C.byref(a), stream_ptr()), "etp_node_pack_fwd")


def step_loss(logits, labels, grad_scale=1.0, ignore_index=-100, want_probs=False):
    """Fused softmax / CE(sum, ignore_index) / dlogits / argmax over node logits [B, N] (ss_trainer_ETP.py:879-900).
    Returns (loss_sum [1], dlogits [B,N], argmax [B], probs [B,N] or None)."""
    B, N = logits.shape
    lg = logits.detach().float().contiguous()
    loss = torch.zeros(1, device=lg.device, dtype=torch.float32)
    dl = torch.empty_like(lg)
    am = torch.empty(B, device=lg.device, dtype=torch.int64)
    pr = torch.empty_like(lg) if want_probs else None
    L = lib()
    L.etp_step_loss.argtypes = [p_void, p_void, i32, i32, C.c_int64, f32, p_void, p_void, p_void, p_void, p_void]
    _check(L.etp_step_loss(ptr(lg), ptr(labels.contiguous().long() if labels is not None else None), B, N, ignore_index,
                           grad_scale, ptr(loss), ptr(dl), ptr(pr), ptr(am), stream_ptr()), "etp_step_loss")
    return loss, dl, am, pr


def sap_tail_fwd(relu_out, gamma, beta, w4, b4, visited, valid, logits, mean=None, rstd=None):
    rows, H = relu_out.shape
    _check(lib().etp_sap_tail_fwd(ptr(relu_out), ptr(gamma), ptr(beta), ptr(w4), ptr(b4), ptr(visited), ptr(valid),
                                  rows, H, ptr(logits), ptr(mean), ptr(rstd), stream_ptr()), "etp_sap_tail_fwd")


def split3(x, form):
    """fp32 [rows, K] -> bf16 [rows, 3K]: form 0 = hi|lo|hi (A operand), form 1 = hi|hi|lo (B operand) of the
    split-bf16 x3 product (etp_split3, high-precision mode)."""
    rows, K = x.shape
    y = torch.empty(rows, 3 * K, dtype=torch.bfloat16, device=x.device)
    L = lib()
    L.etp_split3.argtypes = [p_void, p_void, C.c_int64, i32, i32, p_void]
    _check(L.etp_split3(ptr(x.contiguous()), ptr(y), rows, K, form, stream_ptr()), "etp_split3")
    return y


def attention_f32_fwd(q, k, v, out, *, B, heads, Sq, Sk, scale=0.125, key_valid=None, mask_value=-10000.0, pair=None,
                      pair_w=0.0, pair_b=0.0):
    """fp32 attention of the high-precision mode; q/k/v/out: 2-D fp32 views [B*S, ld], head h at column h*64."""
    a = AttnArgs()
    a.B, a.heads, a.Sq, a.Sk = B, heads, Sq, Sk
    a.q, a.ldq = ptr(q), q.stride(0)
    a.k, a.ldk = ptr(k), k.stride(0)
    a.v, a.ldv = ptr(v), v.stride(0)
    a.scale, a.key_valid, a.mask_value = scale, ptr(key_valid), mask_value
    a.pair, a.pair_w, a.pair_b = ptr(pair), pair_w, pair_b
    a.out, a.ldo = ptr(out), out.stride(0)
    L = lib()
    L.etp_attention_f32_fwd.argtypes = [C.POINTER(AttnArgs), C.c_void_p]
    _check(L.etp_attention_f32_fwd(C.byref(a), stream_ptr()), "etp_attention_f32_fwd")
