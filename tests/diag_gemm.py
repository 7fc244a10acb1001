"""GPU diagnostic (not a pytest): run tcgen05 GEMM variants one by one and print error maps + timings.

    python tests/diag_gemm.py <group>     group in {basic, kk, bmn, amn, perf}
"""
import math
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from etpnav_b200 import lib as L  # noqa: E402

GROUPS = {
    "basic": [(128, 128, 64, False, False, 128, 1)],
    "kk": [(256, 256, 128, False, False, 128, 1), (256, 256, 128, False, False, 256, 1),
           (384, 512, 768, False, False, 256, 1), (200, 200, 72, False, False, 128, 1)],
    "bmn": [(128, 128, 64, False, True, 128, 1), (256, 256, 128, False, True, 128, 1),
            (256, 512, 256, False, True, 256, 1)],
    "amn": [(128, 128, 64, True, False, 128, 1), (128, 128, 64, True, True, 128, 1),
            (256, 256, 128, True, True, 128, 1), (768, 768, 1024, True, True, 128, 4)],
    "perf": [(5120, 768, 768, False, False, 0, 1), (5120, 3072, 768, False, False, 0, 1),
             (5120, 768, 3072, False, False, 0, 1), (12800, 1536, 768, False, False, 0, 1),
             (5120, 2304, 768, False, False, 0, 1), (5120, 768, 768, False, False, 128, 1),
             (5120, 3072, 768, False, False, 128, 1),
             (5120, 768, 3072, False, True, 0, 1), (768, 3072, 5120, True, True, 0, 8),
             (768, 768, 5120, True, True, 128, 4)],
}


def run(case, perf=False):
    M, N, K, a_mn, b_mn, bn, ks = case
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn((K, M) if a_mn else (M, K), generator=g, device="cuda").bfloat16()
    B = torch.randn((K, N) if b_mn else (N, K), generator=g, device="cuda").bfloat16()
    Af = A.float().t() if a_mn else A.float()
    Bf = B.float().t() if b_mn else B.float()
    ref = Af @ Bf.t()
    atomic = ks > 1
    out = torch.zeros(M, N, device="cuda")
    L.gemm(A, B, a_mn=a_mn, b_mn=b_mn, out_f32=out, atomic=atomic, k_splits=ks, block_n=bn)
    torch.cuda.synchronize()
    err = (out - ref).abs()
    tol = 2e-3 * math.sqrt(K / 64) + 1e-6 * K
    ok = bool(err.max() < tol) and bool(torch.isfinite(out).all())
    print(f"case {case}: max_err {err.max().item():.4g} tol {tol:.3g} ref_absmax {ref.abs().max().item():.3g} "
          f"{'OK' if ok else 'FAIL'}", flush=True)
    if not ok:
        R, Cc = min(M, 512), min(N, 512)
        rb, cb = (R + 31) // 32, (Cc + 31) // 32
        print(f"  32x32 block max-err map of the first {R}x{Cc} region (rows = m blocks, cols = n blocks):")
        for r in range(rb):
            print("   " + " ".join(f"{err[r * 32:(r + 1) * 32, c * 32:(c + 1) * 32].max().item():7.2f}" for c in range(cb)))
        print("  out[0,:8]", out[0, :8].tolist())
        print("  ref[0,:8]", ref[0, :8].tolist())
        if M > 128:
            print("  out[128,:8]", out[128, :8].tolist())
            print("  ref[128,:8]", ref[128, :8].tolist())
    if perf:
        for _ in range(3):
            L.gemm(A, B, a_mn=a_mn, b_mn=b_mn, out_f32=out, atomic=atomic, k_splits=ks, block_n=bn)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            L.gemm(A, B, a_mn=a_mn, b_mn=b_mn, out_f32=out, atomic=atomic, k_splits=ks, block_n=bn)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        tf = 2.0 * M * N * K / ms / 1e9
        t0 = time.perf_counter()
        for _ in range(n):
            ref = Af @ Bf.t()
        torch.cuda.synchronize()
        print(f"  time {ms * 1e3:.1f} us  {tf:.1f} TFLOP/s", flush=True)
        Ab, Bb = (A.t().contiguous() if a_mn else A), (B.t().contiguous() if b_mn else B)
        e0.record()
        for _ in range(n):
            torch.matmul(Ab, Bb.t())
        e1.record()
        torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / n
        print(f"  cuBLAS bf16 (torch.matmul) {ms2 * 1e3:.1f} us  {2.0 * M * N * K / ms2 / 1e9:.1f} TFLOP/s", flush=True)
    return ok


def run_epi():
    """The planner's actual epilogues at the c3 shapes: timing only (parity is tests/test_ops_gpu.py)."""
    g = torch.Generator(device="cuda").manual_seed(5)
    M = 5120
    def rb(*shape):
        return torch.randn(*shape, generator=g, device="cuda").bfloat16()
    cases = []
    x768, x3072 = rb(M, 768), rb(M, 3072)
    resid = torch.randn(M, 768, generator=g, device="cuda")
    b768, b3072, b2304 = torch.randn(768, device="cuda"), torch.randn(3072, device="cuda"), torch.randn(2304, device="cuda")
    o32 = torch.empty(M, 768, device="cuda")
    W11, W13, W31, W1q = rb(768, 768), rb(3072, 768), rb(768, 3072), rb(2304, 768)
    cases.append(("proj 768->768 +bias +resid -> f32", 768, 768, lambda: L.gemm(x768, W11, bias=b768, resid=resid, out_f32=o32)))
    h, pre = torch.empty(M, 3072, device="cuda", dtype=torch.bfloat16), torch.empty(M, 3072, device="cuda", dtype=torch.bfloat16)
    qkv = torch.empty(M, 2304, device="cuda", dtype=torch.bfloat16)
    dpre = torch.empty(M, 3072, device="cuda", dtype=torch.bfloat16)
    cases.append(("ffn1 768->3072 +bias gelu -> bf16 + pre", 3072, 768, lambda: L.gemm(x768, W13, bias=b3072, act=1, out_bf16=h, out_pre=pre, pre_mode=1)))
    cases.append(("ffn2 3072->768 +bias +resid -> f32", 768, 3072, lambda: L.gemm(x3072, W31, bias=b768, resid=resid, out_f32=o32)))
    cases.append(("qkv 768->2304 +bias -> bf16", 2304, 768, lambda: L.gemm(x768, W1q, bias=b2304, out_bf16=qkv)))
    cases.append(("dgrad ffn2 768->3072 *gelu'(pre) -> bf16", 3072, 768, lambda: L.gemm(x768, W31, b_mn=True, aux=x3072, aux_mode=3, out_bf16=dpre, colsum=b3072)))
    cases.append(("dgrad ffn1 3072->768 +resid -> f32", 768, 3072, lambda: L.gemm(x3072, W13, b_mn=True, resid=resid, out_f32=o32)))
    for name, N, K, fn in cases:
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"epi {name}: {ms * 1e3:.1f} us  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    L.require_device()
    torch.backends.cuda.matmul.allow_tf32 = False
    grp = sys.argv[1]
    if grp == "epi":
        run_epi()
        sys.exit(0)
    oks = [run(c, perf=(grp == "perf")) for c in GROUPS[grp]]
    print(f"group {grp}: {sum(oks)}/{len(oks)} ok", flush=True)
