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
        e = err[:128, :128]
        blk = e.view(16, 8, 16, 8).amax(dim=(1, 3))
        print("  8x8 block max-err map of the first 128x128 tile (rows = m blocks, cols = n blocks):")
        for r in range(16):
            print("   " + " ".join(f"{blk[r, c].item():7.2f}" for c in range(16)))
        # does the output match a permutation of ref columns / rows?
        o, r_ = out[:128, :128], ref[:128, :128]
        for name, cand in (("ref^T", r_.t()),):
            print(f"  vs {name}: {(o - cand).abs().max().item():.4g}")
        print("  out[0,:8]", out[0, :8].tolist())
        print("  ref[0,:8]", ref[0, :8].tolist())
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


if __name__ == "__main__":
    L.require_device()
    torch.backends.cuda.matmul.allow_tf32 = False
    grp = sys.argv[1]
    oks = [run(c, perf=(grp == "perf")) for c in GROUPS[grp]]
    print(f"group {grp}: {sum(oks)}/{len(oks)} ok", flush=True)
