"""GPU diagnostic (not a pytest): CTA-0 timeline of the tcgen05 attention backward at the c3 shapes."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from etpnav_b200 import lib as L  # noqa: E402

L.require_device()
lib = L.lib()
lib.etp_debug_attention_bwd_timeline.argtypes = [C.c_void_p]
for (B, Sq, Sk, pair_on) in [(64, 80, 200, False), (64, 80, 80, True)]:
    g = torch.Generator(device="cuda").manual_seed(1)
    h = 12
    q = (torch.randn(B * Sq, 768, generator=g, device="cuda") * 0.5).bfloat16()
    kv = (torch.randn(B * Sk, 1536, generator=g, device="cuda") * 0.5).bfloat16()
    k, v = kv[:, :768], kv[:, 768:]
    key_valid = torch.ones(B, Sk, dtype=torch.uint8, device="cuda")
    pair = torch.rand(B, Sq, Sk, generator=g, device="cuda") if pair_on else None
    out = torch.empty(B * Sq, 768, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B, h, Sq, device="cuda")
    L.attention_fwd(q, k, v, out, B=B, heads=h, Sq=Sq, Sk=Sk, key_valid=key_valid, pair=pair, pair_w=0.7, pair_b=-0.1, lse=lse)
    dout = torch.randn(B * Sq, 768, generator=g, device="cuda").bfloat16()
    dq = torch.zeros(B * Sq, 768, device="cuda", dtype=torch.bfloat16)
    dkv = torch.zeros(B * Sk, 1536, device="cuda", dtype=torch.bfloat16)
    dw, db = torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda")
    dbg = torch.zeros(128, dtype=torch.int64, device="cuda")

    def run():
        L.attention_bwd(q, k, v, out, dout, lse, dq, dkv[:, :768], dkv[:, 768:], B=B, heads=h, Sq=Sq, Sk=Sk,
                        key_valid=key_valid, pair=pair, pair_w=0.7, pair_b=-0.1,
                        dpair_w=dw if pair_on else None, dpair_b=db if pair_on else None, impl=2)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    print(f"B{B} Sq{Sq} Sk{Sk} pair={pair_on}: {e0.elapsed_time(e1) * 100:.1f} us per launch")
    lib.etp_debug_attention_bwd_timeline(C.c_void_p(dbg.data_ptr()))
    run()
    torch.cuda.synchronize()
    lib.etp_debug_attention_bwd_timeline(None)
    t = dbg.cpu().tolist()
    t0 = t[0]
    ctl = [(i, t[i] - t0) for i in range(64) if t[i]]
    mth = [(i, t[64 + i] - t0) for i in range(64) if t[64 + i]]
    names_c = {1: "tiles landed", 2: "S/dP issued+prefetch", 3: "P/dS ready", 0: "dV/dK/dQ issued"}
    print(" control (ns since start):")
    for i, dt in ctl[:26]:
        print(f"   step {(i - 1) // 4 if i else 0} {names_c[i % 4] if i else 'start':24s} {dt}")
    names_m = ["step begins", "past CTA barrier", "S/dP ready", "P/dS written", "dV/dK/dQ ready", "read-back stored"]
    print(" math thread 0:")
    for i, dt in mth[:36]:
        print(f"   step {i // 6} {names_m[i % 6]:24s} {dt}")
