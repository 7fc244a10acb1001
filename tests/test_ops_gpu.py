"""GPU: operator-level parity of the sm_100a kernels (through the C ABI) against plain PyTorch fp32
references of the same op on the same bf16-rounded inputs."""
import math

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(120)]


@pytest.fixture(scope="module")
def L():
    from etpnav_b200 import lib
    lib.require_device()
    torch.backends.cuda.matmul.allow_tf32 = False
    return lib


def _rand(shape, g, scale=1.0):
    return (torch.randn(shape, generator=g, device="cuda") * scale)


def _gen(seed):
    return torch.Generator(device="cuda").manual_seed(seed)


GEMM_CASES = [
    # M, N, K, a_mn, b_mn, block_n, k_splits
    (128, 128, 64, False, False, 128, 1),
    (256, 256, 128, False, False, 128, 1),
    (256, 256, 128, False, False, 256, 1),
    (384, 512, 768, False, False, 256, 1),
    (200, 200, 72, False, False, 128, 1),      # ragged M/N/K (TMA zero fill + masked stores)
    (130, 776, 3072, False, False, 256, 1),
    (256, 256, 128, False, True, 128, 1),      # dgrad form: B stored [K, N]
    (300, 768, 3072, False, True, 256, 1),
    (256, 256, 128, True, True, 128, 1),       # wgrad form: A stored [K, M], B stored [K, N]
    (768, 3072, 1000, True, True, 256, 4),     # split-K + atomic accumulate, ragged K
    (256, 128, 192, True, False, 128, 1),
    (5120, 768, 768, False, False, 0, 1),      # x-layer shapes at c3
    (5120, 3072, 768, False, False, 0, 1),
    (12800, 1536, 768, False, False, 0, 1),
]


@pytest.mark.parametrize("M,N,K,a_mn,b_mn,bn,ks", GEMM_CASES)
def test_gemm_plain(L, M, N, K, a_mn, b_mn, bn, ks):
    g = _gen(M * 7 + N * 3 + K)
    A = _rand((K, M) if a_mn else (M, K), g).bfloat16()
    B = _rand((K, N) if b_mn else (N, K), g).bfloat16()
    Af = A.float().t() if a_mn else A.float()
    Bf = B.float().t() if b_mn else B.float()
    ref = Af @ Bf.t()
    atomic = ks > 1
    out = torch.zeros(M, N, device="cuda") if atomic else torch.full((M, N), float("nan"), device="cuda")
    L.gemm(A, B, a_mn=a_mn, b_mn=b_mn, out_f32=out, atomic=atomic, k_splits=ks, block_n=bn)
    torch.cuda.synchronize()
    err = (out - ref).abs().max().item()
    tol = 2e-3 * math.sqrt(K / 64) + 1e-6 * K
    assert torch.isfinite(out).all()
    assert err < tol, f"max abs err {err} (tol {tol})"


def test_gemm_epilogues(L):
    g = _gen(5)
    M, N, K = 300, 776, 256
    A = _rand((M, K), g).bfloat16()
    W = _rand((N, K), g, 0.05).bfloat16()
    bias = _rand((N,), g)
    resid = _rand((M, N), g)
    lin = A.float() @ W.float().t() + bias
    # bias + gelu, bf16 out + pre-activation out
    o16 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    pre = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    L.gemm(A, W, bias=bias, act=1, out_bf16=o16, out_pre=pre)
    ref = torch.nn.functional.gelu(lin)
    assert (o16.float() - ref).abs().max() < 2e-2
    assert (pre.float() - lin).abs().max() < 3e-2
    # bias + residual, fp32 + bf16 out
    o32 = torch.empty(M, N, device="cuda")
    L.gemm(A, W, bias=bias, resid=resid, out_f32=o32, out_bf16=o16)
    assert (o32 - (lin + resid)).abs().max() < 2e-3
    assert (o16.float() - (lin + resid)).abs().max() < 3e-2
    # relu
    L.gemm(A, W, bias=bias, act=2, out_f32=o32)
    assert (o32 - lin.clamp_min(0)).abs().max() < 2e-3
    # alpha + gelu' multiply (dgrad of the FFN) and relu mask
    aux = _rand((M, N), g).bfloat16()
    L.gemm(A, W, alpha=0.5, aux=aux, aux_mode=1, out_f32=o32)
    x = aux.float().double().requires_grad_(True)
    torch.nn.functional.gelu(x).sum().backward()
    ref = 0.5 * (A.float() @ W.float().t()) * x.grad.float()
    assert (o32 - ref).abs().max() < 3e-3
    L.gemm(A, W, aux=aux, aux_mode=2, out_f32=o32)
    ref = (A.float() @ W.float().t()) * (aux.float() > 0)
    assert (o32 - ref).abs().max() < 3e-3
    # gelu that also emits gelu'(pre) for the backward (pre_mode=1), then the backward's plain multiply by the saved
    # derivative (aux_mode=3) with the fused bias gradient (column sums)
    dact = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    L.gemm(A, W, bias=bias, act=1, out_bf16=o16, out_pre=dact, pre_mode=1)
    xd = lin.double().requires_grad_(True)
    torch.nn.functional.gelu(xd).sum().backward()
    assert (o16.float() - torch.nn.functional.gelu(lin)).abs().max() < 2e-2
    assert (dact.float() - xd.grad.float()).abs().max() < 1e-2
    cs = torch.zeros(N, device="cuda")
    L.gemm(A, W, aux=dact, aux_mode=3, out_f32=o32, colsum=cs)
    ref = (A.float() @ W.float().t()) * dact.float()
    assert (o32 - ref).abs().max() < 3e-3
    assert (cs - ref.sum(0)).abs().max() < 2e-2
    # strided output (write into a column slice of a wider buffer)
    wide = torch.zeros(M, 2 * N, device="cuda", dtype=torch.bfloat16)
    L.gemm(A, W, bias=bias, out_bf16=wide[:, N:])
    assert (wide[:, N:].float() - lin).abs().max() < 3e-2
    assert wide[:, :N].abs().max() == 0
    torch.cuda.synchronize()


@pytest.mark.parametrize("rows", [1, 7, 768, 5120])
def test_layernorm_fwd_bwd(L, rows):
    g = _gen(rows)
    H = 768
    x = _rand((rows, H), g, 2.0) + 0.5
    gamma = 1 + 0.1 * _rand((H,), g)
    beta = 0.1 * _rand((H,), g)
    for eps in (1e-12, 1e-5):
        y = torch.empty_like(x)
        y16 = torch.empty(rows, H, device="cuda", dtype=torch.bfloat16)
        mean = torch.empty(rows, device="cuda")
        rstd = torch.empty(rows, device="cuda")
        L.layernorm_fwd(x, gamma, beta, eps, y, y16, mean, rstd)
        xr = x.clone().requires_grad_(True)
        gr = gamma.clone().requires_grad_(True)
        br = beta.clone().requires_grad_(True)
        ref = torch.nn.functional.layer_norm(xr, (H,), gr, br, eps)
        assert (y - ref).abs().max() < 1e-5
        assert (y16.float() - ref).abs().max() < 3e-2
        dy = _rand((rows, H), g)
        ref.backward(dy)
        dx = torch.empty_like(x)
        dx16 = torch.empty(rows, H, device="cuda", dtype=torch.bfloat16)
        dg = torch.zeros(H, device="cuda")
        db = torch.zeros(H, device="cuda")
        L.layernorm_bwd(dy, x, gamma, mean, rstd, dx, False, dx16, dg, db)
        assert (dx - xr.grad).abs().max() < 2e-4 * max(1.0, xr.grad.abs().max().item())
        assert (dg - gr.grad).abs().max() < 1e-3 * max(1.0, gr.grad.abs().max().item())
        assert (db - br.grad).abs().max() < 1e-3 * max(1.0, br.grad.abs().max().item())
        L.layernorm_bwd(dy, x, gamma, mean, rstd, dx, True, None, None, None)
        assert (dx - 2 * xr.grad).abs().max() < 4e-4 * max(1.0, xr.grad.abs().max().item())


def test_colsum_and_cast(L):
    g = _gen(3)
    x = _rand((1000, 776), g)
    out = torch.zeros(776, device="cuda")
    L.colsum(x, out)
    assert (out - x.sum(0)).abs().max() < 1e-3
    xb = x.bfloat16()
    out.zero_()
    L.colsum(xb[:, :770], out)
    assert (out[:770] - xb[:, :770].float().sum(0)).abs().max() < 1e-2
    y = torch.empty(1000 * 776 + 3, device="cuda", dtype=torch.bfloat16)
    xx = _rand((1000 * 776 + 3,), g)
    L.cast_bf16(xx, y)
    assert torch.equal(y, xx.bfloat16())


def _attn_ref(q, k, v, B, h, Sq, Sk, scale, key_valid, mask_value, pair, pw, pb):
    qh = q.float().view(B, Sq, h, 64).permute(0, 2, 1, 3)
    kh = k.float().view(B, Sk, h, 64).permute(0, 2, 1, 3)
    vh = v.float().view(B, Sk, h, 64).permute(0, 2, 1, 3)
    s = qh @ kh.transpose(-1, -2) * scale
    if key_valid is not None:
        s = s + torch.where(key_valid.bool(), 0.0, mask_value)[:, None, None, :]
    if pair is not None:
        s = s + (pair * pw + pb)[:, None]
    p = torch.softmax(s, -1)
    o = (p @ vh).permute(0, 2, 1, 3).reshape(B * Sq, h * 64)
    return o, torch.logsumexp(s, -1)


@pytest.mark.parametrize("impl", [1, 3, 4])   # CUDA-core, tcgen05 two-CTAs-per-SM (attention_tc2.cu), tcgen05 one-CTA (attention_tc.cu)
@pytest.mark.parametrize("B,Sq,Sk,mode", [(3, 12, 12, "pano"), (2, 16, 80, "x"), (4, 40, 40, "self"),
                                           (2, 80, 200, "x"), (2, 80, 80, "self"), (1, 120, 512, "x"),
                                           (2, 150, 150, "self"), (3, 11, 17, "x"), (2, 200, 200, "txt"),
                                           (30, 80, 200, "x"), (30, 80, 80, "self")])   # > 2 x 148 items: every CTA walks several
def test_attention_fwd(L, impl, B, Sq, Sk, mode):
    g = _gen(B * 100 + Sq + Sk)
    h = 12
    q = _rand((B * Sq, 768), g).bfloat16()
    kv = _rand((B * Sk, 1536), g).bfloat16()
    k, v = kv[:, :768], kv[:, 768:]
    lens = torch.randint(max(1, Sk // 2), Sk + 1, (B,), generator=g, device="cuda")
    lens[0] = Sk
    key_valid = (torch.arange(Sk, device="cuda")[None] < lens[:, None])
    mask_value = float("-inf") if mode == "pano" else -10000.0
    pair = _rand((B, Sq, Sk), g).abs() if mode == "self" else None
    out = torch.empty(B * Sq, 768, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B, h, Sq, device="cuda")
    L.attention_fwd(q, k, v, out, B=B, heads=h, Sq=Sq, Sk=Sk, scale=0.125, key_valid=key_valid.to(torch.uint8),
                    mask_value=mask_value, pair=pair, pair_w=0.7, pair_b=-0.1, lse=lse, impl=impl)
    ref, lse_ref = _attn_ref(q, k, v, B, h, Sq, Sk, 0.125, key_valid, mask_value, pair, 0.7, -0.1)
    assert (out.float() - ref).abs().max() < 2e-2
    assert (lse - lse_ref).abs().max() < 1e-3


@pytest.mark.parametrize("impl", [1, 2])
@pytest.mark.parametrize("B,Sq,Sk,mode", [(2, 16, 80, "x"), (3, 40, 40, "self"), (2, 80, 200, "x"), (2, 80, 80, "self"),
                                           (1, 120, 300, "x"), (2, 12, 12, "pano"), (2, 128, 128, "self")])
def test_attention_bwd(L, impl, B, Sq, Sk, mode):
    if impl == 2 and Sq < 32 and Sk < 32:
        pytest.skip("tensor-core backward is not used for tiny problems")
    g = _gen(B * 1000 + Sq + Sk)
    h = 12
    q = (_rand((B * Sq, 768), g) * 0.5).bfloat16()
    kv = (_rand((B * Sk, 1536), g) * 0.5).bfloat16()
    k, v = kv[:, :768], kv[:, 768:]
    lens = torch.randint(max(1, Sk // 2), Sk + 1, (B,), generator=g, device="cuda")
    lens[0] = Sk
    key_valid = (torch.arange(Sk, device="cuda")[None] < lens[:, None])
    mask_value = float("-inf") if mode == "pano" else -10000.0
    pair = _rand((B, Sq, Sk), g).abs() if mode == "self" else None
    out = torch.empty(B * Sq, 768, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B, h, Sq, device="cuda")
    L.attention_fwd(q, k, v, out, B=B, heads=h, Sq=Sq, Sk=Sk, key_valid=key_valid.to(torch.uint8), mask_value=mask_value,
                    pair=pair, pair_w=0.7, pair_b=-0.1, lse=lse, impl=1)
    dout = (_rand((B * Sq, 768), g)).bfloat16()
    dq = torch.zeros(B * Sq, 768, device="cuda", dtype=torch.bfloat16)
    dkv = torch.zeros(B * Sk, 1536, device="cuda", dtype=torch.bfloat16)
    dw = torch.zeros(1, device="cuda")
    db = torch.zeros(1, device="cuda")
    L.attention_bwd(q, k, v, out, dout, lse, dq, dkv[:, :768], dkv[:, 768:], B=B, heads=h, Sq=Sq, Sk=Sk,
                    key_valid=key_valid.to(torch.uint8), mask_value=mask_value, pair=pair, pair_w=0.7, pair_b=-0.1,
                    dpair_w=dw if pair is not None else None, dpair_b=db if pair is not None else None, impl=impl)
    qr, kr, vr = (t.float().clone().requires_grad_(True) for t in (q, k, v))
    pw = torch.tensor(0.7, device="cuda", requires_grad=True)
    pb = torch.tensor(-0.1, device="cuda", requires_grad=True)
    qh = qr.view(B, Sq, h, 64).permute(0, 2, 1, 3)
    kh = kr.view(B, Sk, h, 64).permute(0, 2, 1, 3)
    vh = vr.view(B, Sk, h, 64).permute(0, 2, 1, 3)
    s = qh @ kh.transpose(-1, -2) * 0.125 + torch.where(key_valid, 0.0, mask_value)[:, None, None, :]
    if pair is not None:
        s = s + (pair * pw + pb)[:, None]
    o = (torch.softmax(s, -1) @ vh).permute(0, 2, 1, 3).reshape(B * Sq, 768)
    o.backward(dout.float())
    rel = lambda a, b: ((a.float() - b).norm() / b.norm().clamp_min(1e-9)).item()
    assert rel(dq, qr.grad) < 2e-2, rel(dq, qr.grad)
    assert rel(dkv[:, :768], kr.grad) < 2e-2, rel(dkv[:, :768], kr.grad)
    assert rel(dkv[:, 768:], vr.grad) < 2e-2, rel(dkv[:, 768:], vr.grad)
    if pair is not None:
        assert abs(dw.item() - pw.grad.item()) < 2e-2 * max(1.0, abs(pw.grad.item())), (dw.item(), pw.grad.item())
        assert abs(db.item()) < 0.05 * max(1.0, abs(pw.grad.item()))


def test_step_loss(L):
    """Fused softmax / CE(sum, ignore_index) / gradient / argmax vs torch (ss_trainer_ETP.py:879-900)."""
    g = _gen(77)
    B, N = 37, 83
    logits = _rand((B, N), g, 3.0)
    dead = torch.rand(B, N, generator=g, device="cuda") < 0.3
    dead[:, 0] = False
    logits = logits.masked_fill(dead, float("-inf"))
    labels = torch.randint(0, N, (B,), generator=g, device="cuda")
    labels = torch.where(dead[torch.arange(B, device="cuda"), labels], torch.zeros_like(labels), labels)
    labels[3] = -100
    labels[11] = -100
    ref_in = logits.clone().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(ref_in, labels, reduction="sum", ignore_index=-100)
    (ref / B).backward()
    loss, dl, am, pr = L.step_loss(logits, labels, grad_scale=1.0 / B, want_probs=True)
    assert abs(loss.item() - ref.item()) < 1e-3 * max(1.0, abs(ref.item()))
    assert (dl - torch.nan_to_num(ref_in.grad, nan=0.0)).abs().max() < 1e-6
    assert torch.equal(am, logits.argmax(1))
    assert (pr - torch.softmax(logits, 1)).abs().max() < 1e-6
