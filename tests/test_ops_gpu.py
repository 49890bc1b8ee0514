"""Kernel unit tests (GPU): every C-ABI op against a plain torch restatement of the same op on the same inputs.
GEMM tolerances: 1xTF32 vs the TF32-truncated fp64 product (bit-level operand semantics), 3xTF32 / fp32 vs fp64.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from capdec_b200 import ops as _ops
    _ops.set_precision("tf32")
    return _ops


def trunc_tf32(x):
    return (x.contiguous().view(torch.int32) & -8192).view(torch.float32)


def make_ab(M, N, K, a_major, b_major, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    pad = lambda n: (n + 3) // 4 * 4
    A = (torch.randn(K, pad(M), device="cuda", generator=g)[:, :M] if a_major
         else torch.randn(M, pad(K), device="cuda", generator=g)[:, :K])
    B = (torch.randn(K, pad(N), device="cuda", generator=g)[:, :N] if b_major
         else torch.randn(N, pad(K), device="cuda", generator=g)[:, :K])
    return A, (A.t() if a_major else A), B, (B.t() if b_major else B)


def new_c(M, N):
    return torch.full((M, (N + 3) // 4 * 4), float("nan"), device="cuda")[:, :N]


@pytest.mark.parametrize("a_major,b_major", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("shape", [(128, 256, 32), (200, 300, 100), (1600, 2304, 768), (384, 50257, 64),
                                   (12800, 768, 96), (3000, 1100, 40), (5000, 2304, 64)])
def test_gemm_tf32_matches_truncated_product(ops, shape, a_major, b_major):
    M, N, K = shape
    A, Al, B, Bl = make_ab(M, N, K, a_major, b_major)
    C = new_c(M, N)
    ops.gemm(A, a_major, B, b_major, C, M, N, K)
    ref = trunc_tf32(Al).double() @ trunc_tf32(Bl).double().t()
    err = (C.double() - ref).abs()
    bad = (err > 1e-4 * ref.abs().max()).nonzero()
    assert bad.numel() == 0, f"{bad.shape[0]} bad elements, first {bad[:5].tolist()}, rows {bad[:,0].min().item()}..{bad[:,0].max().item()} cols {bad[:,1].min().item()}..{bad[:,1].max().item()}"


@pytest.mark.parametrize("pair", [0, 1, 2, 3])
@pytest.mark.parametrize("a_major,b_major", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("shape", [(256, 256, 32), (12800, 2304, 768), (1000, 50257, 96), (768, 2304, 3200), (300, 200, 64),
                                   (1300, 768, 3072), (520, 1030, 200)])
def test_gemm_both_tile_engines(ops, shape, a_major, b_major, pair):
    """All tile engines give the same numbers: 0 = cta_group::1 (one CTA per 128-row tile), 1 = cta_group::2 CTA pair,
    2 / 3 = cluster of two pairs with the shared A / B operand TMA-multicast."""
    from capdec_b200 import _lib
    M, N, K = shape
    A, Al, B, Bl = make_ab(M, N, K, a_major, b_major)
    C = new_c(M, N)
    _lib.load().capdec_gemm_debug_force_pair(pair)
    try:
        ops.gemm(A, a_major, B, b_major, C, M, N, K)
        torch.cuda.synchronize()
    finally:
        _lib.load().capdec_gemm_debug_force_pair(-1)
    ref = trunc_tf32(Al).double() @ trunc_tf32(Bl).double().t()
    err = (C.double() - ref).abs()
    bad = (err > 1e-4 * ref.abs().max()).nonzero()
    assert bad.numel() == 0, f"{bad.shape[0]} bad elements, rows {bad[:,0].min().item()}..{bad[:,0].max().item()} cols {bad[:,1].min().item()}..{bad[:,1].max().item()}"


@pytest.mark.parametrize("bn", [64, 128, 256])
def test_gemm_block_n_variants(ops, bn):
    M, N, K = 2500, 1000, 200
    A, Al, B, Bl = make_ab(M, N, K, 0, 1)
    C = new_c(M, N)
    ops.gemm(A, 0, B, 1, C, M, N, K, block_n=bn)
    ref = trunc_tf32(Al).double() @ trunc_tf32(Bl).double().t()
    assert (C.double() - ref).abs().max() < 1e-4 * ref.abs().max()


@pytest.mark.parametrize("act", [0, 1, 2, 3])
def test_gemm_epilogue_bias_act_aux(ops, act):
    M, N, K = 700, 900, 256
    A, Al, B, Bl = make_ab(M, N, K, 0, 0)
    bias = torch.randn(N, device="cuda")
    C, aux = new_c(M, N), new_c(M, N)
    ops.gemm(A, 0, B, 0, C, M, N, K, bias=bias, act=act, aux=aux)
    pre = (trunc_tf32(Al).double() @ trunc_tf32(Bl).double().t() + bias.double())
    f = {0: lambda x: x, 1: lambda x: torch.nn.functional.gelu(x, approximate="tanh"), 2: torch.tanh, 3: torch.relu}[act]
    assert (aux.double() - pre).abs().max() < 1e-3
    assert (C.double() - f(pre)).abs().max() < 1e-3


@pytest.mark.parametrize("split", [0, 1, 3, 8])
def test_gemm_accumulate_splitk(ops, split):
    M, N, K = 768, 2304, 4096
    A, Al, B, Bl = make_ab(M, N, K, 1, 1)
    C0 = torch.randn(M, N, device="cuda")
    C = C0.clone()
    ops.gemm(A, 1, B, 1, C, M, N, K, accumulate=True, split_k=split)
    ref = C0.double() + trunc_tf32(Al).double() @ trunc_tf32(Bl).double().t()
    assert (C.double() - ref).abs().max() < 2e-4 * ref.abs().max()


@pytest.mark.parametrize("mode,tol", [("tf32x3", 2e-5), ("fp32", 3e-6)])
@pytest.mark.parametrize("a_major,b_major", [(0, 0), (0, 1), (1, 1)])
def test_gemm_fp32_grade_modes(ops, mode, tol, a_major, b_major):
    M, N, K = 512, 770, 768
    A, Al, B, Bl = make_ab(M, N, K, a_major, b_major)
    C = new_c(M, N)
    ops.gemm(A, a_major, B, b_major, C, M, N, K, precision=mode)
    ref = Al.double() @ Bl.double().t()
    rel = (C.double() - ref).abs().max() / ref.abs().max()
    assert rel < tol, f"{mode}: rel err {rel:.3e}"


def test_gemm_rejects_bad_arguments(ops):
    from capdec_b200._lib import CapdecError
    A = torch.randn(64, 30, device="cuda")  # pitch 30 floats is not 16-byte aligned
    B = torch.randn(64, 30, device="cuda")
    C = torch.empty(64, 64, device="cuda")
    with pytest.raises(CapdecError):
        ops.gemm(A, 0, B, 0, C, 64, 64, 30)
    with pytest.raises(CapdecError):
        ops.gemm(A.cpu(), 0, B, 0, C, 64, 64, 30)


# ---------------------------------------------------------------------------------------------------------------
def ref_noise_injection(x, variance, noise, offset=None, dont_norm=False):
    """train.py:27-39 with the Gaussian draw passed in."""
    if variance == 0.0:
        return x
    if not dont_norm:
        x = torch.nn.functional.normalize(x, dim=1)
    x = x + noise
    if offset is not None:
        x = x + offset
    return torch.nn.functional.normalize(x, dim=1)


@pytest.mark.parametrize("D", [512, 640])
@pytest.mark.parametrize("dont_norm,with_offset", [(False, False), (True, False), (False, True)])
def test_noise_injection_with_supplied_noise(ops, D, dont_norm, with_offset):
    B = 37
    x = torch.randn(B, D, device="cuda") * 3
    noise = torch.randn(B, D, device="cuda") * math.sqrt(0.016)
    offset = torch.randn(1, D, device="cuda") * 0.1 if with_offset else None
    out = torch.empty_like(x)
    ops.noise_injection(x, out, 0.016, noise=noise, offset=offset, dont_norm=dont_norm)
    ref = ref_noise_injection(x, 0.016, noise, offset, dont_norm)
    assert (out - ref).abs().max() < 2e-6
    # variance == 0 -> identity, NOT normalised (train.py:28-29)
    ops.noise_injection(x, out, 0.0)
    assert torch.equal(out, x)


def test_noise_injection_philox_statistics(ops):
    B, D = 4096, 512
    x = torch.nn.functional.normalize(torch.randn(B, D, device="cuda"), dim=1)
    out = torch.empty_like(x)
    var = 0.016
    ops.noise_injection(x, out, var, seed=ops.make_seed(123), step=5)
    assert torch.allclose(out.norm(dim=1), torch.ones(B, device="cuda"), atol=1e-5)
    # out ~ (x + n)/|x + n| ; E|x+n|^2 = 1 + D*var  => cosine(out, x) ~ 1/sqrt(1 + D var)
    cos = (out * x).sum(1).mean().item()
    assert abs(cos - 1.0 / math.sqrt(1 + D * var)) < 5e-3
    out2 = torch.empty_like(x)
    ops.noise_injection(x, out2, var, seed=ops.make_seed(123), step=5)
    assert torch.equal(out, out2)  # counter-based: reproducible
    ops.noise_injection(x, out2, var, seed=ops.make_seed(123), step=6)
    assert not torch.equal(out, out2)
    # uniform ball: |noise| <= radius
    ops.noise_injection(x, out2, var, uniform_ball=True, dont_norm=True, seed=ops.make_seed(1), step=0)
    assert torch.isfinite(out2).all()


def test_embed_fwd_bwd(ops):
    B, P, L, d, V = 5, 10, 40, 768, 50257
    g = torch.Generator(device="cuda").manual_seed(1)
    tokens = torch.randint(0, V, (B, L), device="cuda", generator=g)
    tokens[0, 30:] = 0
    wte = torch.randn(V, d, device="cuda") * 0.02
    wpe = torch.randn(1024, d, device="cuda") * 0.02
    pp = torch.randn(B, P, d, device="cuda")
    h = torch.empty(B, P + L, d, device="cuda")
    ops.embed_fwd(tokens, pp, wte, wpe, h, B, P, L)
    ref = torch.cat([pp, wte[tokens]], dim=1) + wpe[: P + L]
    assert torch.equal(h, ref)
    dh = torch.randn(B, P + L, d, device="cuda")
    dpp = torch.empty(B, P, d, device="cuda")
    dwte = torch.zeros(V, d, device="cuda")
    dwpe = torch.zeros(1024, d, device="cuda")
    ops.embed_bwd(tokens, dh, dpp, dwte, dwpe, B, P, L, V)
    assert torch.equal(dpp, dh[:, :P])
    ref_wte = torch.zeros(V, d, device="cuda", dtype=torch.float64).index_add_(0, tokens.flatten(), dh[:, P:].reshape(-1, d).double())
    assert (dwte.double() - ref_wte).abs().max() < 1e-5
    assert (dwpe[: P + L].double() - dh.double().sum(0)).abs().max() < 1e-5
    assert dwpe[P + L:].abs().max() == 0


@pytest.mark.parametrize("rows,d", [(1000, 768), (77, 768), (64, 512), (33, 1024)])
def test_add_ln_fwd_bwd(ops, rows, d):
    h = torch.randn(rows, d, device="cuda")
    y = torch.randn(rows, d, device="cuda")
    gamma = torch.randn(d, device="cuda")
    beta = torch.randn(d, device="cuda")
    x = torch.empty(rows, d, device="cuda")
    h_out = torch.empty(rows, d, device="cuda")
    stats = torch.empty(rows, 2, device="cuda")
    ops.add_ln_fwd(h, y, h_out, x, stats, gamma, beta)
    hd_, yd, gd, bd = h.double().requires_grad_(), y.double().requires_grad_(), gamma.double().requires_grad_(), beta.double().requires_grad_()
    r = hd_ + yd
    xr = torch.nn.functional.layer_norm(r, (d,), gd, bd, 1e-5)
    assert (h_out.double() - r).abs().max() < 1e-6
    assert (x.double() - xr).abs().max() < 2e-5
    dx = torch.randn(rows, d, device="cuda")
    dres = torch.randn(rows, d, device="cuda")
    (xr * dx.double()).sum().backward()
    dh_out = torch.empty(rows, d, device="cuda")
    dy = torch.empty(rows, d, device="cuda")
    dg = torch.zeros(d, device="cuda")
    db = torch.zeros(d, device="cuda")
    dbr = torch.zeros(d, device="cuda")
    ops.add_ln_bwd(dx, h_out, stats, gamma, dres, dh_out, dy, dg, db, dbias_branch=dbr)
    ref_dr = hd_.grad + dres.double()
    assert (dbr.double() - ref_dr.sum(0)).abs().max() < 2e-3 * max(1.0, ref_dr.sum(0).abs().max().item())
    assert (dh_out.double() - ref_dr).abs().max() < 1e-4
    assert torch.equal(dy, dh_out)
    assert (dg.double() - gd.grad).abs().max() < 2e-3 * max(1.0, gd.grad.abs().max().item())
    assert (db.double() - bd.grad).abs().max() < 2e-3 * max(1.0, bd.grad.abs().max().item())
    # plain LN (no branch), no residual grad, frozen params
    ops.add_ln_fwd(h, None, None, x, stats, gamma, beta)
    assert (x.double() - torch.nn.functional.layer_norm(h.double(), (d,), gamma.double(), beta.double(), 1e-5)).abs().max() < 2e-5
    ops.add_ln_bwd(dx, h, stats, gamma, None, dh_out, None, None, None)


def test_add_ln_dropout_mask_consistency(ops):
    rows, d, p = 512, 768, 0.1
    seed7 = ops.make_seed(7)
    h = torch.zeros(rows, d, device="cuda")
    y = torch.ones(rows, d, device="cuda")
    gamma, beta = torch.ones(d, device="cuda"), torch.zeros(d, device="cuda")
    x, h_out, stats = torch.empty(rows, d, device="cuda"), torch.empty(rows, d, device="cuda"), torch.empty(rows, 2, device="cuda")
    ops.add_ln_fwd(h, y, h_out, x, stats, gamma, beta, p_drop=p, seed=seed7, stream_id=3)
    keep = h_out != 0
    assert abs(keep.float().mean().item() - (1 - p)) < 5e-3
    assert torch.allclose(h_out[keep], torch.full_like(h_out[keep], 1 / (1 - p)))
    dy, dh_out = torch.empty(rows, d, device="cuda"), torch.empty(rows, d, device="cuda")
    ops.add_ln_bwd(torch.zeros(rows, d, device="cuda"), h_out, stats, gamma, torch.ones(rows, d, device="cuda"), dh_out, dy,
                   None, None, p_drop=p, seed=seed7, stream_id=3)
    assert torch.equal(dy != 0, keep)  # backward regenerates exactly the forward mask
    ops.add_ln_fwd(h, y, h_out, x, stats, gamma, beta, p_drop=p, seed=seed7, stream_id=4)
    assert not torch.equal(h_out != 0, keep)  # a different site draws a different mask


def ref_attention(q, k, v, scale, causal, key_len=None):
    # q [B,H,T,hd], k,v [B,H,S,hd]; HF eager_attention_forward (modeling_gpt2.py:54-72) / train.py:150-167
    s = (q @ k.transpose(-1, -2)) * scale
    T, S = s.shape[-2:]
    if causal:
        m = torch.ones(T, S, dtype=torch.bool, device=q.device).tril(S - T)
        s = s.masked_fill(~m, float("-inf"))
    if key_len is not None:
        km = torch.arange(S, device=q.device)[None, :] < key_len[:, None]
        s = s.masked_fill(~km[:, None, None, :], float("-inf"))
    return torch.softmax(s, dim=-1) @ v


@pytest.mark.parametrize("impl", ["ffma", "tc"])
@pytest.mark.parametrize("B,H,T,hd,causal", [(3, 12, 50, 64, 1), (2, 12, 80, 64, 1), (2, 8, 80, 96, 0), (1, 12, 1, 64, 1),
                                             (2, 12, 128, 64, 1), (2, 12, 17, 64, 1), (1, 8, 50, 96, 0)])
def test_attention_fwd_bwd(ops, B, H, T, hd, causal, impl):
    """exact-fp32 FFMA kernel and mma.sync TF32 tensor-core kernel against an fp64 torch restatement."""
    d = H * hd
    qkv = torch.randn(B, T, 3 * d, device="cuda")
    q, k, v = qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:]
    ctx = torch.empty(B, T, d, device="cuda")
    lse = torch.empty(B, H, T, device="cuda")
    scale = hd ** -0.5
    ops.attention_fwd(q, k, v, ctx, lse, B, H, T, T, hd, T * 3 * d, 3 * d, T * 3 * d, 3 * d, T * d, d, scale, causal, impl=impl)
    qd = qkv.double().requires_grad_()
    split = lambda t: t.view(B, T, H, hd).transpose(1, 2)
    ref = ref_attention(split(qd[..., :d]), split(qd[..., d:2 * d]), split(qd[..., 2 * d:]), scale, causal)
    ref = ref.transpose(1, 2).reshape(B, T, d)
    tol_f, tol_b = (2e-5, 1e-4) if impl == "ffma" else (6e-3, 2e-2)   # TF32 operands: 2^-11 relative per product
    assert (ctx.double() - ref).abs().max() < tol_f * max(1.0, ref.abs().max().item())
    dctx = torch.randn(B, T, d, device="cuda")
    (ref * dctx.double()).sum().backward()
    dqkv = torch.empty(B, T, 3 * d, device="cuda")
    dbias = torch.zeros(3 * d, device="cuda")
    ops.attention_bwd(q, k, v, ctx, dctx, lse, dqkv[..., :d], dqkv[..., d:2 * d], dqkv[..., 2 * d:], B, H, T, T, hd,
                      T * 3 * d, 3 * d, T * 3 * d, 3 * d, T * d, d, scale, causal, dbias_qkv=dbias, impl=impl)
    assert (dqkv.double() - qd.grad).abs().max() < tol_b * max(1.0, qd.grad.abs().max().item())
    assert ((dqkv.double() - qd.grad).norm() / qd.grad.norm()).item() < (1e-5 if impl == "ffma" else 3e-3)
    ref_db = qd.grad.sum(dim=(0, 1))
    assert (dbias.double() - ref_db).abs().max() < (1e-3 if impl == "ffma" else 3e-2) * max(1.0, ref_db.abs().max().item())


@pytest.mark.parametrize("impl", ["ffma", "tc"])
def test_attention_key_padding_and_dropout(ops, impl):
    B, H, T, hd = 4, 12, 50, 64
    d = H * hd
    qkv = torch.randn(B, T, 3 * d, device="cuda")
    q, k, v = qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:]
    key_len = torch.tensor([50, 30, 18, 11], device="cuda", dtype=torch.int32)
    ctx = torch.empty(B, T, d, device="cuda")
    lse = torch.empty(B, H, T, device="cuda")
    args = (B, H, T, T, hd, T * 3 * d, 3 * d, T * 3 * d, 3 * d, T * d, d, hd ** -0.5, 1)
    ops.attention_fwd(q, k, v, ctx, lse, *args, key_len=key_len, impl=impl)
    split = lambda t: t.reshape(B, T, H, hd).transpose(1, 2).double()
    ref = ref_attention(split(q), split(k), split(v), hd ** -0.5, True, key_len).transpose(1, 2).reshape(B, T, d)
    assert (ctx.double() - ref).abs().max() < (2e-5 if impl == "ffma" else 6e-3)
    # padded-key gradients are exactly zero; valid ones match autograd
    qd = qkv.double().requires_grad_()
    sp = lambda t: t.view(B, T, H, hd).transpose(1, 2)
    r2 = ref_attention(sp(qd[..., :d]), sp(qd[..., d:2 * d]), sp(qd[..., 2 * d:]), hd ** -0.5, True, key_len)
    dctx = torch.randn(B, T, d, device="cuda")
    (r2.transpose(1, 2).reshape(B, T, d) * dctx.double()).sum().backward()
    dqkv = torch.empty(B, T, 3 * d, device="cuda")
    ops.attention_bwd(q, k, v, ctx, dctx, lse, dqkv[..., :d], dqkv[..., d:2 * d], dqkv[..., 2 * d:], *args, key_len=key_len,
                      impl=impl)
    assert ((dqkv.double() - qd.grad).norm() / qd.grad.norm()).item() < (1e-5 if impl == "ffma" else 3e-3)
    assert dqkv[1, 30:, d:].abs().max() == 0
    # dropout: E[out] = no-dropout output; check the mean over many heads is unbiased to a few %
    ctx_d = torch.empty_like(ctx)
    seed = ops.make_seed(11)
    ops.attention_fwd(q, k, v, ctx_d, lse, *args, p_drop=0.1, seed=seed, stream_id=2, impl=impl)
    ops.attention_fwd(q, k, v, ctx, lse, *args, impl=impl)
    rel = (ctx_d - ctx).norm() / ctx.norm()
    assert 0.05 < rel < 0.6
    assert abs((ctx_d.mean() - ctx.mean()).item()) < 5e-3
    # both implementations draw the SAME mask (shared (row, col) -> Philox mapping)
    other = torch.empty_like(ctx)
    ops.attention_fwd(q, k, v, other, lse, *args, p_drop=0.1, seed=seed, stream_id=2, impl=("tc" if impl == "ffma" else "ffma"))
    assert (other - ctx_d).abs().max() < 2e-2
    # backward with dropout regenerates the forward mask: compare with autograd through an explicit mask
    vid = torch.zeros(B, T, 3 * d, device="cuda")
    vid[..., :2 * d] = qkv[..., :2 * d]
    eye = torch.eye(T, device="cuda")[:, :hd] if T >= hd else None
    ops.attention_fwd(q, k, v, ctx_d, lse, *args, p_drop=0.1, seed=seed, stream_id=2, impl=impl)
    ops.attention_bwd(q, k, v, ctx_d, dctx, lse, dqkv[..., :d], dqkv[..., d:2 * d], dqkv[..., 2 * d:], *args, p_drop=0.1,
                      seed=seed, stream_id=2, impl=impl)
    assert torch.isfinite(dqkv).all()
    # finite-difference check of dV under the frozen mask (output is linear in V)
    dv_num = torch.zeros(B, T, d, device="cuda")
    probe = torch.randn(B, T, d, device="cuda")
    v2 = (v + 1.0 * probe).contiguous()   # the output is linear in V: a large step keeps TF32 rounding out of the quotient
    qkv2 = qkv.clone(); qkv2[..., 2 * d:] = v2
    c2 = torch.empty_like(ctx)
    ops.attention_fwd(qkv2[..., :d], qkv2[..., d:2 * d], qkv2[..., 2 * d:], c2, lse, *args, p_drop=0.1, seed=seed, stream_id=2, impl=impl)
    lhs = ((c2 - ctx_d) * dctx).sum().item() / 1.0
    rhs = (dqkv[..., 2 * d:] * probe).sum().item()
    assert abs(lhs - rhs) < 3e-2 * max(1.0, abs(rhs)), (lhs, rhs)


@pytest.mark.parametrize("rows,V", [(64, 50257), (10, 1000), (7, 33)])
def test_cross_entropy_fwd_bwd(ops, rows, V):
    ld = (V + 127) // 128 * 128
    logits_full = torch.randn(rows, ld, device="cuda") * 3
    logits = logits_full[:, :V]
    targets = torch.randint(1, V, (rows,), device="cuda")
    targets[::5] = 0
    ref_in = logits.double().clone().requires_grad_()
    ref = torch.nn.functional.cross_entropy(ref_in, targets, ignore_index=0)
    ref.backward()
    n_valid = torch.full((1,), 77.0, device="cuda")
    loss_sum = torch.full((1,), 5.0, device="cuda")
    ops.ce_count(targets, n_valid, loss_sum)
    assert loss_sum.item() == 0.0
    assert n_valid.item() == (targets != 0).sum().item()
    ops.ce_fwd_bwd(logits, targets, V, loss_sum, n_valid=n_valid)
    assert abs(loss_sum.item() / n_valid.item() - ref.item()) < 1e-5 * abs(ref.item())
    assert (logits.double() - ref_in.grad).abs().max() < 1e-7 + 1e-5 * ref_in.grad.abs().max()
    assert (logits[targets == 0] == 0).all()


def test_colsum_actbwd_rows_concat(ops):
    M, N = 1237, 2304
    x = torch.randn(M, N, device="cuda")
    out = torch.ones(N, device="cuda")
    ops.colsum_acc(x, out)
    assert (out.double() - (1 + x.double().sum(0))).abs().max() < 1e-3
    pre = torch.randn(999, 3072, device="cuda").requires_grad_()
    dy = torch.randn(999, 3072, device="cuda")
    torch.nn.functional.gelu(pre, approximate="tanh").backward(dy)
    dx = torch.empty_like(dy)
    dbias = torch.ones(3072, device="cuda")
    ops.act_bwd(dy, pre.detach(), dx, ops.ACT_GELU_NEW, dbias=dbias)
    assert (dx - pre.grad).abs().max() < 1e-5
    assert (dbias.double() - (1 + pre.grad.double().sum(0))).abs().max() < 1e-3
    a = torch.tanh(pre.detach())
    ops.act_bwd(dy, a, dx, ops.ACT_TANH)
    assert (dx - dy * (1 - a * a)).abs().max() < 1e-6
    r = torch.relu(pre.detach())
    ops.act_bwd(dy, r, dx, ops.ACT_RELU)
    assert torch.equal(dx, dy * (r > 0))
    B, T, L, d, off = 6, 50, 40, 768, 9
    src = torch.randn(B, T, d, device="cuda")
    dst = torch.empty(B * L, d, device="cuda")
    ops.rows_gather(src, dst, B, T, L, off)
    assert torch.equal(dst.view(B, L, d), src[:, off:off + L])
    back = torch.zeros(B, T, d, device="cuda")
    ops.rows_scatter(dst, back, B, T, L, off)
    assert torch.equal(back[:, off:off + L], src[:, off:off + L]) and back[:, :off].abs().max() == 0
    C, P = 40, 40
    lin = torch.randn(B, C, d, device="cuda")
    pc = torch.randn(P, d, device="cuda")
    xcat = torch.empty(B, C + P, d, device="cuda")
    ops.mapper_concat_fwd(lin, pc, xcat, B, C, P)
    assert torch.equal(xcat, torch.cat([lin, pc[None].expand(B, P, d)], 1))
    dxc = torch.randn(B, C + P, d, device="cuda")
    dlin = torch.empty(B, C, d, device="cuda")
    dpc = torch.zeros(P, d, device="cuda")
    ops.mapper_concat_bwd(dxc, dlin, dpc, B, C, P)
    assert torch.equal(dlin, dxc[:, :C]) and (dpc - dxc[:, C:].sum(0)).abs().max() < 1e-5


def test_adamw_matches_hf_semantics(ops):
    n = 4 * 1000 + 4
    p = torch.randn(n, device="cuda"); g = torch.randn(n, device="cuda")
    m = torch.zeros(n, device="cuda"); v = torch.zeros(n, device="cuda")
    pr, mr, vr = p.double().clone(), m.double().clone(), v.double().clone()
    lr_dev, t_dev = torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda")
    b1, b2, eps, wd = 0.9, 0.999, 1e-6, 0.01
    denom = torch.full((1,), 4.0, device="cuda")
    for t in range(1, 4):
        lr = 1e-3 * t
        lr_dev.fill_(lr); t_dev.fill_(float(t))
        gt = torch.randn(n, device="cuda")
        g.copy_(gt)
        ops.adamw_step(p, g, m, v, lr_dev, t_dev, b1, b2, eps, wd, grad_denom=denom, zero_grad=True)
        gd = gt.double() / 4.0
        # transformers 4.24 optimization.AdamW.step
        mr = mr * b1 + (1 - b1) * gd
        vr = vr * b2 + (1 - b2) * gd * gd
        step = lr * math.sqrt(1 - b2 ** t) / (1 - b1 ** t)
        pr = pr - step * mr / (vr.sqrt() + eps)
        pr = pr - lr * wd * pr
        assert g.abs().max() == 0
    assert (p.double() - pr).abs().max() < 2e-6


@pytest.mark.parametrize("world", [2, 3, 8])
def test_adamw_peer_step_is_reduce_scatter_update_all_gather(ops, world):
    """csrc/peer.cu on ONE GPU: the 'ranks' are separate local buffers (the kernel only sees pointers).  Every owner's
    launch must match capdec_adamw_step on the rank-order sum of the gradients, restricted to its slice, and leave the
    SAME BITS in every rank's parameter buffer; gradients are not modified (the caller clears them after its fence)."""
    n = 4 * 3 * world * 37
    torch.manual_seed(3)
    p0 = torch.randn(n, device="cuda")
    gs = [torch.randn(n, device="cuda") for _ in range(world)]
    ps = [p0.clone() for _ in range(world)]
    lr_dev, t_dev = torch.full((1,), 1e-3, device="cuda"), torch.full((1,), 2.0, device="cuda")
    denom = torch.full((1,), 7.0, device="cuda")
    b1, b2, eps, wd = 0.9, 0.999, 1e-6, 0.01
    sh = n // world
    m0, v0 = torch.rand(n, device="cuda") * 0.1, torch.rand(n, device="cuda") * 0.01
    # reference: plain fused AdamW on the summed gradient
    gsum = gs[0].clone()
    for r in range(1, world):
        gsum += gs[r]
    pr, mr, vr = p0.clone(), m0.clone(), v0.clone()
    ops.adamw_step(pr, gsum, mr, vr, lr_dev, t_dev, b1, b2, eps, wd, grad_denom=denom, zero_grad=False)
    g_before = [g.clone() for g in gs]
    for owner in range(world):
        lo = owner * sh
        m, v = m0[lo:lo + sh].clone(), v0[lo:lo + sh].clone()
        ops.adamw_peer_step([g[lo:].data_ptr() for g in gs], [q.data_ptr() for q in ps], owner, lo, sh, m, v, lr_dev, t_dev,
                            b1, b2, eps, wd, grad_denom=denom)
        # same arithmetic as adamw_kernel up to the compiler's FMA contraction of the two kernels (1 ulp)
        assert torch.allclose(m, mr[lo:lo + sh], rtol=2e-6, atol=2e-7) and torch.allclose(v, vr[lo:lo + sh], rtol=2e-6, atol=2e-7)
    torch.cuda.synchronize()
    for r in range(world):
        assert torch.equal(ps[r], ps[0]), f"replica {r} differs from replica 0: the all-gather must deliver identical bits"
        assert torch.allclose(ps[r], pr, rtol=2e-6, atol=2e-7), f"replica {r} differs from the single-buffer update"
        assert torch.equal(gs[r], g_before[r])


def test_step_clock_schedule_and_seed(ops):
    seed = ops.make_seed(5)
    step, lr, t = (torch.zeros(1, device="cuda") for _ in range(3))
    lrs = []
    for n in range(6):
        ops.step_clock(seed, step, lr, t, base_lr=2e-5, warmup_steps=2, total_steps=6)
        lrs.append(lr.item())
        assert t.item() == n + 1
    # get_linear_schedule_with_warmup(warmup=2, total=6): 0, .5, 1, .75, .5, .25 (x base)
    assert lrs == pytest.approx([0.0, 1e-5, 2e-5, 1.5e-5, 1e-5, 0.5e-5], rel=1e-6)
    assert seed.item() != 5


def test_compact_targets_and_limited_gemm(ops):
    """Row compaction of the non-ignored targets + GEMMs / CE bounded by a device scalar (LM head on valid rows only)."""
    B, L, P, d, V = 7, 40, 10, 768, 1000
    T = P + L
    g = torch.Generator(device="cuda").manual_seed(3)
    tokens = torch.randint(1, V, (B, L), device="cuda", generator=g)
    lens = torch.randint(3, L + 1, (B,), device="cuda", generator=g)
    tokens[torch.arange(L, device="cuda")[None, :] >= lens[:, None]] = 0
    tokens[2, 5] = 0  # a genuine id-0 token in the middle is ignored too (train.py:350)
    row_src = torch.zeros(B * L, dtype=torch.int32, device="cuda")
    dst_of = torch.zeros(B * T, dtype=torch.int32, device="cuda")
    tc = torch.zeros(B * L, dtype=torch.int64, device="cuda")
    counts = torch.zeros(2, dtype=torch.int32, device="cuda")
    n_valid = torch.zeros(1, device="cuda"); loss_sum = torch.ones(1, device="cuda")
    ops.compact_targets(tokens.reshape(-1), B, L, T, P - 1, row_src, dst_of, tc, counts, n_valid, loss_sum)
    flat = tokens.reshape(-1)
    idx = (flat != 0).nonzero().flatten()
    nv = idx.numel()
    assert counts.tolist() == [nv, (nv + 31) // 32 * 32] and n_valid.item() == nv and loss_sum.item() == 0
    exp_src = (idx // L) * T + (P - 1) + idx % L
    assert torch.equal(row_src[:nv].long(), exp_src) and torch.equal(tc[:nv], flat[idx])
    exp_dst = torch.full((B * T,), -1, dtype=torch.int32, device="cuda")
    exp_dst[exp_src] = torch.arange(nv, dtype=torch.int32, device="cuda")
    assert torch.equal(dst_of, exp_dst)
    xf = torch.randn(B * T, d, device="cuda")
    xsel = torch.full((B * L, d), float("nan"), device="cuda")
    ops.rows_gather_idx(xf, xsel, row_src, counts)
    assert torch.equal(xsel[:nv], xf[exp_src]) and (xsel[nv:counts[1].item()] == 0).all()
    back = torch.empty(B * T, d, device="cuda")
    ops.rows_scatter_idx(xsel, back, dst_of)
    ref = torch.zeros_like(back); ref[exp_src] = xf[exp_src]
    assert torch.equal(back, ref)
    # m_limit: rows beyond the limit's tile are untouched, rows below are correct
    M, N, K = 1200, 520, 96
    A, Al, Bm, Bl = make_ab(M, N, K, 0, 0)
    lim = torch.tensor([300], dtype=torch.int32, device="cuda")
    C = torch.full((M, N), 7.0, device="cuda")
    ops.gemm(A, 0, Bm, 0, C, M, N, K, m_limit=lim)
    refC = trunc_tf32(Al).double() @ trunc_tf32(Bl).double().t()
    assert (C[:300].double() - refC[:300]).abs().max() < 1e-4 * refC.abs().max()
    assert (C[1024:] == 7.0).all()
    # k_limit: reduction stops at the limit (operand rows of the k-tail are zero)
    M, N, K = 700, 768, 640
    A, Al, Bm, Bl = make_ab(M, N, K, 1, 1)
    A[333:352] = 0  # the caller's contract for the K-tail [limit, roundup32(limit))
    kl = torch.tensor([333], dtype=torch.int32, device="cuda")
    C = torch.zeros(M, N, device="cuda")
    ops.gemm(A, 1, Bm, 1, C, M, N, K, accumulate=True, k_limit=kl)
    refC = trunc_tf32(A[:333].t().contiguous()).double() @ trunc_tf32(Bm[:333].t().contiguous()).double().t()
    assert (C.double() - refC).abs().max() < 1e-4 * refC.abs().max()
    # CE with a row limit
    logits = torch.randn(64, 1024, device="cuda")[:, :V]
    tg = torch.randint(1, V, (64,), device="cuda")
    keep = logits.clone()
    rl = torch.tensor([40], dtype=torch.int32, device="cuda")
    ls = torch.zeros(1, device="cuda")
    ops.ce_fwd_bwd(logits, tg, V, ls, row_limit=rl)
    assert torch.equal(logits[40:], keep[40:])
    assert abs(ls.item() - torch.nn.functional.cross_entropy(keep[:40].double(), tg[:40], reduction="sum").item()) < 1e-2


@pytest.mark.parametrize("act", [1, 2, 3])
@pytest.mark.parametrize("layout", ["conv1d", "linear"])
def test_gemm_fused_activation_backward_and_bias_grad(ops, act, layout):
    """dgrad GEMM x act'(.) with the bias-gradient column sums fused in the epilogue == separate dgrad + act_bwd."""
    M, K, N = 1300, 768, 1100   # dx [M,N] = dy [M,K] . W^T, tails in M and N
    g = torch.Generator(device="cuda").manual_seed(5)
    dy = torch.randn(M, K, device="cuda", generator=g)
    W = (torch.randn(N, K, device="cuda", generator=g) if layout == "conv1d" else torch.randn(K, N, device="cuda", generator=g)) * 0.05
    pre = torch.randn(M, N, device="cuda", generator=g)
    act_in = pre if act == 1 else (torch.tanh(pre) if act == 2 else torch.relu(pre))
    dx = torch.empty(M, N, device="cuda")
    db = torch.full((N,), 3.0, device="cuda")
    ops.linear_dgrad_act(dy, W, layout, dx, act_in, act, dbias=db)
    Wl = W if layout == "conv1d" else W.t()
    lin = trunc_tf32(dy).double() @ trunc_tf32(Wl.contiguous()).double().t()
    p64 = pre.double()
    if act == 1:
        p64r = p64.clone().requires_grad_()
        torch.nn.functional.gelu(p64r, approximate="tanh").sum().backward()
        d = p64r.grad
    elif act == 2:
        d = 1 - torch.tanh(p64) ** 2
    else:
        d = (p64 > 0).double()
    ref = lin * d
    assert (dx.double() - ref).abs().max() < 2e-4 * max(1.0, ref.abs().max().item())
    assert (db.double() - (3.0 + ref.sum(0))).abs().max() < 2e-3 * max(1.0, ref.sum(0).abs().max().item())


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
@pytest.mark.parametrize("b_major", [0, 1])
def test_gemm_block_n_192(ops, mode, b_major):
    """192-wide tiles (96 B columns per CTA of a pair) in every engine that supports them."""
    from capdec_b200 import _lib
    M, N, K = 2600, 768, 200
    A, Al, B, Bl = make_ab(M, N, K, 0, b_major)
    C = new_c(M, N)
    lib = _lib.load()
    lib.capdec_gemm_debug_force_pair(mode)
    try:
        if mode == 3 and b_major:
            with pytest.raises(_lib.CapdecError):
                ops.gemm(A, 0, B, b_major, C, M, N, K, block_n=192)
            return
        ops.gemm(A, 0, B, b_major, C, M, N, K, block_n=192)
    finally:
        lib.capdec_gemm_debug_force_pair(-1)
    ref = trunc_tf32(Al).double() @ trunc_tf32(Bl).double().t()
    assert (C.double() - ref).abs().max() < 1e-4 * ref.abs().max()


@pytest.mark.parametrize("mode", [-1, 1, 3])
def test_gemm_mn_major_ragged_extent_inside_padded_pitch(ops, mode):
    """MN-major operands whose extent is not a multiple of 32 but whose pitch covers the rounded extent are fetched with
    3-D boxes (the tied-embedding gradient: V = 50257 in a 50304 pitch); NaNs in the pad columns must not leak."""
    from capdec_b200 import _lib
    M, N, K = 5001, 777, 300
    g = torch.Generator(device="cuda").manual_seed(11)
    Abuf = torch.full((K, 5120), float("nan"), device="cuda")
    Bbuf = torch.full((K, 896), float("nan"), device="cuda")
    Abuf[:, :M] = torch.randn(K, M, device="cuda", generator=g)
    Bbuf[:, :N] = torch.randn(K, N, device="cuda", generator=g)
    A, B = Abuf[:, :M], Bbuf[:, :N]
    C = torch.zeros(M, 800, device="cuda")[:, :N]
    lib = _lib.load()
    lib.capdec_gemm_debug_force_pair(mode)
    try:
        ops.gemm(A, 1, B, 1, C, M, N, K, accumulate=True)
        ops.gemm(A, 1, B, 1, C, M, N, K, accumulate=True)
    finally:
        lib.capdec_gemm_debug_force_pair(-1)
    ref = 2 * (trunc_tf32(A.t().contiguous()).double() @ trunc_tf32(B.t().contiguous()).double().t())
    assert torch.isfinite(C).all()
    assert (C.double() - ref).abs().max() < 1e-4 * ref.abs().max()


@pytest.mark.parametrize("limit", [None, 8820])
def test_gemm_fused_activation_backward_many_tiles(ops, limit):
    """The fused act'-multiply epilogue over several tiles per CTA (its input prefetcher runs across tile boundaries),
    with and without a device-side row limit and row hint."""
    M, K, N = 12800, 96, 3072
    g = torch.Generator(device="cuda").manual_seed(6)
    dy = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) * 0.05
    pre = torch.randn(M, N, device="cuda", generator=g)
    dx = torch.full((M, N), 5.0, device="cuda")
    db = torch.zeros(N, device="cuda")
    rows = None
    if limit is not None:
        rows = torch.tensor([limit], dtype=torch.int32, device="cuda")
        ops.set_row_hint(limit)
    try:
        ops.linear_dgrad_act(dy, W, "conv1d", dx, pre, 1, dbias=db, rows=rows)
    finally:
        ops.set_row_hint(0)
    live = M if limit is None else limit
    lin = trunc_tf32(dy[:live]).double() @ trunc_tf32(W).double().t()
    p64 = pre[:live].double().requires_grad_()
    torch.nn.functional.gelu(p64, approximate="tanh").sum().backward()
    ref = lin * p64.grad
    assert (dx[:live].double() - ref).abs().max() < 5e-4 * max(1.0, ref.abs().max().item())
    assert (db.double() - ref.sum(0)).abs().max() < 5e-3 * max(1.0, ref.sum(0).abs().max().item())
    if limit is not None:
        assert (dx[(live + 511) // 512 * 512:] == 5.0).all()


def test_gemm_autotune_keeps_results_and_remembers_plans(ops):
    """Measured plan selection: whatever plan wins, the product is the same; the plan is remembered per signature."""
    ops.gemm_autotune(-1)
    try:
        assert ops.gemm_autotune(1) == 0
        for (M, N, K, am, bm, acc) in [(2600, 768, 200, 0, 1, False), (768, 2304, 3200, 1, 1, True), (100, 300, 64, 0, 0, False)]:
            A, Al, B, Bl = make_ab(M, N, K, am, bm)
            ref = trunc_tf32(Al).double() @ trunc_tf32(Bl).double().t()
            C = torch.zeros(M, (N + 3) // 4 * 4, device="cuda")[:, :N]
            ops.gemm(A, am, B, bm, C, M, N, K, accumulate=acc)        # measuring call (accumulate output is garbage)
            C.zero_()
            ops.gemm(A, am, B, bm, C, M, N, K, accumulate=acc)        # remembered plan
            assert (C.double() - ref).abs().max() < 1e-4 * ref.abs().max()
        assert ops.gemm_autotune(0) == 3
        # still served from the table after measuring stopped, and under a different hint
        A, Al, B, Bl = make_ab(2600, 768, 200, 0, 1)
        C = new_c(2600, 768)
        ops.gemm(A, 0, B, 1, C, 2600, 768, 200)
        ref = trunc_tf32(Al).double() @ trunc_tf32(Bl).double().t()
        assert (C.double() - ref).abs().max() < 1e-4 * ref.abs().max()
    finally:
        assert ops.gemm_autotune(-1) == 0


def test_gemm_gelu_forward_with_stored_derivative_and_multiply_backward(ops):
    """act 4: C = gelu_new(x W + b), aux = gelu_new'(x W + b) from one tanh; mul_act 4: dx = (dy W2^T) * aux."""
    M, K, N = 700, 256, 1100
    g = torch.Generator(device="cuda").manual_seed(8)
    x = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(K, N, device="cuda", generator=g) * 0.08          # Conv1D layout [in, out]
    b = torch.randn(N, device="cuda", generator=g) * 0.1
    out, aux = new_c(M, N), new_c(M, N)
    ops.linear_fwd(x, W, "conv1d", b, out, act=ops.ACT_GELU_NEW_D, aux=aux)
    pre = (trunc_tf32(x).double() @ trunc_tf32(W.t().contiguous()).double().t() + b.double()).requires_grad_()
    ref = torch.nn.functional.gelu(pre, approximate="tanh")
    ref.sum().backward()
    assert (out.double() - ref).abs().max() < 2e-3          # MUFU.TANH: 2^-11 relative
    assert (aux.double() - pre.grad).abs().max() < 2e-3
    dy = torch.randn(M, 96, device="cuda", generator=g)
    W2 = torch.randn(N, 96, device="cuda", generator=g) * 0.05       # Conv1D [in = N, out = 96]
    dx = new_c(M, N)
    db = torch.zeros(N, device="cuda")
    ops.linear_dgrad_act(dy, W2, "conv1d", dx, aux, ops.ACT_GELU_NEW_D, dbias=db)
    refdx = (trunc_tf32(dy).double() @ trunc_tf32(W2).double().t()) * aux.double()
    assert (dx.double() - refdx).abs().max() < 2e-4 * max(1.0, refdx.abs().max().item())
    assert (db.double() - refdx.sum(0)).abs().max() < 2e-3 * max(1.0, refdx.sum(0).abs().max().item())


@pytest.mark.parametrize("pair", [0, 1])
@pytest.mark.parametrize("a_major,b_major", [(0, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize("bn", [256, 128])
def test_gemm_half_width_tail_tiles(ops, pair, a_major, b_major, bn):
    """A tile count that leaves a small last wave (here: full waves + <= 1/2 wave) is finished with half-width tiles
    (csrc/gemm_tf32.cu `decode`): same numbers as the full-width schedule (CAPDEC_GEMM_TAIL=0 is compiled out here, so the
    reference is the truncated fp64 product), incl. bias + gelu_new + second output, N tail, and a device row limit."""
    from capdec_b200 import _lib
    lib = _lib.load()
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    units = sms // 2 if pair else sms
    tile_m = 256 if pair else 128
    n_tiles = 5
    m_tiles = (units + 6 + n_tiles - 1) // n_tiles + 1          # base tiles = units + a handful -> tail of a few tiles
    M, N, K = m_tiles * tile_m - 37, n_tiles * bn - 19, 160
    A, Al, B, Bl = make_ab(M, N, K, a_major, b_major, seed=5)
    bias = torch.randn(N, device="cuda")
    C, aux = new_c(M, N), new_c(M, N)
    lib.capdec_gemm_debug_force_pair(pair)
    try:
        ops.gemm(A, a_major, B, b_major, C, M, N, K, bias=bias, act=1, aux=aux, block_n=bn)
        pre = trunc_tf32(Al).double() @ trunc_tf32(Bl).double().t() + bias.double()
        assert (aux.double() - pre).abs().max() < 1e-3
        assert (C.double() - torch.nn.functional.gelu(pre, approximate="tanh")).abs().max() < 2e-3
        # device row limit: only whole live tiles are computed, the rest of C stays untouched
        live = M - 3 * tile_m - 5
        lim = torch.tensor([live], device="cuda", dtype=torch.int32)
        C2 = new_c(M, N)
        ops.gemm(A, a_major, B, b_major, C2, M, N, K, bias=bias, block_n=bn, m_limit=lim)
        assert (C2[:live].double() - pre[:live]).abs().max() < 1e-3
        done = (live + tile_m - 1) // tile_m * tile_m
        assert torch.isnan(C2[done:]).all()
    finally:
        lib.capdec_gemm_debug_force_pair(-1)
