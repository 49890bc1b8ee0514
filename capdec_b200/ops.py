"""Torch-tensor wrappers over the C-ABI (include/capdec_b200.h).  Plumbing only: every function hands raw device
pointers of caller-owned tensors to a hand-written sm_100a kernel on torch's current stream.  No CPU fallback.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch

from . import _lib

ACT_NONE, ACT_GELU_NEW, ACT_TANH, ACT_RELU = 0, 1, 2, 3
ACT_GELU_NEW_D = 4   # tcgen05 path only: C = gelu_new(x), aux = gelu_new'(x); backward multiplies by aux (mul_act 4)

# GEMM arithmetic: "tf32" = 1xTF32 tcgen05 (perf), "tf32x3" = 3xTF32 split on tcgen05 (fp32-grade parity mode),
# "fp32" = CUDA-core FFMA verification kernel.
_PRECISION = "tf32"


def set_precision(mode: str) -> None:
    global _PRECISION
    if mode not in ("tf32", "tf32x3", "fp32"):
        raise ValueError(f"unknown precision mode {mode!r}")
    _PRECISION = mode


def get_precision() -> str:
    return _PRECISION


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _chk(t: torch.Tensor, name: str, dtype=torch.float32) -> None:
    if not t.is_cuda:
        raise _lib.CapdecError(f"{name} must be a CUDA tensor: capdec_b200 has no CPU path")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")


def _rowmajor(t: torch.Tensor, name: str) -> int:
    """2-D tensor with unit inner stride; returns its leading dimension (row pitch in elements)."""
    _chk(t, name)
    if t.dim() != 2 or (t.shape[1] > 1 and t.stride(1) != 1):
        raise ValueError(f"{name} must be 2-D with unit inner stride, got shape {tuple(t.shape)} strides {t.stride()}")
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


def is_tc() -> bool:
    """True in the tensor-core modes ("tf32" and the fp32-grade "tf32x3"): fused epilogues, packed rows, device row limits
    and the mma.sync attention are available; "fp32" is the CUDA-core verification mode."""
    return _PRECISION in ("tf32", "tf32x3")


def gemm(A: torch.Tensor, a_major: int, B: torch.Tensor, b_major: int, C: torch.Tensor, M: int, N: int, K: int, *,
         bias: Optional[torch.Tensor] = None, act: int = ACT_NONE, aux: Optional[torch.Tensor] = None,
         accumulate: bool = False, block_n: int = 0, split_k: int = 0, precision: Optional[str] = None,
         m_limit: Optional[torch.Tensor] = None, k_limit: Optional[torch.Tensor] = None) -> None:
    """C[M,N] (+)= act(A . B^T + bias).  a_major/b_major: 0 = stored [M|N, K], 1 = stored [K, M|N].
    m_limit / k_limit: optional int32 device scalars bounding the rows computed / the reduction length at run time
    (tcgen05 modes; the CUDA-core fp32 verification mode computes the full static extent, which is equivalent as long as
    the caller keeps the skipped region's contribution at zero)."""
    lib = _lib.load()
    lda, ldb, ldc = _rowmajor(A, "A"), _rowmajor(B, "B"), _rowmajor(C, "C")
    ea = (K, M) if a_major else (M, K)
    eb = (K, N) if b_major else (N, K)
    if tuple(A.shape) != ea or tuple(B.shape) != eb or tuple(C.shape) != (M, N):
        raise ValueError(f"gemm shape mismatch: A{tuple(A.shape)} expected {ea}, B{tuple(B.shape)} expected {eb}, "
                         f"C{tuple(C.shape)} expected {(M, N)}")
    if bias is not None:
        _chk(bias, "bias")
    if aux is not None and _rowmajor(aux, "aux") != ldc:
        raise ValueError("aux must share C's leading dimension")
    mode = precision or _PRECISION
    if mode == "fp32":
        rc = lib.capdec_gemm_fp32_simt(A.data_ptr(), a_major, lda, B.data_ptr(), b_major, ldb, C.data_ptr(), ldc, M, N,
                                       K, _ptr(bias), act, _ptr(aux), int(accumulate), _stream())
        _lib.check(rc, "gemm_fp32_simt")
        return
    # "tf32x3": same kernel family, operands split into hi/lo TF32 parts inside the shared-memory pipeline
    rc = lib.capdec_gemm_tf32_ex(A.data_ptr(), a_major, lda, B.data_ptr(), b_major, ldb, C.data_ptr(), ldc, M, N, K,
                                 _ptr(bias), act, _ptr(aux), int(accumulate), 1 if mode == "tf32x3" else 0, None, None,
                                 block_n, split_k, _ptr(m_limit), _ptr(k_limit), _stream())
    _lib.check(rc, "gemm_tf32")


def set_row_hint(rows: int) -> None:
    """Tiling hint for the row-limited GEMMs launched next from this thread (0 clears it); never changes results."""
    _lib.load().capdec_gemm_set_row_hint(int(rows))


def gemm_set_schedule(dynamic: bool) -> bool:
    """Tile order of the persistent GEMM kernels launched from now on: dynamic (device-wide tile counter) or static
    round-robin (default).  Returns the previous setting.  See capdec_gemm_set_schedule."""
    return bool(_lib.load().capdec_gemm_set_schedule(int(bool(dynamic))))


def gemm_autotune(enable: int) -> int:
    """1 / 0: start / stop measuring GEMM plans for unseen problems; -1: forget measured plans.  Returns their count."""
    return int(_lib.load().capdec_gemm_autotune(int(enable)))


def gemm_mul(A, a_major, B, b_major, C, M, N, K, mul_in, mul_act, colsum=None, block_n=0, m_limit=None):
    """C = (A . B^T) * act'(mul_in), colsum += column sums of C — dgrad + activation backward + bias gradient in one
    tcgen05 launch (1xTF32, or 3xTF32 with exact derivative arithmetic in the "tf32x3" mode)."""
    lda, ldb, ldc = _rowmajor(A, "A"), _rowmajor(B, "B"), _rowmajor(C, "C")
    if _rowmajor(mul_in, "mul_in") != ldc or tuple(mul_in.shape) != (M, N):
        raise ValueError("mul_in must have C's shape and leading dimension")
    rc = _lib.load().capdec_gemm_tf32_mul_ex(A.data_ptr(), a_major, lda, B.data_ptr(), b_major, ldb, C.data_ptr(), ldc, M,
                                             N, K, mul_in.data_ptr(), mul_act, _ptr(colsum), block_n, _ptr(m_limit),
                                             1 if _PRECISION == "tf32x3" else 0, _stream())
    _lib.check(rc, "gemm_tf32_mul")


def linear_dgrad_act(dy, W, layout, dx, act_in, act, dbias=None, rows=None):
    """dx = (dy . W^T) * act'(act_in) (+ dbias += colsum(dx)).  tcgen05 modes: one fused launch; fp32 verification mode:
    dgrad then act_bwd.  `rows` (everywhere below): device int32 scalar with the live row count of a packed batch
    (tcgen05 modes only)."""
    if is_tc():
        M, K = dy.shape
        gemm_mul(dy, 0, W, 0 if layout == "conv1d" else 1, dx, M, dx.shape[1], K, act_in, act, dbias, m_limit=rows)
    else:
        _no_rows(rows)
        linear_dgrad(dy, W, layout, dx)
        act_bwd(dx, act_in, dx, act, dbias=dbias)


# ---- layer helpers: `layout` is "conv1d" (HF Conv1D weight [in,out]) or "linear" (nn.Linear weight [out,in]) --------
def _no_rows(rows):
    if rows is not None:
        raise _lib.CapdecError("packed rows (device row limits) are implemented for the tcgen05 modes (tf32, tf32x3) only")


def linear_fwd(x, W, layout, bias, out, act=ACT_NONE, aux=None, rows=None, accumulate=False):
    """accumulate=True: out += x W (+ bias once) through TMA reduce-add - the caller has zeroed `out`; lets the planner
    split the reduction (split-K) when the tile count quantises badly onto the SMs (the N = 768 problems)."""
    M, K = x.shape
    N = out.shape[1]
    gemm(x, 0, W, 1 if layout == "conv1d" else 0, out, M, N, K, bias=bias, act=act, aux=aux, m_limit=rows,
         accumulate=accumulate)


def linear_dgrad(dy, W, layout, dx, accumulate=False, rows=None):
    M, K = dy.shape  # K = out features (reduction)
    N = dx.shape[1]
    gemm(dy, 0, W, 0 if layout == "conv1d" else 1, dx, M, N, K, accumulate=accumulate, m_limit=rows)


def linear_wgrad(x, dy, dW, layout, dbias=None, rows=None):
    """dW += x^T dy (conv1d) / dy^T x (linear); dbias (optional) += column sums of dy via a separate pass — the engine
    normally gets bias gradients from the fused producers (add_ln_bwd / act_bwd / attention_bwd) instead."""
    n_rows = x.shape[0]
    if layout == "conv1d":  # dW[in,out] += x^T dy
        gemm(x, 1, dy, 1, dW, x.shape[1], dy.shape[1], n_rows, accumulate=True, k_limit=rows)
    else:  # dW[out,in] += dy^T x
        gemm(dy, 1, x, 1, dW, dy.shape[1], x.shape[1], n_rows, accumulate=True, k_limit=rows)
    if dbias is not None:
        _no_rows(rows)
        colsum_acc(dy, dbias)


def make_seed(value: int, device="cuda") -> torch.Tensor:
    """Device-resident 64-bit Philox seed (int64 storage, read as uint64 by the kernels)."""
    return torch.tensor([value], dtype=torch.int64, device=device)


def _seed_ptr(seed, p_active: bool):
    if seed is None:
        if p_active:
            raise ValueError("a device seed tensor (ops.make_seed) is required when dropout / Philox noise is active")
        return None
    if not (isinstance(seed, torch.Tensor) and seed.is_cuda and seed.dtype == torch.int64 and seed.numel() >= 1):
        raise TypeError("seed must be a CUDA int64 tensor (see ops.make_seed)")
    return seed.data_ptr()


def noise_injection(x, out, variance, noise=None, offset=None, uniform_ball=False, dont_norm=False, seed=None, step=0):
    _chk(x, "x"); _chk(out, "out")
    B, D = x.shape
    rc = _lib.load().capdec_noise_injection(x.data_ptr(), out.data_ptr(), B, D, float(variance), _ptr(noise),
                                            _ptr(offset), int(uniform_ball), int(dont_norm),
                                            _seed_ptr(seed, noise is None and variance > 0), step, _stream())
    _lib.check(rc, "noise_injection")


def embed_fwd(tokens, prefix_proj, wte, wpe, h, B, P, L, p_drop=0.0, seed=None, stream_id=0):
    d = h.shape[-1]
    if tokens is not None:
        _chk(tokens, "tokens", torch.int64)
    rc = _lib.load().capdec_embed_fwd(_ptr(tokens), _ptr(prefix_proj), _ptr(wte), wpe.data_ptr(), h.data_ptr(), B, P, L,
                                      d, wte.shape[0] if wte is not None else 0, float(p_drop), _seed_ptr(seed, p_drop > 0), stream_id,
                                      _stream())
    _lib.check(rc, "embed_fwd")


def embed_bwd(tokens, dh, d_prefix_proj, d_wte, d_wpe, B, P, L, vocab, p_drop=0.0, seed=None, stream_id=0):
    d = dh.shape[-1]
    rc = _lib.load().capdec_embed_bwd(_ptr(tokens), dh.data_ptr(), _ptr(d_prefix_proj), _ptr(d_wte), _ptr(d_wpe), B, P,
                                      L, d, vocab, float(p_drop), _seed_ptr(seed, p_drop > 0), stream_id, _stream())
    _lib.check(rc, "embed_bwd")


def add_ln_fwd(h_in, y, h_out, x, stats, gamma, beta, eps=1e-5, p_drop=0.0, seed=None, stream_id=0, rows=None):
    n_rows, d = x.shape
    rc = _lib.load().capdec_add_ln_fwd(h_in.data_ptr(), _ptr(y), _ptr(h_out), x.data_ptr(), stats.data_ptr(),
                                       gamma.data_ptr(), beta.data_ptr(), n_rows, d, eps, float(p_drop),
                                       _seed_ptr(seed, p_drop > 0), stream_id, _ptr(rows), _stream())
    _lib.check(rc, "add_ln_fwd")


def add_ln_bwd(dx, r, stats, gamma, dh_res, dh_out, dy, dgamma, dbeta, p_drop=0.0, seed=None, stream_id=0,
               dbias_branch=None, rows=None):
    n_rows, d = dx.shape
    rc = _lib.load().capdec_add_ln_bwd(dx.data_ptr(), r.data_ptr(), stats.data_ptr(), gamma.data_ptr(), _ptr(dh_res),
                                       dh_out.data_ptr(), _ptr(dy), _ptr(dgamma), _ptr(dbeta), _ptr(dbias_branch), n_rows,
                                       d, float(p_drop), _seed_ptr(seed, p_drop > 0), stream_id, _ptr(rows), _stream())
    _lib.check(rc, "add_ln_bwd")


def _use_tc(impl):
    """attention implementation: 'tc' = mma.sync tensor-core kernel (1xTF32 in the "tf32" mode, 3xTF32 split in registers
    in the "tf32x3" mode), 'ffma' = exact fp32 CUDA-core kernel; default follows the precision mode (fp32 -> ffma)."""
    if impl is None:
        impl = os.environ.get("CAPDEC_ATTN_IMPL")   # bring-up / bisection switch: "ffma" or "tc" for every un-packed call
    if impl is None:
        return is_tc()
    return impl == "tc"


def attention_fwd(q, k, v, ctx, lse, B, H, T, S, hd, q_bs, q_ts, kv_bs, kv_ts, o_bs, o_ts, scale, causal,
                  key_len=None, p_drop=0.0, seed=None, stream_id=0, impl=None, cu_rows=None):
    if _use_tc(impl) or cu_rows is not None:
        lib = _lib.load()
        fn = lib.capdec_attention_tc_fwd_x3 if _PRECISION == "tf32x3" else lib.capdec_attention_tc_fwd
        rc = fn(q.data_ptr(), k.data_ptr(), v.data_ptr(), ctx.data_ptr(), _ptr(lse), B, H,
                                                 T, S, hd, q_bs, q_ts, kv_bs, kv_ts, o_bs, o_ts, float(scale), int(causal),
                                                 _ptr(key_len), float(p_drop), _seed_ptr(seed, p_drop > 0), stream_id,
                                                 _ptr(cu_rows), _stream())
        if rc != -3 or cu_rows is not None:
            _lib.check(rc, "attention_tc_fwd")
            return
    rc = _lib.load().capdec_attention_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), ctx.data_ptr(), _ptr(lse), B, H, T,
                                          S, hd, q_bs, q_ts, kv_bs, kv_ts, o_bs, o_ts, float(scale), int(causal),
                                          _ptr(key_len), float(p_drop), _seed_ptr(seed, p_drop > 0), stream_id, _stream())
    _lib.check(rc, "attention_fwd")


def attention_bwd(q, k, v, ctx, dctx, lse, dq, dk, dv, B, H, T, S, hd, q_bs, q_ts, kv_bs, kv_ts, o_bs, o_ts, scale,
                  causal, key_len=None, p_drop=0.0, seed=None, stream_id=0, dbias_qkv=None, impl=None, cu_rows=None):
    if _use_tc(impl) or cu_rows is not None:
        lib = _lib.load()
        fn = lib.capdec_attention_tc_bwd_x3 if _PRECISION == "tf32x3" else lib.capdec_attention_tc_bwd
        rc = fn(q.data_ptr(), k.data_ptr(), v.data_ptr(), ctx.data_ptr(), dctx.data_ptr(),
                                                 lse.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(),
                                                 _ptr(dbias_qkv), B, H, T, S, hd, q_bs, q_ts, kv_bs, kv_ts, o_bs, o_ts,
                                                 float(scale), int(causal), _ptr(key_len), float(p_drop),
                                                 _seed_ptr(seed, p_drop > 0), stream_id, _ptr(cu_rows), _stream())
        if rc != -3 or cu_rows is not None:
            _lib.check(rc, "attention_tc_bwd")
            return
    rc = _lib.load().capdec_attention_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), ctx.data_ptr(), dctx.data_ptr(),
                                          lse.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), _ptr(dbias_qkv),
                                          B, H, T, S, hd,
                                          q_bs, q_ts, kv_bs, kv_ts, o_bs, o_ts, float(scale), int(causal),
                                          _ptr(key_len), float(p_drop), _seed_ptr(seed, p_drop > 0), stream_id, _stream())
    _lib.check(rc, "attention_bwd")


def ce_count(targets, n_valid, loss_sum_to_zero=None, ignore_index=0):
    _chk(targets, "targets", torch.int64)
    _lib.check(_lib.load().capdec_ce_count(targets.data_ptr(), targets.numel(), ignore_index, n_valid.data_ptr(),
                                           _ptr(loss_sum_to_zero), _stream()), "ce_count")


def compact_targets(targets, B, L, T, off, row_src, dst_of, targets_c, counts, n_valid, loss_sum_to_zero=None,
                    ignore_index=0, cu_rows=None):
    _chk(targets, "targets", torch.int64)
    rc = _lib.load().capdec_compact_targets(targets.data_ptr(), B, L, T, off, ignore_index, row_src.data_ptr(),
                                            dst_of.data_ptr(), targets_c.data_ptr(), counts.data_ptr(),
                                            n_valid.data_ptr(), _ptr(loss_sum_to_zero), _ptr(cu_rows), _stream())
    _lib.check(rc, "compact_targets")


def rows_gather_idx(src, dst, row_src, counts):
    rc = _lib.load().capdec_rows_gather_idx(src.data_ptr(), dst.data_ptr(), row_src.data_ptr(), counts.data_ptr(),
                                            dst.shape[0], dst.shape[1], _stream())
    _lib.check(rc, "rows_gather_idx")


def rows_scatter_idx(src, dst, dst_of):
    rc = _lib.load().capdec_rows_scatter_idx(src.data_ptr(), dst.data_ptr(), dst_of.data_ptr(), dst.shape[0],
                                             dst.shape[1], _stream())
    _lib.check(rc, "rows_scatter_idx")


def step_clock(seed=None, step_dev=None, lr_dev=None, t_dev=None, base_lr=0.0, warmup_steps=0, total_steps=1):
    _lib.check(_lib.load().capdec_step_clock(_ptr(seed), _ptr(step_dev), _ptr(lr_dev), _ptr(t_dev), float(base_lr),
                                             int(warmup_steps), int(total_steps), _stream()), "step_clock")


def ce_fwd_bwd(logits, targets, V, loss_sum, n_valid=None, grad_scale=1.0, ignore_index=0, write_grad=True,
               row_limit=None):
    ld = _rowmajor(logits, "logits")
    _chk(targets, "targets", torch.int64)
    rc = _lib.load().capdec_ce_fwd_bwd(logits.data_ptr(), ld, targets.data_ptr(), logits.shape[0], V, ignore_index,
                                       _ptr(n_valid), float(grad_scale), loss_sum.data_ptr(), int(write_grad),
                                       _ptr(row_limit), _stream())
    _lib.check(rc, "ce_fwd_bwd")


def colsum_acc(x, out):
    ld = _rowmajor(x, "x")
    M, N = x.shape
    _lib.check(_lib.load().capdec_colsum_acc(x.data_ptr(), ld, out.data_ptr(), M, N, _stream()), "colsum_acc")


def act_bwd(dy, pre, dx, act, dbias=None):
    M, N = dy.shape
    if not (dy.is_contiguous() and pre.is_contiguous() and dx.is_contiguous()):
        raise ValueError("act_bwd expects contiguous [M, N] tensors")
    _lib.check(_lib.load().capdec_act_bwd(dy.data_ptr(), pre.data_ptr(), dx.data_ptr(), _ptr(dbias), M, N, act, _stream()),
               "act_bwd")


def rows_gather(src, dst, B, T, L, off):
    d = src.shape[-1]
    _lib.check(_lib.load().capdec_rows_gather(src.data_ptr(), dst.data_ptr(), B, T, L, off, d, _stream()), "rows_gather")


def rows_scatter(src, dst, B, T, L, off):
    d = src.shape[-1]
    _lib.check(_lib.load().capdec_rows_scatter(src.data_ptr(), dst.data_ptr(), B, T, L, off, d, _stream()),
               "rows_scatter")


def mapper_concat_fwd(lin, prefix_const, x, B, C, P):
    d = x.shape[-1]
    _lib.check(_lib.load().capdec_mapper_concat_fwd(lin.data_ptr(), prefix_const.data_ptr(), x.data_ptr(), B, C, P, d,
                                                    _stream()), "mapper_concat_fwd")


def mapper_concat_bwd(dx, dlin, dprefix_const, B, C, P):
    d = dx.shape[-1]
    _lib.check(_lib.load().capdec_mapper_concat_bwd(dx.data_ptr(), dlin.data_ptr(), _ptr(dprefix_const), B, C, P, d,
                                                    _stream()), "mapper_concat_bwd")


def adamw_step(p, g, m, v, lr_dev, t_dev, beta1=0.9, beta2=0.999, eps=1e-6, weight_decay=0.0, grad_denom=None,
               zero_grad=False):
    rc = _lib.load().capdec_adamw_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(),
                                       lr_dev.data_ptr(), t_dev.data_ptr(), beta1, beta2, eps, weight_decay,
                                       _ptr(grad_denom), int(zero_grad), _stream())
    _lib.check(rc, "adamw_step")


# ---- data-parallel update over NVLink peer memory (peer.cu) ------------------------------------------------------
def peer_export(t: torch.Tensor):
    """(64-byte CUDA-IPC handle, byte offset) of tensor `t`'s first element, for another process of this node."""
    h = ctypes.create_string_buffer(64)
    off = ctypes.c_int64(0)
    rc = _lib.load().capdec_peer_export(t.data_ptr(), ctypes.cast(h, ctypes.c_void_p), ctypes.byref(off))
    _lib.check(rc, "peer_export")
    return bytes(h.raw), int(off.value)


def peer_open(handle: bytes, offset: int) -> int:
    """Map a buffer exported by another process; returns the device address of its first element."""
    out = ctypes.c_void_p(0)
    rc = _lib.load().capdec_peer_open(ctypes.c_char_p(handle), int(offset), ctypes.byref(out))
    _lib.check(rc, "peer_open")
    return int(out.value)


def peer_close(ptr: int, offset: int) -> None:
    _lib.check(_lib.load().capdec_peer_close(ptr, int(offset)), "peer_close")


def copy_async(dst_ptr: int, src_ptr: int, nbytes: int) -> None:
    """cudaMemcpyAsync between two device addresses (peer addresses included) on the current stream; a copy-engine
    memcpy node when the stream is being captured."""
    _lib.check(_lib.load().capdec_copy_async(dst_ptr, src_ptr, int(nbytes), _stream()), "copy_async")


def adamw_peer_step(g_slices, p_ptrs, rank, lo, n, m, v, lr_dev, t_dev, beta1=0.9, beta2=0.999, eps=1e-6,
                    weight_decay=0.0, grad_denom=None):
    """Fused reduce-scatter + HF-AdamW + all-gather: `g_slices[r]` = device address of this rank's slice [lo, lo + n) of
    the gradients rank r computed (a peer buffer or local staging), `p_ptrs[r]` = device address of rank r's parameter
    buffer (ints, rank order)."""
    world = len(g_slices)
    arr = ctypes.c_void_p * world
    rc = _lib.load().capdec_adamw_peer_step(arr(*g_slices), arr(*p_ptrs), world, rank, int(lo), int(n), m.data_ptr(),
                                            v.data_ptr(), lr_dev.data_ptr(), t_dev.data_ptr(), beta1, beta2, eps,
                                            weight_decay, _ptr(grad_denom), _stream())
    _lib.check(rc, "adamw_peer_step")


# ---- KV-cached beam search (decode.cu) ---------------------------------------------------------------------------
def beam_init(st, n_img, beam, P, Tmax):
    rc = _lib.load().capdec_beam_init(st.step.data_ptr(), st.scores.data_ptr(), st.seq_len.data_ptr(),
                                      st.stopped.data_ptr(), st.src.data_ptr(), st.img_done.data_ptr(),
                                      st.ticket.data_ptr(), n_img, beam, P, Tmax, _stream())
    _lib.check(rc, "beam_init")


def kv_prefill(qkv, kcache, vcache, n_img, beam, P, Tmax, d):
    _chk(qkv, "qkv")
    rc = _lib.load().capdec_kv_prefill(qkv.data_ptr(), kcache.data_ptr(), vcache.data_ptr(), n_img, beam, P, Tmax, d,
                                       _stream())
    _lib.check(rc, "kv_prefill")


def decode_embed(st, wte, wpe, x, P):
    rc = _lib.load().capdec_decode_embed(st.step.data_ptr(), st.hist_tok.data_ptr(), wte.data_ptr(), wpe.data_ptr(),
                                         x.data_ptr(), x.shape[0], P, x.shape[1], _stream())
    _lib.check(rc, "decode_embed")


def decode_attention(qkv, kcache, vcache, st, ctx, H, hd, P, Tmax, scale):
    rc = _lib.load().capdec_decode_attention(qkv.data_ptr(), kcache.data_ptr(), vcache.data_ptr(), st.src.data_ptr(),
                                             st.step.data_ptr(), ctx.data_ptr(), qkv.shape[0], H, hd, P, Tmax,
                                             float(scale), _stream())
    _lib.check(rc, "decode_attention")


def row_topk(logits, V, temperature, k, cand_val, cand_idx, row_lse):
    ld = _rowmajor(logits, "logits")
    rc = _lib.load().capdec_row_topk(logits.data_ptr(), ld, logits.shape[0], V, float(temperature), k,
                                     cand_val.data_ptr(), cand_idx.data_ptr(), row_lse.data_ptr(), _stream())
    _lib.check(rc, "row_topk")


def beam_select(st, n_img, beam, P, Tmax, V, stop_token):
    rc = _lib.load().capdec_beam_select(st.cand_val.data_ptr(), st.cand_idx.data_ptr(), st.row_lse.data_ptr(),
                                        st.step.data_ptr(), st.scores.data_ptr(), st.seq_len.data_ptr(),
                                        st.stopped.data_ptr(), st.src.data_ptr(), st.hist_tok.data_ptr(),
                                        st.hist_parent.data_ptr(), st.img_done.data_ptr(), st.ticket.data_ptr(), n_img,
                                        beam, P, Tmax, V, int(stop_token), _stream())
    _lib.check(rc, "beam_select")


# ---- packed-row execution (packed.cu) ------------------------------------------------------------------------------
def pack_plan(tokens, P, cu, rows, row_bt):
    _chk(tokens, "tokens", torch.int64)
    B, L = tokens.shape
    _lib.check(_lib.load().capdec_pack_plan(tokens.data_ptr(), B, L, P, cu.data_ptr(), rows.data_ptr(), row_bt.data_ptr(),
                                            _stream()), "pack_plan")


def embed_fwd_packed(tokens, prefix_proj, wte, wpe, h, row_bt, rows, B, P, L, p_drop=0.0, seed=None, stream_id=0):
    d = h.shape[1]
    rc = _lib.load().capdec_embed_fwd_packed(tokens.data_ptr(), prefix_proj.data_ptr(), wte.data_ptr(), wpe.data_ptr(),
                                             h.data_ptr(), row_bt.data_ptr(), rows.data_ptr(), B, P, L, d, wte.shape[0],
                                             float(p_drop), _seed_ptr(seed, p_drop > 0), stream_id, _stream())
    _lib.check(rc, "embed_fwd_packed")


def embed_bwd_packed(tokens, dh, d_prefix_proj, d_wte, d_wpe, cu, B, P, L, vocab, p_drop=0.0, seed=None, stream_id=0):
    d = dh.shape[1]
    rc = _lib.load().capdec_embed_bwd_packed(tokens.data_ptr(), dh.data_ptr(), _ptr(d_prefix_proj), _ptr(d_wte), _ptr(d_wpe),
                                             cu.data_ptr(), B, P, L, d, vocab, float(p_drop), _seed_ptr(seed, p_drop > 0),
                                             stream_id, _stream())
    _lib.check(rc, "embed_bwd_packed")


def zero_tail_rows(buf, rows):
    _lib.check(_lib.load().capdec_zero_tail_rows(buf.data_ptr(), _rowmajor(buf, "buf"), rows.data_ptr(), _stream()),
               "zero_tail_rows")


def batch_gather(tokens_all, cap2emb, table, idx, tokens, prefix, P, mask=None, normalize=False):
    """One launch assembles a batch from the device-resident caption table (csrc/datafeed.cu; train.py:52-72)."""
    _chk(tokens_all, "tokens_all", torch.int32)
    _chk(cap2emb, "cap2emb", torch.int32)
    _chk(idx, "idx", torch.int64)
    _chk(tokens, "tokens", torch.int64)
    _chk(prefix, "prefix")
    if table.dtype not in (torch.float32, torch.float16) or not table.is_cuda:
        raise TypeError("the CLIP embedding table must be a CUDA fp32 or fp16 tensor")
    B, L = tokens.shape
    D = table.shape[1]
    if idx.numel() != B or tuple(prefix.shape) != (B, D) or tokens_all.shape[1] != L:
        raise ValueError("batch_gather: shape mismatch")
    if mask is not None and tuple(mask.shape) != (B, P + L):
        raise ValueError("batch_gather: mask must be [B, P + L]")
    for t in (tokens_all, cap2emb, table, idx, tokens, prefix):
        if not t.is_contiguous():
            raise ValueError("batch_gather expects contiguous tensors")
    rc = _lib.load().capdec_batch_gather(tokens_all.data_ptr(), cap2emb.data_ptr(), table.data_ptr(),
                                         int(table.dtype == torch.float16), idx.data_ptr(), tokens.data_ptr(), _ptr(mask),
                                         prefix.data_ptr(), B, L, P, D, int(bool(normalize)), _stream())
    _lib.check(rc, "batch_gather")


def zero_fill(t: torch.Tensor) -> None:
    """t <- 0 through the library (cudaMemsetAsync on the current stream; t must be contiguous)."""
    if not t.is_cuda or not t.is_contiguous():
        raise ValueError("zero_fill expects a contiguous CUDA tensor")
    _lib.check(_lib.load().capdec_zero_fill(t.data_ptr(), t.numel() * t.element_size(), _stream()), "zero_fill")
