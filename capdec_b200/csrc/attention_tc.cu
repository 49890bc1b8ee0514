// capdec_b200 — single-tile attention on the warp-level tensor-core path (mma.sync m16n8k8, TF32 in / FP32 accumulate),
// forward and backward.  Used when the GEMM precision mode is "tf32" (the benchmarked mode); the exact-fp32 FFMA
// kernels in attention.cu remain the parity-mode implementation and share the same C ABI and dropout mapping.
//
//   GPT-2:  softmax(q k^T / 8 + causal (+ key padding)) -> dropout -> . v   (HF:modeling_gpt2.py:54-72,185-191)
//   mapper: softmax(q k^T * 96^-0.5) . v, 8 heads x 96                      (train.py:150-167)
//
// One CTA (4 warps) per (batch, head); T, S <= 128 so Q/K/V (and dO, P, dS in backward) live in shared memory and
// scores never touch HBM.  Each warp owns 16-row tiles: S = QK^T -> masked softmax (+dropout) in registers -> P through
// shared memory -> O = PV.  Backward: phase A per query tile (P, dP = dO V^T, dS, dQ = dS K), phase B per key tile
// (dK = dS^T Q, dV = P^T dO), plus the fused c_attn bias gradient (column sums of dq|dk|dv).
// tcgen05 is reserved for the GEMMs (attention is 0.7 % of the step FLOPs; a 50x50x64 tile cannot feed a 128-row UMMA).
#include <stdlib.h>

#include "../../include/capdec_b200.h"
#include "common.cuh"

namespace capdec {

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// X3 = fp32-grade arithmetic on the same tensor-core path ("tf32x3" mode): every fragment is split in registers into
// hi = RN_tf32(x) and lo = x - hi (exact), and the product is accumulated as lo*hi + hi*lo + hi*hi (3xTF32); exp / log use
// the accurate library functions instead of the MUFU approximations.
__device__ __forceinline__ void split_tf32(uint32_t x, uint32_t& hi, uint32_t& lo) {
  hi = (x + 0x1000u) & 0xFFFFE000u;
  lo = __float_as_uint(__uint_as_float(x) - __uint_as_float(hi));
}
template <bool X3>
__device__ __forceinline__ void mma_t(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  if constexpr (!X3) {
    mma_tf32(d, a, b);
  } else {
    uint32_t ah[4], al[4], bh[2], bl[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_tf32(a[i], ah[i], al[i]);
#pragma unroll
    for (int i = 0; i < 2; ++i) split_tf32(b[i], bh[i], bl[i]);
    mma_tf32(d, al, bh);
    mma_tf32(d, ah, bl);
    mma_tf32(d, ah, bh);
  }
}
template <bool X3> __device__ __forceinline__ float exp_t(float x) { return X3 ? expf(x) : __expf(x); }
template <bool X3> __device__ __forceinline__ float log_t(float x) { return X3 ? logf(x) : __logf(x); }

// 16-byte asynchronous global -> shared copy (LDGSTS); !valid zero-fills the destination without touching `g`
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* g, bool valid) {
  const uint32_t sz = valid ? 16u : 0u;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(g), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// A fragment (16x8, row-major) of X[row0.., k0..] with row stride `ld` (floats)
__device__ __forceinline__ void lda_frag(uint32_t (&a)[4], const float* X, int ld, int row0, int k0, int g, int t) {
  const float* p = X + (size_t)(row0 + g) * ld + k0 + t;
  a[0] = __float_as_uint(p[0]);
  a[1] = __float_as_uint(p[8 * ld]);
  a[2] = __float_as_uint(p[4]);
  a[3] = __float_as_uint(p[8 * ld + 4]);
}
// A fragment of X^T: element (row r, k) = X[k0 + k][row0 + r]
__device__ __forceinline__ void lda_frag_t(uint32_t (&a)[4], const float* X, int ld, int row0, int k0, int g, int t) {
  const float* p = X + (size_t)(k0 + t) * ld + row0 + g;
  a[0] = __float_as_uint(p[0]);
  a[1] = __float_as_uint(p[8]);
  a[2] = __float_as_uint(p[4 * ld]);
  a[3] = __float_as_uint(p[4 * ld + 8]);
}
// B fragment (8x8, "col-major": B[k][n]) where B[k][n] = X[n0 + n][k0 + k]  (X stored [n][k])
__device__ __forceinline__ void ldb_frag_nk(uint32_t (&b)[2], const float* X, int ld, int n0, int k0, int g, int t) {
  const float* p = X + (size_t)(n0 + g) * ld + k0 + t;
  b[0] = __float_as_uint(p[0]);
  b[1] = __float_as_uint(p[4]);
}
// B fragment where B[k][n] = X[k0 + k][n0 + n]  (X stored [k][n])
__device__ __forceinline__ void ldb_frag_kn(uint32_t (&b)[2], const float* X, int ld, int n0, int k0, int g, int t) {
  const float* p = X + (size_t)(k0 + t) * ld + n0 + g;
  b[0] = __float_as_uint(p[0]);
  b[1] = __float_as_uint(p[4 * ld]);
}

// Dropout keep-scales for the two accumulator columns (8n+2t, 8n+2t+1) of tiles n = 2*np, 2*np+1 of probability row
// `grow` (= (b*H+h)*T + i): one Philox call yields all four.  Mapping (row, col) -> (quad, component):
//   quad = grow*32 + (col>>4)*4 + ((col&7)>>1), comp = ((col>>3)&1)*2 + (col&1)      (attention.cu uses the same)
__device__ __forceinline__ void attn_drop4(uint64_t seed, uint32_t stream_id, uint64_t grow, int np, int t, float p,
                                           float inv_keep, float (&s)[4]) {
  uint32_t r[4];
  Philox::gen(seed, stream_id, grow * 32 + (uint64_t)(np * 4 + t), r);
  const uint32_t thr = (uint32_t)(p * 4294967296.0f);
#pragma unroll
  for (int i = 0; i < 4; ++i) s[i] = (r[i] >= thr) ? inv_keep : 0.0f;
}

template <int HD, int NT_MAX, bool X3>
__global__ void __launch_bounds__(128) attention_tc_fwd_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                               const float* __restrict__ v, float* __restrict__ ctx,
                                                               float* __restrict__ lse, int H, int T_arg, int S_arg, int64_t q_bs,
                                                               int64_t q_ts, int64_t kv_bs, int64_t kv_ts, int64_t o_bs,
                                                               int64_t o_ts, float scale, int causal,
                                                               const int32_t* __restrict__ key_len, float p_drop,
                                                               const uint64_t* seed_dev, uint32_t stream_id, const int32_t* __restrict__ cu_rows) {
  pdl_trigger();   // lets a programmatically launched successor (the GEMMs) set itself up while this grid drains
  const uint64_t seed = seed_dev ? *seed_dev : 0ull;
  const int bh = blockIdx.x, b = bh / H, h = bh % H;
  // dense: sequence b owns rows [b*T, (b+1)*T).  packed (cu_rows != NULL): rows [cu_rows[b], cu_rows[b+1]) — only the
  // live positions of each caption exist; T = S = that count.  lse / dropout keep the dense pitch TL.
  int T = T_arg, S = S_arg;
  const int TL = T_arg;
  size_t q_b = (size_t)b * q_bs, kv_b = (size_t)b * kv_bs, o_b = (size_t)b * o_bs;
  if (cu_rows) {
    const int row_lo = cu_rows[b];
    T = S = cu_rows[b + 1] - row_lo;
    q_b = (size_t)row_lo * q_ts; kv_b = (size_t)row_lo * kv_ts; o_b = (size_t)row_lo * o_ts;
  }
  extern __shared__ __align__(16) float smem[];
  const int Tp = (T + 15) & ~15, Sp = (S + 7) & ~7;
  constexpr int LQ = HD + 4, LV = HD + 8;
  const int LP = Sp + 4;
  float* sQ = smem;
  float* sK = sQ + (size_t)Tp * LQ;
  float* sV = sK + (size_t)Sp * LQ;
  float* sP = sV + (size_t)Sp * LV;  // [4 warps][16][LP]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  constexpr int hd4 = HD / 4;
  // the whole head goes global -> shared with 16-byte async copies: every load of the tile is in flight at once (the
  // register-staged loop exposed one HBM round trip per iteration); rows >= T / S are zero-filled
  for (int i = threadIdx.x; i < Tp * hd4; i += 128) {
    const int r = i / hd4, c = i % hd4;
    const bool ok = r < T;
    cp_async16(sQ + (size_t)r * LQ + 4 * c, ok ? q + q_b + (size_t)r * q_ts + (size_t)h * HD + 4 * c : q, ok);
  }
  for (int i = threadIdx.x; i < Sp * hd4; i += 128) {
    const int r = i / hd4, c = i % hd4;
    const bool ok = r < S;
    const size_t gofs = ok ? kv_b + (size_t)r * kv_ts + (size_t)h * HD + 4 * c : 0;
    cp_async16(sK + (size_t)r * LQ + 4 * c, k + gofs, ok);
    cp_async16(sV + (size_t)r * LV + 4 * c, v + gofs, ok);
  }
  const int klen = key_len ? min(S, (int)key_len[b]) : S;
  const int nt_all = Sp / 8;
  const float inv_keep = 1.0f / (1.0f - p_drop);
  float* myP = sP + (size_t)warp * 16 * LP;
  cp_async_wait_all();
  __syncthreads();

  for (int m0 = warp * 16; m0 < Tp; m0 += 64) {
    // key tiles that can hold an unmasked column for this query tile
    int nt = nt_all;
    if (causal) nt = min(nt, (m0 + 15 + (S - T)) / 8 + 1);
    nt = min(nt, (klen + 7) / 8);
    if (nt < 1) nt = 1;
    float s[NT_MAX][4];
#pragma unroll
    for (int n = 0; n < NT_MAX; ++n) { s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < HD / 8; ++kk) {
      uint32_t a[4];
      lda_frag(a, sQ, LQ, m0, kk * 8, g, t);
#pragma unroll
      for (int n = 0; n < NT_MAX; ++n) {
        if (n < nt) {
          uint32_t bb[2];
          ldb_frag_nk(bb, sK, LQ, n * 8, kk * 8, g, t);
          mma_t<X3>(s[n], a, bb);
        }
      }
    }
    // masked softmax over the row pair (m0+g, m0+g+8); this thread holds columns 8n+2t, 8n+2t+1
    const int r0 = m0 + g, r1 = r0 + 8;
    const int lim0 = min(klen, causal ? r0 + 1 + (S - T) : S), lim1 = min(klen, causal ? r1 + 1 + (S - T) : S);
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int n = 0; n < NT_MAX; ++n) {
      if (n < nt) {
        const int c0 = n * 8 + 2 * t;
        s[n][0] = (c0 < lim0) ? s[n][0] * scale : -INFINITY;
        s[n][1] = (c0 + 1 < lim0) ? s[n][1] * scale : -INFINITY;
        s[n][2] = (c0 < lim1) ? s[n][2] * scale : -INFINITY;
        s[n][3] = (c0 + 1 < lim1) ? s[n][3] * scale : -INFINITY;
        mx0 = fmaxf(mx0, fmaxf(s[n][0], s[n][1]));
        mx1 = fmaxf(mx1, fmaxf(s[n][2], s[n][3]));
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    if (mx0 == -INFINITY) mx0 = 0.f;
    if (mx1 == -INFINITY) mx1 = 0.f;
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int n = 0; n < NT_MAX; ++n) {
      if (n < nt) {
        s[n][0] = exp_t<X3>(s[n][0] - mx0); s[n][1] = exp_t<X3>(s[n][1] - mx0);
        s[n][2] = exp_t<X3>(s[n][2] - mx1); s[n][3] = exp_t<X3>(s[n][3] - mx1);
        sum0 += s[n][0] + s[n][1];
        sum1 += s[n][2] + s[n][3];
      }
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
    const float inv0 = sum0 > 0.f ? 1.0f / sum0 : 0.f, inv1 = sum1 > 0.f ? 1.0f / sum1 : 0.f;
    if (t == 0 && lse) {
      if (r0 < T) lse[(size_t)bh * TL + r0] = mx0 + log_t<X3>(sum0);
      if (r1 < T) lse[(size_t)bh * TL + r1] = mx1 + log_t<X3>(sum1);
    }
    __syncwarp();  // previous tile's P reads are done
#pragma unroll
    for (int np = 0; np < NT_MAX / 2; ++np) {
      if (2 * np < nt) {
        float d0[4] = {1.f, 1.f, 1.f, 1.f}, d1[4] = {1.f, 1.f, 1.f, 1.f};
        if (p_drop > 0.f) {
          attn_drop4(seed, stream_id, (uint64_t)bh * TL + r0, np, t, p_drop, inv_keep, d0);
          attn_drop4(seed, stream_id, (uint64_t)bh * TL + r1, np, t, p_drop, inv_keep, d1);
        }
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int n = 2 * np + e;
          if (n < nt) {
            const int c0 = n * 8 + 2 * t;
            *reinterpret_cast<float2*>(myP + (size_t)g * LP + c0) =
                make_float2(s[n][0] * inv0 * d0[2 * e], s[n][1] * inv0 * d0[2 * e + 1]);
            *reinterpret_cast<float2*>(myP + (size_t)(g + 8) * LP + c0) =
                make_float2(s[n][2] * inv1 * d1[2 * e], s[n][3] * inv1 * d1[2 * e + 1]);
          }
        }
      }
    }
    __syncwarp();
    // O = P V
    float o[HD / 8][4];
#pragma unroll
    for (int n = 0; n < HD / 8; ++n) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }
    for (int kk = 0; kk < nt; ++kk) {
      uint32_t a[4];
      lda_frag(a, myP, LP, 0, kk * 8, g, t);
#pragma unroll
      for (int n = 0; n < HD / 8; ++n) {
        uint32_t bb[2];
        ldb_frag_kn(bb, sV, LV, n * 8, kk * 8, g, t);
        mma_t<X3>(o[n], a, bb);
      }
    }
#pragma unroll
    for (int n = 0; n < HD / 8; ++n) {
      const int c0 = n * 8 + 2 * t;
      if (r0 < T) *reinterpret_cast<float2*>(ctx + o_b + (size_t)r0 * o_ts + (size_t)h * HD + c0) = make_float2(o[n][0], o[n][1]);
      if (r1 < T) *reinterpret_cast<float2*>(ctx + o_b + (size_t)r1 * o_ts + (size_t)h * HD + c0) = make_float2(o[n][2], o[n][3]);
    }
  }
}

template <int HD, int NT_MAX, bool X3>
__global__ void __launch_bounds__(128) attention_tc_bwd_kernel(
    const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v, const float* __restrict__ ctx,
    const float* __restrict__ dctx, const float* __restrict__ lse, float* __restrict__ dq, float* __restrict__ dk,
    float* __restrict__ dv, float* __restrict__ dbias_qkv, int H, int T_arg, int S_arg, int64_t q_bs, int64_t q_ts, int64_t kv_bs,
    int64_t kv_ts, int64_t o_bs, int64_t o_ts, float scale, int causal, const int32_t* __restrict__ key_len, float p_drop,
    const uint64_t* seed_dev, uint32_t stream_id, const int32_t* __restrict__ cu_rows) {
  const uint64_t seed = seed_dev ? *seed_dev : 0ull;
  const int bh = blockIdx.x, b = bh / H, h = bh % H;
  // dense: sequence b owns rows [b*T, (b+1)*T).  packed (cu_rows != NULL): rows [cu_rows[b], cu_rows[b+1]) — only the
  // live positions of each caption exist; T = S = that count.  lse / dropout keep the dense pitch TL.
  int T = T_arg, S = S_arg;
  const int TL = T_arg;
  size_t q_b = (size_t)b * q_bs, kv_b = (size_t)b * kv_bs, o_b = (size_t)b * o_bs;
  if (cu_rows) {
    const int row_lo = cu_rows[b];
    T = S = cu_rows[b + 1] - row_lo;
    q_b = (size_t)row_lo * q_ts; kv_b = (size_t)row_lo * kv_ts; o_b = (size_t)row_lo * o_ts;
  }
  extern __shared__ __align__(16) float smem[];
  const int Tp = (T + 15) & ~15, Sp = (S + 15) & ~15;
  constexpr int LQ = HD + 4;
  const int LS = Sp + 4;
  float* sQ = smem;
  float* sK = sQ + (size_t)Tp * LQ;
  float* sV = sK + (size_t)Sp * LQ;
  float* sdO = sV + (size_t)Sp * LQ;
  float* sdS = sdO + (size_t)Tp * LQ;       // [Tp][LS]: P, then dS * scale
  float* sPd = sdS + (size_t)Tp * LS;       // [Tp][LS]: dropout(P)
  float* sDb = sPd + (size_t)Tp * LS;       // [3][HD] bias-gradient partial sums
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  constexpr int hd4 = HD / 4;
  for (int i = threadIdx.x; i < 3 * HD; i += 128) sDb[i] = 0.f;
  for (int i = threadIdx.x; i < Tp * hd4; i += 128) {
    const int r = i / hd4, c = i % hd4;
    float4 qv = make_float4(0.f, 0.f, 0.f, 0.f), dov = qv;
    if (r < T) {
      qv = *reinterpret_cast<const float4*>(q + q_b + (size_t)r * q_ts + (size_t)h * HD + 4 * c);
      dov = *reinterpret_cast<const float4*>(dctx + o_b + (size_t)r * o_ts + (size_t)h * HD + 4 * c);
    }
    *reinterpret_cast<float4*>(sQ + (size_t)r * LQ + 4 * c) = qv;
    *reinterpret_cast<float4*>(sdO + (size_t)r * LQ + 4 * c) = dov;
  }
  for (int i = threadIdx.x; i < Sp * hd4; i += 128) {
    const int r = i / hd4, c = i % hd4;
    float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
    if (r < S) {
      const size_t gofs = kv_b + (size_t)r * kv_ts + (size_t)h * HD + 4 * c;
      kv = *reinterpret_cast<const float4*>(k + gofs);
      vv = *reinterpret_cast<const float4*>(v + gofs);
    }
    *reinterpret_cast<float4*>(sK + (size_t)r * LQ + 4 * c) = kv;
    *reinterpret_cast<float4*>(sV + (size_t)r * LQ + 4 * c) = vv;
  }
  __syncthreads();
  const int klen = key_len ? min(S, (int)key_len[b]) : S;
  const int nt_all = Sp / 8;
  const float inv_keep = 1.0f / (1.0f - p_drop);

  // ================= phase A: per query tile — P, dP, dS (to smem), dQ =================
  for (int m0 = warp * 16; m0 < Tp; m0 += 64) {
    int nt = nt_all;
    if (causal) nt = min(nt, (m0 + 15 + (S - T)) / 8 + 1);
    nt = min(nt, (klen + 7) / 8);
    if (nt < 1) nt = 1;
    const int r0 = m0 + g, r1 = r0 + 8;
    const float l0 = (r0 < T) ? lse[(size_t)bh * TL + r0] : 0.f, l1 = (r1 < T) ? lse[(size_t)bh * TL + r1] : 0.f;
    const int lim0 = (r0 < T) ? min(klen, causal ? r0 + 1 + (S - T) : S) : 0;
    const int lim1 = (r1 < T) ? min(klen, causal ? r1 + 1 + (S - T) : S) : 0;
    float acc[NT_MAX][4];
    // ---- S = Q K^T -> P ----
#pragma unroll
    for (int n = 0; n < NT_MAX; ++n) { acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < HD / 8; ++kk) {
      uint32_t a[4];
      lda_frag(a, sQ, LQ, m0, kk * 8, g, t);
#pragma unroll
      for (int n = 0; n < NT_MAX; ++n) {
        if (n < nt) {
          uint32_t bb[2];
          ldb_frag_nk(bb, sK, LQ, n * 8, kk * 8, g, t);
          mma_t<X3>(acc[n], a, bb);
        }
      }
    }
#pragma unroll
    for (int np = 0; np < NT_MAX / 2; ++np) {
      float d0[4] = {1.f, 1.f, 1.f, 1.f}, d1[4] = {1.f, 1.f, 1.f, 1.f};
      if (p_drop > 0.f && 2 * np < nt) {
        attn_drop4(seed, stream_id, (uint64_t)bh * TL + r0, np, t, p_drop, inv_keep, d0);
        attn_drop4(seed, stream_id, (uint64_t)bh * TL + r1, np, t, p_drop, inv_keep, d1);
      }
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int n = 2 * np + e;
        if (n < nt_all) {   // also zero-fill the tiles this query tile never touches (phase B reads whole columns)
          const int c0 = n * 8 + 2 * t;
          float p00 = 0.f, p01 = 0.f, p10 = 0.f, p11 = 0.f;
          if (n < nt) {
            p00 = (c0 < lim0) ? exp_t<X3>(acc[n][0] * scale - l0) : 0.f;
            p01 = (c0 + 1 < lim0) ? exp_t<X3>(acc[n][1] * scale - l0) : 0.f;
            p10 = (c0 < lim1) ? exp_t<X3>(acc[n][2] * scale - l1) : 0.f;
            p11 = (c0 + 1 < lim1) ? exp_t<X3>(acc[n][3] * scale - l1) : 0.f;
          }
          *reinterpret_cast<float2*>(sdS + (size_t)r0 * LS + c0) = make_float2(p00, p01);
          *reinterpret_cast<float2*>(sdS + (size_t)r1 * LS + c0) = make_float2(p10, p11);
          *reinterpret_cast<float2*>(sPd + (size_t)r0 * LS + c0) = make_float2(p00 * d0[2 * e], p01 * d0[2 * e + 1]);
          *reinterpret_cast<float2*>(sPd + (size_t)r1 * LS + c0) = make_float2(p10 * d1[2 * e], p11 * d1[2 * e + 1]);
        }
      }
    }
    // ---- dP = dO V^T ----
#pragma unroll
    for (int n = 0; n < NT_MAX; ++n) { acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < HD / 8; ++kk) {
      uint32_t a[4];
      lda_frag(a, sdO, LQ, m0, kk * 8, g, t);
#pragma unroll
      for (int n = 0; n < NT_MAX; ++n) {
        if (n < nt) {
          uint32_t bb[2];
          ldb_frag_nk(bb, sV, LQ, n * 8, kk * 8, g, t);
          mma_t<X3>(acc[n], a, bb);
        }
      }
    }
    // ---- D_i = sum_j Pd_ij dP_ij: the row sum of softmax's backward exactly as autograd forms it, from the SAME P and dP
    // that enter dS below.  (The flash-attention shortcut D_i = dO_i . O_i equals it only in exact arithmetic: with
    // tensor-core products the two sides of dP_ij - D_i carry different rounding biases and their difference - the
    // to_queries / c_attn query gradient - lost two digits, profiles/r2_attention_rowsum.md.)
    float D0 = 0.f, D1 = 0.f;
#pragma unroll
    for (int n = 0; n < NT_MAX; ++n) {
      if (n < nt) {
        const int c0 = n * 8 + 2 * t;
        const float2 pd0 = *reinterpret_cast<const float2*>(sPd + (size_t)r0 * LS + c0);
        const float2 pd1 = *reinterpret_cast<const float2*>(sPd + (size_t)r1 * LS + c0);
        D0 += pd0.x * acc[n][0] + pd0.y * acc[n][1];
        D1 += pd1.x * acc[n][2] + pd1.y * acc[n][3];
      }
    }
    D0 += __shfl_xor_sync(0xffffffffu, D0, 1); D0 += __shfl_xor_sync(0xffffffffu, D0, 2);
    D1 += __shfl_xor_sync(0xffffffffu, D1, 1); D1 += __shfl_xor_sync(0xffffffffu, D1, 2);
    // ---- dS = Pd * dP - P * D (each thread re-reads the P / Pd values it wrote), stored pre-multiplied by `scale` ----
#pragma unroll
    for (int n = 0; n < NT_MAX; ++n) {
      if (n < nt) {
        const int c0 = n * 8 + 2 * t;
        float2* ps0 = reinterpret_cast<float2*>(sdS + (size_t)r0 * LS + c0);
        float2* ps1 = reinterpret_cast<float2*>(sdS + (size_t)r1 * LS + c0);
        const float2 p0 = *ps0, p1 = *ps1;
        const float2 pd0 = *reinterpret_cast<const float2*>(sPd + (size_t)r0 * LS + c0);
        const float2 pd1 = *reinterpret_cast<const float2*>(sPd + (size_t)r1 * LS + c0);
        *ps0 = make_float2((pd0.x * acc[n][0] - p0.x * D0) * scale, (pd0.y * acc[n][1] - p0.y * D0) * scale);
        *ps1 = make_float2((pd1.x * acc[n][2] - p1.x * D1) * scale, (pd1.y * acc[n][3] - p1.y * D1) * scale);
      }
    }
    __syncwarp();
    // ---- dQ = dS K ----
    float o[HD / 8][4];
#pragma unroll
    for (int n = 0; n < HD / 8; ++n) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }
    for (int kk = 0; kk < nt; ++kk) {
      uint32_t a[4];
      lda_frag(a, sdS, LS, m0, kk * 8, g, t);
#pragma unroll
      for (int n = 0; n < HD / 8; ++n) {
        uint32_t bb[2];
        ldb_frag_kn(bb, sK, LQ, n * 8, kk * 8, g, t);
        mma_t<X3>(o[n], a, bb);
      }
    }
#pragma unroll
    for (int n = 0; n < HD / 8; ++n) {
      const int c0 = n * 8 + 2 * t;
      if (r0 < T) *reinterpret_cast<float2*>(dq + q_b + (size_t)r0 * q_ts + (size_t)h * HD + c0) = make_float2(o[n][0], o[n][1]);
      if (r1 < T) *reinterpret_cast<float2*>(dq + q_b + (size_t)r1 * q_ts + (size_t)h * HD + c0) = make_float2(o[n][2], o[n][3]);
      if (dbias_qkv) {  // rows >= T contribute exact zeros (dS rows are zero there)
        float c_even = o[n][0] + o[n][2], c_odd = o[n][1] + o[n][3];
#pragma unroll
        for (int off = 4; off < 32; off <<= 1) {
          c_even += __shfl_xor_sync(0xffffffffu, c_even, off);
          c_odd += __shfl_xor_sync(0xffffffffu, c_odd, off);
        }
        if (g == 0) { atomicAdd(&sDb[c0], c_even); atomicAdd(&sDb[c0 + 1], c_odd); }
      }
    }
  }
  __syncthreads();
  // ================= phase B: per key tile — dK = dS^T Q, dV = Pd^T dO =================
  for (int j0 = warp * 16; j0 < Sp; j0 += 64) {
    float ok[HD / 8][4], ov[HD / 8][4];
#pragma unroll
    for (int n = 0; n < HD / 8; ++n) { ok[n][0] = ok[n][1] = ok[n][2] = ok[n][3] = 0.f; ov[n][0] = ov[n][1] = ov[n][2] = ov[n][3] = 0.f; }
    // queries that can see any key of this tile: i >= j0 - (S - T) under the causal mask
    int kk0 = 0;
    if (causal) kk0 = max(0, (j0 - (S - T))) / 8;
    for (int kk = kk0; kk < Tp / 8; ++kk) {
      uint32_t as_[4], ap[4];
      lda_frag_t(as_, sdS, LS, j0, kk * 8, g, t);
      lda_frag_t(ap, sPd, LS, j0, kk * 8, g, t);
#pragma unroll
      for (int n = 0; n < HD / 8; ++n) {
        uint32_t bq[2], bo[2];
        ldb_frag_kn(bq, sQ, LQ, n * 8, kk * 8, g, t);
        ldb_frag_kn(bo, sdO, LQ, n * 8, kk * 8, g, t);
        mma_t<X3>(ok[n], as_, bq);
        mma_t<X3>(ov[n], ap, bo);
      }
    }
    const int r0 = j0 + g, r1 = r0 + 8;
#pragma unroll
    for (int n = 0; n < HD / 8; ++n) {
      const int c0 = n * 8 + 2 * t;
      if (r0 < S) {
        const size_t gofs = kv_b + (size_t)r0 * kv_ts + (size_t)h * HD + c0;
        *reinterpret_cast<float2*>(dk + gofs) = make_float2(ok[n][0], ok[n][1]);
        *reinterpret_cast<float2*>(dv + gofs) = make_float2(ov[n][0], ov[n][1]);
      }
      if (r1 < S) {
        const size_t gofs = kv_b + (size_t)r1 * kv_ts + (size_t)h * HD + c0;
        *reinterpret_cast<float2*>(dk + gofs) = make_float2(ok[n][2], ok[n][3]);
        *reinterpret_cast<float2*>(dv + gofs) = make_float2(ov[n][2], ov[n][3]);
      }
      if (dbias_qkv) {  // padded key rows hold exact zeros (their dS / Pd columns are zero)
        float ke = ok[n][0] + ok[n][2], ko = ok[n][1] + ok[n][3], ve = ov[n][0] + ov[n][2], vo = ov[n][1] + ov[n][3];
#pragma unroll
        for (int off = 4; off < 32; off <<= 1) {
          ke += __shfl_xor_sync(0xffffffffu, ke, off); ko += __shfl_xor_sync(0xffffffffu, ko, off);
          ve += __shfl_xor_sync(0xffffffffu, ve, off); vo += __shfl_xor_sync(0xffffffffu, vo, off);
        }
        if (g == 0) {
          atomicAdd(&sDb[HD + c0], ke); atomicAdd(&sDb[HD + c0 + 1], ko);
          atomicAdd(&sDb[2 * HD + c0], ve); atomicAdd(&sDb[2 * HD + c0 + 1], vo);
        }
      }
    }
  }
  if (dbias_qkv) {
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * HD; i += 128)
      atomicAdd(dbias_qkv + (size_t)(i / HD) * H * HD + (size_t)h * HD + (i % HD), sDb[i]);
  }
}

// Backward, 8-warp variant (default): same arithmetic, dropout mapping and shared-memory tiles as
// attention_tc_bwd_kernel, but TWO warps share every 16-row tile so that a 105 KB (batch, head) CTA brings 16 resident
// warps per SM instead of 8 (profiles/r1_ncu_nongemm.md: 12 % warps active, latency-bound at 1.4 TB/s):
//   phase A  warp (wq, hh): query tile wq, key-tile PAIRS np with (np & 1) == hh for S / P / dP / dS (so each Philox call is
//            still made exactly once), then - after a 64-thread named barrier - half of the head-dim columns of dQ = dS K;
//   phase B  warp (wq, hh): key tile wq, hh == 0 -> dK = dS^T Q, hh == 1 -> dV = Pd^T dO.
// Q / K / V / dO arrive by cp.async; the LSE rows are fetched from global memory while those copies are in flight.  `ctx`
// is no longer read: D_i is the row sum of Pd * dP (see the 4-warp kernel), not dO_i . O_i.
template <int HD, int NT_MAX, bool X3>
__global__ void __launch_bounds__(256, X3 ? 1 : 2) attention_tc_bwd8_kernel(
    const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v, const float* __restrict__ ctx,
    const float* __restrict__ dctx, const float* __restrict__ lse, float* __restrict__ dq, float* __restrict__ dk,
    float* __restrict__ dv, float* __restrict__ dbias_qkv, int H, int T_arg, int S_arg, int64_t q_bs, int64_t q_ts, int64_t kv_bs,
    int64_t kv_ts, int64_t o_bs, int64_t o_ts, float scale, int causal, const int32_t* __restrict__ key_len, float p_drop,
    const uint64_t* seed_dev, uint32_t stream_id, const int32_t* __restrict__ cu_rows) {
  pdl_trigger();   // lets a programmatically launched successor (the GEMMs) set itself up while this grid drains
  const uint64_t seed = seed_dev ? *seed_dev : 0ull;
  const int bh = blockIdx.x, b = bh / H, h = bh % H;
  int T = T_arg, S = S_arg;
  const int TL = T_arg;
  size_t q_b = (size_t)b * q_bs, kv_b = (size_t)b * kv_bs, o_b = (size_t)b * o_bs;
  if (cu_rows) {
    const int row_lo = cu_rows[b];
    T = S = cu_rows[b + 1] - row_lo;
    q_b = (size_t)row_lo * q_ts; kv_b = (size_t)row_lo * kv_ts; o_b = (size_t)row_lo * o_ts;
  }
  extern __shared__ __align__(16) float smem[];
  const int Tp = (T + 15) & ~15, Sp = (S + 15) & ~15;
  constexpr int LQ = HD + 4;
  const int LS = Sp + 4;
  float* sQ = smem;
  float* sK = sQ + (size_t)Tp * LQ;
  float* sV = sK + (size_t)Sp * LQ;
  float* sdO = sV + (size_t)Sp * LQ;
  float* sdS = sdO + (size_t)Tp * LQ;       // [Tp][LS]: P, then dS * scale
  float* sPd = sdS + (size_t)Tp * LS;       // [Tp][LS]: dropout(P)
  float* sDb = sPd + (size_t)Tp * LS;       // [3][HD] bias-gradient partial sums
  float* sD = sDb + 3 * HD;                 // [2][Tp] partial row sums of Pd * dP (one per warp of the tile's pair)
  float* sL = sD + 2 * Tp;                  // [Tp] log-sum-exp of row i
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int wq = warp & 3, hh = warp >> 2;
  constexpr int hd4 = HD / 4;
  for (int i = tid; i < Tp * hd4; i += 256) {
    const int r = i / hd4, c = i % hd4;
    const bool ok = r < T;
    cp_async16(sQ + (size_t)r * LQ + 4 * c, ok ? q + q_b + (size_t)r * q_ts + (size_t)h * HD + 4 * c : q, ok);
    cp_async16(sdO + (size_t)r * LQ + 4 * c, ok ? dctx + o_b + (size_t)r * o_ts + (size_t)h * HD + 4 * c : dctx, ok);
  }
  for (int i = tid; i < Sp * hd4; i += 256) {
    const int r = i / hd4, c = i % hd4;
    const bool ok = r < S;
    const size_t gofs = ok ? kv_b + (size_t)r * kv_ts + (size_t)h * HD + 4 * c : 0;
    cp_async16(sK + (size_t)r * LQ + 4 * c, k + gofs, ok);
    cp_async16(sV + (size_t)r * LQ + 4 * c, v + gofs, ok);
  }
  for (int i = tid; i < 3 * HD; i += 256) sDb[i] = 0.f;
  // LSE rows while the tile copies are in flight
  for (int row = tid; row < Tp; row += 256) sL[row] = (row < T) ? lse[(size_t)bh * TL + row] : 0.f;
  const int klen = key_len ? min(S, (int)key_len[b]) : S;
  const int nt_all = Sp / 8;
  const float inv_keep = 1.0f / (1.0f - p_drop);
  cp_async_wait_all();
  __syncthreads();

  // ================= phase A: per query tile - P, dP, dS (to smem), dQ =================
  constexpr int NL = NT_MAX / 2;            // key tiles owned by one warp of the pair
  for (int m0 = wq * 16; m0 < Tp; m0 += 64) {
    int nt = nt_all;
    if (causal) nt = min(nt, (m0 + 15 + (S - T)) / 8 + 1);
    nt = min(nt, (klen + 7) / 8);
    if (nt < 1) nt = 1;
    const int r0 = m0 + g, r1 = r0 + 8;
    const float l0 = sL[r0], l1 = sL[r1];
    const int lim0 = (r0 < T) ? min(klen, causal ? r0 + 1 + (S - T) : S) : 0;
    const int lim1 = (r1 < T) ? min(klen, causal ? r1 + 1 + (S - T) : S) : 0;
    float acc[NL][4];
    // local tile lt -> key tile n = 4*(lt>>1) + 2*hh + (lt&1)   (tile pair np = 2*(lt>>1) + hh)
    // ---- S = Q K^T -> P ----
#pragma unroll
    for (int lt = 0; lt < NL; ++lt) { acc[lt][0] = acc[lt][1] = acc[lt][2] = acc[lt][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < HD / 8; ++kk) {
      uint32_t a[4];
      lda_frag(a, sQ, LQ, m0, kk * 8, g, t);
#pragma unroll
      for (int lt = 0; lt < NL; ++lt) {
        const int n = 4 * (lt >> 1) + 2 * hh + (lt & 1);
        if (n < nt) {
          uint32_t bb[2];
          ldb_frag_nk(bb, sK, LQ, n * 8, kk * 8, g, t);
          mma_t<X3>(acc[lt], a, bb);
        }
      }
    }
#pragma unroll
    for (int lp = 0; lp < NL / 2; ++lp) {
      const int np = 2 * lp + hh;
      float d0[4] = {1.f, 1.f, 1.f, 1.f}, d1[4] = {1.f, 1.f, 1.f, 1.f};
      if (p_drop > 0.f && 2 * np < nt) {
        attn_drop4(seed, stream_id, (uint64_t)bh * TL + r0, np, t, p_drop, inv_keep, d0);
        attn_drop4(seed, stream_id, (uint64_t)bh * TL + r1, np, t, p_drop, inv_keep, d1);
      }
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int lt = 2 * lp + e;
        const int n = 2 * np + e;
        if (n < nt_all) {   // also zero-fill the tiles this query tile never touches (phase B reads whole columns)
          const int c0 = n * 8 + 2 * t;
          float p00 = 0.f, p01 = 0.f, p10 = 0.f, p11 = 0.f;
          if (n < nt) {
            p00 = (c0 < lim0) ? exp_t<X3>(acc[lt][0] * scale - l0) : 0.f;
            p01 = (c0 + 1 < lim0) ? exp_t<X3>(acc[lt][1] * scale - l0) : 0.f;
            p10 = (c0 < lim1) ? exp_t<X3>(acc[lt][2] * scale - l1) : 0.f;
            p11 = (c0 + 1 < lim1) ? exp_t<X3>(acc[lt][3] * scale - l1) : 0.f;
          }
          *reinterpret_cast<float2*>(sdS + (size_t)r0 * LS + c0) = make_float2(p00, p01);
          *reinterpret_cast<float2*>(sdS + (size_t)r1 * LS + c0) = make_float2(p10, p11);
          *reinterpret_cast<float2*>(sPd + (size_t)r0 * LS + c0) = make_float2(p00 * d0[2 * e], p01 * d0[2 * e + 1]);
          *reinterpret_cast<float2*>(sPd + (size_t)r1 * LS + c0) = make_float2(p10 * d1[2 * e], p11 * d1[2 * e + 1]);
        }
      }
    }
    // ---- dP = dO V^T ----
#pragma unroll
    for (int lt = 0; lt < NL; ++lt) { acc[lt][0] = acc[lt][1] = acc[lt][2] = acc[lt][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < HD / 8; ++kk) {
      uint32_t a[4];
      lda_frag(a, sdO, LQ, m0, kk * 8, g, t);
#pragma unroll
      for (int lt = 0; lt < NL; ++lt) {
        const int n = 4 * (lt >> 1) + 2 * hh + (lt & 1);
        if (n < nt) {
          uint32_t bb[2];
          ldb_frag_nk(bb, sV, LQ, n * 8, kk * 8, g, t);
          mma_t<X3>(acc[lt], a, bb);
        }
      }
    }
    // ---- D_i = sum_j Pd_ij dP_ij over ALL key tiles of the row (see attention_tc_bwd_kernel): this warp's key tiles,
    // then the partner warp's share through shared memory ----
    float D0 = 0.f, D1 = 0.f;
#pragma unroll
    for (int lt = 0; lt < NL; ++lt) {
      const int n = 4 * (lt >> 1) + 2 * hh + (lt & 1);
      if (n < nt) {
        const int c0 = n * 8 + 2 * t;
        const float2 pd0 = *reinterpret_cast<const float2*>(sPd + (size_t)r0 * LS + c0);
        const float2 pd1 = *reinterpret_cast<const float2*>(sPd + (size_t)r1 * LS + c0);
        D0 += pd0.x * acc[lt][0] + pd0.y * acc[lt][1];
        D1 += pd1.x * acc[lt][2] + pd1.y * acc[lt][3];
      }
    }
    D0 += __shfl_xor_sync(0xffffffffu, D0, 1); D0 += __shfl_xor_sync(0xffffffffu, D0, 2);
    D1 += __shfl_xor_sync(0xffffffffu, D1, 1); D1 += __shfl_xor_sync(0xffffffffu, D1, 2);
    if (t == 0) { sD[hh * Tp + r0] = D0; sD[hh * Tp + r1] = D1; }
    named_bar_sync(1 + wq, 64);   // both warps of the tile have published their partial row sums
    D0 = sD[r0] + sD[Tp + r0];    // same order in both warps: bit-identical D
    D1 = sD[r1] + sD[Tp + r1];
    // ---- dS = Pd * dP - P * D (each thread re-reads the P / Pd values it wrote), stored pre-multiplied by `scale` ----
#pragma unroll
    for (int lt = 0; lt < NL; ++lt) {
      const int n = 4 * (lt >> 1) + 2 * hh + (lt & 1);
      if (n < nt) {
        const int c0 = n * 8 + 2 * t;
        float2* ps0 = reinterpret_cast<float2*>(sdS + (size_t)r0 * LS + c0);
        float2* ps1 = reinterpret_cast<float2*>(sdS + (size_t)r1 * LS + c0);
        const float2 p0 = *ps0, p1 = *ps1;
        const float2 pd0 = *reinterpret_cast<const float2*>(sPd + (size_t)r0 * LS + c0);
        const float2 pd1 = *reinterpret_cast<const float2*>(sPd + (size_t)r1 * LS + c0);
        *ps0 = make_float2((pd0.x * acc[lt][0] - p0.x * D0) * scale, (pd0.y * acc[lt][1] - p0.y * D0) * scale);
        *ps1 = make_float2((pd1.x * acc[lt][2] - p1.x * D1) * scale, (pd1.y * acc[lt][3] - p1.y * D1) * scale);
      }
    }
    named_bar_sync(1 + wq, 64);   // both warps of the tile have published their dS columns
    // ---- dQ = dS K: this warp's half of the head-dim column tiles ----
    constexpr int NH = HD / 16;
    float o[NH][4];
#pragma unroll
    for (int n = 0; n < NH; ++n) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }
    for (int kk = 0; kk < nt; ++kk) {
      uint32_t a[4];
      lda_frag(a, sdS, LS, m0, kk * 8, g, t);
#pragma unroll
      for (int n = 0; n < NH; ++n) {
        uint32_t bb[2];
        ldb_frag_kn(bb, sK, LQ, (hh * NH + n) * 8, kk * 8, g, t);
        mma_t<X3>(o[n], a, bb);
      }
    }
#pragma unroll
    for (int n = 0; n < NH; ++n) {
      const int c0 = (hh * NH + n) * 8 + 2 * t;
      if (r0 < T) *reinterpret_cast<float2*>(dq + q_b + (size_t)r0 * q_ts + (size_t)h * HD + c0) = make_float2(o[n][0], o[n][1]);
      if (r1 < T) *reinterpret_cast<float2*>(dq + q_b + (size_t)r1 * q_ts + (size_t)h * HD + c0) = make_float2(o[n][2], o[n][3]);
      if (dbias_qkv) {  // rows >= T contribute exact zeros (dS rows are zero there)
        float c_even = o[n][0] + o[n][2], c_odd = o[n][1] + o[n][3];
#pragma unroll
        for (int off = 4; off < 32; off <<= 1) {
          c_even += __shfl_xor_sync(0xffffffffu, c_even, off);
          c_odd += __shfl_xor_sync(0xffffffffu, c_odd, off);
        }
        if (g == 0) { atomicAdd(&sDb[c0], c_even); atomicAdd(&sDb[c0 + 1], c_odd); }
      }
    }
  }
  __syncthreads();
  // ================= phase B: per key tile - hh == 0: dK = dS^T Q, hh == 1: dV = Pd^T dO =================
  {
    const float* sA = hh ? sPd : sdS;
    const float* sB = hh ? sdO : sQ;
    float* outp = hh ? dv : dk;
    float* sDbo = sDb + (hh ? 2 * HD : HD);
    for (int j0 = wq * 16; j0 < Sp; j0 += 64) {
      float oa[HD / 8][4];
#pragma unroll
      for (int n = 0; n < HD / 8; ++n) { oa[n][0] = oa[n][1] = oa[n][2] = oa[n][3] = 0.f; }
      // queries that can see any key of this tile: i >= j0 - (S - T) under the causal mask
      int kk0 = 0;
      if (causal) kk0 = max(0, (j0 - (S - T))) / 8;
      for (int kk = kk0; kk < Tp / 8; ++kk) {
        uint32_t a[4];
        lda_frag_t(a, sA, LS, j0, kk * 8, g, t);
#pragma unroll
        for (int n = 0; n < HD / 8; ++n) {
          uint32_t bb[2];
          ldb_frag_kn(bb, sB, LQ, n * 8, kk * 8, g, t);
          mma_t<X3>(oa[n], a, bb);
        }
      }
      const int r0 = j0 + g, r1 = r0 + 8;
#pragma unroll
      for (int n = 0; n < HD / 8; ++n) {
        const int c0 = n * 8 + 2 * t;
        if (r0 < S) *reinterpret_cast<float2*>(outp + kv_b + (size_t)r0 * kv_ts + (size_t)h * HD + c0) = make_float2(oa[n][0], oa[n][1]);
        if (r1 < S) *reinterpret_cast<float2*>(outp + kv_b + (size_t)r1 * kv_ts + (size_t)h * HD + c0) = make_float2(oa[n][2], oa[n][3]);
        if (dbias_qkv) {  // padded key rows hold exact zeros (their dS / Pd columns are zero)
          float ce = oa[n][0] + oa[n][2], co = oa[n][1] + oa[n][3];
#pragma unroll
          for (int off = 4; off < 32; off <<= 1) {
            ce += __shfl_xor_sync(0xffffffffu, ce, off);
            co += __shfl_xor_sync(0xffffffffu, co, off);
          }
          if (g == 0) { atomicAdd(&sDbo[c0], ce); atomicAdd(&sDbo[c0 + 1], co); }
        }
      }
    }
  }
  if (dbias_qkv) {
    __syncthreads();
    for (int i = tid; i < 3 * HD; i += 256)
      atomicAdd(dbias_qkv + (size_t)(i / HD) * H * HD + (size_t)h * HD + (i % HD), sDb[i]);
  }
}

size_t attn_tc_fwd_smem(int T, int S, int hd) {
  const int Tp = (T + 15) & ~15, Sp = (S + 7) & ~7;
  return ((size_t)Tp * (hd + 4) + (size_t)Sp * (hd + 4) + (size_t)Sp * (hd + 8) + (size_t)4 * 16 * (Sp + 4)) * sizeof(float);
}
size_t attn_tc_bwd_smem(int T, int S, int hd) {
  const int Tp = (T + 15) & ~15, Sp = (S + 15) & ~15;
  return ((size_t)2 * Tp * (hd + 4) + (size_t)2 * Sp * (hd + 4) + (size_t)2 * Tp * (Sp + 4) + 3 * hd + 3 * Tp) * sizeof(float);
}

template <int HD, int NT, bool X3>
static int launch_fwd(const float* q, const float* k, const float* v, float* ctx, float* lse, int B, int H, int T, int S,
                      int64_t q_bs, int64_t q_ts, int64_t kv_bs, int64_t kv_ts, int64_t o_bs, int64_t o_ts, float scale,
                      int causal, const int32_t* key_len, float p_drop, const uint64_t* seed_dev, uint32_t stream_id,
                      const int32_t* cu_rows, cudaStream_t stream) {
  static bool set = false;
  if (!set) { cudaFuncSetAttribute(attention_tc_fwd_kernel<HD, NT, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); set = true; }
  attention_tc_fwd_kernel<HD, NT, X3><<<B * H, 128, attn_tc_fwd_smem(T, S, HD), stream>>>(
      q, k, v, ctx, lse, H, T, S, q_bs, q_ts, kv_bs, kv_ts, o_bs, o_ts, scale, causal, key_len, p_drop, seed_dev, stream_id, cu_rows);
  g_launches.fetch_add(1);
  CAPDEC_LAUNCH_CHECK("attention_tc_fwd_kernel");
  return CAPDEC_OK;
}
template <int HD, int NT, bool X3>
static int launch_bwd(const float* q, const float* k, const float* v, const float* ctx, const float* dctx, const float* lse,
                      float* dq, float* dk, float* dv, float* dbias, int B, int H, int T, int S, int64_t q_bs, int64_t q_ts,
                      int64_t kv_bs, int64_t kv_ts, int64_t o_bs, int64_t o_ts, float scale, int causal,
                      const int32_t* key_len, float p_drop, const uint64_t* seed_dev, uint32_t stream_id,
                      const int32_t* cu_rows, cudaStream_t stream) {
  static bool set = false;
  if (!set) {
    cudaFuncSetAttribute(attention_tc_bwd_kernel<HD, NT, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(attention_tc_bwd8_kernel<HD, NT, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    set = true;
  }
  static const char* env8 = getenv("CAPDEC_ATTN_BWD8");   // bring-up switch: 0 = the 4-warp kernel
  if (env8 && env8[0] == '0') {
    attention_tc_bwd_kernel<HD, NT, X3><<<B * H, 128, attn_tc_bwd_smem(T, S, HD), stream>>>(
        q, k, v, ctx, dctx, lse, dq, dk, dv, dbias, H, T, S, q_bs, q_ts, kv_bs, kv_ts, o_bs, o_ts, scale, causal, key_len,
        p_drop, seed_dev, stream_id, cu_rows);
  } else {
    attention_tc_bwd8_kernel<HD, NT, X3><<<B * H, 256, attn_tc_bwd_smem(T, S, HD), stream>>>(
        q, k, v, ctx, dctx, lse, dq, dk, dv, dbias, H, T, S, q_bs, q_ts, kv_bs, kv_ts, o_bs, o_ts, scale, causal, key_len,
        p_drop, seed_dev, stream_id, cu_rows);
  }
  g_launches.fetch_add(1);
  CAPDEC_LAUNCH_CHECK("attention_tc_bwd_kernel");
  return CAPDEC_OK;
}

}  // namespace capdec

using namespace capdec;

// Same contract as capdec_attention_fwd / _bwd (attention.cu); returns CAPDEC_ERR_UNSUPPORTED (-3) when the shape does
// not fit the tensor-core kernel's shared-memory budget so that the caller can use the FFMA kernel instead.
static int attention_tc_fwd_impl(bool x3, const float* q, const float* k, const float* v, float* ctx, float* lse, int B,
                                 int H, int T, int S, int hd, int64_t q_bs, int64_t q_ts, int64_t kv_bs,
                                 int64_t kv_ts, int64_t o_bs, int64_t o_ts, float scale, int causal,
                                 const int32_t* key_len, float p_drop, const uint64_t* seed_dev,
                                 uint32_t stream_id, const int32_t* cu_rows, cudaStream_t stream) {
  CAPDEC_REQUIRE(q && k && v && ctx, "attention_tc_fwd: null argument");
  CAPDEC_REQUIRE(!cu_rows || (T == S && causal), "attention_tc_fwd: packed rows are for causal self-attention");
  CAPDEC_REQUIRE(B > 0 && H > 0 && T > 0 && S > 0 && T <= 128 && S <= 128, "attention_tc_fwd: T,S must be in 1..128");
  CAPDEC_REQUIRE(hd == 64 || hd == 96, "attention_tc_fwd: head_dim must be 64 or 96");
  CAPDEC_REQUIRE(q_ts % 4 == 0 && kv_ts % 4 == 0 && o_ts % 2 == 0, "attention_tc_fwd: misaligned strides");
  if (attn_tc_fwd_smem(T, S, hd) > 227 * 1024) { set_last_error("attention_tc_fwd: tile does not fit shared memory"); return CAPDEC_ERR_UNSUPPORTED; }
#define FWD_ARGS q, k, v, ctx, lse, B, H, T, S, q_bs, q_ts, kv_bs, kv_ts, o_bs, o_ts, scale, causal, key_len, p_drop, seed_dev, stream_id, cu_rows, stream
  if (x3) {
    if (hd == 64) return S <= 64 ? launch_fwd<64, 8, true>(FWD_ARGS) : launch_fwd<64, 16, true>(FWD_ARGS);
    return S <= 64 ? launch_fwd<96, 8, true>(FWD_ARGS) : launch_fwd<96, 16, true>(FWD_ARGS);
  }
  if (hd == 64) return S <= 64 ? launch_fwd<64, 8, false>(FWD_ARGS) : launch_fwd<64, 16, false>(FWD_ARGS);
  return S <= 64 ? launch_fwd<96, 8, false>(FWD_ARGS) : launch_fwd<96, 16, false>(FWD_ARGS);
#undef FWD_ARGS
}
#define ATTN_FWD_PARAMS const float* q, const float* k, const float* v, float* ctx, float* lse, int B, int H, int T, int S,     \
                        int hd, int64_t q_bs, int64_t q_ts, int64_t kv_bs, int64_t kv_ts, int64_t o_bs, int64_t o_ts,          \
                        float scale, int causal, const int32_t* key_len, float p_drop, const uint64_t* seed_dev,               \
                        uint32_t stream_id, const int32_t* cu_rows, capdec_stream_t stream_
#define ATTN_FWD_FORWARD q, k, v, ctx, lse, B, H, T, S, hd, q_bs, q_ts, kv_bs, kv_ts, o_bs, o_ts, scale, causal, key_len, p_drop, \
                         seed_dev, stream_id, cu_rows, reinterpret_cast<cudaStream_t>(stream_)
extern "C" int capdec_attention_tc_fwd(ATTN_FWD_PARAMS) { return attention_tc_fwd_impl(false, ATTN_FWD_FORWARD); }
// fp32-grade variant (3xTF32 split in registers, accurate exp / log): the "tf32x3" mode's attention
extern "C" int capdec_attention_tc_fwd_x3(ATTN_FWD_PARAMS) { return attention_tc_fwd_impl(true, ATTN_FWD_FORWARD); }

static int attention_tc_bwd_impl(bool x3, const float* q, const float* k, const float* v, const float* ctx,
                                 const float* dctx, const float* lse, float* dq, float* dk, float* dv,
                                 float* dbias_qkv, int B, int H, int T, int S, int hd, int64_t q_bs, int64_t q_ts,
                                 int64_t kv_bs, int64_t kv_ts, int64_t o_bs, int64_t o_ts, float scale,
                                 int causal, const int32_t* key_len, float p_drop, const uint64_t* seed_dev,
                                 uint32_t stream_id, const int32_t* cu_rows, cudaStream_t stream) {
  CAPDEC_REQUIRE(q && k && v && ctx && dctx && lse && dq && dk && dv, "attention_tc_bwd: null argument");
  CAPDEC_REQUIRE(!cu_rows || (T == S && causal), "attention_tc_bwd: packed rows are for causal self-attention");
  CAPDEC_REQUIRE(B > 0 && H > 0 && T > 0 && S > 0 && T <= 128 && S <= 128, "attention_tc_bwd: T,S must be in 1..128");
  CAPDEC_REQUIRE(hd == 64 || hd == 96, "attention_tc_bwd: head_dim must be 64 or 96");
  CAPDEC_REQUIRE(q_ts % 4 == 0 && kv_ts % 4 == 0 && o_ts % 4 == 0, "attention_tc_bwd: misaligned strides");
  if (attn_tc_bwd_smem(T, S, hd) > 227 * 1024) { set_last_error("attention_tc_bwd: tile does not fit shared memory"); return CAPDEC_ERR_UNSUPPORTED; }
#define BWD_ARGS q, k, v, ctx, dctx, lse, dq, dk, dv, dbias_qkv, B, H, T, S, q_bs, q_ts, kv_bs, kv_ts, o_bs, o_ts, scale, causal, key_len, p_drop, seed_dev, stream_id, cu_rows, stream
  if (x3) {
    if (hd == 64) return S <= 64 ? launch_bwd<64, 8, true>(BWD_ARGS) : launch_bwd<64, 16, true>(BWD_ARGS);
    return S <= 64 ? launch_bwd<96, 8, true>(BWD_ARGS) : launch_bwd<96, 16, true>(BWD_ARGS);
  }
  if (hd == 64) return S <= 64 ? launch_bwd<64, 8, false>(BWD_ARGS) : launch_bwd<64, 16, false>(BWD_ARGS);
  return S <= 64 ? launch_bwd<96, 8, false>(BWD_ARGS) : launch_bwd<96, 16, false>(BWD_ARGS);
#undef BWD_ARGS
}
#define ATTN_BWD_PARAMS const float* q, const float* k, const float* v, const float* ctx, const float* dctx, const float* lse, \
                        float* dq, float* dk, float* dv, float* dbias_qkv, int B, int H, int T, int S, int hd, int64_t q_bs,   \
                        int64_t q_ts, int64_t kv_bs, int64_t kv_ts, int64_t o_bs, int64_t o_ts, float scale, int causal,       \
                        const int32_t* key_len, float p_drop, const uint64_t* seed_dev, uint32_t stream_id,                    \
                        const int32_t* cu_rows, capdec_stream_t stream_
#define ATTN_BWD_FORWARD q, k, v, ctx, dctx, lse, dq, dk, dv, dbias_qkv, B, H, T, S, hd, q_bs, q_ts, kv_bs, kv_ts, o_bs, o_ts,   \
                         scale, causal, key_len, p_drop, seed_dev, stream_id, cu_rows, reinterpret_cast<cudaStream_t>(stream_)
extern "C" int capdec_attention_tc_bwd(ATTN_BWD_PARAMS) { return attention_tc_bwd_impl(false, ATTN_BWD_FORWARD); }
extern "C" int capdec_attention_tc_bwd_x3(ATTN_BWD_PARAMS) { return attention_tc_bwd_impl(true, ATTN_BWD_FORWARD); }
