// capdec_b200 — single-tile multi-head attention, forward and backward, fp32 on CUDA cores.
//   GPT-2:   softmax(q k^T / sqrt(64) + causal mask (+ key padding mask)) -> dropout -> . v
//            (HF:modeling_gpt2.py:54-72 eager_attention_forward, q|k|v split :185-191)
//   mapper:  softmax(q k^T * 96^-0.5) . v, 8 heads x 96, no mask   (train.py:150-167)
// T, S <= 128 on this path (train: P+L <= 80; decode: <= P+67), so a whole (batch, head) lives in one CTA's
// shared memory: scores never touch HBM (the reference materialises [B,H,T,T] scores).
// Attention is 0.7 % of the step FLOPs, so it stays exact fp32 FFMA (bit-faithful softmax for the parity tests);
// the tcgen05 budget goes to the GEMMs.  Each query row is owned by 4 lanes (head_dim/4 each, 2 shuffles per dot).
#include "../../include/capdec_b200.h"
#include "common.cuh"

namespace capdec {

__device__ __forceinline__ float group4_sum(float v, uint32_t mask) {
  v += __shfl_xor_sync(mask, v, 1);
  v += __shfl_xor_sync(mask, v, 2);
  return v;
}

// dropout keep-scale of probability element (grow = (b*H+h)*T + i, column j); same (row, col) -> Philox mapping as the
// tensor-core kernels (attention_tc.cu): quad = grow*32 + (j>>4)*4 + ((j&7)>>1), component = ((j>>3)&1)*2 + (j&1)
__device__ __forceinline__ float drop_scale_elem(uint64_t seed, uint32_t stream_id, uint64_t grow, int j, float p, float inv_keep) {
  uint32_t r[4];
  Philox::gen(seed, stream_id, grow * 32 + (uint64_t)((j >> 4) * 4 + ((j & 7) >> 1)), r);
  const uint32_t thr = (uint32_t)(p * 4294967296.0f);
  return (r[((j >> 3) & 1) * 2 + (j & 1)] >= thr) ? inv_keep : 0.0f;
}

template <int HD>
__global__ void __launch_bounds__(512) attention_fwd_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                     const float* __restrict__ v, float* __restrict__ ctx, float* __restrict__ lse,
                                     int H, int T, int S, int64_t q_bs, int64_t q_ts, int64_t kv_bs, int64_t kv_ts,
                                     int64_t o_bs, int64_t o_ts, float scale, int causal,
                                     const int32_t* __restrict__ key_len, float p_drop, const uint64_t* seed_dev,
                                     uint32_t stream_id) {
  const uint64_t seed = seed_dev ? *seed_dev : 0ull;  // device-resident: fresh masks under CUDA-graph replay
  constexpr int DPL = HD / 4;   // dims per lane
  constexpr int V4 = DPL / 4;   // float4 per lane
  extern __shared__ __align__(16) float smem[];
  float* sK = smem;
  float* sV = smem + (size_t)S * HD;
  const int bh = blockIdx.x, b = bh / H, h = bh % H;
  // cooperative, coalesced load of this head's K and V slices
  const int hd4 = HD / 4;
  for (int i = threadIdx.x; i < S * hd4; i += blockDim.x) {
    const int j = i / hd4, c = i % hd4;
    const size_t g = (size_t)b * kv_bs + (size_t)j * kv_ts + (size_t)h * HD + 4 * c;
    reinterpret_cast<float4*>(sK)[i] = *reinterpret_cast<const float4*>(k + g);
    reinterpret_cast<float4*>(sV)[i] = *reinterpret_cast<const float4*>(v + g);
  }
  __syncthreads();
  const int row = threadIdx.x >> 2, sub = threadIdx.x & 3;
  if (row >= T) return;
  const uint32_t gmask = 0xFu << ((threadIdx.x & 31) & ~3);
  float qv[DPL], acc[DPL];
  {
    const float* qp = q + (size_t)b * q_bs + (size_t)row * q_ts + (size_t)h * HD + sub * DPL;
#pragma unroll
    for (int i = 0; i < V4; ++i) {
      const float4 t = *reinterpret_cast<const float4*>(qp + 4 * i);
      qv[4 * i] = t.x * scale; qv[4 * i + 1] = t.y * scale; qv[4 * i + 2] = t.z * scale; qv[4 * i + 3] = t.w * scale;
    }
  }
#pragma unroll
  for (int i = 0; i < DPL; ++i) acc[i] = 0.f;
  int jmax = causal ? min(S, row + 1 + (S - T)) : S;
  if (key_len) jmax = min(jmax, key_len[b]);
  float m = -INFINITY, l = 0.f;
  const float inv_keep = 1.0f / (1.0f - p_drop);
  const uint64_t grow = (uint64_t)bh * T + row;
  for (int j = 0; j < jmax; ++j) {
    const float4* kp = reinterpret_cast<const float4*>(sK + (size_t)j * HD + sub * DPL);
    float part = 0.f;
#pragma unroll
    for (int i = 0; i < V4; ++i) {
      const float4 t = kp[i];
      part = fmaf(qv[4 * i], t.x, part); part = fmaf(qv[4 * i + 1], t.y, part);
      part = fmaf(qv[4 * i + 2], t.z, part); part = fmaf(qv[4 * i + 3], t.w, part);
    }
    const float s = group4_sum(part, gmask);
    const float m_new = fmaxf(m, s);
    const float alpha = __expf(m - m_new);  // exp(-inf) = 0 on the first key
    const float pe = __expf(s - m_new);
    l = l * alpha + pe;
    m = m_new;
    float pv = pe;
    if (p_drop > 0.f) pv *= drop_scale_elem(seed, stream_id, grow, j, p_drop, inv_keep);
    const float4* vp = reinterpret_cast<const float4*>(sV + (size_t)j * HD + sub * DPL);
#pragma unroll
    for (int i = 0; i < V4; ++i) {
      const float4 t = vp[i];
      acc[4 * i] = fmaf(acc[4 * i], alpha, pv * t.x); acc[4 * i + 1] = fmaf(acc[4 * i + 1], alpha, pv * t.y);
      acc[4 * i + 2] = fmaf(acc[4 * i + 2], alpha, pv * t.z); acc[4 * i + 3] = fmaf(acc[4 * i + 3], alpha, pv * t.w);
    }
  }
  const float inv_l = l > 0.f ? 1.0f / l : 0.f;
  float* op = ctx + (size_t)b * o_bs + (size_t)row * o_ts + (size_t)h * HD + sub * DPL;
#pragma unroll
  for (int i = 0; i < V4; ++i)
    *reinterpret_cast<float4*>(op + 4 * i) =
        make_float4(acc[4 * i] * inv_l, acc[4 * i + 1] * inv_l, acc[4 * i + 2] * inv_l, acc[4 * i + 3] * inv_l);
  if (sub == 0 && lse) lse[(size_t)bh * T + row] = m + logf(l);
}

// Backward.  Pass 1 (4 lanes per query row): D_i = dO_i . O_i, dQ_i.  Pass 2 (4 lanes per key row): dK_j, dV_j.
template <int HD>
__global__ void __launch_bounds__(512) attention_bwd_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                     const float* __restrict__ v, const float* __restrict__ ctx,
                                     const float* __restrict__ dctx, const float* __restrict__ lse,
                                     float* __restrict__ dq, float* __restrict__ dk, float* __restrict__ dv,
                                     float* __restrict__ dbias_qkv, int H, int T, int S, int64_t q_bs, int64_t q_ts, int64_t kv_bs, int64_t kv_ts,
                                     int64_t o_bs, int64_t o_ts, float scale, int causal,
                                     const int32_t* __restrict__ key_len, float p_drop, const uint64_t* seed_dev,
                                     uint32_t stream_id) {
  const uint64_t seed = seed_dev ? *seed_dev : 0ull;  // device-resident: fresh masks under CUDA-graph replay
  constexpr int DPL = HD / 4;
  constexpr int V4 = DPL / 4;
  extern __shared__ __align__(16) float smem[];
  float* sK = smem;
  float* sV = sK + (size_t)S * HD;
  float* sQ = sV + (size_t)S * HD;
  float* sdO = sQ + (size_t)T * HD;
  float* sLse = sdO + (size_t)T * HD;
  float* sD = sLse + T;
  float* sDb = sD + T;  // [3][HD] column sums of dq | dk | dv (bias gradient of the fused QKV projection)
  const int bh = blockIdx.x, b = bh / H, h = bh % H;
  const int hd4 = HD / 4;
  for (int i = threadIdx.x; i < 3 * HD; i += blockDim.x) sDb[i] = 0.f;
  for (int i = threadIdx.x; i < S * hd4; i += blockDim.x) {
    const int j = i / hd4, c = i % hd4;
    const size_t g = (size_t)b * kv_bs + (size_t)j * kv_ts + (size_t)h * HD + 4 * c;
    reinterpret_cast<float4*>(sK)[i] = *reinterpret_cast<const float4*>(k + g);
    reinterpret_cast<float4*>(sV)[i] = *reinterpret_cast<const float4*>(v + g);
  }
  for (int i = threadIdx.x; i < T * hd4; i += blockDim.x) {
    const int t = i / hd4, c = i % hd4;
    const size_t gq = (size_t)b * q_bs + (size_t)t * q_ts + (size_t)h * HD + 4 * c;
    const size_t go = (size_t)b * o_bs + (size_t)t * o_ts + (size_t)h * HD + 4 * c;
    float4 qq = *reinterpret_cast<const float4*>(q + gq);
    qq.x *= scale; qq.y *= scale; qq.z *= scale; qq.w *= scale;  // sQ holds q * scale
    reinterpret_cast<float4*>(sQ)[i] = qq;
    reinterpret_cast<float4*>(sdO)[i] = *reinterpret_cast<const float4*>(dctx + go);
  }
  for (int i = threadIdx.x; i < T; i += blockDim.x) sLse[i] = lse[(size_t)bh * T + i];
  __syncthreads();

  const int r = threadIdx.x >> 2, sub = threadIdx.x & 3;
  const uint32_t gmask = 0xFu << ((threadIdx.x & 31) & ~3);
  const float inv_keep = 1.0f / (1.0f - p_drop);
  const int klen = key_len ? min(S, (int)key_len[b]) : S;

  // ---------------- pass 1: query rows ----------------
  if (r < T) {
    float qv[DPL], dov[DPL], dqa[DPL];
    const float* qp = sQ + (size_t)r * HD + sub * DPL;
    const float* dop = sdO + (size_t)r * HD + sub * DPL;
    const float* op = ctx + (size_t)b * o_bs + (size_t)r * o_ts + (size_t)h * HD + sub * DPL;
    float dpart = 0.f;
#pragma unroll
    for (int i = 0; i < DPL; ++i) {
      qv[i] = qp[i];
      dov[i] = dop[i];
      dqa[i] = 0.f;
      dpart = fmaf(dov[i], op[i], dpart);
    }
    const float Di = group4_sum(dpart, gmask);
    if (sub == 0) sD[r] = Di;
    const float lse_i = sLse[r];
    int jmax = causal ? min(S, r + 1 + (S - T)) : S;
    jmax = min(jmax, klen);
    const uint64_t grow = (uint64_t)bh * T + r;
    for (int j = 0; j < jmax; ++j) {
      const float* kp = sK + (size_t)j * HD + sub * DPL;
      const float* vp = sV + (size_t)j * HD + sub * DPL;
      float sp = 0.f, dpp = 0.f;
#pragma unroll
      for (int i = 0; i < DPL; ++i) { sp = fmaf(qv[i], kp[i], sp); dpp = fmaf(dov[i], vp[i], dpp); }
      const float s = group4_sum(sp, gmask);
      float dp = group4_sum(dpp, gmask);
      const float p = __expf(s - lse_i);
      if (p_drop > 0.f) dp *= drop_scale_elem(seed, stream_id, grow, j, p_drop, inv_keep);
      const float ds = p * (dp - Di);
#pragma unroll
      for (int i = 0; i < DPL; ++i) dqa[i] = fmaf(ds, kp[i], dqa[i]);
    }
    float* dqp = dq + (size_t)b * q_bs + (size_t)r * q_ts + (size_t)h * HD + sub * DPL;
#pragma unroll
    for (int i = 0; i < V4; ++i)
      *reinterpret_cast<float4*>(dqp + 4 * i) =
          make_float4(dqa[4 * i] * scale, dqa[4 * i + 1] * scale, dqa[4 * i + 2] * scale, dqa[4 * i + 3] * scale);
    if (dbias_qkv) {
#pragma unroll
      for (int i = 0; i < DPL; ++i) atomicAdd(&sDb[sub * DPL + i], dqa[i] * scale);
    }
  }
  __syncthreads();
  // ---------------- pass 2: key rows ----------------
  if (r < S) {
    const int j = r;
    float kv_[DPL], vv[DPL], dka[DPL], dva[DPL];
    const float* kp = sK + (size_t)j * HD + sub * DPL;
    const float* vp = sV + (size_t)j * HD + sub * DPL;
#pragma unroll
    for (int i = 0; i < DPL; ++i) { kv_[i] = kp[i]; vv[i] = vp[i]; dka[i] = 0.f; dva[i] = 0.f; }
    int i0 = causal ? max(0, j - (S - T)) : 0;
    if (j >= klen) i0 = T;  // masked key: no gradient
    for (int i = i0; i < T; ++i) {
      const float* qp = sQ + (size_t)i * HD + sub * DPL;
      const float* dop = sdO + (size_t)i * HD + sub * DPL;
      float sp = 0.f, dpp = 0.f;
#pragma unroll
      for (int c = 0; c < DPL; ++c) { sp = fmaf(qp[c], kv_[c], sp); dpp = fmaf(dop[c], vv[c], dpp); }
      const float s = group4_sum(sp, gmask);
      float dp = group4_sum(dpp, gmask);
      const float p = __expf(s - sLse[i]);
      float pd = p;
      if (p_drop > 0.f) {
        const float sc = drop_scale_elem(seed, stream_id, (uint64_t)bh * T + i, j, p_drop, inv_keep);
        pd *= sc;
        dp *= sc;
      }
      const float ds = p * (dp - sD[i]);
#pragma unroll
      for (int c = 0; c < DPL; ++c) { dva[c] = fmaf(pd, dop[c], dva[c]); dka[c] = fmaf(ds, qp[c], dka[c]); }
    }
    float* dkp = dk + (size_t)b * kv_bs + (size_t)j * kv_ts + (size_t)h * HD + sub * DPL;
    float* dvp = dv + (size_t)b * kv_bs + (size_t)j * kv_ts + (size_t)h * HD + sub * DPL;
#pragma unroll
    for (int i = 0; i < V4; ++i) {
      // sQ already carries `scale`, so dK needs no extra factor
      *reinterpret_cast<float4*>(dkp + 4 * i) = make_float4(dka[4 * i], dka[4 * i + 1], dka[4 * i + 2], dka[4 * i + 3]);
      *reinterpret_cast<float4*>(dvp + 4 * i) = make_float4(dva[4 * i], dva[4 * i + 1], dva[4 * i + 2], dva[4 * i + 3]);
    }
    if (dbias_qkv) {
#pragma unroll
      for (int i = 0; i < DPL; ++i) {
        atomicAdd(&sDb[HD + sub * DPL + i], dka[i]);
        atomicAdd(&sDb[2 * HD + sub * DPL + i], dva[i]);
      }
    }
  }
  if (dbias_qkv) {  // one global atomic per (head, column): dbias laid out [q | k | v] x [H*HD] like c_attn.bias
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * HD; i += blockDim.x)
      atomicAdd(dbias_qkv + (size_t)(i / HD) * H * HD + (size_t)h * HD + (i % HD), sDb[i]);
  }
}

static int check_attn_args(int B, int H, int T, int S, int hd, int64_t q_ts, int64_t kv_ts, int64_t o_ts) {
  CAPDEC_REQUIRE(B > 0 && H > 0 && T > 0 && S > 0, "attention: bad shape");
  CAPDEC_REQUIRE(T <= 128 && S <= 128, "attention: single-tile kernel supports T,S <= 128 (got T=%d S=%d)", T, S);
  CAPDEC_REQUIRE(hd == 64 || hd == 96, "attention: head_dim must be 64 (GPT-2) or 96 (mapper), got %d", hd);
  CAPDEC_REQUIRE(q_ts % 4 == 0 && kv_ts % 4 == 0 && o_ts % 4 == 0, "attention: strides must be multiples of 4 floats");
  return CAPDEC_OK;
}

}  // namespace capdec

using namespace capdec;

extern "C" int capdec_attention_fwd(const float* q, const float* k, const float* v, float* ctx, float* lse, int B,
                                    int H, int T, int S, int hd, int64_t q_bs, int64_t q_ts, int64_t kv_bs,
                                    int64_t kv_ts, int64_t o_bs, int64_t o_ts, float scale, int causal,
                                    const int32_t* key_len, float p_drop, const uint64_t* seed_dev, uint32_t stream_id,
                                    capdec_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CAPDEC_REQUIRE(q && k && v && ctx, "attention_fwd: null argument");
  int rc = check_attn_args(B, H, T, S, hd, q_ts, kv_ts, o_ts);
  if (rc) return rc;
  const int threads = ((4 * T + 31) / 32) * 32;
  const size_t smem = (size_t)2 * S * hd * sizeof(float);
  if (hd == 64) {
    static bool set64 = false;
    if (!set64) { cudaFuncSetAttribute(attention_fwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024); set64 = true; }
    attention_fwd_kernel<64><<<B * H, threads, smem, stream>>>(q, k, v, ctx, lse, H, T, S, q_bs, q_ts, kv_bs, kv_ts, o_bs, o_ts,
                                                               scale, causal, key_len, p_drop, seed_dev, stream_id);
  } else {
    static bool set96 = false;
    if (!set96) { cudaFuncSetAttribute(attention_fwd_kernel<96>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024); set96 = true; }
    attention_fwd_kernel<96><<<B * H, threads, smem, stream>>>(q, k, v, ctx, lse, H, T, S, q_bs, q_ts, kv_bs, kv_ts, o_bs, o_ts,
                                                               scale, causal, key_len, p_drop, seed_dev, stream_id);
  }
  g_launches.fetch_add(1);
  CAPDEC_LAUNCH_CHECK("attention_fwd_kernel");
  return CAPDEC_OK;
}

extern "C" int capdec_attention_bwd(const float* q, const float* k, const float* v, const float* ctx,
                                    const float* dctx, const float* lse, float* dq, float* dk, float* dv, float* dbias_qkv, int B,
                                    int H, int T, int S, int hd, int64_t q_bs, int64_t q_ts, int64_t kv_bs, int64_t kv_ts,
                                    int64_t o_bs, int64_t o_ts, float scale, int causal, const int32_t* key_len,
                                    float p_drop, const uint64_t* seed_dev, uint32_t stream_id, capdec_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CAPDEC_REQUIRE(q && k && v && ctx && dctx && lse && dq && dk && dv, "attention_bwd: null argument");
  int rc = check_attn_args(B, H, T, S, hd, q_ts, kv_ts, o_ts);
  if (rc) return rc;
  const int mx = T > S ? T : S;
  const int threads = ((4 * mx + 31) / 32) * 32;
  const size_t smem = ((size_t)2 * S * hd + (size_t)2 * T * hd + 2 * T + 3 * hd) * sizeof(float);
  if (hd == 64) {
    static bool set64 = false;
    if (!set64) { cudaFuncSetAttribute(attention_bwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); set64 = true; }
    attention_bwd_kernel<64><<<B * H, threads, smem, stream>>>(q, k, v, ctx, dctx, lse, dq, dk, dv, dbias_qkv, H, T, S, q_bs, q_ts, kv_bs,
                                                               kv_ts, o_bs, o_ts, scale, causal, key_len, p_drop, seed_dev, stream_id);
  } else {
    static bool set96 = false;
    if (!set96) { cudaFuncSetAttribute(attention_bwd_kernel<96>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); set96 = true; }
    attention_bwd_kernel<96><<<B * H, threads, smem, stream>>>(q, k, v, ctx, dctx, lse, dq, dk, dv, dbias_qkv, H, T, S, q_bs, q_ts, kv_bs,
                                                               kv_ts, o_bs, o_ts, scale, causal, key_len, p_drop, seed_dev, stream_id);
  }
  g_launches.fetch_add(1);
  CAPDEC_LAUNCH_CHECK("attention_bwd_kernel");
  return CAPDEC_OK;
}
