// capdec_b200 — fused (residual add + dropout +) LayerNorm, forward and backward.
// Replaces nn.LayerNorm(768, eps=1e-5) (HF:modeling_gpt2.py:252-254,505; train.py:184-188), the residual adds
// (HF:modeling_gpt2.py:282,307; train.py:177-178) and resid_dropout (HF:modeling_gpt2.py:224,242).
// HBM-bound: one warp per row, 128-bit loads, the row lives in registers (d <= 1024), two shuffle reductions.
#include "../../include/capdec_b200.h"
#include "common.cuh"

namespace capdec {

constexpr int kMaxV = 8;  // float4 per lane -> d <= 1024

template <int NV>
__global__ void __launch_bounds__(256) add_ln_fwd_kernel(const float* __restrict__ h_in, const float* __restrict__ y,
                                                         float* __restrict__ h_out, float* __restrict__ x,
                                                         float* __restrict__ stats, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, int rows, int d, float eps,
                                                         float p_drop, const uint64_t* seed_dev, uint32_t stream_id) {
  const uint64_t seed = seed_dev ? *seed_dev : 0ull;  // device-resident: fresh masks under CUDA-graph replay
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= rows) return;
  const int d4 = d >> 2;
  const float4* hin4 = reinterpret_cast<const float4*>(h_in) + (size_t)row * d4;
  const float4* y4 = y ? reinterpret_cast<const float4*>(y) + (size_t)row * d4 : nullptr;
  float4 r[NV];
  const float inv_keep = 1.0f / (1.0f - p_drop);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    r[i] = ld_stream(hin4 + c);
    if (y4) {
      float4 yv = ld_stream(y4 + c);
      if (p_drop > 0.0f) {
        float s[4];
        dropout_scale4(seed, stream_id, (uint64_t)row * d4 + c, p_drop, inv_keep, s);
        yv.x *= s[0]; yv.y *= s[1]; yv.z *= s[2]; yv.w *= s[3];
      }
      r[i].x += yv.x; r[i].y += yv.y; r[i].z += yv.z; r[i].w += yv.w;
    }
  }
  if (y4 && h_out) {
    float4* ho4 = reinterpret_cast<float4*>(h_out) + (size_t)row * d4;
#pragma unroll
    for (int i = 0; i < NV; ++i) ho4[lane + 32 * i] = r[i];
  }
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += (r[i].x + r[i].y) + (r[i].z + r[i].w);
  const float mean = warp_sum(s) / (float)d;
  float v = 0.0f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = r[i].x - mean, b = r[i].y - mean, c = r[i].z - mean, e = r[i].w - mean;
    v += (a * a + b * b) + (c * c + e * e);
  }
  const float rstd = rsqrtf(warp_sum(v) / (float)d + eps);
  if (lane == 0) {
    stats[2 * (size_t)row] = mean;
    stats[2 * (size_t)row + 1] = rstd;
  }
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
  float4* x4 = reinterpret_cast<float4*>(x) + (size_t)row * d4;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    const float4 g = __ldg(g4 + c), bb = __ldg(b4 + c);
    float4 o;
    o.x = (r[i].x - mean) * rstd * g.x + bb.x;
    o.y = (r[i].y - mean) * rstd * g.y + bb.y;
    o.z = (r[i].z - mean) * rstd * g.z + bb.z;
    o.w = (r[i].w - mean) * rstd * g.w + bb.w;
    x4[c] = o;
  }
}

// persistent: each warp strides over rows.  The column reductions (dgamma, dbeta, bias gradient of the branch) are
// accumulated with shared-memory float atomics (conflict-free: a warp instruction touches 32 distinct banks) instead of
// per-lane register accumulators: that keeps the kernel at ~80 registers -> 3 CTAs/SM, which is what a streaming
// kernel needs to cover HBM latency (the register-accumulator version ran at 2.6 TB/s, this one is HBM-bound).
template <int NV>
__global__ void __launch_bounds__(256, 3) add_ln_bwd_kernel(const float* __restrict__ dx, const float* __restrict__ r,
                                                            const float* __restrict__ stats,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ dh_res, float* __restrict__ dh_out,
                                                            float* __restrict__ dy, float* __restrict__ dgamma,
                                                            float* __restrict__ dbeta, float* __restrict__ dbias_branch,
                                                            int rows, int d, float p_drop,
                                                            const uint64_t* seed_dev, uint32_t stream_id) {
  const uint64_t seed = seed_dev ? *seed_dev : 0ull;  // device-resident: fresh masks under CUDA-graph replay
  __shared__ float s_acc[3][NV * 128];  // dgamma | dbeta | dbias_branch partial sums of this CTA
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d4 = d >> 2;
  for (int i = threadIdx.x; i < 3 * NV * 128; i += 256) (&s_acc[0][0])[i] = 0.f;
  __syncthreads();
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float inv_keep = 1.0f / (1.0f - p_drop);
  const float inv_d = 1.0f / (float)d;
  for (int row = blockIdx.x * 8 + warp; row < rows; row += gridDim.x * 8) {
    const float4* dx4 = reinterpret_cast<const float4*>(dx) + (size_t)row * d4;
    const float4* r4 = reinterpret_cast<const float4*>(r) + (size_t)row * d4;
    const float mean = stats[2 * (size_t)row], rstd = stats[2 * (size_t)row + 1];
    float4 g[NV], xh[NV];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      const float4 dv = ld_stream(dx4 + c), rv = ld_stream(r4 + c);
      const float4 gm = __ldg(g4 + c);
      xh[i].x = (rv.x - mean) * rstd; xh[i].y = (rv.y - mean) * rstd;
      xh[i].z = (rv.z - mean) * rstd; xh[i].w = (rv.w - mean) * rstd;
      if (dgamma) {
        float* a0 = &s_acc[0][4 * c];
        float* a1 = &s_acc[1][4 * c];
        atomicAdd(a0 + 0, dv.x * xh[i].x); atomicAdd(a0 + 1, dv.y * xh[i].y);
        atomicAdd(a0 + 2, dv.z * xh[i].z); atomicAdd(a0 + 3, dv.w * xh[i].w);
        atomicAdd(a1 + 0, dv.x); atomicAdd(a1 + 1, dv.y); atomicAdd(a1 + 2, dv.z); atomicAdd(a1 + 3, dv.w);
      }
      g[i].x = dv.x * gm.x; g[i].y = dv.y * gm.y; g[i].z = dv.z * gm.z; g[i].w = dv.w * gm.w;
      c1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
      c2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
    }
    c1 = warp_sum(c1) * inv_d;
    c2 = warp_sum(c2) * inv_d;
    const float4* res4 = dh_res ? reinterpret_cast<const float4*>(dh_res) + (size_t)row * d4 : nullptr;
    float4* out4 = reinterpret_cast<float4*>(dh_out) + (size_t)row * d4;
    float4* dy4 = dy ? reinterpret_cast<float4*>(dy) + (size_t)row * d4 : nullptr;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      float4 o;
      o.x = rstd * (g[i].x - c1 - xh[i].x * c2);
      o.y = rstd * (g[i].y - c1 - xh[i].y * c2);
      o.z = rstd * (g[i].z - c1 - xh[i].z * c2);
      o.w = rstd * (g[i].w - c1 - xh[i].w * c2);
      if (res4) {
        const float4 rr = ld_stream(res4 + c);
        o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
      }
      out4[c] = o;
      if (dy4 || dbias_branch) {
        if (p_drop > 0.0f) {
          float s[4];
          dropout_scale4(seed, stream_id, (uint64_t)row * d4 + c, p_drop, inv_keep, s);
          o.x *= s[0]; o.y *= s[1]; o.z *= s[2]; o.w *= s[3];
        }
        if (dy4) dy4[c] = o;
        if (dbias_branch) {  // bias gradient of the branch's last Linear
          float* a2 = &s_acc[2][4 * c];
          atomicAdd(a2 + 0, o.x); atomicAdd(a2 + 1, o.y); atomicAdd(a2 + 2, o.z); atomicAdd(a2 + 3, o.w);
        }
      }
    }
  }
  if (dgamma || dbias_branch) {
    __syncthreads();
    for (int i = threadIdx.x; i < d; i += 256) {
      if (dgamma) { atomicAdd(dgamma + i, s_acc[0][i]); atomicAdd(dbeta + i, s_acc[1][i]); }
      if (dbias_branch) atomicAdd(dbias_branch + i, s_acc[2][i]);
    }
  }
}

}  // namespace capdec

using namespace capdec;

#define DISPATCH_NV(nv, ...)                                   \
  switch (nv) {                                                \
    case 1: { constexpr int NV = 1; __VA_ARGS__; } break;      \
    case 2: { constexpr int NV = 2; __VA_ARGS__; } break;      \
    case 4: { constexpr int NV = 4; __VA_ARGS__; } break;      \
    case 6: { constexpr int NV = 6; __VA_ARGS__; } break;      \
    case 8: { constexpr int NV = 8; __VA_ARGS__; } break;      \
    default:                                                   \
      set_last_error("layernorm: unsupported width d=%d (need d in {128,256,512,768,1024})", d); \
      return CAPDEC_ERR_UNSUPPORTED;                           \
  }

extern "C" int capdec_add_ln_fwd(const float* h_in, const float* y, float* h_out, float* x, float* stats,
                                 const float* gamma, const float* beta, int rows, int d, float eps, float p_drop,
                                 const uint64_t* seed_dev, uint32_t stream_id, capdec_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CAPDEC_REQUIRE(h_in && x && stats && gamma && beta && rows > 0, "add_ln_fwd: null argument");
  CAPDEC_REQUIRE(d % 128 == 0 && d <= 128 * kMaxV, "add_ln_fwd: d=%d must be a multiple of 128 and <= 1024", d);
  CAPDEC_REQUIRE(p_drop >= 0.0f && p_drop < 1.0f, "add_ln_fwd: bad dropout p");
  const int nv = d / 128;
  const int grid = (rows + 7) / 8;
  DISPATCH_NV(nv, (add_ln_fwd_kernel<NV><<<grid, 256, 0, stream>>>(h_in, y, h_out, x, stats, gamma, beta, rows, d, eps,
                                                                   p_drop, seed_dev, stream_id)));
  g_launches.fetch_add(1);
  CAPDEC_LAUNCH_CHECK("add_ln_fwd_kernel");
  return CAPDEC_OK;
}

extern "C" int capdec_add_ln_bwd(const float* dx, const float* r, const float* stats, const float* gamma,
                                 const float* dh_res, float* dh_out, float* dy, float* dgamma, float* dbeta,
                                 float* dbias_branch, int rows, int d, float p_drop, const uint64_t* seed_dev, uint32_t stream_id, capdec_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CAPDEC_REQUIRE(dx && r && stats && gamma && dh_out && rows > 0, "add_ln_bwd: null argument");
  CAPDEC_REQUIRE(d % 128 == 0 && d <= 128 * kMaxV, "add_ln_bwd: d=%d must be a multiple of 128 and <= 1024", d);
  CAPDEC_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), "add_ln_bwd: dgamma and dbeta must both be set or both NULL");
  const int nv = d / 128;
  int grid = (rows + 7) / 8;
  const int cap = num_sms() * 3;
  if (grid > cap) grid = cap;
  DISPATCH_NV(nv, (add_ln_bwd_kernel<NV><<<grid, 256, 0, stream>>>(dx, r, stats, gamma, dh_res, dh_out, dy, dgamma,
                                                                   dbeta, dbias_branch, rows, d, p_drop, seed_dev, stream_id)));
  g_launches.fetch_add(1);
  CAPDEC_LAUNCH_CHECK("add_ln_bwd_kernel");
  return CAPDEC_OK;
}
