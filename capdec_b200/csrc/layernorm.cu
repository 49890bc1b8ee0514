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
                                                         float p_drop, const uint64_t* seed_dev, uint32_t stream_id,
                                                         const int32_t* __restrict__ rows_dev) {
  const uint64_t seed = seed_dev ? *seed_dev : 0ull;  // device-resident: fresh masks under CUDA-graph replay
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (rows_dev) rows = min(rows, __ldg(rows_dev));     // packed rows: the live row count of this step
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

// Backward, column-parallel: thread c of a CTA owns float4 column c of every row the CTA processes (d/4 threads), so
// the column reductions (dgamma, dbeta, bias gradient of the branch) are 12 private registers per thread and need no
// atomics until the very end; the two ROW reductions (mean of g, mean of g*xhat) are done for kR rows at a time with
// one shuffle tree + one shared-memory exchange.  ~4x fewer instructions than a warp-per-row kernel with shared-memory
// atomics (17.5 M -> ~4.6 M warp instructions at C2) and enough CTAs per SM to stream at HBM speed.
constexpr int kR = 4;  // rows per batch

__global__ void __launch_bounds__(256) add_ln_bwd_kernel(const float* __restrict__ dx, const float* __restrict__ r,
                                                         const float* __restrict__ stats,
                                                         const float* __restrict__ gamma,
                                                         const float* __restrict__ dh_res, float* __restrict__ dh_out,
                                                         float* __restrict__ dy, float* __restrict__ dgamma,
                                                         float* __restrict__ dbeta, float* __restrict__ dbias_branch,
                                                         int rows, int d, int rows_per_cta, float p_drop,
                                                         const uint64_t* seed_dev, uint32_t stream_id,
                                                         const int32_t* __restrict__ rows_dev) {
  const uint64_t seed = seed_dev ? *seed_dev : 0ull;  // device-resident: fresh masks under CUDA-graph replay
  if (rows_dev) {  // packed rows: re-balance the static grid over the live rows of this step
    const int rows_max = rows;
    rows = min(rows, __ldg(rows_dev));
    rows_per_cta = (((rows + (int)gridDim.x - 1) / (int)gridDim.x + kR - 1) / kR) * kR;
    if (rows_per_cta < kR) rows_per_cta = kR;
    // rows [live, next multiple of 32) are read by the K-limited weight-gradient GEMMs: they must be exact zeros
    if (blockIdx.x == gridDim.x - 1) {
      const int tail1 = min(rows_max, (rows + 31) & ~31);
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int row = rows; row < tail1; ++row) {
        reinterpret_cast<float4*>(dh_out)[(size_t)row * (d >> 2) + threadIdx.x] = z;
        if (dy) reinterpret_cast<float4*>(dy)[(size_t)row * (d >> 2) + threadIdx.x] = z;
      }
    }
  }
  __shared__ float s_part[2][8][2 * kR];              // [batch parity][warp][c1_0..c1_{R-1}, c2_0..c2_{R-1}]
  const int c = threadIdx.x;                          // float4 column
  const int d4 = d >> 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + c);
  const float inv_keep = 1.0f / (1.0f - p_drop);
  const float inv_d = 1.0f / (float)d;
  float4 dg = make_float4(0.f, 0.f, 0.f, 0.f), db = dg, dbr = dg;
  const int row0 = blockIdx.x * rows_per_cta, row1 = min(rows, row0 + rows_per_cta);
  int parity = 0;
  for (int rb = row0; rb < row1; rb += kR, parity ^= 1) {
    float4 g[kR], xh[kR], res[kR];
    float part[2 * kR];
    // issue every global load of the batch up front (dx, r, the residual gradient and the row statistics): the row index
    // is clamped instead of predicated so that nothing stops the compiler from hoisting all 12 16-byte loads
    float4 dvv[kR], rvv[kR];
    float mean[kR], rstd[kR];
#pragma unroll
    for (int i = 0; i < kR; ++i) {
      const size_t rowc = (size_t)min(rb + i, row1 - 1);
      res[i] = dh_res ? ld_stream(reinterpret_cast<const float4*>(dh_res) + rowc * d4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      dvv[i] = ld_stream(reinterpret_cast<const float4*>(dx) + rowc * d4 + c);
      rvv[i] = ld_stream(reinterpret_cast<const float4*>(r) + rowc * d4 + c);
      mean[i] = __ldg(stats + 2 * rowc);
      rstd[i] = __ldg(stats + 2 * rowc + 1);
    }
#pragma unroll
    for (int i = 0; i < kR; ++i) {
      const bool live = rb + i < row1;
      const float4 dv = live ? dvv[i] : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 rv = rvv[i];
      xh[i].x = (rv.x - mean[i]) * rstd[i]; xh[i].y = (rv.y - mean[i]) * rstd[i];
      xh[i].z = (rv.z - mean[i]) * rstd[i]; xh[i].w = (rv.w - mean[i]) * rstd[i];
      dg.x += dv.x * xh[i].x; dg.y += dv.y * xh[i].y; dg.z += dv.z * xh[i].z; dg.w += dv.w * xh[i].w;
      db.x += dv.x; db.y += dv.y; db.z += dv.z; db.w += dv.w;
      g[i].x = dv.x * gm.x; g[i].y = dv.y * gm.y; g[i].z = dv.z * gm.z; g[i].w = dv.w * gm.w;
      part[i] = (g[i].x + g[i].y) + (g[i].z + g[i].w);
      part[kR + i] = (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
    }
#pragma unroll
    for (int j = 0; j < 2 * kR; ++j) part[j] = warp_sum(part[j]);
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < 2 * kR; ++j) s_part[parity][warp][j] = part[j];
    }
    __syncthreads();  // one barrier per batch: the partials are double-buffered by batch parity
#pragma unroll
    for (int j = 0; j < 2 * kR; ++j) {
      float t = 0.f;
      for (int w = 0; w < nwarps; ++w) t += s_part[parity][w][j];
      part[j] = t * inv_d;
    }
#pragma unroll
    for (int i = 0; i < kR; ++i) {
      const int row = rb + i;
      if (row >= row1) break;
      const float c1 = part[i], c2 = part[kR + i];
      float4 o;
      o.x = rstd[i] * (g[i].x - c1 - xh[i].x * c2);
      o.y = rstd[i] * (g[i].y - c1 - xh[i].y * c2);
      o.z = rstd[i] * (g[i].z - c1 - xh[i].z * c2);
      o.w = rstd[i] * (g[i].w - c1 - xh[i].w * c2);
      o.x += res[i].x; o.y += res[i].y; o.z += res[i].z; o.w += res[i].w;
      reinterpret_cast<float4*>(dh_out)[(size_t)row * d4 + c] = o;
      if (dy || dbias_branch) {
        if (p_drop > 0.0f) {
          float sc[4];
          dropout_scale4(seed, stream_id, (uint64_t)row * d4 + c, p_drop, inv_keep, sc);
          o.x *= sc[0]; o.y *= sc[1]; o.z *= sc[2]; o.w *= sc[3];
        }
        if (dy) reinterpret_cast<float4*>(dy)[(size_t)row * d4 + c] = o;
        dbr.x += o.x; dbr.y += o.y; dbr.z += o.z; dbr.w += o.w;  // bias gradient of the branch's last Linear
      }
    }
  }
  if (dgamma) {
    float* a = dgamma + 4 * c;
    float* b = dbeta + 4 * c;
    atomicAdd(a + 0, dg.x); atomicAdd(a + 1, dg.y); atomicAdd(a + 2, dg.z); atomicAdd(a + 3, dg.w);
    atomicAdd(b + 0, db.x); atomicAdd(b + 1, db.y); atomicAdd(b + 2, db.z); atomicAdd(b + 3, db.w);
  }
  if (dbias_branch) {
    float* a = dbias_branch + 4 * c;
    atomicAdd(a + 0, dbr.x); atomicAdd(a + 1, dbr.y); atomicAdd(a + 2, dbr.z); atomicAdd(a + 3, dbr.w);
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
                                 const uint64_t* seed_dev, uint32_t stream_id, const int32_t* rows_dev,
                                 capdec_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CAPDEC_REQUIRE(h_in && x && stats && gamma && beta && rows > 0, "add_ln_fwd: null argument");
  CAPDEC_REQUIRE(d % 128 == 0 && d <= 128 * kMaxV, "add_ln_fwd: d=%d must be a multiple of 128 and <= 1024", d);
  CAPDEC_REQUIRE(p_drop >= 0.0f && p_drop < 1.0f, "add_ln_fwd: bad dropout p");
  const int nv = d / 128;
  const int grid = (rows + 7) / 8;
  DISPATCH_NV(nv, (add_ln_fwd_kernel<NV><<<grid, 256, 0, stream>>>(h_in, y, h_out, x, stats, gamma, beta, rows, d, eps,
                                                                   p_drop, seed_dev, stream_id, rows_dev)));
  g_launches.fetch_add(1);
  CAPDEC_LAUNCH_CHECK("add_ln_fwd_kernel");
  return CAPDEC_OK;
}

extern "C" int capdec_add_ln_bwd(const float* dx, const float* r, const float* stats, const float* gamma,
                                 const float* dh_res, float* dh_out, float* dy, float* dgamma, float* dbeta,
                                 float* dbias_branch, int rows, int d, float p_drop, const uint64_t* seed_dev, uint32_t stream_id,
                                 const int32_t* rows_dev, capdec_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CAPDEC_REQUIRE(dx && r && stats && gamma && dh_out && rows > 0, "add_ln_bwd: null argument");
  CAPDEC_REQUIRE(d % 128 == 0 && d <= 128 * kMaxV, "add_ln_bwd: d=%d must be a multiple of 128 and <= 1024", d);
  CAPDEC_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), "add_ln_bwd: dgamma and dbeta must both be set or both NULL");
  const int threads = d / 4;                     // one float4 column per thread (d = 768 -> 192 threads)
  int ctas = num_sms() * 4;
  int rpc = (rows + ctas - 1) / ctas;
  rpc = ((rpc + kR - 1) / kR) * kR;
  if (rpc < kR) rpc = kR;
  ctas = (rows + rpc - 1) / rpc;
  add_ln_bwd_kernel<<<ctas, threads, 0, stream>>>(dx, r, stats, gamma, dh_res, dh_out, dy, dgamma, dbeta, dbias_branch, rows,
                                                  d, rpc, p_drop, seed_dev, stream_id, rows_dev);
  g_launches.fetch_add(1);
  CAPDEC_LAUNCH_CHECK("add_ln_bwd_kernel");
  return CAPDEC_OK;
}
