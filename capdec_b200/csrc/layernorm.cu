// capdec_b200 — fused (residual add + dropout +) LayerNorm, forward and backward.
// Replaces nn.LayerNorm(768, eps=1e-5) (HF:modeling_gpt2.py:252-254,505; train.py:184-188), the residual adds
// (HF:modeling_gpt2.py:282,307; train.py:177-178) and resid_dropout (HF:modeling_gpt2.py:224,242).
// HBM-bound: one warp per row, 128-bit loads, the row lives in registers (d <= 1024), two shuffle reductions.
#include <stdlib.h>

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
  pdl_trigger();   // lets a programmatically launched successor (the GEMMs) set itself up while this grid drains
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

// Backward, bulk-copy pipelined variant (default).  Same column-parallel arithmetic as add_ln_bwd_kernel, but the rows are
// STAGED IN SHARED MEMORY by the TMA engine instead of being loaded into registers: a batch of kR consecutive rows of
// dx / r / dh_res is three contiguous 12 KB spans, so one elected thread issues three `cp.async.bulk` copies (+ one for
// the row statistics) per batch and keeps kLnStages batches in flight per CTA (2 CTAs per SM -> up to 216 KB of loads
// outstanding per SM, independent of occupancy and register pressure).  The register-staged kernel reached 2.8 TB/s
// (48 us for 135 MB at C2, profiles/r1_ncu_nongemm.md: 17 % warps active, every batch exposed a full HBM round trip).
// Per batch: pass 1 (row partial sums + dgamma/dbeta from smem) -> one __syncthreads -> refill of the stage released one
// batch earlier -> pass 2 (re-read the rows from smem, write dh_out / dy, accumulate the branch bias gradient).
constexpr int kLnStages = 3;

__global__ void __launch_bounds__(256, 2) add_ln_bwd_pipe_kernel(
    const float* __restrict__ dx, const float* __restrict__ r, const float* __restrict__ stats,
    const float* __restrict__ gamma, const float* dh_res, float* dh_out, float* __restrict__ dy,
    float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias_branch, int rows, int d,
    int rows_per_cta, float p_drop, const uint64_t* seed_dev, uint32_t stream_id, const int32_t* __restrict__ rows_dev) {
  pdl_trigger();   // lets a programmatically launched successor (the GEMMs) set itself up while this grid drains
  extern __shared__ __align__(128) uint8_t ln_smem[];
  __shared__ __align__(16) float s_part[2][8][2 * kR];   // [batch parity][warp][c1_0..c1_{R-1}, c2_0..c2_{R-1}]
  __shared__ __align__(8) uint64_t full_bar[kLnStages];
  const uint64_t seed = seed_dev ? *seed_dev : 0ull;  // device-resident: fresh masks under CUDA-graph replay
  const int d4 = d >> 2;
  const int c = threadIdx.x;                          // float4 column
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
        reinterpret_cast<float4*>(dh_out)[(size_t)row * d4 + c] = z;
        if (dy) reinterpret_cast<float4*>(dy)[(size_t)row * d4 + c] = z;
      }
    }
  }
  const int row0 = blockIdx.x * rows_per_cta, row1 = min(rows, row0 + rows_per_cta);
  const int nb = row1 > row0 ? (row1 - row0 + kR - 1) / kR : 0;
  const int narr = dh_res ? 3 : 2;
  const uint32_t arr_bytes = (uint32_t)kR * (uint32_t)d * 4u;            // one array's rows of a batch
  const uint32_t stage_bytes = narr * arr_bytes + 128u;                   // + the batch's (mean, rstd) pairs
  if (threadIdx.x == 0) {
    for (int s_ = 0; s_ < kLnStages; ++s_) mbar_init(&full_bar[s_], 1);
    fence_barrier_init();
  }
  __syncthreads();
  auto issue = [&](int j) {   // thread 0: stage batch j
    const int s_ = j % kLnStages;
    const int rb = row0 + j * kR;
    const int nr = min(kR, row1 - rb);
    uint8_t* st = ln_smem + (size_t)s_ * stage_bytes;
    const uint32_t bytes = (uint32_t)nr * (uint32_t)d * 4u;
    const bool stats_bulk = (nr == kR);   // a ragged tail batch reads its statistics with plain loads (16-byte granularity)
    mbar_arrive_expect_tx(&full_bar[s_], narr * bytes + (stats_bulk ? (uint32_t)(kR * 8) : 0u));
    bulk_load_1d(st, dx + (size_t)rb * d, bytes, &full_bar[s_]);
    bulk_load_1d(st + arr_bytes, r + (size_t)rb * d, bytes, &full_bar[s_]);
    if (dh_res) bulk_load_1d(st + 2 * arr_bytes, dh_res + (size_t)rb * d, bytes, &full_bar[s_]);
    if (stats_bulk) bulk_load_1d(st + narr * arr_bytes, stats + 2 * (size_t)rb, kR * 8, &full_bar[s_]);
  };
  if (threadIdx.x == 0) {
    for (int j = 0; j < kLnStages && j < nb; ++j) issue(j);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + c);
  const float inv_keep = 1.0f / (1.0f - p_drop);
  const float inv_d = 1.0f / (float)d;
  float4 dg = make_float4(0.f, 0.f, 0.f, 0.f), db = dg, dbr = dg;
  for (int i = 0; i < nb; ++i) {
    const int s_ = i % kLnStages, parity = i & 1;
    const int rb = row0 + i * kR;
    const int nr = min(kR, row1 - rb);
    const uint8_t* st = ln_smem + (size_t)s_ * stage_bytes;
    const float4* s_dx = reinterpret_cast<const float4*>(st);
    const float4* s_r = reinterpret_cast<const float4*>(st + arr_bytes);
    const float4* s_res = reinterpret_cast<const float4*>(st + 2 * arr_bytes);
    const float* s_stats = reinterpret_cast<const float*>(st + narr * arr_bytes);
    mbar_wait(&full_bar[s_], (uint32_t)((i / kLnStages) & 1), 7);
    float mean[kR], rstd[kR];
#pragma unroll
    for (int k = 0; k < kR; ++k) {
      if (nr == kR) { mean[k] = s_stats[2 * k]; rstd[k] = s_stats[2 * k + 1]; }
      else {
        const size_t rowc = (size_t)min(rb + k, row1 - 1);
        mean[k] = __ldg(stats + 2 * rowc); rstd[k] = __ldg(stats + 2 * rowc + 1);
      }
    }
    // ---- pass 1: row partial sums (c1 = sum g, c2 = sum g*xhat) and the column sums dgamma / dbeta ----
    float part[2 * kR];
#pragma unroll
    for (int k = 0; k < kR; ++k) {
      float4 dv = make_float4(0.f, 0.f, 0.f, 0.f), rv = dv;
      if (k < nr) { dv = s_dx[k * d4 + c]; rv = s_r[k * d4 + c]; }
      float4 xh, g;
      xh.x = (rv.x - mean[k]) * rstd[k]; xh.y = (rv.y - mean[k]) * rstd[k];
      xh.z = (rv.z - mean[k]) * rstd[k]; xh.w = (rv.w - mean[k]) * rstd[k];
      dg.x += dv.x * xh.x; dg.y += dv.y * xh.y; dg.z += dv.z * xh.z; dg.w += dv.w * xh.w;
      db.x += dv.x; db.y += dv.y; db.z += dv.z; db.w += dv.w;
      g.x = dv.x * gm.x; g.y = dv.y * gm.y; g.z = dv.z * gm.z; g.w = dv.w * gm.w;
      part[k] = (g.x + g.y) + (g.z + g.w);
      part[kR + k] = (g.x * xh.x + g.y * xh.y) + (g.z * xh.z + g.w * xh.w);
    }
#pragma unroll
    for (int j = 0; j < 2 * kR; ++j) part[j] = warp_sum(part[j]);
    if (lane == 0) {
      float4* sp = reinterpret_cast<float4*>(&s_part[parity][warp][0]);
      sp[0] = make_float4(part[0], part[1], part[2], part[3]);
      sp[1] = make_float4(part[4], part[5], part[6], part[7]);
    }
    __syncthreads();  // one barrier per batch (partials double-buffered by parity); everyone is past pass 2 of batch i-1
    if (threadIdx.x == 0 && i >= 1 && i - 1 + kLnStages < nb) issue(i - 1 + kLnStages);   // refill the stage of batch i-1
    {
      float4 t0 = make_float4(0.f, 0.f, 0.f, 0.f), t1 = t0;
      for (int w = 0; w < nwarps; ++w) {
        const float4* sp = reinterpret_cast<const float4*>(&s_part[parity][w][0]);
        const float4 a = sp[0], b = sp[1];
        t0.x += a.x; t0.y += a.y; t0.z += a.z; t0.w += a.w;
        t1.x += b.x; t1.y += b.y; t1.z += b.z; t1.w += b.w;
      }
      part[0] = t0.x * inv_d; part[1] = t0.y * inv_d; part[2] = t0.z * inv_d; part[3] = t0.w * inv_d;
      part[4] = t1.x * inv_d; part[5] = t1.y * inv_d; part[6] = t1.z * inv_d; part[7] = t1.w * inv_d;
    }
    // ---- pass 2: re-read the rows from shared memory and write the outputs ----
#pragma unroll
    for (int k = 0; k < kR; ++k) {
      if (k >= nr) break;
      const int row = rb + k;
      const float4 dv = s_dx[k * d4 + c], rv = s_r[k * d4 + c];
      const float c1 = part[k], c2 = part[kR + k];
      float4 o;
      {
        const float x0 = (rv.x - mean[k]) * rstd[k], x1 = (rv.y - mean[k]) * rstd[k];
        const float x2 = (rv.z - mean[k]) * rstd[k], x3 = (rv.w - mean[k]) * rstd[k];
        o.x = rstd[k] * (dv.x * gm.x - c1 - x0 * c2);
        o.y = rstd[k] * (dv.y * gm.y - c1 - x1 * c2);
        o.z = rstd[k] * (dv.z * gm.z - c1 - x2 * c2);
        o.w = rstd[k] * (dv.w * gm.w - c1 - x3 * c2);
      }
      if (dh_res) {
        const float4 rs = s_res[k * d4 + c];
        o.x += rs.x; o.y += rs.y; o.z += rs.z; o.w += rs.w;
      }
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
  // ~300 CTAs finish together and reduce into the same 3 x d floats: 128-bit vector reductions quarter the number of
  // same-address L2 atomic operations in that tail
  if (nb > 0) {
    if (dgamma) {
      red_add_v4(dgamma + 4 * c, dg);
      red_add_v4(dbeta + 4 * c, db);
    }
    if (dbias_branch) red_add_v4(dbias_branch + 4 * c, dbr);
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
  static const char* env_pipe = getenv("CAPDEC_LN_BWD_PIPE");   // bring-up switch: 0 = register-staged kernel
  const bool aligned = ((uintptr_t)dx % 16 == 0) && ((uintptr_t)r % 16 == 0) && ((uintptr_t)stats % 16 == 0) &&
                       (!dh_res || (uintptr_t)dh_res % 16 == 0) && ((uintptr_t)dgamma % 16 == 0) &&
                       ((uintptr_t)dbeta % 16 == 0) && ((uintptr_t)dbias_branch % 16 == 0);
  if (!(env_pipe && env_pipe[0] == '0') && aligned) {
    const size_t smem = (size_t)kLnStages * ((dh_res ? 3 : 2) * (size_t)kR * d * 4 + 128);
    static bool attr_set = false;
    if (!attr_set) {
      cudaError_t e = cudaFuncSetAttribute(add_ln_bwd_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
      if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(add_ln_bwd_pipe_kernel)");
      attr_set = true;
    }
    int ctas = num_sms() * 2;                    // two resident CTAs per SM, one contiguous row range each
    int rpc = (rows + ctas - 1) / ctas;
    rpc = ((rpc + kR - 1) / kR) * kR;
    if (rpc < kR) rpc = kR;
    ctas = (rows + rpc - 1) / rpc;
    add_ln_bwd_pipe_kernel<<<ctas, threads, smem, stream>>>(dx, r, stats, gamma, dh_res, dh_out, dy, dgamma, dbeta,
                                                            dbias_branch, rows, d, rpc, p_drop, seed_dev, stream_id, rows_dev);
    g_launches.fetch_add(1);
    CAPDEC_LAUNCH_CHECK("add_ln_bwd_pipe_kernel");
    return CAPDEC_OK;
  }
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
