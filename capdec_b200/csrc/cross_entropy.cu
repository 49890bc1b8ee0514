// capdec_b200 — masked cross entropy over the caption tokens (train.py:349-350):
//   loss = mean_{targets != ignore} ( logsumexp(logits_row) - logits_row[target] ), plus dlogits in the same kernel.
// HBM-bound: one CTA per logits row (V = 50257 fp32 = 201 KB: pass 2 re-reads the row from L2), 128-bit loads,
// online max / sum-exp, warp-shuffle + smem block reduction.
#include "../../include/capdec_b200.h"
#include "common.cuh"

namespace capdec {

// single block: WRITES the count of non-ignored targets to *out (+= when accumulate) and zeroes *loss_sum
__global__ void __launch_bounds__(1024) ce_count_kernel(const int64_t* __restrict__ targets, int64_t n, int64_t ignore,
                                                        float* __restrict__ out, float* __restrict__ loss_sum) {
  __shared__ int s[32];
  int c = 0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) c += (targets[i] != ignore);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x < 32) {
    c = s[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (threadIdx.x == 0) {
      *out = (float)c;
      if (loss_sum) *loss_sum = 0.0f;
    }
  }
}

// Compaction of the consumed logits rows (train.py:349-350: positions whose target is the ignore index contribute
// nothing, so their logits never need to exist).  Single CTA, block-wide exclusive scan over B*L <= 32768 targets.
//   row_src[r]   = hidden-state row (b*T + off + j) of compact row r, r < n_valid
//   dst_of[b*T+t] = compact row fed by hidden row (b,t), or -1
//   targets_c[r] = the target of compact row r;  counts[0] = n_valid, counts[1] = n_valid rounded up to 32
__global__ void __launch_bounds__(1024) compact_targets_kernel(const int64_t* __restrict__ targets, int B, int L, int T,
                                                               int off, int64_t ignore, int32_t* __restrict__ row_src,
                                                               int32_t* __restrict__ dst_of,
                                                               int64_t* __restrict__ targets_c,
                                                               int32_t* __restrict__ counts, float* __restrict__ n_valid,
                                                               float* __restrict__ loss_sum,
                                                               const int32_t* __restrict__ cu) {
  __shared__ int s_warp[32];
  __shared__ int s_base;
  const int n = B * L;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_base = 0;
  for (int i = threadIdx.x; i < B * T; i += 1024) dst_of[i] = -1;
  __syncthreads();
  for (int start = 0; start < n; start += 1024) {
    const int i = start + threadIdx.x;
    const int64_t tg = (i < n) ? targets[i] : ignore;
    const int valid = (i < n) && (tg != ignore);
    const unsigned ballot = __ballot_sync(0xffffffffu, valid);
    const int wpre = __popc(ballot & ((1u << lane) - 1));
    if (lane == 0) s_warp[warp] = __popc(ballot);
    __syncthreads();
    int woff = 0, total = 0;
    for (int w = 0; w < 32; ++w) { const int c = s_warp[w]; if (w < warp) woff += c; total += c; }
    const int base = s_base;
    if (valid) {
      const int r = base + woff + wpre;
      const int b = i / L, j = i % L;
      const int hrow = (cu ? cu[b] : b * T) + off + j;   // packed rows: caption b starts at cu[b]
      row_src[r] = hrow;
      dst_of[hrow] = r;
      targets_c[r] = tg;
    }
    __syncthreads();
    if (threadIdx.x == 0) s_base = base + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int nv = s_base;
    counts[0] = nv;
    counts[1] = (nv + 31) & ~31;
    *n_valid = (float)nv;
    if (loss_sum) *loss_sum = 0.0f;
  }
}

// xsel[r] = xf[row_src[r]] for r < n_valid; zero rows up to the next multiple of 32 (so a K-limited wgrad sees zeros)
__global__ void __launch_bounds__(256) rows_gather_idx_kernel(const float4* __restrict__ src, float4* __restrict__ dst,
                                                              const int32_t* __restrict__ row_src,
                                                              const int32_t* __restrict__ counts, int d4, int max_rows) {
  pdl_trigger();   // lets a programmatically launched successor (the GEMMs) set itself up while this grid drains
  const int nv = counts[0], nvp = min(counts[1], max_rows);
  const int64_t total = (int64_t)nvp * d4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / d4), c = (int)(i % d4);
    dst[i] = (r < nv) ? src[(int64_t)row_src[r] * d4 + c] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
// dst[row] = (dst_of[row] >= 0) ? src[dst_of[row]] : 0 for every hidden row
__global__ void __launch_bounds__(256) rows_scatter_idx_kernel(const float4* __restrict__ src, float4* __restrict__ dst,
                                                               const int32_t* __restrict__ dst_of, int rows, int d4) {
  pdl_trigger();   // lets a programmatically launched successor (the GEMMs) set itself up while this grid drains
  const int64_t total = (int64_t)rows * d4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / d4), c = (int)(i % d4);
    const int s = dst_of[r];
    dst[i] = (s >= 0) ? src[(int64_t)s * d4 + c] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

constexpr int kCeThreads = 1024;  // 2 CTAs/SM: 296 rows x 201 KB = 60 MB in flight, so pass 2 re-reads its row from L2 (512 threads thrashed the 126 MB L2)

__device__ __forceinline__ void online_update(float& m, float& s, float x) {
  if (x > m) { s = s * __expf(m - x) + 1.0f; m = x; } else { s += __expf(x - m); }
}

__global__ void __launch_bounds__(kCeThreads) ce_kernel(float* __restrict__ logits, int64_t ld,
                                                        const int64_t* __restrict__ targets, int V, int64_t ignore,
                                                        const float* __restrict__ n_valid, float grad_scale,
                                                        float* __restrict__ loss_sum, int write_grad,
                                                        const int32_t* __restrict__ row_limit) {
  pdl_trigger();   // lets a programmatically launched successor (the GEMMs) set itself up while this grid drains
  __shared__ float s_m[kCeThreads / 32], s_s[kCeThreads / 32];
  __shared__ float s_bm, s_bs;
  const int row = blockIdx.x;
  if (row_limit && row >= *row_limit) return;  // compacted rows: nothing beyond the number of valid targets
  float* x = logits + (size_t)row * ld;
  const int64_t tgt = targets[row];
  const int v4 = V >> 2;
  float4* x4 = reinterpret_cast<float4*>(x);
  if (tgt == ignore || tgt < 0 || tgt >= V) {  // ignored row: zero gradient, no loss (ignore_index semantics)
    if (write_grad) {
      for (int i = threadIdx.x; i < v4; i += kCeThreads) x4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int i = (v4 << 2) + threadIdx.x; i < V; i += kCeThreads) x[i] = 0.f;
    }
    return;
  }
  float m = -INFINITY, s = 0.f;
  for (int i = threadIdx.x; i < v4; i += kCeThreads) {
    const float4 v = x4[i];
    online_update(m, s, v.x); online_update(m, s, v.y); online_update(m, s, v.z); online_update(m, s, v.w);
  }
  for (int i = (v4 << 2) + threadIdx.x; i < V; i += kCeThreads) online_update(m, s, x[i]);
  // combine (m, s) pairs across the warp, then across warps
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o), os = __shfl_xor_sync(0xffffffffu, s, o);
    const float nm = fmaxf(m, om);
    s = (nm == -INFINITY) ? 0.f : s * __expf(m - nm) + os * __expf(om - nm);
    m = nm;
  }
  if ((threadIdx.x & 31) == 0) { s_m[threadIdx.x >> 5] = m; s_s[threadIdx.x >> 5] = s; }
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < kCeThreads / 32 ? s_m[threadIdx.x] : -INFINITY;
    s = threadIdx.x < kCeThreads / 32 ? s_s[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, m, o), os = __shfl_xor_sync(0xffffffffu, s, o);
      const float nm = fmaxf(m, om);
      s = (nm == -INFINITY) ? 0.f : s * __expf(m - nm) + os * __expf(om - nm);
      m = nm;
    }
    if (threadIdx.x == 0) { s_bm = m; s_bs = s; }
  }
  __syncthreads();
  const float bm = s_bm, bs = s_bs;
  const float xt = x[tgt];
  __syncthreads();  // everyone has read x[tgt] before pass 2 overwrites it
  if (threadIdx.x == 0) atomicAdd(loss_sum, (bm + logf(bs)) - xt);
  if (!write_grad) return;
  const float sc = grad_scale / (n_valid ? *n_valid : 1.0f);
  const float inv = sc / bs;
  for (int i = threadIdx.x; i < v4; i += kCeThreads) {
    float4 v = x4[i];
    v.x = __expf(v.x - bm) * inv; v.y = __expf(v.y - bm) * inv; v.z = __expf(v.z - bm) * inv; v.w = __expf(v.w - bm) * inv;
    const int base = i << 2;
    if ((int)tgt >= base && (int)tgt < base + 4) {
      float* pv = reinterpret_cast<float*>(&v);
      pv[(int)tgt - base] -= sc;
    }
    x4[i] = v;
  }
  for (int i = (v4 << 2) + threadIdx.x; i < V; i += kCeThreads) x[i] = __expf(x[i] - bm) * inv - (i == (int)tgt ? sc : 0.f);
}

}  // namespace capdec

using namespace capdec;

extern "C" int capdec_ce_count(const int64_t* targets, int64_t n, int64_t ignore_index, float* n_valid,
                               float* loss_sum_to_zero, capdec_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CAPDEC_REQUIRE(targets && n_valid && n > 0, "ce_count: bad arguments");
  ce_count_kernel<<<1, 1024, 0, stream>>>(targets, n, ignore_index, n_valid, loss_sum_to_zero);
  g_launches.fetch_add(1);
  CAPDEC_LAUNCH_CHECK("ce_count_kernel");
  return CAPDEC_OK;
}

extern "C" int capdec_ce_fwd_bwd(float* logits, int64_t ld, const int64_t* targets, int rows, int V,
                                 int64_t ignore_index, const float* n_valid, float grad_scale, float* loss_sum,
                                 int write_grad, const int32_t* row_limit_dev, capdec_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CAPDEC_REQUIRE(logits && targets && loss_sum && rows > 0 && V > 0, "ce: bad arguments");
  CAPDEC_REQUIRE(ld >= V && ld % 4 == 0 && ((uintptr_t)logits % 16) == 0, "ce: logits pitch must be a multiple of 4 floats and 16-byte aligned");
  ce_kernel<<<rows, kCeThreads, 0, stream>>>(logits, ld, targets, V, ignore_index, n_valid, grad_scale, loss_sum, write_grad, row_limit_dev);
  g_launches.fetch_add(1);
  CAPDEC_LAUNCH_CHECK("ce_kernel");
  return CAPDEC_OK;
}

extern "C" int capdec_compact_targets(const int64_t* targets, int B, int L, int T, int off, int64_t ignore_index,
                                      int32_t* row_src, int32_t* dst_of, int64_t* targets_c, int32_t* counts,
                                      float* n_valid, float* loss_sum_to_zero, const int32_t* cu_rows,
                                      capdec_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CAPDEC_REQUIRE(targets && row_src && dst_of && targets_c && counts && n_valid, "compact_targets: null argument");
  CAPDEC_REQUIRE(B > 0 && L > 0 && off >= 0 && off + L <= T, "compact_targets: bad shape");
  compact_targets_kernel<<<1, 1024, 0, stream>>>(targets, B, L, T, off, ignore_index, row_src, dst_of, targets_c, counts,
                                                 n_valid, loss_sum_to_zero, cu_rows);
  g_launches.fetch_add(1);
  CAPDEC_LAUNCH_CHECK("compact_targets_kernel");
  return CAPDEC_OK;
}

extern "C" int capdec_rows_gather_idx(const float* src, float* dst, const int32_t* row_src, const int32_t* counts,
                                      int max_rows, int d, capdec_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CAPDEC_REQUIRE(src && dst && row_src && counts && max_rows > 0 && d % 4 == 0, "rows_gather_idx: bad arguments");
  int64_t blocks = ((int64_t)max_rows * (d / 4) + 255) / 256;
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  rows_gather_idx_kernel<<<(int)blocks, 256, 0, stream>>>(reinterpret_cast<const float4*>(src), reinterpret_cast<float4*>(dst),
                                                         row_src, counts, d / 4, max_rows);
  g_launches.fetch_add(1);
  CAPDEC_LAUNCH_CHECK("rows_gather_idx_kernel");
  return CAPDEC_OK;
}

extern "C" int capdec_rows_scatter_idx(const float* src, float* dst, const int32_t* dst_of, int rows, int d,
                                       capdec_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CAPDEC_REQUIRE(src && dst && dst_of && rows > 0 && d % 4 == 0, "rows_scatter_idx: bad arguments");
  int64_t blocks = ((int64_t)rows * (d / 4) + 255) / 256;
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  rows_scatter_idx_kernel<<<(int)blocks, 256, 0, stream>>>(reinterpret_cast<const float4*>(src), reinterpret_cast<float4*>(dst),
                                                          dst_of, rows, d / 4);
  g_launches.fetch_add(1);
  CAPDEC_LAUNCH_CHECK("rows_scatter_idx_kernel");
  return CAPDEC_OK;
}
