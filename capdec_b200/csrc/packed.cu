// capdec_b200 — packed-row execution of the caption batch.
//
// ClipCocoDataset right-pads every caption to max_seq_len (train.py:55-63) and the reference runs GPT-2 over all
// B x (P+L) positions.  Rows at or after a caption's last non-zero token are dead: their targets are ignored
// (train.py:350 ignore_index=0), the logits slice drops the final position (train.py:349) and under the causal mask no
// live row attends to them (the padding mask has exactly zero effect on the loss).  The packed path therefore
// keeps, per caption b with last non-zero token index e_b (len_b = e_b + 1), only positions 0 .. P + len_b - 2 — the
// prefix and every token that is the INPUT of a non-ignored target — laid out back to back.  All row-wise kernels
// (GEMMs, LayerNorm, activations) then run on `rows[0]` rows read from device memory; attention and the embedding
// kernels use the plan below.  Results are identical to the dense path (tests/test_packed_gpu.py).
#include "../../include/capdec_b200.h"
#include "common.cuh"

namespace capdec {

// single CTA: cu[b] = first packed row of caption b (cu[B] = live rows), rows = {live, live rounded up to 32 (capped)},
// row_bt[r] = (b << 8) | t for every live row.
__global__ void __launch_bounds__(1024) pack_plan_kernel(const int64_t* __restrict__ tokens, int B, int L, int P,
                                                         int32_t* __restrict__ cu, int32_t* __restrict__ rows,
                                                         int32_t* __restrict__ row_bt) {
  __shared__ int s_warp[32];
  __shared__ int s_base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int start = 0; start < B; start += 1024) {
    const int b = start + threadIdx.x;
    int n = 0;
    if (b < B) {
      int len = 0;
      for (int j = L - 1; j >= 0; --j)
        if (tokens[(size_t)b * L + j] != 0) { len = j + 1; break; }
      n = P + (len > 0 ? len - 1 : 0);
    }
    // inclusive warp scan, then the per-warp totals
    int x = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    int woff = 0, total = 0;
    for (int w = 0; w < 32; ++w) { const int c = s_warp[w]; if (w < warp) woff += c; total += c; }
    const int base = s_base;
    if (b < B) cu[b] = base + woff + x - n;
    __syncthreads();
    if (threadIdx.x == 0) s_base = base + total;
    __syncthreads();
  }
  const int live = s_base;
  if (threadIdx.x == 0) {
    cu[B] = live;
    rows[0] = live;
    rows[1] = min(B * (P + L), (live + 31) & ~31);
  }
  __syncthreads();
  const int T = P + L;
  for (int i = threadIdx.x; i < B * T; i += 1024) {
    const int b = i / T, t = i % T;
    const int lo = cu[b], n = cu[b + 1] - lo;
    if (t < n) row_bt[lo + t] = (b << 8) | t;
  }
}

// one warp per live row: h[r] = dropout(prefix_proj[b,t] or wte[tokens[b,t-P]]  +  wpe[t])     (train.py:253-255)
__global__ void __launch_bounds__(256) embed_fwd_packed_kernel(const int64_t* __restrict__ tokens,
                                                               const float* __restrict__ prefix_proj,
                                                               const float* __restrict__ wte, const float* __restrict__ wpe,
                                                               float* __restrict__ h, const int32_t* __restrict__ row_bt,
                                                               const int32_t* __restrict__ rows, int P, int L, int d,
                                                               int vocab, float p_drop, const uint64_t* seed_dev,
                                                               uint32_t stream_id) {
  const uint64_t seed = seed_dev ? *seed_dev : 0ull;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= rows[0]) return;
  const int bt = row_bt[row], b = bt >> 8, t = bt & 255;
  const int d4 = d >> 2;
  const float4* src;
  if (t < P) {
    src = reinterpret_cast<const float4*>(prefix_proj) + ((size_t)b * P + t) * d4;
  } else {
    int64_t tok = tokens[(size_t)b * L + (t - P)];
    tok = tok < 0 ? 0 : (tok >= vocab ? vocab - 1 : tok);
    src = reinterpret_cast<const float4*>(wte) + (size_t)tok * d4;
  }
  const float4* pe = reinterpret_cast<const float4*>(wpe) + (size_t)t * d4;
  float4* dst = reinterpret_cast<float4*>(h) + (size_t)row * d4;
  const float inv_keep = 1.0f / (1.0f - p_drop);
  for (int c = lane; c < d4; c += 32) {
    float4 a = __ldg(src + c);
    const float4 e = __ldg(pe + c);
    a.x += e.x; a.y += e.y; a.z += e.z; a.w += e.w;
    if (p_drop > 0.0f) {
      float s[4];
      dropout_scale4(seed, stream_id, (uint64_t)row * d4 + c, p_drop, inv_keep, s);
      a.x *= s[0]; a.y *= s[1]; a.z *= s[2]; a.w *= s[3];
    }
    dst[c] = a;
  }
}

// grid (T, ceil(B/32)), block d/4 threads: same reduction structure as the dense embed_bwd_kernel; positions a caption
// does not have are skipped (their d_prefix_proj rows cannot occur: t < P is always live).
__global__ void embed_bwd_packed_kernel(const int64_t* __restrict__ tokens, const float* __restrict__ dh,
                                        float* __restrict__ d_prefix_proj, float* __restrict__ d_wte,
                                        float* __restrict__ d_wpe, const int32_t* __restrict__ cu, int B, int P, int L,
                                        int d, int vocab, float p_drop, const uint64_t* seed_dev, uint32_t stream_id) {
  const uint64_t seed = seed_dev ? *seed_dev : 0ull;
  const int t = blockIdx.x;
  const int b0 = blockIdx.y * 32, b1 = min(B, b0 + 32);
  const int d4 = d >> 2;
  const int c = threadIdx.x;
  if (c >= d4) return;
  const float inv_keep = 1.0f / (1.0f - p_drop);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int b = b0; b < b1; ++b) {
    const int lo = cu[b];
    if (t >= cu[b + 1] - lo) continue;
    const size_t row = (size_t)lo + t;
    float4 g = reinterpret_cast<const float4*>(dh)[row * d4 + c];
    if (p_drop > 0.0f) {
      float s[4];
      dropout_scale4(seed, stream_id, (uint64_t)row * d4 + c, p_drop, inv_keep, s);
      g.x *= s[0]; g.y *= s[1]; g.z *= s[2]; g.w *= s[3];
    }
    acc.x += g.x; acc.y += g.y; acc.z += g.z; acc.w += g.w;
    if (t < P) {
      if (d_prefix_proj) reinterpret_cast<float4*>(d_prefix_proj)[((size_t)b * P + t) * d4 + c] = g;
    } else if (d_wte) {
      int64_t tok = tokens[(size_t)b * L + (t - P)];
      tok = tok < 0 ? 0 : (tok >= vocab ? vocab - 1 : tok);
      red_add_v4(d_wte + (size_t)tok * d + 4 * c, g);
    }
  }
  if (d_wpe) red_add_v4(d_wpe + (size_t)t * d + 4 * c, acc);
}

// rows [rows[0], rows[1]) of a [*, ld] buffer <- 0 (the K-limited weight-gradient GEMMs read whole 32-row k-blocks)
__global__ void zero_tail_rows_kernel(float4* __restrict__ buf, int ld4, const int32_t* __restrict__ rows) {
  const int lo = rows[0], hi = rows[1];
  const int64_t total = (int64_t)(hi - lo) * ld4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    buf[(int64_t)lo * ld4 + i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

}  // namespace capdec

using namespace capdec;

extern "C" int capdec_pack_plan(const int64_t* tokens, int B, int L, int P, int32_t* cu, int32_t* rows, int32_t* row_bt,
                                capdec_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CAPDEC_REQUIRE(tokens && cu && rows && row_bt, "pack_plan: null argument");
  CAPDEC_REQUIRE(B > 0 && L > 0 && P > 0 && P + L <= 256 && B < (1 << 23), "pack_plan: need P >= 1, P + L <= 256");
  pack_plan_kernel<<<1, 1024, 0, stream>>>(tokens, B, L, P, cu, rows, row_bt);
  g_launches.fetch_add(1);
  CAPDEC_LAUNCH_CHECK("pack_plan_kernel");
  return CAPDEC_OK;
}

extern "C" int capdec_embed_fwd_packed(const int64_t* tokens, const float* prefix_proj, const float* wte, const float* wpe,
                                       float* h, const int32_t* row_bt, const int32_t* rows, int B, int P, int L, int d,
                                       int vocab, float p_drop, const uint64_t* seed_dev, uint32_t stream_id,
                                       capdec_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CAPDEC_REQUIRE(tokens && prefix_proj && wte && wpe && h && row_bt && rows, "embed_fwd_packed: null argument");
  CAPDEC_REQUIRE(B > 0 && P > 0 && L > 0 && d % 4 == 0, "embed_fwd_packed: bad shape");
  const int max_rows = B * (P + L);
  embed_fwd_packed_kernel<<<(max_rows + 7) / 8, 256, 0, stream>>>(tokens, prefix_proj, wte, wpe, h, row_bt, rows, P, L, d,
                                                                 vocab, p_drop, seed_dev, stream_id);
  g_launches.fetch_add(1);
  CAPDEC_LAUNCH_CHECK("embed_fwd_packed_kernel");
  return CAPDEC_OK;
}

extern "C" int capdec_embed_bwd_packed(const int64_t* tokens, const float* dh, float* d_prefix_proj, float* d_wte,
                                       float* d_wpe, const int32_t* cu, int B, int P, int L, int d, int vocab, float p_drop,
                                       const uint64_t* seed_dev, uint32_t stream_id, capdec_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CAPDEC_REQUIRE(dh && cu && B > 0 && P + L > 0 && d % 4 == 0 && d / 4 <= 1024, "embed_bwd_packed: bad arguments");
  CAPDEC_REQUIRE(!d_wte || tokens, "embed_bwd_packed: tokens required for the wte gradient");
  dim3 grid(P + L, (B + 31) / 32);
  const int threads = ((d / 4 + 31) / 32) * 32;
  embed_bwd_packed_kernel<<<grid, threads, 0, stream>>>(tokens, dh, d_prefix_proj, d_wte, d_wpe, cu, B, P, L, d, vocab, p_drop,
                                                        seed_dev, stream_id);
  g_launches.fetch_add(1);
  CAPDEC_LAUNCH_CHECK("embed_bwd_packed_kernel");
  return CAPDEC_OK;
}

extern "C" int capdec_zero_tail_rows(float* buf, int64_t ld, const int32_t* rows, capdec_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CAPDEC_REQUIRE(buf && rows && ld % 4 == 0 && ((uintptr_t)buf % 16) == 0, "zero_tail_rows: bad arguments");
  zero_tail_rows_kernel<<<32, 256, 0, stream>>>(reinterpret_cast<float4*>(buf), (int)(ld / 4), rows);
  g_launches.fetch_add(1);
  CAPDEC_LAUNCH_CHECK("zero_tail_rows_kernel");
  return CAPDEC_OK;
}
