// capdec_b200 — device-resident data feed: one launch assembles a training batch from the tokenised caption table.
// Replaces ClipCocoDataset.pad_tokens / __getitem__ (train.py:52-72: pad to max_seq_len with -1 -> mask = tokens >= 0 ->
// pad ids set to 0 -> mask prefixed with prefix_length ones; prefix = prefixes[caption2embedding[i]] (.float(), optionally
// / its L2 norm, no epsilon)), the DataLoader's default collate and the per-batch H2D copies of train.py:346.
// The caption table is stored pre-padded as int32 [N, L] with -1 in the padding, so tokens and mask are exact; the CLIP
// table keeps its pickled dtype (fp32 or fp16) and is widened in the kernel.  HBM-bound and tiny: B x (L x 4 + D x 2..4) bytes.
#include <cuda_fp16.h>

#include "../../include/capdec_b200.h"
#include "common.cuh"

namespace capdec {

template <typename TP>
__global__ void __launch_bounds__(128) batch_gather_kernel(const int32_t* __restrict__ tokens_all,
                                                           const int32_t* __restrict__ cap2emb, const TP* __restrict__ table,
                                                           const int64_t* __restrict__ idx, int64_t* __restrict__ tokens,
                                                           float* __restrict__ mask, float* __restrict__ prefix, int L,
                                                           int P, int D, int normalize) {
  const int b = blockIdx.x;
  const int64_t item = idx[b];
  const int32_t* src = tokens_all + item * L;
  for (int j = threadIdx.x; j < L; j += blockDim.x) {
    const int32_t t = src[j];
    tokens[(size_t)b * L + j] = t >= 0 ? (int64_t)t : 0;                 // train.py:60-61
    if (mask) mask[(size_t)b * (P + L) + P + j] = t >= 0 ? 1.0f : 0.0f;   // train.py:60,62
  }
  if (mask) {
    for (int j = threadIdx.x; j < P; j += blockDim.x) mask[(size_t)b * (P + L) + j] = 1.0f;  // train.py:63
  }
  const TP* row = table + (size_t)cap2emb[item] * D;
  float ss = 0.f;
  for (int j = threadIdx.x; j < D; j += blockDim.x) {
    const float x = (float)row[j];
    ss += x * x;
  }
  __shared__ float s_part[4];
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = ss;
  __syncthreads();
  float inv = 1.0f;
  if (normalize) inv = 1.0f / sqrtf((s_part[0] + s_part[1]) + (s_part[2] + s_part[3]));   // train.py:71 (no epsilon)
  for (int j = threadIdx.x; j < D; j += blockDim.x) {
    const float x = (float)row[j];
    prefix[(size_t)b * D + j] = normalize ? x * inv : x;
  }
}

}  // namespace capdec

using namespace capdec;

extern "C" int capdec_batch_gather(const int32_t* tokens_all, const int32_t* cap2emb, const void* table, int table_fp16,
                                   const int64_t* idx, int64_t* tokens, float* mask, float* prefix, int B, int L, int P, int D,
                                   int normalize, capdec_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CAPDEC_REQUIRE(tokens_all && cap2emb && table && idx && tokens && prefix, "batch_gather: null argument");
  CAPDEC_REQUIRE(B > 0 && L > 0 && P >= 0 && D > 0, "batch_gather: bad shape B=%d L=%d P=%d D=%d", B, L, P, D);
  if (table_fp16)
    batch_gather_kernel<__half><<<B, 128, 0, stream>>>(tokens_all, cap2emb, reinterpret_cast<const __half*>(table), idx, tokens,
                                                       mask, prefix, L, P, D, normalize);
  else
    batch_gather_kernel<float><<<B, 128, 0, stream>>>(tokens_all, cap2emb, reinterpret_cast<const float*>(table), idx, tokens,
                                                      mask, prefix, L, P, D, normalize);
  g_launches.fetch_add(1);
  CAPDEC_LAUNCH_CHECK("batch_gather_kernel");
  return CAPDEC_OK;
}

// Zero-fill of a caller-owned buffer on the caller's stream (a memset node under CUDA-graph capture): accumulate targets of
// the split-K GEMMs are cleared through the library instead of a framework fill kernel.
extern "C" int capdec_zero_fill(void* p, int64_t bytes, capdec_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CAPDEC_REQUIRE(p && bytes >= 0, "zero_fill: bad arguments");
  return check_cuda(cudaMemsetAsync(p, 0, (size_t)bytes, stream), "cudaMemsetAsync");
}
