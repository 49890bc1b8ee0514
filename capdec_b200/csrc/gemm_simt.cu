// capdec_b200 — fp32 CUDA-core GEMM with the same contract as capdec_gemm_tf32 (verification kernel).
// Exact fp32 FMA accumulation, no tensor cores: used by the parity tests to separate "TF32 rounding" from
// "kernel bug", and as the bit-faithful fp32 path for small shapes.  Never used by the perf path.
#include "../../include/capdec_b200.h"
#include "common.cuh"

namespace capdec {
constexpr int TS = 64;   // tile 64x64, 256 threads, 4x4 micro-tile per thread
constexpr int TK = 16;

__device__ __forceinline__ float simt_act(float x, int act) {
  switch (act) {
    case 1: return gelu_new_fwd(x);
    case 2: return tanhf(x);
    case 3: return fmaxf(x, 0.0f);
    default: return x;
  }
}

// A(m,k) = A[m*sam + k*sak], B(n,k) = B[n*sbn + k*sbk]
__global__ void __launch_bounds__(256) gemm_simt_kernel(const float* __restrict__ A, int64_t sam, int64_t sak,
                                                        const float* __restrict__ B, int64_t sbn, int64_t sbk,
                                                        float* __restrict__ C, int64_t ldc, int M, int N, int K,
                                                        const float* __restrict__ bias, int act,
                                                        float* __restrict__ aux, int accumulate) {
  __shared__ float sA[TK][TS + 1];
  __shared__ float sB[TK][TS + 1];
  const int m0 = blockIdx.y * TS, n0 = blockIdx.x * TS;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  for (int k0 = 0; k0 < K; k0 += TK) {
    for (int i = threadIdx.x; i < TS * TK; i += 256) {
      int r, kk;
      if (sak == 1) { kk = i % TK; r = i / TK; } else { r = i % TS; kk = i / TS; }
      const int m = m0 + r, k = k0 + kk;
      sA[kk][r] = (m < M && k < K) ? A[m * sam + k * sak] : 0.0f;
    }
    for (int i = threadIdx.x; i < TS * TK; i += 256) {
      int r, kk;
      if (sbk == 1) { kk = i % TK; r = i / TK; } else { r = i % TS; kk = i / TS; }
      const int n = n0 + r, k = k0 + kk;
      sB[kk][r] = (n < N && k < K) ? B[n * sbn + k * sbk] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sA[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sB[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.0f);
      if (aux) aux[m * ldc + n] = v;
      v = simt_act(v, act);
      if (accumulate) C[m * ldc + n] += v; else C[m * ldc + n] = v;
    }
  }
}

__device__ __forceinline__ float rn_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

__global__ void split_tf32_kernel(const float4* __restrict__ x, float4* __restrict__ hi, float4* __restrict__ lo,
                                  int64_t n4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    // hi = RN_tf32(x), lo = RN_tf32(x - hi): both are exact TF32 values, so the tensor core's operand truncation
    // is a no-op and the split error is unbiased (|x - hi - lo| <= 2^-22 |x|).
    float4 v = x[i], h, l;
    h.x = rn_tf32(v.x); l.x = rn_tf32(v.x - h.x);
    h.y = rn_tf32(v.y); l.y = rn_tf32(v.y - h.y);
    h.z = rn_tf32(v.z); l.z = rn_tf32(v.z - h.z);
    h.w = rn_tf32(v.w); l.w = rn_tf32(v.w - h.w);
    hi[i] = h;
    lo[i] = l;
  }
}
}  // namespace capdec

using namespace capdec;

extern "C" int capdec_gemm_fp32_simt(const float* A, int a_major, int64_t lda, const float* B, int b_major,
                                     int64_t ldb, float* C, int64_t ldc, int M, int N, int K, const float* bias,
                                     int act, float* aux, int accumulate, capdec_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CAPDEC_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0, "gemm_simt: bad arguments");
  const int64_t sam = a_major ? 1 : lda, sak = a_major ? lda : 1;
  const int64_t sbn = b_major ? 1 : ldb, sbk = b_major ? ldb : 1;
  dim3 grid((N + TS - 1) / TS, (M + TS - 1) / TS);
  gemm_simt_kernel<<<grid, 256, 0, stream>>>(A, sam, sak, B, sbn, sbk, C, ldc, M, N, K, bias, act, aux, accumulate);
  g_launches.fetch_add(1);
  CAPDEC_LAUNCH_CHECK("gemm_simt_kernel");
  return CAPDEC_OK;
}

extern "C" int capdec_split_tf32(const float* x, float* hi, float* lo, int64_t n, capdec_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CAPDEC_REQUIRE(x && hi && lo && n > 0 && (n % 4) == 0, "split_tf32: n must be a positive multiple of 4");
  const int64_t n4 = n / 4;
  int blocks = (int)((n4 + 255) / 256);
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  split_tf32_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(hi),
                                                reinterpret_cast<float4*>(lo), n4);
  g_launches.fetch_add(1);
  CAPDEC_LAUNCH_CHECK("split_tf32_kernel");
  return CAPDEC_OK;
}
