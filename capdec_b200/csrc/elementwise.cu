// capdec_b200 — bandwidth-bound kernels of the CapDec hot path (warp-shuffle reductions, 128-bit HBM accesses):
//   noise injection (train.py:18-39), embedding assembly (train.py:253-255 + HF:modeling_gpt2.py:579-585,612) and its
//   backward, activation backward, bias-gradient column sums, row gather/scatter for the logits slice
//   (train.py:349), TransformerMapper input assembly (train.py:230-233), HF-semantics AdamW (train.py:326,352).
#include "../../include/capdec_b200.h"
#include "common.cuh"

namespace capdec {

// ------------------------------------------------------------------------------------------------------------
// noise injection
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void gauss4(uint64_t seed, uint32_t stream, uint64_t idx, float (&z)[4]) {
  uint32_t r[4];
  Philox::gen(seed, stream, idx, r);
  // Box-Muller on (0,1] uniforms
  const float u0 = ((float)r[0] + 1.0f) * 2.3283064365386963e-10f, u1 = (float)r[1] * 2.3283064365386963e-10f;
  const float u2 = ((float)r[2] + 1.0f) * 2.3283064365386963e-10f, u3 = (float)r[3] * 2.3283064365386963e-10f;
  const float ra = sqrtf(-2.0f * logf(u0)), rb = sqrtf(-2.0f * logf(u2));
  float s, c;
  sincospif(2.0f * u1, &s, &c);
  z[0] = ra * c; z[1] = ra * s;
  sincospif(2.0f * u3, &s, &c);
  z[2] = rb * c; z[3] = rb * s;
}

// one warp per row; D % 4 == 0, D <= 1024
__global__ void __launch_bounds__(256) noise_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int D,
                                                    float variance, const float* __restrict__ noise,
                                                    const float* __restrict__ offset, int uniform_ball, int dont_norm,
                                                    const uint64_t* seed_dev, uint64_t step) {
  const uint64_t seed = seed_dev ? *seed_dev : 0ull;  // device-resident: fresh masks under CUDA-graph replay
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= B) return;
  const int d4 = D >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(x) + (size_t)row * d4;
  float4* o4 = reinterpret_cast<float4*>(out) + (size_t)row * d4;
  float4 v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = lane + 32 * i;
    v[i] = (c < d4) ? x4[c] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (variance == 0.0f) {  // train.py:28-29: identity, no normalisation
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = lane + 32 * i;
      if (c < d4) o4[c] = v[i];
    }
    return;
  }
  const float stdv = sqrtf(variance);
  if (!dont_norm) {  // train.py:31-32
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) ss += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    const float inv = 1.0f / fmaxf(sqrtf(warp_sum(ss)), 1e-12f);
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i].x *= inv; v[i].y *= inv; v[i].z *= inv; v[i].w *= inv; }
  }
  if (noise) {  // parity mode: caller-provided noise tensor (already scaled)
    const float4* n4 = reinterpret_cast<const float4*>(noise) + (size_t)row * d4;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = lane + 32 * i;
      if (c < d4) { const float4 n = n4[c]; v[i].x += n.x; v[i].y += n.y; v[i].z += n.z; v[i].w += n.w; }
    }
  } else {
    float4 z[8];
    float zz = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = lane + 32 * i;
      float g[4] = {0.f, 0.f, 0.f, 0.f};
      if (c < d4) gauss4(seed, 0x4e01u, step * ((uint64_t)1 << 32) + (uint64_t)row * d4 + c, g);
      z[i] = make_float4(g[0], g[1], g[2], g[3]);
      zz += (g[0] * g[0] + g[1] * g[1]) + (g[2] * g[2] + g[3] * g[3]);
    }
    float mul = stdv;  // train.py:36: x + randn * std
    if (uniform_ball) {  // train.py:18-24: direction g/|g|, radius std * u^(1/D)
      uint32_t r[4];
      Philox::gen(seed, 0x4e02u, step * ((uint64_t)1 << 32) + row, r);
      const float u = ((float)r[0] + 1.0f) * 2.3283064365386963e-10f;
      mul = stdv * powf(u, 1.0f / (float)D) / fmaxf(sqrtf(warp_sum(zz)), 1e-12f);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i].x += z[i].x * mul; v[i].y += z[i].y * mul; v[i].z += z[i].z * mul; v[i].w += z[i].w * mul; }
  }
  if (offset) {  // train.py:37-38
    const float4* f4 = reinterpret_cast<const float4*>(offset);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = lane + 32 * i;
      if (c < d4) { const float4 f = __ldg(f4 + c); v[i].x += f.x; v[i].y += f.y; v[i].z += f.z; v[i].w += f.w; }
    }
  }
  float ss = 0.f;  // train.py:39
#pragma unroll
  for (int i = 0; i < 8; ++i) ss += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
  const float inv = 1.0f / fmaxf(sqrtf(warp_sum(ss)), 1e-12f);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = lane + 32 * i;
    if (c < d4) o4[c] = make_float4(v[i].x * inv, v[i].y * inv, v[i].z * inv, v[i].w * inv);
  }
}

// ------------------------------------------------------------------------------------------------------------
// embedding assembly: one warp per (b,t) row
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) embed_fwd_kernel(const int64_t* __restrict__ tokens,
                                                        const float* __restrict__ prefix_proj,
                                                        const float* __restrict__ wte, const float* __restrict__ wpe,
                                                        float* __restrict__ h, int B, int P, int L, int d, int vocab,
                                                        float p_drop, const uint64_t* seed_dev, uint32_t stream_id) {
  const uint64_t seed = seed_dev ? *seed_dev : 0ull;  // device-resident: fresh masks under CUDA-graph replay
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = P + L;
  const int row = blockIdx.x * 8 + warp;
  if (row >= B * T) return;
  const int b = row / T, t = row % T;
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

// grid (T, ceil(B/32)); block = d/4 threads (one float4 column each). wpe gradient reduced in registers over the
// block's 32 batch rows, then one atomicAdd per column; token rows scatter-added into the (tied) wte gradient.
__global__ void embed_bwd_kernel(const int64_t* __restrict__ tokens, const float* __restrict__ dh,
                                 float* __restrict__ d_prefix_proj, float* __restrict__ d_wte,
                                 float* __restrict__ d_wpe, int B, int P, int L, int d, int vocab, float p_drop,
                                 const uint64_t* seed_dev, uint32_t stream_id) {
  const uint64_t seed = seed_dev ? *seed_dev : 0ull;  // device-resident: fresh masks under CUDA-graph replay
  const int T = P + L;
  const int t = blockIdx.x;
  const int b0 = blockIdx.y * 32, b1 = min(B, b0 + 32);
  const int d4 = d >> 2;
  const int c = threadIdx.x;
  if (c >= d4) return;
  const float inv_keep = 1.0f / (1.0f - p_drop);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int b = b0; b < b1; ++b) {
    const size_t row = (size_t)b * T + t;
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
      red_add_v4(d_wte + (size_t)tok * d + 4 * c, g);   // one 128-bit reduction per thread instead of four scalar atomics
    }
  }
  if (d_wpe) red_add_v4(d_wpe + (size_t)t * d + 4 * c, acc);
}

// ------------------------------------------------------------------------------------------------------------
// column sums (bias gradients): block = 32 float4-columns x 8 row lanes
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, int64_t ld, float* __restrict__ out,
                                                     int M, int N, int rows_per_block) {
  __shared__ float4 s[8][32];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + cl) * 4;
  const int m0 = blockIdx.y * rows_per_block, m1 = min(M, m0 + rows_per_block);
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col < N) {  // N % 4 == 0 guaranteed by the launcher
    for (int m = m0 + rl; m < m1; m += 8) {
      const float4 v = ld_stream(reinterpret_cast<const float4*>(x + (size_t)m * ld + col));
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
  }
  s[rl][cl] = a;
  __syncthreads();
  if (rl == 0 && col < N) {
#pragma unroll
    for (int w = 1; w < 8; ++w) { const float4 t = s[w][cl]; a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w; }
    atomicAdd(out + col + 0, a.x); atomicAdd(out + col + 1, a.y); atomicAdd(out + col + 2, a.z); atomicAdd(out + col + 3, a.w);
  }
}

// ------------------------------------------------------------------------------------------------------------
// activation backward (+ fused bias gradient: dbias[n] += sum_m dx[m,n]); same thread layout as colsum_kernel
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float act_grad(float g, float p, int act) {
  if (act == 1) return g * gelu_new_grad(p);
  if (act == 2) return g * (1.f - p * p);
  return p > 0.f ? g : 0.f;
}

__global__ void __launch_bounds__(256) act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ pre,
                                                      float* __restrict__ dx, float* __restrict__ dbias, int M, int N,
                                                      int rows_per_block, int act) {
  __shared__ float4 s[8][32];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + cl) * 4;
  const int m0 = blockIdx.y * rows_per_block, m1 = min(M, m0 + rows_per_block);
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col < N) {
    for (int m = m0 + rl; m < m1; m += 8) {
      const size_t off = (size_t)m * N + col;
      const float4 g = ld_stream(reinterpret_cast<const float4*>(dy + off));
      const float4 p = ld_stream(reinterpret_cast<const float4*>(pre + off));
      float4 o;
      o.x = act_grad(g.x, p.x, act); o.y = act_grad(g.y, p.y, act);
      o.z = act_grad(g.z, p.z, act); o.w = act_grad(g.w, p.w, act);
      *reinterpret_cast<float4*>(dx + off) = o;
      a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w;
    }
  }
  if (!dbias) return;
  s[rl][cl] = a;
  __syncthreads();
  if (rl == 0 && col < N) {
#pragma unroll
    for (int w = 1; w < 8; ++w) { const float4 t = s[w][cl]; a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w; }
    atomicAdd(dbias + col + 0, a.x); atomicAdd(dbias + col + 1, a.y); atomicAdd(dbias + col + 2, a.z); atomicAdd(dbias + col + 3, a.w);
  }
}

// ------------------------------------------------------------------------------------------------------------
// row gather / scatter (logits slice [:, P-1:-1] expressed on hidden states)
// ------------------------------------------------------------------------------------------------------------
// gather: dst[B*L, d] <- src[B, T, d] rows (b, off + j)
// scatter: dst[B, T, d] <- src[B*L, d] at rows (b, off + j), every other row of dst is written as zero
__global__ void __launch_bounds__(256) rows_copy_kernel(const float4* __restrict__ src, float4* __restrict__ dst, int B,
                                                        int T, int L, int off, int d4, int scatter) {
  if (!scatter) {
    const int64_t total = (int64_t)B * L * d4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
      const int c = (int)(i % d4);
      const int64_t rj = i / d4;
      const int j = (int)(rj % L);
      const int64_t b = rj / L;
      dst[i] = src[((b * T) + off + j) * d4 + c];
    }
  } else {
    const int64_t total = (int64_t)B * T * d4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
      const int c = (int)(i % d4);
      const int64_t rt = i / d4;
      const int t = (int)(rt % T);
      const int64_t b = rt / T;
      const int j = t - off;
      dst[i] = (j >= 0 && j < L) ? src[((b * L) + j) * d4 + c] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

__global__ void __launch_bounds__(256) mapper_concat_fwd_kernel(const float4* __restrict__ lin,
                                                                const float4* __restrict__ pc, float4* __restrict__ x,
                                                                int B, int C, int P, int d4) {
  const int S = C + P;
  const int64_t total = (int64_t)B * S * d4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % d4);
    const int64_t rs = i / d4;
    const int s = (int)(rs % S);
    const int64_t b = rs / S;
    x[i] = (s < C) ? lin[(b * C + s) * d4 + c] : __ldg(pc + (size_t)(s - C) * d4 + c);
  }
}
// dlin[b,s<C] = dx[b,s]; dprefix_const[s] += sum_b dx[b,C+s]  (grid.x = P*d4/256 column chunks handled by loop over b)
__global__ void __launch_bounds__(256) mapper_concat_bwd_kernel(const float4* __restrict__ dx, float4* __restrict__ dlin,
                                                                float* __restrict__ dpc, int B, int C, int P, int d4) {
  const int S = C + P;
  const int64_t total_lin = (int64_t)B * C * d4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total_lin; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % d4);
    const int64_t rs = i / d4;
    const int s = (int)(rs % C);
    const int64_t b = rs / C;
    dlin[i] = dx[(b * S + s) * d4 + c];
  }
  if (dpc) {
    const int64_t total_pc = (int64_t)P * d4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total_pc; i += (int64_t)gridDim.x * blockDim.x) {
      const int c = (int)(i % d4);
      const int s = (int)(i / d4);
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int b = 0; b < B; ++b) {
        const float4 v = dx[((int64_t)b * S + C + s) * d4 + c];
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
      }
      float* dst = dpc + (size_t)i * 4;
      dst[0] += a.x; dst[1] += a.y; dst[2] += a.z; dst[3] += a.w;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// AdamW (HuggingFace 4.24 semantics): eps added to sqrt(v) BEFORE bias correction, decoupled decay after the update
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) adamw_kernel(float4* __restrict__ p, float4* __restrict__ g,
                                                    float4* __restrict__ m, float4* __restrict__ v, int64_t n4,
                                                    const float* __restrict__ lr_dev, const float* __restrict__ t_dev,
                                                    float b1, float b2, float eps, float wd,
                                                    const float* __restrict__ denom_dev, int zero_grad) {
  const float lr = *lr_dev, t = *t_dev;
  const float gs = denom_dev ? 1.0f / *denom_dev : 1.0f;
  const float bc1 = 1.0f - powf(b1, t), bc2 = 1.0f - powf(b2, t);
  const float step = lr * sqrtf(bc2) / bc1;
  const float decay = 1.0f - lr * wd;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 pp = p[i], gg = g[i], mm = m[i], vv = v[i];
    gg.x *= gs; gg.y *= gs; gg.z *= gs; gg.w *= gs;
    mm.x = b1 * mm.x + (1.f - b1) * gg.x; mm.y = b1 * mm.y + (1.f - b1) * gg.y;
    mm.z = b1 * mm.z + (1.f - b1) * gg.z; mm.w = b1 * mm.w + (1.f - b1) * gg.w;
    vv.x = b2 * vv.x + (1.f - b2) * gg.x * gg.x; vv.y = b2 * vv.y + (1.f - b2) * gg.y * gg.y;
    vv.z = b2 * vv.z + (1.f - b2) * gg.z * gg.z; vv.w = b2 * vv.w + (1.f - b2) * gg.w * gg.w;
    pp.x = (pp.x - step * mm.x / (sqrtf(vv.x) + eps)) * decay; pp.y = (pp.y - step * mm.y / (sqrtf(vv.y) + eps)) * decay;
    pp.z = (pp.z - step * mm.z / (sqrtf(vv.z) + eps)) * decay; pp.w = (pp.w - step * mm.w / (sqrtf(vv.w) + eps)) * decay;
    p[i] = pp; m[i] = mm; v[i] = vv;
    if (zero_grad) g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// device-side training clock: advances the RNG seed and evaluates get_linear_schedule_with_warmup
// (HF:optimization.py:101-104 as used at train.py:328-330,353) so that a whole step replays as one CUDA graph.
__global__ void step_clock_kernel(uint64_t* seed_dev, float* step_dev, float* lr_dev, float* t_dev, float base_lr,
                                  float warmup, float total) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    if (seed_dev) *seed_dev += 0x9E3779B97F4A7C15ull;
    if (step_dev) {
      const float n = *step_dev;  // number of optimizer steps already taken
      float f = (n < warmup) ? n / fmaxf(1.0f, warmup) : fmaxf(0.0f, (total - n) / fmaxf(1.0f, total - warmup));
      *lr_dev = base_lr * f;
      *t_dev = n + 1.0f;  // Adam's bias-correction step number for this update
      *step_dev = n + 1.0f;
    }
  }
}

static inline int grid_for(int64_t n_items, int threads, int per_sm = 8) {
  int64_t b = (n_items + threads - 1) / threads;
  const int64_t cap = (int64_t)num_sms() * per_sm;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace capdec

using namespace capdec;

extern "C" int capdec_noise_injection(const float* x, float* out, int B, int D, float variance, const float* noise,
                                      const float* offset, int uniform_ball, int dont_norm, const uint64_t* seed_dev,
                                      uint64_t step, capdec_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CAPDEC_REQUIRE(x && out && B > 0, "noise_injection: null argument");
  CAPDEC_REQUIRE(D % 4 == 0 && D > 0 && D <= 1024, "noise_injection: D=%d must be a multiple of 4 and <= 1024", D);
  CAPDEC_REQUIRE(variance >= 0.0f, "noise_injection: negative variance");
  noise_kernel<<<(B + 7) / 8, 256, 0, stream>>>(x, out, B, D, variance, noise, offset, uniform_ball, dont_norm, seed_dev, step);
  g_launches.fetch_add(1);
  CAPDEC_LAUNCH_CHECK("noise_kernel");
  return CAPDEC_OK;
}

extern "C" int capdec_embed_fwd(const int64_t* tokens, const float* prefix_proj, const float* wte, const float* wpe,
                                float* h, int B, int P, int L, int d, int vocab, float p_drop, const uint64_t* seed_dev,
                                uint32_t stream_id, capdec_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CAPDEC_REQUIRE(wpe && h && B > 0 && P >= 0 && L >= 0 && P + L > 0 && d % 4 == 0, "embed_fwd: bad arguments");
  CAPDEC_REQUIRE((L == 0 || (tokens && wte)) && (P == 0 || prefix_proj), "embed_fwd: missing tokens/wte/prefix_proj");
  const int rows = B * (P + L);
  embed_fwd_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(tokens, prefix_proj, wte, wpe, h, B, P, L, d, vocab, p_drop, seed_dev, stream_id);
  g_launches.fetch_add(1);
  CAPDEC_LAUNCH_CHECK("embed_fwd_kernel");
  return CAPDEC_OK;
}

extern "C" int capdec_embed_bwd(const int64_t* tokens, const float* dh, float* d_prefix_proj, float* d_wte,
                                float* d_wpe, int B, int P, int L, int d, int vocab, float p_drop, const uint64_t* seed_dev,
                                uint32_t stream_id, capdec_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CAPDEC_REQUIRE(dh && B > 0 && P + L > 0 && d % 4 == 0 && d / 4 <= 1024, "embed_bwd: bad arguments");
  CAPDEC_REQUIRE(!d_wte || tokens, "embed_bwd: tokens required for the wte gradient");
  dim3 grid(P + L, (B + 31) / 32);
  const int threads = ((d / 4 + 31) / 32) * 32;
  embed_bwd_kernel<<<grid, threads, 0, stream>>>(tokens, dh, d_prefix_proj, d_wte, d_wpe, B, P, L, d, vocab, p_drop, seed_dev, stream_id);
  g_launches.fetch_add(1);
  CAPDEC_LAUNCH_CHECK("embed_bwd_kernel");
  return CAPDEC_OK;
}

extern "C" int capdec_colsum_acc(const float* x, int64_t ld, float* out, int M, int N, capdec_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CAPDEC_REQUIRE(x && out && M > 0 && N > 0 && N % 4 == 0 && ld % 4 == 0, "colsum: N and ld must be multiples of 4");
  const int gx = (N + 127) / 128;
  int gy = (num_sms() * 4 + gx - 1) / gx;
  int rpb = (M + gy - 1) / gy;
  rpb = ((rpb + 7) / 8) * 8;
  if (rpb < 8) rpb = 8;
  gy = (M + rpb - 1) / rpb;
  colsum_kernel<<<dim3(gx, gy), 256, 0, stream>>>(x, ld, out, M, N, rpb);
  g_launches.fetch_add(1);
  CAPDEC_LAUNCH_CHECK("colsum_kernel");
  return CAPDEC_OK;
}

extern "C" int capdec_act_bwd(const float* dy, const float* pre, float* dx, float* dbias, int M, int N, int act,
                              capdec_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CAPDEC_REQUIRE(dy && pre && dx && M > 0 && N > 0 && N % 4 == 0 && act >= 1 && act <= 3, "act_bwd: bad arguments");
  const int gx = (N + 127) / 128;
  int gy = (num_sms() * 8 + gx - 1) / gx;
  int rpb = (M + gy - 1) / gy;
  rpb = ((rpb + 7) / 8) * 8;
  if (rpb < 8) rpb = 8;
  gy = (M + rpb - 1) / rpb;
  act_bwd_kernel<<<dim3(gx, gy), 256, 0, stream>>>(dy, pre, dx, dbias, M, N, rpb, act);
  g_launches.fetch_add(1);
  CAPDEC_LAUNCH_CHECK("act_bwd_kernel");
  return CAPDEC_OK;
}

static int rows_copy(const float* src, float* dst, int B, int T, int L, int off, int d, int scatter, cudaStream_t stream) {
  CAPDEC_REQUIRE(src && dst && B > 0 && L > 0 && off >= 0 && off + L <= T && d % 4 == 0, "rows_gather/scatter: bad arguments");
  rows_copy_kernel<<<grid_for((int64_t)B * (scatter ? T : L) * (d / 4), 256), 256, 0, stream>>>(
      reinterpret_cast<const float4*>(src), reinterpret_cast<float4*>(dst), B, T, L, off, d / 4, scatter);
  g_launches.fetch_add(1);
  CAPDEC_LAUNCH_CHECK("rows_copy_kernel");
  return CAPDEC_OK;
}
extern "C" int capdec_rows_gather(const float* src, float* dst, int B, int T, int L, int off, int d, capdec_stream_t s) {
  return rows_copy(src, dst, B, T, L, off, d, 0, reinterpret_cast<cudaStream_t>(s));
}
extern "C" int capdec_rows_scatter(const float* src, float* dst, int B, int T, int L, int off, int d, capdec_stream_t s) {
  return rows_copy(src, dst, B, T, L, off, d, 1, reinterpret_cast<cudaStream_t>(s));
}

extern "C" int capdec_mapper_concat_fwd(const float* lin, const float* prefix_const, float* x, int B, int C, int P,
                                        int d, capdec_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CAPDEC_REQUIRE(lin && prefix_const && x && B > 0 && C > 0 && P > 0 && d % 4 == 0, "mapper_concat_fwd: bad arguments");
  mapper_concat_fwd_kernel<<<grid_for((int64_t)B * (C + P) * (d / 4), 256), 256, 0, stream>>>(
      reinterpret_cast<const float4*>(lin), reinterpret_cast<const float4*>(prefix_const), reinterpret_cast<float4*>(x), B, C, P, d / 4);
  g_launches.fetch_add(1);
  CAPDEC_LAUNCH_CHECK("mapper_concat_fwd_kernel");
  return CAPDEC_OK;
}
extern "C" int capdec_mapper_concat_bwd(const float* dx, float* dlin, float* dprefix_const, int B, int C, int P, int d,
                                        capdec_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CAPDEC_REQUIRE(dx && dlin && B > 0 && C > 0 && P > 0 && d % 4 == 0, "mapper_concat_bwd: bad arguments");
  mapper_concat_bwd_kernel<<<grid_for((int64_t)B * C * (d / 4), 256), 256, 0, stream>>>(
      reinterpret_cast<const float4*>(dx), reinterpret_cast<float4*>(dlin), dprefix_const, B, C, P, d / 4);
  g_launches.fetch_add(1);
  CAPDEC_LAUNCH_CHECK("mapper_concat_bwd_kernel");
  return CAPDEC_OK;
}

extern "C" int capdec_adamw_step(float* p, float* g, float* m, float* v, int64_t n, const float* lr_dev,
                                 const float* t_dev, float beta1, float beta2, float eps, float weight_decay,
                                 const float* grad_denom_dev, int zero_grad, capdec_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CAPDEC_REQUIRE(p && g && m && v && lr_dev && t_dev && n > 0 && n % 4 == 0, "adamw: n must be a positive multiple of 4");
  adamw_kernel<<<grid_for(n / 4, 256, 16), 256, 0, stream>>>(reinterpret_cast<float4*>(p), reinterpret_cast<float4*>(g),
                                                            reinterpret_cast<float4*>(m), reinterpret_cast<float4*>(v), n / 4,
                                                            lr_dev, t_dev, beta1, beta2, eps, weight_decay, grad_denom_dev, zero_grad);
  g_launches.fetch_add(1);
  CAPDEC_LAUNCH_CHECK("adamw_kernel");
  return CAPDEC_OK;
}

extern "C" int capdec_step_clock(uint64_t* seed_dev, float* step_dev, float* lr_dev, float* t_dev, float base_lr,
                                 int warmup_steps, int total_steps, capdec_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CAPDEC_REQUIRE(seed_dev || step_dev, "step_clock: nothing to do");
  CAPDEC_REQUIRE(!step_dev || (lr_dev && t_dev), "step_clock: lr_dev and t_dev required with step_dev");
  step_clock_kernel<<<1, 32, 0, stream>>>(seed_dev, step_dev, lr_dev, t_dev, base_lr, (float)warmup_steps, (float)total_steps);
  g_launches.fetch_add(1);
  CAPDEC_LAUNCH_CHECK("step_clock_kernel");
  return CAPDEC_OK;
}
