// capdec_b200 — KV-cached batched beam search (replaces gpt2_prefix_eval.py:50-115 generate_beam, which
// re-runs the whole GPT-2 forward over the growing sequence for every generated token).
//
// Layout (all fp32 unless noted), R = n_img * beam physical rows, Tmax = cache pitch in positions:
//   K/V cache per layer : [R][Tmax][d]           (position-major; a head's 64 floats of one position are contiguous)
//   src table           : int32 [2][R][Tmax]      lineage map: position t of logical beam b lives in physical row
//                                                 src[c&1][b][t]; re-pointed (never copied) when beams are re-ordered
//   state               : step c (int32, number of selections done), scores[R], seq_len[R], stopped[R] (int32),
//                         hist_tok / hist_parent int32 [max_sel][R], img_done int32 [n_img]
// Everything a decode step needs (position, table parity, tokens) is read from device memory keyed by `c`, so the
// whole step is one replayable CUDA graph.
#include "../../include/capdec_b200.h"
#include "common.cuh"

#include <math_constants.h>

namespace capdec {

constexpr int kMaxBeam = 8;

// ----------------------------------------------------------------------------------------------------------------
// init: state reset + prefix lineage (every beam of an image reads the prefill row img*beam for t < P)
// ----------------------------------------------------------------------------------------------------------------
__global__ void beam_init_kernel(int* __restrict__ step, float* __restrict__ scores, float* __restrict__ seq_len,
                                 int* __restrict__ stopped, int* __restrict__ src, int* __restrict__ img_done, int n_img,
                                 int beam, int P, int Tmax) {
  const int R = n_img * beam;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < R * Tmax; i += gridDim.x * blockDim.x) {
    const int b = i / Tmax, t = i % Tmax;
    src[i] = (t < P) ? (b / beam) * beam : b;
    src[R * Tmax + i] = b;
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < R; i += gridDim.x * blockDim.x) {
    scores[i] = 0.f;
    seq_len[i] = 1.f;  // gpt2_prefix_eval.py:59 seq_lengths = ones(beam_size)
    stopped[i] = 0;
    if (i < n_img) img_done[i] = 0;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *step = 0;
}

// prefill K/V (rows of the fused QKV projection [n_img*P, 3d]) -> cache rows img*beam, positions 0..P-1
__global__ void kv_prefill_kernel(const float4* __restrict__ qkv, float4* __restrict__ kc, float4* __restrict__ vc,
                                  int n_img, int beam, int P, int Tmax, int d4) {
  const int64_t total = (int64_t)n_img * P * d4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % d4);
    const int64_t row = i / d4;  // img*P + t
    const int img = (int)(row / P), t = (int)(row % P);
    const int64_t dst = ((int64_t)(img * beam) * Tmax + t) * d4 + c;
    kc[dst] = qkv[row * (3 * d4) + d4 + c];
    vc[dst] = qkv[row * (3 * d4) + 2 * d4 + c];
  }
}

// x[b,:] = wte[tok_b] + wpe[P + c - 1], tok_b = the token selected for beam b by the previous selection
__global__ void decode_embed_kernel(const int* __restrict__ step, const int* __restrict__ hist_tok,
                                    const float4* __restrict__ wte, const float4* __restrict__ wpe, float4* __restrict__ x,
                                    int R, int P, int d4) {
  const int c = *step;
  const int pos = P + c - 1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < R * d4; i += gridDim.x * blockDim.x) {
    const int b = i / d4, k = i % d4;
    const int tok = hist_tok[(c - 1) * R + b];
    const float4 e = wte[(int64_t)tok * d4 + k], p = wpe[(int64_t)pos * d4 + k];
    x[i] = make_float4(e.x + p.x, e.y + p.y, e.z + p.z, e.w + p.w);
  }
}

// ----------------------------------------------------------------------------------------------------------------
// single-query attention over the lineage-indexed cache.  One warp per (row, head), head_dim 64.
// ----------------------------------------------------------------------------------------------------------------
constexpr int kDecWarps = 4;
constexpr int kDecMaxT = 128;

__global__ void __launch_bounds__(kDecWarps * 32)
decode_attention_kernel(const float* __restrict__ qkv, float* __restrict__ kc, float* __restrict__ vc,
                        const int* __restrict__ src2, const int* __restrict__ step, float* __restrict__ ctx, int R, int H,
                        int P, int Tmax, float scale) {
  __shared__ float sq[kDecWarps][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x;
  const int h = blockIdx.y * kDecWarps + warp;
  if (h >= H) return;
  const int d = H * 64;
  const int c = *step;
  const int pos = P + c - 1;  // position of the token being processed == number of cached positions before it
  const int* src = src2 + (int64_t)(c & 1) * R * Tmax + (int64_t)b * Tmax;

  const float* qrow = qkv + (int64_t)b * 3 * d + h * 64;
  const float2 q2 = *reinterpret_cast<const float2*>(qrow + 2 * lane);
  const float2 k2 = *reinterpret_cast<const float2*>(qrow + d + 2 * lane);
  const float2 v2 = *reinterpret_cast<const float2*>(qrow + 2 * d + 2 * lane);
  sq[warp][2 * lane] = q2.x;
  sq[warp][2 * lane + 1] = q2.y;
  // append this token's K/V (physical row b, position pos)
  const int64_t own = ((int64_t)b * Tmax + pos) * d + h * 64 + 2 * lane;
  *reinterpret_cast<float2*>(kc + own) = k2;
  *reinterpret_cast<float2*>(vc + own) = v2;
  __syncwarp();

  // scores: lane owns positions lane, lane+32, ...; the current position comes from registers (warp dot)
  float s[kDecMaxT / 32];
  float mx = -CUDART_INF_F;
#pragma unroll
  for (int j = 0; j < kDecMaxT / 32; ++j) {
    const int t = j * 32 + lane;
    s[j] = -CUDART_INF_F;
    if (t < pos) {
      const float4* kr = reinterpret_cast<const float4*>(kc + ((int64_t)src[t] * Tmax + t) * d + h * 64);
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float4 kk = kr[i];
        const float4 qq = *reinterpret_cast<const float4*>(&sq[warp][4 * i]);
        acc = fmaf(kk.x, qq.x, acc);
        acc = fmaf(kk.y, qq.y, acc);
        acc = fmaf(kk.z, qq.z, acc);
        acc = fmaf(kk.w, qq.w, acc);
      }
      s[j] = acc * scale;
    }
    mx = fmaxf(mx, s[j]);
  }
  const float s_cur = warp_sum(q2.x * k2.x + q2.y * k2.y) * scale;
  mx = fmaxf(warp_max(mx), s_cur);
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < kDecMaxT / 32; ++j) {
    s[j] = (j * 32 + lane < pos) ? __expf(s[j] - mx) : 0.f;
    sum += s[j];
  }
  const float p_cur = __expf(s_cur - mx);
  sum = warp_sum(sum) + p_cur;
  const float inv = 1.f / sum;

  float2 acc = make_float2(p_cur * v2.x, p_cur * v2.y);
#pragma unroll
  for (int j = 0; j < kDecMaxT / 32; ++j) {
    if (j * 32 >= pos) break;
    const int tmax = min(32, pos - j * 32);
    for (int tt = 0; tt < tmax; ++tt) {
      const float p = __shfl_sync(0xffffffffu, s[j], tt);
      const int t = j * 32 + tt;
      const float2 vv = *reinterpret_cast<const float2*>(vc + ((int64_t)src[t] * Tmax + t) * d + h * 64 + 2 * lane);
      acc.x = fmaf(p, vv.x, acc.x);
      acc.y = fmaf(p, vv.y, acc.y);
    }
  }
  *reinterpret_cast<float2*>(ctx + (int64_t)b * d + h * 64 + 2 * lane) = make_float2(acc.x * inv, acc.y * inv);
}

// ----------------------------------------------------------------------------------------------------------------
// per-row log-sum-exp + top-k of logits/temperature.  (score+logp)/len is monotone in the logit inside one row, so the
// image-level top-k over beam*V candidates (gpt2_prefix_eval.py:95) is contained in the per-row top-k lists.
// ----------------------------------------------------------------------------------------------------------------
constexpr int kTopThreads = 512;

__device__ __forceinline__ bool cand_better(float v, int i, float w, int j) { return v > w || (v == w && i < j); }

__global__ void __launch_bounds__(kTopThreads)
row_topk_kernel(const float* __restrict__ logits, int64_t ld, int V, float inv_temp, int k, float* __restrict__ cand_val,
                int* __restrict__ cand_idx, float* __restrict__ row_lse) {
  const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* x = logits + (int64_t)row * ld;
  float tv[kMaxBeam];
  int ti[kMaxBeam];
#pragma unroll
  for (int j = 0; j < kMaxBeam; ++j) { tv[j] = -CUDART_INF_F; ti[j] = 0x7fffffff; }
  float mx = -CUDART_INF_F, sum = 0.f;
  for (int v = tid; v < V; v += kTopThreads) {
    const float val = x[v] * inv_temp;
    if (val > mx) { sum = sum * __expf(mx - val) + 1.f; mx = val; }
    else sum += __expf(val - mx);
    if (cand_better(val, v, tv[kMaxBeam - 1], ti[kMaxBeam - 1])) {
      tv[kMaxBeam - 1] = val; ti[kMaxBeam - 1] = v;
#pragma unroll
      for (int j = kMaxBeam - 1; j > 0; --j) {
        if (cand_better(tv[j], ti[j], tv[j - 1], ti[j - 1])) {
          const float fv = tv[j]; tv[j] = tv[j - 1]; tv[j - 1] = fv;
          const int fi = ti[j]; ti[j] = ti[j - 1]; ti[j - 1] = fi;
        }
      }
    }
  }
  __shared__ float s_mx[kTopThreads / 32], s_sum[kTopThreads / 32], s_v[kTopThreads / 32];
  __shared__ int s_i[kTopThreads / 32], s_owner[kTopThreads / 32];
  __shared__ int s_win;
  // log-sum-exp
  const float wm = warp_max(mx);
  const float ws = warp_sum(mx == -CUDART_INF_F ? 0.f : sum * __expf(mx - wm));
  if (lane == 0) { s_mx[warp] = wm; s_sum[warp] = ws; }
  __syncthreads();
  if (tid == 0) {
    float m = -CUDART_INF_F, s = 0.f;
    for (int w = 0; w < kTopThreads / 32; ++w) m = fmaxf(m, s_mx[w]);
    for (int w = 0; w < kTopThreads / 32; ++w) s += s_sum[w] * __expf(s_mx[w] - m);
    row_lse[row] = m + logf(s);
  }
  // k rounds of block arg-max over the heads of the per-thread sorted lists
  for (int r = 0; r < k; ++r) {
    float v = tv[0];
    int i = ti[0], owner = tid;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, v, o);
      const int oi = __shfl_xor_sync(0xffffffffu, i, o);
      const int oo = __shfl_xor_sync(0xffffffffu, owner, o);
      if (cand_better(ov, oi, v, i)) { v = ov; i = oi; owner = oo; }
    }
    if (lane == 0) { s_v[warp] = v; s_i[warp] = i; s_owner[warp] = owner; }
    __syncthreads();
    if (tid == 0) {
      int best = 0;
      for (int w = 1; w < kTopThreads / 32; ++w)
        if (cand_better(s_v[w], s_i[w], s_v[best], s_i[best])) best = w;
      cand_val[row * kMaxBeam + r] = s_v[best];
      cand_idx[row * kMaxBeam + r] = s_i[best];
      s_win = s_owner[best];
    }
    __syncthreads();
    if (tid == s_win) {
#pragma unroll
      for (int j = 0; j < kMaxBeam - 1; ++j) { tv[j] = tv[j + 1]; ti[j] = ti[j + 1]; }
      tv[kMaxBeam - 1] = -CUDART_INF_F; ti[kMaxBeam - 1] = 0x7fffffff;
    }
    __syncthreads();
  }
}

// ----------------------------------------------------------------------------------------------------------------
// image-level selection + state advance (gpt2_prefix_eval.py:80-108).  One warp per image.
// ----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
beam_select_kernel(const float* __restrict__ cand_val, const int* __restrict__ cand_idx, const float* __restrict__ row_lse,
                   int* __restrict__ step, float* __restrict__ scores, float* __restrict__ seq_len,
                   int* __restrict__ stopped, int* __restrict__ src2, int* __restrict__ hist_tok,
                   int* __restrict__ hist_parent, int* __restrict__ img_done, int* __restrict__ ticket, int n_img, int beam,
                   int P, int Tmax, int V, int stop_token) {
  const int img = blockIdx.x, lane = threadIdx.x;
  const int R = n_img * beam;
  const int c = *step;
  const bool first = (c == 0);
  const int rows_in = first ? 1 : beam;
  const int ncand = rows_in * beam;  // <= 64: two candidate slots per lane

  float av[2], nl[2];
  int atok[2], apar[2];
  long long aflat[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int ci = lane + 32 * u;
    av[u] = -CUDART_INF_F; nl[u] = 1.f; atok[u] = 0; apar[u] = 0; aflat[u] = 0x7fffffffffffffffLL;
    if (ci < ncand) {
      const int r = ci / beam, j = ci % beam;
      // candidate lists: first selection reads row `img` of an [n_img]-row list, later ones rows img*beam + r
      const int lrow = first ? img : img * beam + r;
      const int g = img * beam + r;
      if (!first && stopped[g]) {
        // :90-91 a finished beam offers exactly one continuation: token 0 with log-prob 0 (length unchanged)
        if (j == 0) {
          nl[u] = seq_len[g];
          av[u] = scores[g] / nl[u];
          atok[u] = 0; apar[u] = r; aflat[u] = (long long)r * V;
        }
      } else {
        const float logp = cand_val[lrow * kMaxBeam + j] - row_lse[lrow];
        const int tok = cand_idx[lrow * kMaxBeam + j];
        nl[u] = first ? 1.f : seq_len[g] + 1.f;           // :93
        av[u] = first ? logp : (scores[g] + logp) / nl[u];  // :92,94
        atok[u] = tok; apar[u] = r; aflat[u] = (long long)r * V + tok;
      }
    }
  }
  __syncwarp();  // all reads of the old state are done before any lane writes the new one

  int all_stop = 1;
  float new_len = 1.f, new_score = 0.f;
  int new_stop = 0;
  for (int o = 0; o < beam; ++o) {
    // warp arg-max over the remaining candidates (ties: lower flat index, i.e. the order topk sees them)
    int bu = (av[1] > av[0] || (av[1] == av[0] && aflat[1] < aflat[0])) ? 1 : 0;
    float v = av[bu];
    long long f = aflat[bu];
    int who = lane * 2 + bu;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, v, off);
      const long long of = __shfl_xor_sync(0xffffffffu, f, off);
      const int ow = __shfl_xor_sync(0xffffffffu, who, off);
      if (ov > v || (ov == v && of < f)) { v = ov; f = of; who = ow; }
    }
    const int wl = who >> 1, wu = who & 1;
    const float len = __shfl_sync(0xffffffffu, wu ? nl[1] : nl[0], wl);
    const int tok = __shfl_sync(0xffffffffu, wu ? atok[1] : atok[0], wl);
    const int par = __shfl_sync(0xffffffffu, wu ? apar[1] : apar[0], wl);
    if (lane == wl) { av[wu] = -CUDART_INF_F; aflat[wu] = 0x7fffffffffffffffLL; }
    const int gnew = img * beam + o, gpar = img * beam + par;
    int was = 0;
    if (lane == 0) was = first ? 0 : stopped[gpar];
    was = __shfl_sync(0xffffffffu, was, 0);
    const int st = was | (tok == stop_token);
    all_stop &= st;
    // lineage: inherit the parent's map for the cur_len cached positions, own row for the next one
    const int cur_len = P + c;
    const int* sin = src2 + (int64_t)(c & 1) * R * Tmax + (int64_t)gpar * Tmax;
    int* sout = src2 + (int64_t)((c + 1) & 1) * R * Tmax + (int64_t)gnew * Tmax;
    for (int t = lane; t < cur_len; t += 32) sout[t] = sin[t];
    if (lane == 0) {
      if (cur_len < Tmax) sout[cur_len] = gnew;
      hist_tok[c * R + gnew] = tok;
      hist_parent[c * R + gnew] = par;
    }
    // new state is staged in registers of lane o and written after the loop (parents are still being read)
    if (lane == o) { new_len = len; new_score = v * len; new_stop = st; }  // :100 scores = average * seq_lengths
  }
  __syncwarp();
  if (lane < beam) {
    const int g = img * beam + lane;
    seq_len[g] = new_len;
    scores[g] = new_score;
    stopped[g] = new_stop;
  }
  if (lane == 0) {
    img_done[img] = all_stop;
    __threadfence();
    const int done = atomicAdd(ticket, 1);
    if (done == n_img - 1) {  // every image has read `c`: advance the step
      *ticket = 0;
      *step = c + 1;
    }
  }
}

static inline int grid_for(int64_t n, int threads) {
  int64_t b = (n + threads - 1) / threads;
  const int64_t cap = (int64_t)num_sms() * 8;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace capdec

using namespace capdec;

extern "C" int capdec_beam_init(int32_t* step, float* scores, float* seq_len, int32_t* stopped, int32_t* src,
                                int32_t* img_done, int32_t* ticket, int n_img, int beam, int P, int Tmax,
                                capdec_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CAPDEC_REQUIRE(step && scores && seq_len && stopped && src && img_done && ticket, "beam_init: null argument");
  CAPDEC_REQUIRE(n_img > 0 && beam > 0 && beam <= kMaxBeam && P > 0 && P < Tmax && Tmax <= kDecMaxT,
                 "beam_init: need 0 < beam <= %d, 0 < P < Tmax <= %d", kMaxBeam, kDecMaxT);
  cudaError_t e = cudaMemsetAsync(ticket, 0, sizeof(int32_t), stream);
  if (e != cudaSuccess) return check_cuda(e, "beam_init memset");
  beam_init_kernel<<<grid_for((int64_t)n_img * beam * Tmax, 256), 256, 0, stream>>>(step, scores, seq_len, stopped, src,
                                                                                  img_done, n_img, beam, P, Tmax);
  g_launches.fetch_add(1);
  CAPDEC_LAUNCH_CHECK("beam_init_kernel");
  return CAPDEC_OK;
}

extern "C" int capdec_kv_prefill(const float* qkv, float* kcache, float* vcache, int n_img, int beam, int P, int Tmax,
                                 int d, capdec_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CAPDEC_REQUIRE(qkv && kcache && vcache && n_img > 0 && beam > 0 && P > 0 && P <= Tmax && d % 4 == 0,
                 "kv_prefill: bad arguments");
  kv_prefill_kernel<<<grid_for((int64_t)n_img * P * (d / 4), 256), 256, 0, stream>>>(
      reinterpret_cast<const float4*>(qkv), reinterpret_cast<float4*>(kcache), reinterpret_cast<float4*>(vcache), n_img, beam,
      P, Tmax, d / 4);
  g_launches.fetch_add(1);
  CAPDEC_LAUNCH_CHECK("kv_prefill_kernel");
  return CAPDEC_OK;
}

extern "C" int capdec_decode_embed(const int32_t* step, const int32_t* hist_tok, const float* wte, const float* wpe, float* x,
                                   int rows, int P, int d, capdec_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CAPDEC_REQUIRE(step && hist_tok && wte && wpe && x && rows > 0 && d % 4 == 0, "decode_embed: bad arguments");
  decode_embed_kernel<<<grid_for((int64_t)rows * (d / 4), 256), 256, 0, stream>>>(
      step, hist_tok, reinterpret_cast<const float4*>(wte), reinterpret_cast<const float4*>(wpe),
      reinterpret_cast<float4*>(x), rows, P, d / 4);
  g_launches.fetch_add(1);
  CAPDEC_LAUNCH_CHECK("decode_embed_kernel");
  return CAPDEC_OK;
}

extern "C" int capdec_decode_attention(const float* qkv, float* kcache, float* vcache, const int32_t* src,
                                       const int32_t* step, float* ctx, int rows, int H, int head_dim, int P, int Tmax,
                                       float scale, capdec_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CAPDEC_REQUIRE(qkv && kcache && vcache && src && step && ctx, "decode_attention: null argument");
  CAPDEC_REQUIRE(head_dim == 64 && rows > 0 && H > 0 && Tmax <= kDecMaxT && P > 0 && P < Tmax,
                 "decode_attention: head_dim must be 64 and P < Tmax <= %d", kDecMaxT);
  dim3 grid(rows, (H + kDecWarps - 1) / kDecWarps);
  decode_attention_kernel<<<grid, kDecWarps * 32, 0, stream>>>(qkv, kcache, vcache, src, step, ctx, rows, H, P, Tmax, scale);
  g_launches.fetch_add(1);
  CAPDEC_LAUNCH_CHECK("decode_attention_kernel");
  return CAPDEC_OK;
}

extern "C" int capdec_row_topk(const float* logits, int64_t ld, int rows, int V, float temperature, int k, float* cand_val,
                               int32_t* cand_idx, float* row_lse, capdec_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CAPDEC_REQUIRE(logits && cand_val && cand_idx && row_lse && rows > 0 && V > 0 && k > 0 && k <= kMaxBeam,
                 "row_topk: bad arguments (k <= %d)", kMaxBeam);
  const float inv_temp = 1.f / (temperature > 0.f ? temperature : 1.f);  // gpt2_prefix_eval.py:77
  row_topk_kernel<<<rows, kTopThreads, 0, stream>>>(logits, ld, V, inv_temp, k, cand_val, cand_idx, row_lse);
  g_launches.fetch_add(1);
  CAPDEC_LAUNCH_CHECK("row_topk_kernel");
  return CAPDEC_OK;
}

extern "C" int capdec_beam_select(const float* cand_val, const int32_t* cand_idx, const float* row_lse, int32_t* step,
                                  float* scores, float* seq_len, int32_t* stopped, int32_t* src, int32_t* hist_tok,
                                  int32_t* hist_parent, int32_t* img_done, int32_t* ticket, int n_img, int beam, int P,
                                  int Tmax, int V, int stop_token, capdec_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CAPDEC_REQUIRE(cand_val && cand_idx && row_lse && step && scores && seq_len && stopped && src && hist_tok && hist_parent &&
                     img_done && ticket,
                 "beam_select: null argument");
  CAPDEC_REQUIRE(n_img > 0 && beam > 0 && beam <= kMaxBeam && P > 0 && P < Tmax, "beam_select: bad shape");
  beam_select_kernel<<<n_img, 32, 0, stream>>>(cand_val, cand_idx, row_lse, step, scores, seq_len, stopped, src, hist_tok,
                                               hist_parent, img_done, ticket, n_img, beam, P, Tmax, V, stop_token);
  g_launches.fetch_add(1);
  CAPDEC_LAUNCH_CHECK("beam_select_kernel");
  return CAPDEC_OK;
}
