// capdec_b200 — persistent warp-specialised TF32 GEMM on tcgen05 (sm_100a).
//
//   C[M,N] = act( sum_k A(m,k) * B(n,k) + bias[n] )
//
// Replaces nn.Linear (train.py:113-118,124-126,144-147,241), HF Conv1D (HF:pytorch_utils.py:97-123),
// the tied lm_head (HF:modeling_gpt2.py:703-706) and every dgrad / wgrad product autograd derives from them
// (train.py:351).  One kernel serves all operand orientations:
//   K-major  operand: TMA box {32 fp32 (=128 B, one swizzle row), rows} with SWIZZLE_128B,
//                     UMMA descriptor SWIZZLE_128B, SBO = 1024 B, k-step = +32 B inside the swizzle atom;
//   MN-major operand: TMA boxes {32 fp32 along M/N, 32 rows of K} with SWIZZLE_128B_ATOM_32B (the only layout
//                     tcgen05 accepts for 32-bit MN-major data), UMMA descriptor SWIZZLE_128B_BASE32B,
//                     LBO = 4096 B (next 32-wide M/N slab), SBO = 512 B (next 4 k-rows), k-step = +1024 B.
// Pipeline: warp 0 = TMA producer, warp 1 = single-thread tcgen05.mma issuer, warp 2 = TMEM allocator,
// warps 4-7 = epilogue (tcgen05.ld -> bias/activation -> swizzled smem -> TMA store / TMA reduce-add).
// Two TMEM accumulator stages (2 x 256 columns) let the epilogue of tile i overlap the mainloop of tile i+1.
// 3xTF32 (fp32-grade) mode, template kSplit: the fp32 operand tiles are split INSIDE the pipeline.  Four converter warps
// turn every landed stage into hi = RN_tf32(x) (in place) and lo = x - hi (second tile, same swizzled layout - the split
// is element-wise, so it is layout-agnostic); the issuer then runs three MMAs per k-step (lo*hi, hi*lo, hi*hi) into the
// same TMEM accumulator.  Operands are read from HBM/L2 once, exactly as in the 1xTF32 mode; no host-side hi/lo buffers.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <map>
#include <mutex>
#include <vector>

#include "../../include/capdec_b200.h"
#include "common.cuh"

namespace capdec {

// ---------------------------------------------------------------------------------------------------------------
// library-wide helpers (defined once here)
// ---------------------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return CAPDEC_OK;
  set_last_error("%s: %s", what, cudaGetErrorString(e));
  return CAPDEC_ERR_CUDA;
}
constexpr int kMaxDevices = 64;
static int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }
  return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}
int num_sms() {   // cached per device: a process may drive several GPUs
  static std::atomic<int> n[kMaxDevices];
  const int dev = current_device();
  int v = n[dev].load(std::memory_order_relaxed);
  if (v == 0) {
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) cudaGetLastError();
    if (v <= 0) v = 148;
    n[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}

// ---------------------------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------------------------
constexpr int kBlockM = 128;
constexpr int kBlockK = 32;  // fp32 elements per k-block = 128 bytes = one swizzle row
constexpr int kUmmaK = 8;    // tf32
constexpr int kThreads = 256;
constexpr int kSplitThreads = 384;  // 3xTF32: + warps 8..11 = operand converters (hi/lo split in shared memory)
constexpr int kConvThreads = 128;
constexpr int kEpiThreads = 128;
constexpr int kABytes = kBlockM * kBlockK * 4;  // 16 KB
constexpr int kStagingBytes = 128 * 128;        // 128 rows x 32 fp32
constexpr int kWarpStagingBytes = 32 * 128;     // one epilogue warp's box: 32 rows x 32 fp32 (4 per 16 KB slot)
constexpr int kMaxStages = 8;
constexpr int kMulDepth = 4;  // epilogue-input boxes in flight per epilogue warp (C = acc * act'(input))
constexpr int kSchedDepth = 4;  // tile ids the cluster's scheduler thread may run ahead of the slowest role
constexpr int kSchedSlots = 1024;  // launch slots of the device-side tile counters (one per GEMM launch in flight)
constexpr int kTmemCols = 512;
constexpr int kAccCols = 256;  // columns per accumulator stage
constexpr int kSmemLimit = 226 * 1024;  // of 228 KB per SM: 1 KB is reserved per CTA, and > 1 KB stays free for a co-resident CTA

// device-side tile counters of the dynamic scheduler: one 128-byte line per launch slot, {next tile, clusters done, pad...}
// zero at module load; every launch leaves its slot zeroed again (see the scheduler thread)
__device__ int g_sched_slots[32 * kSchedSlots];

struct __align__(64) GemmDev {
  CUtensorMap tmA;
  CUtensorMap tmB;
  CUtensorMap tmBh;    // B with half the box rows / slabs: the half-width tiles of the last, partly filled wave
  CUtensorMap tmC;
  CUtensorMap tmAux;
  CUtensorMap tmMul;   // optional epilogue INPUT (same shape as C): C = acc * act'(mul)
  const float* bias;
  int M, N, K;
  int block_n, stages;
  int m_tiles, n_tiles, splits, kb_per_split, kb_total;
  int exact;                        // 3xTF32 mode: exact tanhf in the activation epilogues (1xTF32: MUFU.TANH)
  int act, has_aux, accumulate;
  int a_mn, b_mn;
  int a_3d, b_3d;                   // MN-major operand fetched as ONE 3-D TMA box {32, 32 k-rows, slabs} per stage
  uint32_t idesc;
  uint32_t idesc_h;                 // instruction descriptor of a half-width tile (N = block_n / 2)
  int tail_ok;                      // half-width tail tiles allowed for this launch (engine, tile width, no split-K)
  uint32_t adesc_hi, bdesc_hi;      // upper 32 bits of the smem descriptors (SBO, version, layout)
  uint32_t adesc_lo16, bdesc_lo16;  // LBO field (bits 16..29 of the low word), pre-shifted
  uint32_t a_kstep, b_kstep;        // start-address increment per UMMA_K step, in 16-byte units
  float* colsum;                    // optional [N]: += column sums of the stored C (bias gradient), mul epilogue only
  int mul_act;                      // 0 none; 1 gelu_new'(pre-activation), 2 tanh' = 1 - a^2 (activated a), 3 relu' (activated a)
  const int* m_limit;               // optional device scalar: rows >= *m_limit are not computed (whole tiles skipped)
  const int* k_limit;               // optional device scalar: reduction stops at *k_limit (rounded up to a k-block)
  int* sched;                       // dynamic tile scheduler: {next tile, clusters done} of this launch (NULL = static round-robin)
  float* c_ptr;                     // C / aux and their pitch for the LSU store path of the epilogue (lsu_store != 0)
  float* aux_ptr;
  long long ldc;
  int lsu_store;                    // epilogue stores with coalesced st.global from the staging box instead of TMA stores
  long long* trace;                 // bring-up: cycle stamps of CTA 0's roles (capdec_gemm_debug_trace), NULL = off
  int c_depth;                      // epilogue staging boxes per warp and output (2..4): TMA stores in flight before the warp must wait
  int raster;                       // tile order inside a wave: 0 = n fastest (neighbouring clusters share A rows), 1 = m fastest (share B columns)
  uint32_t dbg;                     // bring-up switches (CAPDEC_GEMM_DBG): 1 skip TMA loads, 2 skip MMA, 4 skip stores, 8 skip epilogue, 16 no proxy fence, 32 no staging writes, 64 no stage handshake, 128 no TMEM reads
};

// cycle stamp of CTA 0 into trace[role * 64 + idx]: role 0 = CTA, 1 = producer, 2 = MMA issuer, 3 = epilogue warp 4
#define CAPDEC_TRACE(role, idx)                                                                                      \
  do {                                                                                                               \
    if (p.trace && blockIdx.x == 0 && (idx) < 64) p.trace[(role) * 64 + (idx)] = clock64();                          \
  } while (0)

__device__ __forceinline__ float tanh_fast(float x) {  // MUFU.TANH, max rel. error 2^-11: same grade as a TF32 operand
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_grad_fast(float x) {  // d/dx gelu_new with MUFU.TANH (1xTF32 mode only)
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  const float x2 = x * x;
  const float t = tanh_fast(k0 * (x + k1 * x * x2));
  return 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * k0 * (1.0f + 3.0f * k1 * x2);
}
// Activation over a 32-column register chunk.  The `act` / `exact` dispatch is hoisted OUT of the element loop: with
// the switch inside the unrolled loop the compiler emitted a branch tree per element and the c_fc epilogue took 3x its
// mainloop (313 vs 110 us, profiles/r1_gemm_pipeline_experiments.md).  `exact` (3xTF32 parity mode) uses tanhf; the
// 1xTF32 mode uses MUFU.TANH (2^-11 relative error, the grade of a TF32 operand).
__device__ __forceinline__ void apply_act32(float (&v)[32], int act, bool exact) {
  if (act == 1) {
    if (exact) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = gelu_new_fwd(v[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float x = v[j];
        const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
        v[j] = 0.5f * x * (1.0f + tanh_fast(u));
      }
    }
  } else if (act == 2) {
    if (exact) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = tanhf(v[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = tanh_fast(v[j]);
    }
  } else if (act == 3) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
  }
}

// act == 4: gelu_new forward AND its derivative from one MUFU.TANH (v <- gelu_new(v), dv <- gelu_new'(v)).  The forward
// c_fc epilogue has slack under its mainloop; storing the derivative instead of the pre-activation turns the fused
// GELU-backward of the dgrad epilogue (which is NOT hidden: N = 3072, K = 768) into a plain multiply (mul_act == 4).
template <bool kExact>
__device__ __forceinline__ void gelu_fwd_and_grad32_impl(float (&v)[32], float (&dv)[32]) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float x = v[j];
    const float x2 = x * x;
    const float u = k0 * (x + k1 * x * x2);
    const float t = kExact ? tanhf(u) : tanh_fast(u);
    const float h = 0.5f * (1.0f + t);
    v[j] = x * h;
    dv[j] = h + 0.5f * x * (1.0f - t * t) * k0 * (1.0f + 3.0f * k1 * x2);
  }
}
// the exact / fast choice is made ONCE per chunk, outside the element loop (a per-element select between tanhf and
// MUFU.TANH kept both code paths in the unrolled body and cost the forward c_fc GEMM 27 us: 89 -> 116 us)
__device__ __forceinline__ void gelu_fwd_and_grad32(float (&v)[32], float (&dv)[32], bool exact) {
  if (exact) gelu_fwd_and_grad32_impl<true>(v, dv);
  else gelu_fwd_and_grad32_impl<false>(v, dv);
}

// v *= act'(u) over a 32-column register chunk, dispatch hoisted out of the element loop for the same reason: the
// 32 independent dependency chains (MUFU.TANH + ~15 FMAs each) must interleave, the epilogue has one warp per scheduler.
__device__ __forceinline__ void apply_mul32(float (&v)[32], const float (&u)[32], int mul_act, bool exact) {
  if (mul_act == 1) {
    if (exact) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= gelu_new_grad(u[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= gelu_grad_fast(u[j]);
    }
  } else if (mul_act == 2) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= 1.f - u[j] * u[j];
  } else if (mul_act == 4) {   // u already holds the derivative (forward ran with act == 4)
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= u[j];
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = u[j] > 0.f ? v[j] : 0.f;
  }
}

// Tile engines (template kMode):
//   0  one CTA per 128 x block_n tile (tcgen05 cta_group::1).
//   1  a CTA PAIR (cluster of 2, cta_group::2) per 256 x block_n tile.  CTA r of the pair TMA-loads its own 128 A rows
//      and its own block_n/2 B columns into ITS shared memory; the leader issues one M=256 MMA per k-step that reads
//      both CTAs' smem and writes both CTAs' TMEM.  Operand bytes per CTA drop from (128 + block_n) to
//      (128 + block_n/2) rows per k-block.
//   2  a QUAD (cluster of 4 = two pairs side by side in N) per 256 x (2*block_n) tile: the two pairs need the SAME A
//      rows, so each CTA loads only HALF of its A slab and TMA-multicasts it to its twin in the other pair.
//   3  a QUAD stacked in M (512 x block_n): the pairs need the same B columns; B halves are multicast instead.
// Why: with fp32 operands these GEMMs are bound by L2->SM bandwidth (~8.3 TB/s measured: loads-only experiment in
// profiles/r1_gemm_pipeline_experiments.md), not by the tensor pipe; FLOP per L2 byte is 43 / 64 / 87 for modes 0/1/2-3.
// Register / shared-memory budget: every variant is compiled for <= 168 registers per thread (the bound of the 384-thread
// 3xTF32 variant; no spills) and leaves > 1 KB of the SM's shared memory free, so that a GEMM CTA does not lock the SM:
// a small streaming kernel (the fused AdamW on its side stream) can be co-resident and use the HBM bandwidth the
// tensor-bound GEMM leaves idle.  At 244 registers x 256 threads + 227 KB nothing else could ever share the SM.
// kEpi: the epilogue variant the instantiation is compiled for - 0 = any (run-time switches), 1 = plain: bias (+ accumulate)
// only, 2 = gelu_new forward with its derivative as second output (c_fc forward), 3 = multiply by a stored derivative (+ the
// bias-gradient column sums; the dgrad that feeds c_fc's backward).  1-3 cover every GEMM launch of the GPT-2 trunk's train
// step.  They get their own instantiations because the general
// epilogue's per-chunk code is ~5000 instructions (every activation / derivative variant unrolled 32x) of which the plain
// path executes ~150, scattered: its instruction-cache misses cost ~1000 clocks per 32-column chunk, i.e. most of the
// epilogue's time and of each launch's un-overlapped tail (profiles/r2_gemm_timeline.md).
template <int kMode, bool kSplit, int kEpi>
__global__ void __launch_bounds__(kSplitThreads, 1) gemm_tf32_kernel(const __grid_constant__ GemmDev p) {
  constexpr bool kPair = kMode >= 1;
  constexpr bool kQuad = kMode >= 2;
  constexpr bool kShareB = kMode == 3;
  static_assert(!(kSplit && kQuad), "the in-pipeline 3xTF32 split runs on the single-CTA and CTA-pair engines");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve-up (all tile buffers 1024-byte aligned for the 128B swizzle patterns); identical in every CTA of a cluster
  // (the __align__(1024) on the dynamic array places it on a 1024-byte boundary of the CTA's shared window: no slack bytes)
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem_raw) & 1023u) != 0u) {
    if (threadIdx.x == 0) printf("capdec gemm: dynamic shared memory is not 1024-byte aligned (%u)\n", smem_u32(smem_raw));
    __trap();
  }
  const int bn_local = kPair ? p.block_n / 2 : p.block_n;  // B columns this CTA stages
  const int b_bytes = bn_local * kBlockK * 4;
  const int stage_bytes = kABytes + b_bytes;
  uint8_t* sA = smem;
  uint8_t* sB = smem + p.stages * kABytes;
  uint8_t* sAlo = smem + p.stages * stage_bytes;            // kSplit only: the lo = x - RN_tf32(x) tiles
  uint8_t* sBlo = sAlo + p.stages * kABytes;
  uint8_t* sStage = smem + (kSplit ? 2 : 1) * p.stages * stage_bytes;  // epilogue staging: C ping-pong [+ aux ping-pong]
  const int c_depth = p.c_depth;
  const int n_staging = c_depth + (p.mul_act ? kMulDepth : (p.has_aux ? c_depth : 0));
  float* sBias = reinterpret_cast<float*>(sStage + n_staging * kStagingBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sBias + 256);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tmem_full_bar = empty_bar + kMaxStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint64_t* mul_bar = tmem_empty_bar + 2;           // TMA loads of the epilogue input boxes (kMulDepth per epilogue warp)
  uint64_t* conv_bar = mul_bar + 4 * kMulDepth;     // kSplit: stage converted (hi in place, lo written) in every CTA of the pair
  uint64_t* sched_full = conv_bar + kMaxStages;     // tile id of ring slot i published (by the cluster's scheduler thread)
  uint64_t* sched_empty = sched_full + kSchedDepth; // ring slot i read by every role of every CTA of the cluster (rank-0 CTA's copy)
  int* sched_tile = reinterpret_cast<int*>(sched_empty + kSchedDepth);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sched_tile + kSchedDepth);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = kPair ? cluster_ctarank() : 0u;
  const uint32_t half = cta_rank & 1u;          // which 128 rows / which B half inside the pair
  const uint32_t pair_idx = cta_rank >> 1;      // which pair inside the quad
  const uint32_t pair_leader = cta_rank & ~1u;  // cluster rank of this pair's MMA-issuing CTA
  const bool leader = (half == 0);
  constexpr int kCtasPerPair = kPair ? 2 : 1;
  constexpr int kPairsPerCluster = kQuad ? 2 : 1;
  constexpr int kCtasPerCluster = kCtasPerPair * kPairsPerCluster;

  pdl_trigger();   // the next kernel of the stream may take this SM as soon as this CTA has left it
  if (threadIdx.x == 0) CAPDEC_TRACE(0, 0);
  if (threadIdx.x == 32) CAPDEC_TRACE(0, 6);   // a second warp's entry
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&p.tmA);
    prefetch_tensormap(&p.tmB);
    if (p.tail_ok) prefetch_tensormap(&p.tmBh);
    prefetch_tensormap(&p.tmC);
    if (p.has_aux) prefetch_tensormap(&p.tmAux);
    if (p.mul_act) prefetch_tensormap(&p.tmMul);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.stages; ++i) {
      // 1xTF32 pair: the leader's barrier collects the leader's arrive.expect_tx + the peer's remote arrive and the bytes
      // of both CTAs.  3xTF32: every CTA tracks its OWN loads (its converter warps wait on them locally).
      mbar_init(&full_bar[i], kSplit ? 1 : kCtasPerPair);
      mbar_init(&empty_bar[i], kPairsPerCluster);  // quad: both pairs' MMAs must have retired (multicast writes my smem)
      mbar_init(&conv_bar[i], kCtasPerPair * (kConvThreads / 32));  // one arrive per converter warp of each CTA
    }
    for (int i = 0; i < 4 * kMulDepth; ++i) mbar_init(&mul_bar[i], 1);
    for (int i = 0; i < kSchedDepth; ++i) {
      mbar_init(&sched_full[i], 1);
      // readers of a tile id per cluster: every CTA's producer thread + 4 epilogue warps (+ 4 converter warps), and the
      // MMA thread of every pair leader
      mbar_init(&sched_empty[i], kCtasPerCluster * (1 + 4 + (kSplit ? kConvThreads / 32 : 0)) + kPairsPerCluster);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], kCtasPerPair * kEpiThreads);  // pair: both CTAs' epilogues report to the leader
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (kPair) { tmem_alloc_pair(tmem_slot, kTmemCols); tmem_relinquish_pair(); }
    else { tmem_alloc(tmem_slot, kTmemCols); tmem_relinquish(); }
  }
  // (the opening cluster sync comes AFTER the scalar set-up below: the device-side row / reduction limits are global loads,
  // and their latency - like the reciprocals and tile counts derived from them - now overlaps the barrier initialisation and
  // the TMEM allocation of warps 1 and 2 instead of following them)

  // Everything above touched only this CTA's shared / tensor memory and the kernel parameters, so under programmatic
  // dependent launch it ran while the previous kernel of the stream was still draining its last wave; from here on the
  // kernel reads what its predecessors wrote.
  pdl_wait();
  // data-dependent extents (LM head over the non-ignored caption tokens only): every role skips the same tiles
  const int m_lim = p.m_limit ? __ldg(p.m_limit) : p.M;
  const int kb_lim = p.k_limit ? min(p.kb_total, (__ldg(p.k_limit) + kBlockK - 1) / kBlockK) : p.kb_total;
  const int kb_per_split = p.k_limit ? max(1, (kb_lim + p.splits - 1) / p.splits) : p.kb_per_split;  // rebalance the splits
  // cluster-level tile grid: m_tiles x n_tiles cluster tiles (each kCM x kCN*block_n), times split-K
  const int tile0 = (int)(blockIdx.x / kCtasPerCluster);
  const int tile_step = (int)(gridDim.x / kCtasPerCluster);
  constexpr int kPairM = kPair ? 2 * kBlockM : kBlockM;                       // rows per pair tile
  constexpr int kTileM = (kQuad && kShareB) ? 2 * kPairM : kPairM;            // rows per cluster tile
  // only the LIVE row tiles are enumerated, so that the round-robin over cluster tiles stays balanced under split-K
  const int m_tiles = min(p.m_tiles, (m_lim + kTileM - 1) / kTileM);
  // Wave quantisation: `base` tiles on `tile_step` clusters run floor(base / tile_step) full waves plus a tail of R tiles
  // that keeps only R clusters busy.  When 0 < 2R <= tile_step the tail tiles are cut in two along N (tile ids
  // F + 2r, F + 2r + 1 = the halves of tile F + r): twice as many clusters share the last wave, which then lasts about
  // 0.6 of a tile time.  Everything is derived from the LIVE row count, identically in every role of every CTA.
  const int base_tiles = m_tiles * p.n_tiles * p.splits;
  int tail_first = base_tiles;                 // first tile id that denotes a half tile
  if (!kQuad && p.tail_ok) {
    const int full = (base_tiles / tile_step) * tile_step, rest = base_tiles - full;
    if (full > 0 && rest > 0 && 2 * rest <= tile_step) tail_first = full;
  }
  const int total_tiles = base_tiles + (base_tiles - tail_first);
  // tile id -> linear (m, n, split) index, tile width and column offset inside the full-width tile
  auto decode = [&](int tile, int& lin, int& bn_t, int& n_off) {
    if (tile < tail_first) { lin = tile; bn_t = p.block_n; n_off = 0; }
    else { const int j = tile - tail_first; lin = tail_first + (j >> 1); bn_t = p.block_n >> 1; n_off = (j & 1) * bn_t; }
  };
  // linear index -> (row tile, column tile, reduction split)
  // (integer division through a float reciprocal + correction: exact for the < 2^22 tile ids here, and a fraction of the
  // ~150-clock cost of a hardware-sequenced 32-bit division - three of them sat on every role's path at every tile boundary)
  const int mn_tiles = p.n_tiles * m_tiles;
  const float rcp_mn = __frcp_rn((float)mn_tiles), rcp_n = __frcp_rn((float)p.n_tiles), rcp_m = __frcp_rn((float)m_tiles);
  auto fdiv = [](int x, int d, float rcp) {
    int q = __float2int_rz(__int2float_rz(x) * rcp);
    const int r = x - q * d;
    if (r < 0) --q;
    else if (r >= d) ++q;
    return q;
  };
  auto tile_mn = [&](int lin, int& m_blk, int& n_blk, int& split) {
    split = (p.splits > 1) ? fdiv(lin, mn_tiles, rcp_mn) : 0;
    const int r = lin - split * mn_tiles;
    if (p.raster) { n_blk = fdiv(r, m_tiles, rcp_m); m_blk = r - n_blk * m_tiles; }
    else { m_blk = fdiv(r, p.n_tiles, rcp_n); n_blk = r - m_blk * p.n_tiles; }
  };
  const int tile_n = (kQuad && !kShareB) ? 2 * p.block_n : p.block_n;         // columns per cluster tile
  // this CTA's pair tile inside the cluster tile
  const int pm_off = (kQuad && kShareB) ? (int)pair_idx * kPairM : 0;
  const int pn_off = (kQuad && !kShareB) ? (int)pair_idx * p.block_n : 0;
  const uint16_t mc_mask = (uint16_t)((1u << half) | (1u << (half + 2)));   // me and my twin in the other pair
  const uint16_t commit_empty_mask = kQuad ? (uint16_t)0xF : (uint16_t)0x3;
  const uint16_t commit_pair_mask = (uint16_t)(0x3u << (2 * pair_idx));

  // ---- tile order ------------------------------------------------------------------------------------------------
  // static (p.sched == NULL): cluster c takes tiles c, c + #clusters, ... ; dynamic: the scheduler thread of the cluster
  // (warp 3 of its rank-0 CTA) draws tile ids from a device-wide counter and publishes them, kSchedDepth ahead at most, to
  // a small ring in EVERY CTA of the cluster; each role walks the ring at its own pace.  Dynamic order keeps the kernel
  // efficient when it does not own the whole GPU (a collective or the optimizer running beside it): a CTA that starts late
  // simply draws fewer tiles, instead of serialising its fixed share of the problem behind everybody else.
  const bool dyn = p.sched != nullptr;
  auto next_tile_thread = [&](int& tile, uint32_t& cit) -> bool {      // for a single-thread role
    if (!dyn) { tile = (cit == 0) ? tile0 : tile + tile_step; ++cit; return tile < total_tiles; }
    const int slot = (int)(cit % kSchedDepth);
    mbar_wait_cluster(&sched_full[slot], (cit / kSchedDepth) & 1, 9);
    tile = *reinterpret_cast<volatile int*>(&sched_tile[slot]);
    // plain arrive, issued only once the id sits in a register (operand dependency): a cluster-scope RELEASE here made the
    // role wait for all of its outstanding memory traffic at every tile and cost 4-6 % per GEMM (profiles/r2_gemm_scheduler.md)
    if constexpr (kPair) mbar_arrive_remote_after(&sched_empty[slot], 0, (uint32_t)tile); else mbar_arrive(&sched_empty[slot]);
    ++cit;
    return tile >= 0;
  };
  auto next_tile_warp = [&](int& tile, uint32_t& cit) -> bool {        // for a whole warp (converged)
    if (!dyn) { tile = (cit == 0) ? tile0 : tile + tile_step; ++cit; return tile < total_tiles; }
    const int slot = (int)(cit % kSchedDepth);
    mbar_wait_cluster(&sched_full[slot], (cit / kSchedDepth) & 1, 10);
    tile = *reinterpret_cast<volatile int*>(&sched_tile[slot]);
    __syncwarp();                                                       // every lane holds the id before the slot is released
    if (lane == 0) {
      if constexpr (kPair) mbar_arrive_remote_after(&sched_empty[slot], 0, (uint32_t)tile); else mbar_arrive(&sched_empty[slot]);
    }
    ++cit;
    return tile >= 0;
  };

  tc_fence_before();
  if (threadIdx.x == 32) CAPDEC_TRACE(0, 7);   // barriers initialised, about to join the opening cluster sync
  if constexpr (kPair) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) CAPDEC_TRACE(0, 1);   // set-up done (barriers, TMEM, cluster sync)

  if (warp == 3) {
    // ============================== tile scheduler (one thread per cluster) ========================================
    if (dyn && lane == 0 && cta_rank == 0) {
      for (uint32_t it = 0;; ++it) {
        const int slot = (int)(it % kSchedDepth);
        mbar_wait_cluster(&sched_empty[slot], ((it / kSchedDepth) & 1) ^ 1, 8);
        int tile = atomicAdd(p.sched, 1);
        if (tile >= total_tiles) tile = -1;
#pragma unroll
        for (uint32_t c = 0; c < (uint32_t)kCtasPerCluster; ++c) {
          if constexpr (kPair) {
            st_shared_remote_u32(&sched_tile[slot], c, (uint32_t)tile);
            mbar_arrive_remote_release(&sched_full[slot], c);
          } else {
            sched_tile[slot] = tile;
            mbar_arrive(&sched_full[slot]);
          }
        }
        if (tile < 0) {
          // the last cluster to run dry re-arms the counters for the next launch that uses this slot (a CUDA-graph
          // replay launches with the same slot); every other cluster has made its final draw before counting itself done
          __threadfence();
          const int done = atomicAdd(p.sched + 1, 1);
          if (done == (int)(gridDim.x / kCtasPerCluster) - 1) {
            atomicExch(p.sched, 0);
            atomicExch(p.sched + 1, 0);
          }
          break;
        }
      }
    }
    __syncwarp();
  } else if (warp == 0) {
    // ============================== TMA producer (every CTA feeds its own smem, and its twin's when multicasting) ====
    // The WHOLE warp walks the loop converged and one elected lane issues: inside `if (lane == 0)` the compiler cannot
    // know that a single lane is active, moves every operand of the (uniform-datapath) TMA instructions through R2UR and
    // wraps each in an ELECT / BRA.U.ANY loop - 580 clk per k-block of dependent issue latency, which made this thread, not
    // L2 or the tensor pipe, the limiter of the mainloop (profiles/r2_gemm_timeline.md).
    {
      if (lane == 0) CAPDEC_TRACE(0, 4);   // producer enters its role
      int stage = 0;
      uint32_t phase = 0;
      int tile = 0;
      uint32_t cit = 0;
      while (next_tile_warp(tile, cit)) {
        int lin, bn_t, n_off;
        decode(tile, lin, bn_t, n_off);
        if (cit == 1 && lane == 0) CAPDEC_TRACE(0, 5);   // first tile id decoded
        const bool half_tile = bn_t != p.block_n;
        const int bnl_t = kPair ? bn_t / 2 : bn_t;                                  // B columns this CTA stages for this tile
        int m_blk, n_blk, split;
        tile_mn(lin, m_blk, n_blk, split);
        const int m0 = m_blk * kTileM + pm_off + (int)half * kBlockM;               // my 128 A rows
        const int n0 = n_blk * tile_n + pn_off + n_off + (kPair ? (int)half * bnl_t : 0);  // my B columns
        const int kb0 = split * kb_per_split;
        const int kb1 = min(kb0 + kb_per_split, kb_lim);
        if (m_blk * kTileM >= m_lim || kb0 >= kb1) continue;
        if (p.dbg & 64u) continue;   // timing experiment: raw tcgen05.mma issue rate, no stage handshake at all
        const CUtensorMap* mapA = &p.tmA;
        const CUtensorMap* mapB = half_tile ? &p.tmBh : &p.tmB;
        const uint32_t tx_bytes = (uint32_t)(kABytes + bnl_t * kBlockK * 4);
        if (lane == 0) CAPDEC_TRACE(1, (int)cit - 1);   // producer starts tile #cit-1
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1, 1);
          if (elect_one()) {
          const int k0 = kb * kBlockK;
          uint8_t* a_dst = sA + stage * kABytes;
          uint8_t* b_dst = sB + stage * b_bytes;
          uint64_t* fb = &full_bar[stage];
          constexpr bool kLocalBar = !kPair || kSplit;   // the loads are credited to THIS CTA's barrier
          if (p.dbg & 1u) {  // timing experiment: no loads, just hand the (stale) stage on
            if constexpr (kLocalBar) mbar_arrive(fb);
            else { if (leader) mbar_arrive(fb); else mbar_arrive_remote(fb, pair_leader); }
          } else {
          if constexpr (kLocalBar) mbar_arrive_expect_tx(fb, tx_bytes);
          auto load = [&](void* dst, const CUtensorMap* m, int x, int y) {
            if constexpr (kLocalBar) tma_load_2d(dst, m, fb, x, y); else tma_load_2d_pair(dst, m, fb, x, y);
          };
          auto load_mc = [&](void* dst, const CUtensorMap* m, int x, int y) { tma_load_2d_pair_mc(dst, m, fb, x, y, mc_mask); };
          auto load3 = [&](void* dst, const CUtensorMap* m, int y, int z) {   // {32 floats, 32 k-rows from y, slabs from z}
            if constexpr (kLocalBar) tma_load_3d(dst, m, fb, 0, y, z); else tma_load_3d_pair(dst, m, fb, 0, y, z);
          };
          auto load3_mc = [&](void* dst, const CUtensorMap* m, int y, int z) { tma_load_3d_pair_mc(dst, m, fb, 0, y, z, mc_mask); };
          // ---- A ----
          if constexpr (kQuad && !kShareB) {   // my half of the shared A slab -> me + twin
            if (!p.a_mn) load_mc(a_dst + pair_idx * (kABytes / 2), mapA, k0, m0 + (int)pair_idx * (kBlockM / 2));
            else if (p.a_3d) load3_mc(a_dst + pair_idx * (kABytes / 2), mapA, k0, m0 / 32 + (int)pair_idx * (kBlockM / 64));
            else {
#pragma unroll
              for (int i = 0; i < kBlockM / 64; ++i) {
                const int sl = (int)pair_idx * (kBlockM / 64) + i;
                load_mc(a_dst + sl * 4096, mapA, m0 + 32 * sl, k0);
              }
            }
          } else {
            if (!p.a_mn) load(a_dst, mapA, k0, m0);
            else if (p.a_3d) load3(a_dst, mapA, k0, m0 / 32);
            else {
#pragma unroll
              for (int i = 0; i < kBlockM / 32; ++i) load(a_dst + i * 4096, mapA, m0 + 32 * i, k0);
            }
          }
          // ---- B ----
          if constexpr (kQuad && kShareB) {    // my half of the shared B slab -> me + twin
            if (!p.b_mn) load_mc(b_dst + pair_idx * (b_bytes / 2), mapB, k0, n0 + (int)pair_idx * (bn_local / 2));
            else if (p.b_3d) load3_mc(b_dst + pair_idx * (b_bytes / 2), mapB, k0, n0 / 32 + (int)pair_idx * (bn_local / 64));
            else {
              const int ns = bn_local / 64;
              for (int i = 0; i < ns; ++i) {
                const int sl = (int)pair_idx * ns + i;
                load_mc(b_dst + sl * 4096, mapB, n0 + 32 * sl, k0);
              }
            }
          } else {
            if (!p.b_mn) load(b_dst, mapB, k0, n0);
            else if (p.b_3d) load3(b_dst, mapB, k0, n0 / 32);
            else {
              for (int i = 0; i < bnl_t / 32; ++i) load(b_dst + i * 4096, &p.tmB, n0 + 32 * i, k0);   // 32-wide slab boxes
            }
          }
          if constexpr (!kLocalBar) {
            if (leader) mbar_arrive_expect_tx(fb, 2u * tx_bytes);  // bytes landing in BOTH CTAs of my pair
            else mbar_arrive_remote(fb, pair_leader);
          }
          if (cit == 1 && kb - kb0 < 8) CAPDEC_TRACE(0, 8 + (kb - kb0));   // the first k-blocks requested
          }   // loads
          }   // elected lane
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ============================== MMA issuer (one thread of each pair's leader CTA) ==============================
    if (leader) {   // whole warp, converged; one elected lane issues (see the producer)
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint64_t adesc_base = ((uint64_t)p.adesc_hi << 32) | (uint64_t)p.adesc_lo16;
      const uint64_t bdesc_base = ((uint64_t)p.bdesc_hi << 32) | (uint64_t)p.bdesc_lo16;
      int tile = 0;
      uint32_t cit = 0;
      while (next_tile_warp(tile, cit)) {
        int lin, bn_t, n_off;
        decode(tile, lin, bn_t, n_off);
        const uint32_t idesc = (bn_t != p.block_n) ? p.idesc_h : p.idesc;
        int m_blk, n_blk, split;
        tile_mn(lin, m_blk, n_blk, split);
        const int kb0 = split * kb_per_split;
        const int kb1 = min(kb0 + kb_per_split, kb_lim);
        if (m_blk * kTileM >= m_lim || kb0 >= kb1) continue;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1, 2);
        tc_fence_after();
        if (lane == 0) CAPDEC_TRACE(2, 2 * ((int)cit - 1));       // accumulator free, tile's first k-block may be issued
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kAccCols);
        for (int kb = kb0; kb < kb1; ++kb) {
          uint32_t accumulate = (kb != kb0) ? 1u : 0u;   // (per k-block, in every lane: whichever lane is elected sees it)
          if (p.dbg & 64u) {}                                                     // (experiment: no wait)
          else if constexpr (kSplit) mbar_wait_cluster(&conv_bar[stage], phase, 6);   // hi/lo tiles of both CTAs are in place
          else mbar_wait(&full_bar[stage], phase, 3);
          tc_fence_after();
          if (cit == 1 && kb - kb0 < 8 && lane == 0) CAPDEC_TRACE(0, 16 + (kb - kb0));  // the first k-blocks landed (seen by the MMA issuer)
          if (elect_one()) {
          const uint32_t a_start = smem_u32(sA + stage * kABytes) >> 4;
          const uint32_t b_start = smem_u32(sB + stage * b_bytes) >> 4;
          auto mma = [&](uint64_t adesc, uint64_t bdesc) {
            if (!(p.dbg & 2u)) {
              if constexpr (kPair) umma_tf32_pair(d_tmem, adesc, bdesc, idesc, accumulate);
              else umma_tf32(d_tmem, adesc, bdesc, idesc, accumulate);
            }
            accumulate = 1;
          };
          if constexpr (kSplit) {
            const uint32_t alo_start = smem_u32(sAlo + stage * kABytes) >> 4;
            const uint32_t blo_start = smem_u32(sBlo + stage * b_bytes) >> 4;
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k) {
              const uint64_t ah = adesc_base | (uint64_t)((a_start + k * p.a_kstep) & 0x3FFFu);
              const uint64_t bh = bdesc_base | (uint64_t)((b_start + k * p.b_kstep) & 0x3FFFu);
              const uint64_t al = adesc_base | (uint64_t)((alo_start + k * p.a_kstep) & 0x3FFFu);
              const uint64_t bl = bdesc_base | (uint64_t)((blo_start + k * p.b_kstep) & 0x3FFFu);
              mma(al, bh);   // the two cross terms first, then the leading term
              mma(ah, bl);
              mma(ah, bh);
            }
          } else {
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k) {
              mma(adesc_base | (uint64_t)((a_start + k * p.a_kstep) & 0x3FFFu),
                  bdesc_base | (uint64_t)((b_start + k * p.b_kstep) & 0x3FFFu));
            }
          }
          // smem slot reusable (in every CTA that writes into my pair's smem) once these MMAs retire
          if constexpr (kPair) umma_commit_mc(&empty_bar[stage], commit_empty_mask); else umma_commit(&empty_bar[stage]);
          }   // elected lane
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        // accumulator complete -> epilogue(s) of my pair
        if (elect_one()) {
          if constexpr (kPair) umma_commit_mc(&tmem_full_bar[acc], commit_pair_mask); else umma_commit(&tmem_full_bar[acc]);
        }
        __syncwarp();
        if (lane == 0) CAPDEC_TRACE(2, 2 * ((int)cit - 1) + 1);   // tile's last k-block issued
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (kSplit && warp >= 8) {
    // ============================== 3xTF32 operand converters (4 warps, every CTA converts what IT staged) =========
    // hi = RN_tf32(x): add half a TF32 ulp to the magnitude bits and clear the 13 low mantissa bits (= cvt.rna.tf32.f32,
    // on the integer pipe); lo = x - hi is exact in fp32 and the tensor core keeps its leading 11 bits.  Element-wise, so
    // the swizzled TMA layout of the tile carries over to the lo tile unchanged.
    const int ct = threadIdx.x - kThreads;
    int stage = 0;
    uint32_t phase = 0;
    auto split4 = [](float4& x, float4& lo) {
      float* xv = reinterpret_cast<float*>(&x);
      float* lv = reinterpret_cast<float*>(&lo);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float hi = __uint_as_float((__float_as_uint(xv[j]) + 0x1000u) & 0xFFFFE000u);
        lv[j] = xv[j] - hi;
        xv[j] = hi;
      }
    };
    int tile = 0;
    uint32_t cit = 0;
    while (next_tile_warp(tile, cit)) {
      int lin, bn_t, n_off;
      decode(tile, lin, bn_t, n_off);
      int m_blk, n_blk, split;
      tile_mn(lin, m_blk, n_blk, split);
      const int kb0 = split * kb_per_split;
      const int kb1 = min(kb0 + kb_per_split, kb_lim);
      if (m_blk * kTileM >= m_lim || kb0 >= kb1) continue;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase, 7);
        float4* a4 = reinterpret_cast<float4*>(sA + stage * kABytes);
        float4* al4 = reinterpret_cast<float4*>(sAlo + stage * kABytes);
        float4* b4 = reinterpret_cast<float4*>(sB + stage * b_bytes);
        float4* bl4 = reinterpret_cast<float4*>(sBlo + stage * b_bytes);
#pragma unroll
        for (int i = 0; i < kABytes / 16 / kConvThreads; ++i) {
          float4 x = a4[ct + i * kConvThreads], lo;
          split4(x, lo);
          a4[ct + i * kConvThreads] = x;
          al4[ct + i * kConvThreads] = lo;
        }
        const int nb4 = (kPair ? bn_t / 2 : bn_t) * kBlockK * 4 / 16;   // a half tile stages half the B rows
#pragma unroll 4
        for (int i = ct; i < nb4; i += kConvThreads) {
          float4 x = b4[i], lo;
          split4(x, lo);
          b4[i] = x;
          bl4[i] = lo;
        }
        fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core's (async proxy) operand reads
        __syncwarp();
        if (lane == 0) {
          // (plain arrive: the tensor core reads this CTA's tiles through the async proxy, which the fence above has
          // ordered after the conversion; a cluster-scope release would stall the warp on all of its outstanding traffic)
          if (kPair && !leader) mbar_arrive_remote(&conv_bar[stage], pair_leader);
          else mbar_arrive(&conv_bar[stage]);
        }
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ============================== epilogue (4 warps = this CTA's 128 TMEM lanes) ==============================
    // Every warp runs its OWN pipeline over its 32 rows: tcgen05.ld -> bias / activation -> swizzled 4 KB staging box ->
    // TMA store of a {32 cols, 32 rows} box.  No cross-warp barrier inside the chunk loop (two per tile remain, for the
    // bias tile and the column sums); staging is ping-pong per warp, gated by the warp leader's bulk-group counter.
    const int q = warp & 3;                   // TMEM lane quadrant this warp may access
    const int epi_tid = threadIdx.x - 4 * 32;
    const int e_act = kEpi == 0 ? p.act : (kEpi == 2 ? 4 : 0);
    const int e_mul = kEpi == 0 ? p.mul_act : (kEpi == 3 ? 4 : 0);
    const bool e_aux = kEpi == 0 ? (p.has_aux != 0) : (kEpi == 2);
    float* const e_colsum = (kEpi == 0 || kEpi == 3) ? p.colsum : nullptr;
    const bool e_lsu = kEpi == 0 ? (p.lsu_store != 0) : false;
    uint8_t* wst = sStage + q * (n_staging * kWarpStagingBytes);   // this warp's staging: C[0], C[1], (X[0], X[1])
    uint64_t* my_mul_bar = mul_bar + kMulDepth * q;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t store_idx = 0;  // running chunk counter: staging buffers alternate ACROSS tiles too
    uint32_t mul_idx = 0;    // running count of epilogue-input chunks consumed (buffer = idx % kMulDepth, phase = (idx / kMulDepth) & 1)
    // this warp's (tile, chunk) stream: first row of the warp's 32-row box, the pair's first column, chunks to store
    auto tile_coords = [&](int tile, int& m0w, int& n0w, int& nch, int& bn_t, int& split) -> bool {
      int lin, n_off;
      decode(tile, lin, bn_t, n_off);
      int m_blk, n_blk;
      tile_mn(lin, m_blk, n_blk, split);
      if (m_blk * kTileM >= m_lim || split * kb_per_split >= kb_lim) return false;
      m0w = m_blk * kTileM + pm_off + (int)half * kBlockM + q * 32;
      n0w = n_blk * tile_n + pn_off + n_off;   // the tile's columns (each CTA stores its 128 rows x bn_t)
      nch = (p.N > n0w) ? min(bn_t, p.N - n0w + 31) / 32 : 0;  // skip chunks (or whole tiles) past N
      if (p.dbg & 8u) nch = 0;
      return true;
    };
    // Epilogue-input prefetcher (lane 0): runs kMulDepth chunks AHEAD of the consumer inside a tile; the first kMulDepth
    // boxes of a tile are requested as soon as its id is known, i.e. while the tile's MMAs are still running, so that the
    // HBM latency of the 4 KB input boxes never sits on the epilogue's critical path (one-chunk-ahead prefetch left the
    // fused GELU-backward dgrad at 315 TF/s vs 650 for the plain dgrad).
    int la_c = 0, la_nch = 0, la_m0 = 0, la_n0 = 0;
    uint32_t mul_issued = 0;
    auto mul_prefetch = [&]() {
      if (la_c >= la_nch) return;
      const uint32_t slot = mul_issued % kMulDepth;
      uint64_t* mb = &my_mul_bar[slot];
      mbar_arrive_expect_tx(mb, kWarpStagingBytes);
      tma_load_2d(wst + (c_depth + slot) * kWarpStagingBytes, &p.tmMul, mb, la_n0 + la_c * 32, la_m0);
      ++mul_issued; ++la_c;
    };
    int tile = 0;
    uint32_t cit = 0;
    while (next_tile_warp(tile, cit)) {
      int m0, n0, n_chunks, bn_t, split;
      if (!tile_coords(tile, m0, n0, n_chunks, bn_t, split)) continue;
      if (e_mul && lane == 0) {   // every box of the previous tile has been consumed: all kMulDepth slots are free
        la_c = 0; la_nch = n_chunks; la_m0 = m0; la_n0 = n0;
        for (int i = 0; i < kMulDepth; ++i) mul_prefetch();
      }
      const bool use_bias = (p.bias != nullptr) && (split == 0);
      if (use_bias) {
        named_bar_sync(1, kEpiThreads);  // the warps run independently: nobody may still be reading the previous bias tile
        for (int i = epi_tid; i < bn_t; i += kEpiThreads) sBias[i] = (n0 + i < p.N) ? __ldg(p.bias + n0 + i) : 0.0f;
      }
      mbar_wait(&tmem_full_bar[acc], acc_phase, 4);
      tc_fence_after();
      if (q == 0 && lane == 0) CAPDEC_TRACE(3, 3 * ((int)cit - 1));       // accumulator complete
      if (use_bias) named_bar_sync(1, kEpiThreads);  // bias tile visible to all 4 warps
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * kAccCols);
      if (n_chunks == 0) {  // nothing to store (or dbg 8): release the accumulator right away
        tc_fence_before();
        if constexpr (kPair) mbar_arrive_remote(&tmem_empty_bar[acc], pair_leader); else mbar_arrive(&tmem_empty_bar[acc]);
        n_chunks = 0;
      }
      // One 32-column chunk: bias / fused multiply / activation -> swizzled staging box -> TMA store.  The staging boxes
      // rotate c_depth deep (per warp), so up to c_depth - 1 stores are still being read by the TMA engine - which also
      // serves the mainloop's loads and answers late - while the warp fills the next box.
      auto process = [&](float (&v)[32], int c) {
        if (q == 0 && lane == 0) CAPDEC_TRACE(3, 32 + 4 * c);        // chunk c: accumulator columns in registers
        if (use_bias) {
          const float4* b4 = reinterpret_cast<const float4*>(sBias + c * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bb = b4[j];
            v[4 * j] += bb.x; v[4 * j + 1] += bb.y; v[4 * j + 2] += bb.z; v[4 * j + 3] += bb.w;
          }
        }
        if (e_mul) {  // v *= act'(input box), read back from the 128B-swizzled TMA layout (row = lane)
          const uint32_t slot = mul_idx % kMulDepth;
          mbar_wait(&my_mul_bar[slot], (mul_idx / kMulDepth) & 1, 5);
          const float4* u4 = reinterpret_cast<const float4*>(wst + (c_depth + slot) * kWarpStagingBytes + lane * 128);
          ++mul_idx;
          float uu[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 u = u4[j ^ (lane & 7)];
            uu[4 * j] = u.x; uu[4 * j + 1] = u.y; uu[4 * j + 2] = u.z; uu[4 * j + 3] = u.w;
          }
          apply_mul32(v, uu, e_mul, p.exact != 0);
        }
        const uint32_t pp = store_idx++ % (uint32_t)c_depth;
        uint8_t* buf0 = wst + pp * kWarpStagingBytes;                // C box
        uint8_t* buf1 = wst + (c_depth + pp) * kWarpStagingBytes;    // aux (pre-activation / derivative) box
        // (elect.sync names the same leader for the same member mask every time: the lane that waits here is the lane that
        // committed the bulk groups below; `if (lane == 0)` would wrap each TMA instruction in an ELECT / BRA.U.ANY loop)
        if (!e_lsu && elect_one()) {  // the group this lane committed c_depth chunks ago has released buf[pp]
          if (c_depth == 2) tma_store_wait_read<1>();
          else if (c_depth == 3) tma_store_wait_read<2>();
          else tma_store_wait_read<3>();
        }
        __syncwarp();
        if (e_mul && lane == 0) mul_prefetch();  // every lane has read the input box just consumed: refill its slot
        if (e_act == 4) {      // aux <- gelu_new'(pre-activation), C <- gelu_new(pre-activation)
          float dv[32];
          gelu_fwd_and_grad32(v, dv, p.exact != 0);
          if (e_aux) {
            float4* d1 = reinterpret_cast<float4*>(buf1 + lane * 128);
#pragma unroll
            for (int j = 0; j < 8; ++j) d1[j ^ (lane & 7)] = make_float4(dv[4 * j], dv[4 * j + 1], dv[4 * j + 2], dv[4 * j + 3]);
          }
        } else {
          if (e_aux) {     // aux <- pre-activation
            float4* d1 = reinterpret_cast<float4*>(buf1 + lane * 128);
#pragma unroll
            for (int j = 0; j < 8; ++j) d1[j ^ (lane & 7)] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
          if (e_act != 0) apply_act32(v, e_act, p.exact != 0);
        }
        if (!(p.dbg & 32u)) {
          float4* d0 = reinterpret_cast<float4*>(buf0 + lane * 128);
#pragma unroll
          for (int j = 0; j < 8; ++j) d0[j ^ (lane & 7)] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        if (q == 0 && lane == 0) CAPDEC_TRACE(3, 32 + 4 * c + 1);    // staged
        if (!e_lsu && !(p.dbg & 16u)) fence_proxy_async_smem();
        __syncwarp();
        if (q == 0 && lane == 0) CAPDEC_TRACE(3, 32 + 4 * c + 2);    // fenced
        if (e_colsum) {  // bias gradient: lane = column of this chunk, summed over this warp's live staged rows and sent
          // straight to global memory as one coalesced 128-byte reduction per chunk (no cross-warp exchange, no barrier:
          // the four epilogue warps stay independent; shared-memory atomics + two named barriers per tile cost 20 us here)
          const int nrows = min(32, m_lim - m0);
          const int col = n0 + c * 32 + lane;
          float cs0 = 0.f, cs1 = 0.f;
          const uint8_t* colp = buf0 + ((lane & 3) << 2);
          const int l4 = lane >> 2;
          if (nrows >= 32) {
#pragma unroll
            for (int rr = 0; rr < 32; rr += 2) {
              cs0 += *reinterpret_cast<const float*>(colp + rr * 128 + ((l4 ^ (rr & 7)) << 4));
              cs1 += *reinterpret_cast<const float*>(colp + (rr + 1) * 128 + ((l4 ^ ((rr + 1) & 7)) << 4));
            }
          } else {
            for (int rr = 0; rr < nrows; ++rr) cs0 += *reinterpret_cast<const float*>(colp + rr * 128 + ((l4 ^ (rr & 7)) << 4));
          }
          if (nrows > 0 && col < p.N) atomicAdd(e_colsum + col, cs0 + cs1);
        }
        if (e_lsu) {
          // experiment (off by default, see launch_plan): the staging box is read back row-wise - 8 lanes cover one 128-byte
          // row, a warp instruction four rows - and stored with coalesced 16-byte st.global instead of a TMA store
          if (!(p.dbg & 4u)) {
            const int chunk = lane & 7, rsub = lane >> 3;
            const int col = n0 + c * 32 + chunk * 4;
            if (col < p.N) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int row = 4 * i + rsub;
                if (m0 + row < p.M) {
                  const int off = row * 128 + ((chunk ^ (row & 7)) << 4);
                  const size_t g = (size_t)(m0 + row) * (size_t)p.ldc + (size_t)col;
                  st_stream(reinterpret_cast<float4*>(p.c_ptr + g), *reinterpret_cast<const float4*>(buf0 + off));
                  if (e_aux) st_stream(reinterpret_cast<float4*>(p.aux_ptr + g), *reinterpret_cast<const float4*>(buf1 + off));
                }
              }
            }
          }
        } else if (!(p.dbg & 4u)) {
          if (elect_one()) {
            if (p.accumulate) tma_reduce_add_2d(&p.tmC, buf0, n0 + c * 32, m0);
            else tma_store_2d(&p.tmC, buf0, n0 + c * 32, m0);
            if (e_aux) tma_store_2d(&p.tmAux, buf1, n0 + c * 32, m0);
            tma_store_commit();
          }
        }
        if (q == 0 && lane == 0) CAPDEC_TRACE(3, 32 + 4 * c + 3);    // store issued
      };
      auto release_acc = [&]() {   // all TMEM reads of this accumulator stage are done -> hand it back to the (leader's) MMA warp
        if (q == 0 && lane == 0) CAPDEC_TRACE(3, 3 * ((int)cit - 1) + 1);
        tc_fence_before();
        if constexpr (kPair) mbar_arrive_remote(&tmem_empty_bar[acc], pair_leader); else mbar_arrive(&tmem_empty_bar[acc]);
      };
      // TMEM reads run one chunk AHEAD of the arithmetic: the load of chunk c + 1 is in flight while chunk c is processed
      float va[32], vb[32];
      const bool no_tmem = (p.dbg & 128u) != 0;   // timing experiment: the epilogue without its TMEM reads
      if (no_tmem) {
#pragma unroll
        for (int j = 0; j < 32; ++j) { va[j] = 0.f; vb[j] = 0.f; }
      }
      if (n_chunks > 0 && !no_tmem) tmem_ld32(t_row, va);
      for (int c = 0; c < n_chunks; c += 2) {
        if (!no_tmem) tmem_ld_wait_for(va);
        if (c + 1 < n_chunks) { if (!no_tmem) tmem_ld32(t_row + (uint32_t)((c + 1) * 32), vb); }
        else release_acc();
        process(va, c);
        if (c + 1 < n_chunks) {
          if (!no_tmem) tmem_ld_wait_for(vb);
          if (c + 2 < n_chunks) { if (!no_tmem) tmem_ld32(t_row + (uint32_t)((c + 2) * 32), va); }
          else release_acc();
          process(vb, c + 1);
        }
      }
      if (q == 0 && lane == 0) CAPDEC_TRACE(3, 3 * ((int)cit - 1) + 2);   // tile's last chunk handed to the TMA engine
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    __syncwarp();
    // the staging boxes must have been READ before this CTA's shared memory goes away; the writes themselves are complete
    // (and visible) at grid completion like any other store
    if (elect_one()) tma_store_wait_read<0>();
    if (q == 0 && lane == 0) CAPDEC_TRACE(0, 2);   // every store of this warp has reached global memory
  }

  tc_fence_before();
  if constexpr (kPair) cluster_sync_all(); else __syncthreads();  // nobody exits while a peer may still touch its smem/barriers
  if (threadIdx.x == 0) CAPDEC_TRACE(0, 3);     // after the closing cluster sync
  if (warp == 2) {
    tc_fence_after();
    if constexpr (kPair) tmem_dealloc_pair(tmem_base, kTmemCols); else tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

// L2 promotion of every tensor map (CAPDEC_GEMM_L2PROMO = 0 none, 1 64 B, 2 128 B, 3 256 B; bring-up / A-B switch)
static CUtensorMapL2promotion l2_promotion() {
  static const char* env = getenv("CAPDEC_GEMM_L2PROMO");
  const int v = env ? atoi(env) : 3;
  return v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : v == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
       : v == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
}

// 2-D fp32 tensor map: dim0 = contiguous extent, dim1 = rows with pitch `pitch_elems`.
static int make_map(CUtensorMap* m, const void* ptr, uint64_t dim0, uint64_t dim1, uint64_t pitch_elems,
                    uint32_t box0, uint32_t box1, CUtensorMapSwizzle swz) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_last_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return CAPDEC_ERR_CUDA;
  }
  cuuint64_t dims[2] = {dim0, dim1};
  cuuint64_t strides[1] = {pitch_elems * 4};
  cuuint32_t box[2] = {box0, box1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, l2_promotion(), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed (%d): ptr=%p dims=(%llu,%llu) pitch=%llu box=(%u,%u) swz=%d", (int)r,
                   ptr, (unsigned long long)dim0, (unsigned long long)dim1, (unsigned long long)pitch_elems, box0,
                   box1, (int)swz);
    return CAPDEC_ERR_CUDA;
  }
  return CAPDEC_OK;
}

// 3-D view of an MN-major operand: [MN/32 slabs][K rows][32 floats]; one box = {32, 32 k-rows, nslabs} lands in shared
// memory slab after slab (4 KB each), i.e. exactly the layout the UMMA descriptor expects.  Needs MN % 32 == 0, or a row
// pitch that covers the 32-rounded extent (the tied-embedding gradient: MN = V = 50257 inside a 50304-float pitch): the pad
// columns only feed accumulator rows / columns >= M / N, which the bounds of the C tensor map clip at the store.
static int make_map_mn3d(CUtensorMap* m, const void* ptr, uint64_t mn, uint64_t k_rows, uint64_t pitch_elems,
                         uint32_t nslabs, CUtensorMapSwizzle swz) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return CAPDEC_ERR_CUDA;
  cuuint64_t dims[3] = {32, k_rows, (mn + 31) / 32};   // a ragged last slab reads (finite or not) pad columns < pitch
  cuuint64_t strides[2] = {pitch_elems * 4, 128};
  cuuint32_t box[3] = {32, kBlockK, nslabs};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, l2_promotion(), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? CAPDEC_OK : CAPDEC_ERR_UNSUPPORTED;
}

// bring-up overrides for the MN-major encoding (see capdec_gemm_debug_mn_encoding)
static int g_mn_layout = -1, g_mn_lbo = -1, g_mn_sbo = -1, g_mn_swz = -1;

struct OperandEnc {
  uint32_t desc_hi, desc_lo16, kstep;
  CUtensorMapSwizzle swz;
};
static OperandEnc operand_encoding(bool mn_major) {
  OperandEnc e;
  uint32_t layout, lbo, sbo;
  if (!mn_major) {
    layout = 2;  // SWIZZLE_128B
    lbo = 16;    // unused for swizzled K-major
    sbo = 1024;  // 8 rows x 128 B
    e.kstep = (kUmmaK * 4) >> 4;
    e.swz = CU_TENSOR_MAP_SWIZZLE_128B;
  } else {
    layout = g_mn_layout >= 0 ? (uint32_t)g_mn_layout : 1;  // SWIZZLE_128B_BASE32B
    lbo = g_mn_lbo >= 0 ? (uint32_t)g_mn_lbo : 4096;        // next 32-wide MN slab (32 k-rows x 128 B)
    sbo = g_mn_sbo >= 0 ? (uint32_t)g_mn_sbo : 512;         // next group of 4 k-rows
    e.kstep = (kUmmaK * 128) >> 4;                          // 8 k-rows x 128 B
    e.swz = g_mn_swz >= 0 ? (CUtensorMapSwizzle)g_mn_swz : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  }
  e.desc_lo16 = ((lbo >> 4) & 0x3FFFu) << 16;
  e.desc_hi = ((sbo >> 4) & 0x3FFFu) | (1u << 14) /* version = 1 (sm_100) */ | (layout << 29);
  return e;
}

// tile order of the persistent kernels: 0 = static round-robin (default: fastest while a GEMM owns the GPU, which is the
// single-GPU step), 1 = dynamic (device-wide tile counter: robust when collectives / the optimizer share the SMs)
static std::atomic<int> g_sched_dynamic{0};
static long long* g_trace = nullptr;   // capdec_gemm_debug_trace
static int g_force_pair = -1;  // bring-up override: -1 auto, 0 never, 1 always (pairs), 2/3 force quad modes
static thread_local int t_m_hint = 0;   // expected live rows of the next row-limited GEMMs (capdec_gemm_set_row_hint)

// pick block_n and split-K for a given engine: `units` = concurrently running tile owners (SMs, pairs or quads),
// tile_m rows and nmul*block_n columns per cluster tile
static void choose_tiling(int M, int N, int kb_total, int accumulate, int units, int tile_m, int nmul, int min_bn,
                          bool allow192, int& block_n, int& splits) {
  const int m_tiles = (M + tile_m - 1) / tile_m;
  auto cost = [&](int bn, int s) {
    const long tiles = (long)m_tiles * ((N + nmul * bn - 1) / (nmul * bn)) * s;
    const long waves = (tiles + units - 1) / units;
    const double kb = (double)((kb_total + s - 1) / s);
    // per-tile time ~ k-blocks x (operand bytes per k-block, the L2 feed is the limiter) + epilogue
    // (measured with the CTA timeline, profiles/r2_gemm_timeline.md: a 192-wide tile issues its k-blocks in 13.8 k clk against
    // 14.0 k for a 256-wide one - the tensor pipe does not get proportionally faster below N = 256)
    const double per_kb = (bn >= 256) ? 1.0 : (bn == 192 ? 0.95 : (bn == 128 ? 0.62 : 0.40));
    const double epi = (bn / 256.0) * 5.0;
    return waves * (kb * per_kb + epi + 1.5);
  };
  int best_bn = block_n, best_s = splits;
  double best = 1e30;
  const int bns[4] = {256, 192, 128, 64};
  for (int bi = 0; bi < 4; ++bi) {
    const int bn = bns[bi];
    if (block_n > 0 && bn != block_n) continue;
    if (bn < min_bn) continue;
    if (bn == 192 && !allow192 && block_n != 192) continue;
    const int smax = (splits > 0) ? splits : (accumulate ? 32 : 1);
    for (int s = (splits > 0 ? splits : 1); s <= smax; ++s) {
      if (s > 1 && kb_total / s < 4) break;
      const double c = cost(bn, s);
      if (c < best * 0.999) { best = c; best_bn = bn; best_s = s; }
    }
  }
  block_n = best_bn > 0 ? best_bn : (min_bn > 128 ? min_bn : 128);
  splits = best_s > 0 ? best_s : 1;
}

// programmatic dependent launch of the GEMM kernels (CAPDEC_GEMM_PDL=0 turns it off: bring-up / A-B switch)
static bool gemm_pdl() {
  static const char* env = getenv("CAPDEC_GEMM_PDL");
  return !(env && env[0] == '0');
}

template <int kMode, bool kSplit, int kEpi>
static int max_clusters(int cluster_size, int smem_bytes) {
  static std::atomic<int> cached_dev[kMaxDevices];   // per instantiation and per device
  std::atomic<int>& cached = cached_dev[current_device()];
  if (cached.load(std::memory_order_relaxed) > 0) return cached.load(std::memory_order_relaxed);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(cluster_size * 64);
  cfg.blockDim = dim3(kSplit ? kSplitThreads : kThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = cluster_size;
  attr.val.clusterDim.y = 1;
  attr.val.clusterDim.z = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, gemm_tf32_kernel<kMode, kSplit, kEpi>, &cfg) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    n = num_sms() / cluster_size;
  }
  cached.store(n, std::memory_order_relaxed);
  return n;
}

template <int kMode, bool kSplit, int kEpi>
static int launch_clustered(const GemmDev& p, int cluster_size, int total_tiles, int smem_bytes, cudaStream_t stream) {
  const int maxc = max_clusters<kMode, kSplit, kEpi>(cluster_size, smem_bytes);
  const int nc = total_tiles < maxc ? total_tiles : maxc;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(cluster_size * nc);
  cfg.blockDim = dim3(kSplit ? kSplitThreads : kThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster_size;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = gemm_pdl() ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_tf32_kernel<kMode, kSplit, kEpi>, p);
  if (e != cudaSuccess) return check_cuda(e, "cudaLaunchKernelEx(gemm_tf32_kernel)");
  return CAPDEC_OK;
}

}  // namespace capdec

using namespace capdec;

extern "C" const char* capdec_last_error(void) { return g_err; }
extern "C" int capdec_version(void) { return 100; }
extern "C" int64_t capdec_launch_count(void) { return g_launches.load(); }

extern "C" void capdec_gemm_debug_force_pair(int mode) { g_force_pair = mode; }  // -1 auto, 0..3 engine
// bring-up: every following GEMM launch writes cycle stamps (clock64 of CTA 0's SM) of its roles into trace[4][64]
// (int64, device memory; NULL = off): [0] CTA: entry, set-up done, stores drained, closing sync; [1] producer: start of its
// i-th tile; [2] MMA issuer: 2i = accumulator free, 2i+1 = tile's last k-block issued; [3] epilogue warp 4: 3i = accumulator
// complete, 3i+1 = last TMEM read, 3i+2 = last chunk handed to the TMA engine
extern "C" void capdec_gemm_debug_trace(void* trace_dev) { g_trace = static_cast<long long*>(trace_dev); }
extern "C" int capdec_gemm_set_schedule(int dynamic) { return g_sched_dynamic.exchange(dynamic ? 1 : 0); }

// Tiling hint for GEMMs launched with a device-side row limit (packed caption batches): the row count is data dependent
// and unknown to the host, so engine / tile width / wave quantisation are chosen for `rows` (0 = the static M).  Results
// do not depend on the hint.
extern "C" void capdec_gemm_set_row_hint(int rows) { t_m_hint = rows > 0 ? rows : 0; }

extern "C" void capdec_gemm_debug_mn_encoding(int layout_type, int lbo_bytes, int sbo_bytes, int tma_swizzle) {
  g_mn_layout = layout_type;
  g_mn_lbo = lbo_bytes;
  g_mn_sbo = sbo_bytes;
  g_mn_swz = tma_swizzle;
}

extern "C" int capdec_gemm_tf32_ex(const float* A, int a_major, int64_t lda, const float* B, int b_major, int64_t ldb,
                                   float* C, int64_t ldc, int M, int N, int K, const float* bias, int act, float* aux,
                                   int accumulate, int precision, const float* a_lo, const float* b_lo, int block_n,
                                   int split_k, const int32_t* m_limit_dev, const int32_t* k_limit_dev,
                                   capdec_stream_t stream_);

extern "C" int capdec_gemm_tf32(const float* A, int a_major, int64_t lda, const float* B, int b_major, int64_t ldb,
                                float* C, int64_t ldc, int M, int N, int K, const float* bias, int act, float* aux,
                                int accumulate, int precision, const float* a_lo, const float* b_lo, int block_n,
                                int split_k, capdec_stream_t stream_) {
  return capdec_gemm_tf32_ex(A, a_major, lda, B, b_major, ldb, C, ldc, M, N, K, bias, act, aux, accumulate, precision,
                             a_lo, b_lo, block_n, split_k, nullptr, nullptr, stream_);
}

// ---------------------------------------------------------------------------------------------------------------
// plan (engine, tile width, split-K)  ->  launch;  measured plan selection ("autotune") on top of the heuristic
// ---------------------------------------------------------------------------------------------------------------
namespace capdec {

struct GemmArgs {
  const float *A, *B; float* C;
  int a_major, b_major; int64_t lda, ldb, ldc;
  int M, N, K;
  const float* bias; int act; float* aux; int accumulate, precision;
  const float *a_lo, *b_lo;
  const int32_t *m_limit, *k_limit;
  const float* mul_in; int mul_act; float* colsum;
};
struct GemmPlan { int mode, bn, splits; };

static inline int plan_units(int mode) { return mode >= 2 ? 33 : num_sms() / (mode == 0 ? 1 : 2); }  // quads: 132 of 148 SMs
static inline int plan_tile_m(int mode) { return mode == 0 ? kBlockM : (mode == 3 ? 4 * kBlockM : 2 * kBlockM); }
static inline bool plan_ok192(int mode, int b_major) { return !(mode == 3 && b_major); }

// heuristic plan (also the fallback whenever nothing was measured for this problem)
static GemmPlan heuristic_plan(const GemmArgs& a, int block_n, int split_k, int forced) {
  const int M = a.M, N = a.N;
  int mode = ((M > kBlockM) && (N >= 128) && (block_n == 0 || block_n >= 128)) ? 1 : 0;
  if (forced == 0) mode = 0;
  const int m_live = (a.m_limit && t_m_hint > 0 && t_m_hint < M) ? t_m_hint : M;
  if (mode >= 1) {
    const bool quad_ok = (M > 2 * kBlockM || N > 256);
    if (a.precision) {
      // 3xTF32 is tensor-bound (three MMAs per staged byte): the operand-sharing quads have nothing to win
      if (forced == 0) mode = 0;
    } else if (forced < 0 && quad_ok) {
      // Measured on B200 (profiles/r1_gemm_modes.md): stacking the two pairs in M and multicasting B (mode 3) wins
      // 10-16 % at the dense C2 extent (M = 12800) when B is MN-major; sharing A (mode 2) loses to wave quantisation on
      // every shape of this model, and few-row problems (wgrad, M = 768) stay on plain pairs.
      // Round 2, after the pipeline work (profiles/r2_gemm_timeline.md): at dense extents plain pairs are at least as fast
      // on every shape (roofline anchor 72.6 us = 0.759 of peak on pairs, 74.3 us = 0.742 on quads, 148 vs 132 SMs), so the
      // quad is proposed for row-limited launches only (packed qkv: 56 vs 61 us) - where the Trainer measures the plan anyway.
      const int mt = (m_live + 255) / 256;
      const double waste_b = (double)(((mt + 1) / 2) * 2) / mt;
      if (a.m_limit && mt >= 8 && waste_b <= 1.15 && a.b_major && block_n != 192) mode = 3;
    } else if (forced == 2 || forced == 3) {
      mode = forced;
    }
  }
  GemmPlan pl;
  pl.mode = mode;
  pl.bn = block_n;
  pl.splits = a.accumulate ? split_k : 1;
  static const char* env_192 = getenv("CAPDEC_GEMM_BN192");   // bring-up switch: 0 keeps the automatic choice to 256/128/64
  const bool allow192 = plan_ok192(mode, a.b_major) && !(env_192 && env_192[0] == '0');
  const int kb_total = (a.K + kBlockK - 1) / kBlockK;
  choose_tiling(m_live, N, kb_total, a.accumulate, plan_units(mode), plan_tile_m(mode), mode == 2 ? 2 : 1, mode == 0 ? 64 : 128,
                allow192, pl.bn, pl.splits);
  if (a.precision) {
    // 3xTF32: two tiles (hi, lo) per operand and stage -> the single-CTA engine fits 128 columns at most
    if (mode == 0 && pl.bn > 128 && block_n == 0) pl.bn = 128;
    // the tensor core adds into its fp32 accumulator with truncation, so the error of one accumulation chain grows with
    // its length: reductions that may be split (accumulate = 1, reduce-add in L2 rounds to nearest) keep chains short
    static const char* env_chain = getenv("CAPDEC_X3_CHAIN");
    const int chain = env_chain ? atoi(env_chain) : 2048;
    if (chain > 0 && a.accumulate && split_k == 0 && a.act == 0 && !a.aux) {
      const int want = (a.K + chain - 1) / chain;
      if (pl.splits < want) pl.splits = want;
    }
  }
  return pl;
}

static int launch_plan(const GemmArgs& a, const GemmPlan& pl, cudaStream_t stream) {
  const int M = a.M, N = a.N, K = a.K, mode = pl.mode, bn = pl.bn;
  CAPDEC_REQUIRE(bn == 64 || bn == 128 || bn == 192 || bn == 256, "gemm: bad tile width %d", bn);
  CAPDEC_REQUIRE(mode == 0 || bn >= 128, "gemm: CTA-pair engines need block_n >= 128");
  CAPDEC_REQUIRE(bn != 192 || plan_ok192(mode, a.b_major),
                 "gemm: block_n=192 is not available for the B-sharing quad with an MN-major B");
  CAPDEC_REQUIRE(pl.splits == 1 || (a.accumulate && a.act == 0 && !a.aux), "gemm: split-K needs accumulate=1, act=0, no aux");
  const bool split3 = a.precision != 0;   // 3xTF32: operands split into hi/lo inside the pipeline
  CAPDEC_REQUIRE(!split3 || mode <= 1, "gemm: the 3xTF32 mode runs on the single-CTA and CTA-pair engines only");
  GemmDev p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K;
  p.kb_total = (K + kBlockK - 1) / kBlockK;
  const int tile_m = plan_tile_m(mode);                                                 // rows per cluster tile
  const int pair_m = (mode == 0) ? kBlockM : 2 * kBlockM;                               // rows per MMA (instruction M)
  const int nmul = (mode == 2) ? 2 : 1;
  int splits = pl.splits < 1 ? 1 : pl.splits;
  p.block_n = bn;
  p.splits = splits;
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  p.splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;  // drop empty splits
  p.m_tiles = (M + tile_m - 1) / tile_m;
  p.n_tiles = (N + nmul * bn - 1) / (nmul * bn);
  p.exact = split3 ? 1 : 0;
  p.act = a.act;
  p.has_aux = a.aux ? 1 : 0;
  p.accumulate = a.accumulate ? 1 : 0;
  p.a_mn = a.a_major ? 1 : 0;
  p.b_mn = a.b_major ? 1 : 0;
  p.bias = a.bias;
  p.m_limit = a.m_limit;
  p.mul_act = a.mul_act;
  p.colsum = a.colsum;
  p.k_limit = a.k_limit;
  p.trace = g_trace;
  // "1": plain outputs leave through coalesced st.global from the staging box instead of TMA stores.  Measured SLOWER
  // (qkv 96 vs 80 us: the epilogue warps stall on the store queue while the TMA engine keeps its stores asynchronous), so
  // this stays an A-B switch (profiles/r2_gemm_timeline.md)
  static const char* env_lsu = getenv("CAPDEC_GEMM_LSU_STORE");
  p.c_ptr = a.C; p.aux_ptr = a.aux; p.ldc = (long long)a.ldc;
  p.lsu_store = (!a.accumulate && (N % 4) == 0 && env_lsu && env_lsu[0] == '1') ? 1 : 0;
  // Tile order inside a wave.  n-fastest (0) lets the clusters of a wave share A rows and read every B panel once per row
  // of tiles: right while B (a trunk weight, <= 9.4 MB) lives in L2.  The tied LM head's B is wte, 154 MB > L2: walking n
  // first re-streams it from HBM for each of the 24-40 row tiles (3.7 GB per launch - the GEMM was HBM-bound at 560 TF/s);
  // m-fastest (1) keeps three B panels hot per wave and reads wte once.  CAPDEC_GEMM_RASTER forces either (A-B switch).
  static const char* env_raster = getenv("CAPDEC_GEMM_RASTER");
  const double b_bytes_all = (double)N * (double)K * 4.0, a_bytes_all = (double)((a.m_limit && t_m_hint > 0 && t_m_hint < M) ? t_m_hint : M) * (double)K * 4.0;
  p.raster = env_raster ? atoi(env_raster) : ((b_bytes_all > 48e6 && b_bytes_all > a_bytes_all) ? 1 : 0);
  static const char* env_dbg = getenv("CAPDEC_GEMM_DBG");
  p.dbg = env_dbg ? (uint32_t)atoi(env_dbg) : 0u;
  static const char* env_sched = getenv("CAPDEC_GEMM_SCHED");   // "dynamic" / "static": overrides capdec_gemm_set_schedule
  const bool dynamic = env_sched ? env_sched[0] == 'd' : g_sched_dynamic.load(std::memory_order_relaxed) != 0;
  if (dynamic) {
    static std::atomic<int*> base_dev[kMaxDevices];
    static std::atomic<unsigned> next_slot{0};
    const int dev = current_device();
    int* base = base_dev[dev].load(std::memory_order_acquire);
    if (!base) {
      void* sym = nullptr;
      cudaError_t e = cudaGetSymbolAddress(&sym, g_sched_slots);
      if (e != cudaSuccess) return check_cuda(e, "cudaGetSymbolAddress(g_sched_slots)");
      base = static_cast<int*>(sym);
      base_dev[dev].store(base, std::memory_order_release);
    }
    // one slot per launch; a slot is reused 1024 launches later (or by the replay of the CUDA graph that captured it), long
    // after the launch that owned it has re-armed it
    p.sched = base + 32 * (next_slot.fetch_add(1, std::memory_order_relaxed) % kSchedSlots);
  }

  const int bn_local = (mode == 0) ? bn : bn / 2;
  const int b_bytes = bn_local * kBlockK * 4;
  const int per_stage = (kABytes + b_bytes) * (split3 ? 2 : 1);   // 3xTF32 keeps a lo tile beside every operand tile
  // Shared-memory split between pipeline stages and epilogue staging.  Measured (profiles/r2_gemm_timeline.md): four
  // stages feed the tensor core as well as six, while the epilogue - whose TMA stores queue behind the mainloop's loads in
  // the SM's one TMA engine - needs more than two staging boxes per warp to keep storing.  So: the deepest staging (<= 4
  // boxes per warp and output) that still leaves four stages.
  static const char* env_depth = getenv("CAPDEC_GEMM_CDEPTH");   // bring-up / A-B switch: force 2..4
  auto fixed_for = [&](int depth) {
    return (depth + (a.mul_act ? kMulDepth : (a.aux ? depth : 0))) * kStagingBytes + 256 * 4 +
           (3 * kMaxStages + 4 + 4 * kMulDepth + 2 * kSchedDepth) * 8 + kSchedDepth * 4 + 32;
  };
  int c_depth = 2;
  for (int d = 4; d > 2; --d)
    if ((kSmemLimit - fixed_for(d)) / per_stage >= 4) { c_depth = d; break; }
  if (env_depth && atoi(env_depth) >= 2 && atoi(env_depth) <= 4) c_depth = atoi(env_depth);
  p.c_depth = c_depth;
  const int fixed = fixed_for(c_depth);
  int stages = (kSmemLimit - fixed) / per_stage;
  if (stages > kMaxStages) stages = kMaxStages;
  static const char* env_stages = getenv("CAPDEC_GEMM_STAGES");   // bring-up / A-B switch: cap the pipeline depth
  if (env_stages && atoi(env_stages) >= 2 && atoi(env_stages) < stages) stages = atoi(env_stages);
  CAPDEC_REQUIRE(stages >= 2, "gemm: tile width %d leaves fewer than two pipeline stages in shared memory (3xTF32: use block_n <= 128 on the single-CTA engine)", bn);
  p.stages = stages;
  const int smem_bytes = stages * per_stage + fixed;

  const OperandEnc ea = operand_encoding(a.a_major != 0), eb = operand_encoding(a.b_major != 0);
  p.adesc_hi = ea.desc_hi; p.adesc_lo16 = ea.desc_lo16; p.a_kstep = ea.kstep;
  p.bdesc_hi = eb.desc_hi; p.bdesc_lo16 = eb.desc_lo16; p.b_kstep = eb.kstep;
  // instruction descriptor: D=F32 (bits 4-5 = 1), A/B = TF32 (2) at bits 7-9 / 10-12, majors at 15/16,
  // N>>3 at bits 17-22, M>>4 at bits 24-28
  p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a.a_major ? 1 : 0) << 15) |
            ((uint32_t)(a.b_major ? 1 : 0) << 16) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(pair_m >> 4) << 24);

  // TMA boxes: K-major operands move {32 fp32, rows}; a quad that shares the operand moves half the rows per CTA
  const uint32_t a_rows = (mode == 2) ? kBlockM / 2 : kBlockM;
  const uint32_t b_rows = (mode == 3) ? (uint32_t)bn_local / 2 : (uint32_t)bn_local;
  static const char* env_3d = getenv("CAPDEC_GEMM_3D");   // bring-up switch: 0 disables the 3-D MN-major boxes
  const bool allow3d = !(env_3d && env_3d[0] == '0') && g_mn_swz < 0;
  const uint32_t a_slabs = (mode == 2) ? kBlockM / 64 : kBlockM / 32;
  const uint32_t b_slabs = (mode == 3) ? (uint32_t)bn_local / 64 : (uint32_t)bn_local / 32;
  p.a_3d = (a.a_major && allow3d && ((M % 32) == 0 || a.lda >= (int64_t)((M + 31) / 32) * 32)) ? 1 : 0;
  p.b_3d = (a.b_major && allow3d && ((N % 32) == 0 || a.ldb >= (int64_t)((N + 31) / 32) * 32)) ? 1 : 0;
  int rc;
  {
    const float* pa = a.A;
    const float* pb = a.B;
    if (!a.a_major) rc = make_map(&p.tmA, pa, (uint64_t)K, (uint64_t)M, (uint64_t)a.lda, kBlockK, a_rows, ea.swz);
    else {
      rc = p.a_3d ? make_map_mn3d(&p.tmA, pa, (uint64_t)M, (uint64_t)K, (uint64_t)a.lda, a_slabs, ea.swz) : CAPDEC_ERR_UNSUPPORTED;
      if (rc) { p.a_3d = 0; rc = make_map(&p.tmA, pa, (uint64_t)M, (uint64_t)K, (uint64_t)a.lda, 32, kBlockK, ea.swz); }
    }
    if (rc) return rc;
    if (!a.b_major) rc = make_map(&p.tmB, pb, (uint64_t)K, (uint64_t)N, (uint64_t)a.ldb, kBlockK, b_rows, eb.swz);
    else {
      rc = p.b_3d ? make_map_mn3d(&p.tmB, pb, (uint64_t)N, (uint64_t)K, (uint64_t)a.ldb, b_slabs, eb.swz) : CAPDEC_ERR_UNSUPPORTED;
      if (rc) { p.b_3d = 0; rc = make_map(&p.tmB, pb, (uint64_t)N, (uint64_t)K, (uint64_t)a.ldb, 32, kBlockK, eb.swz); }
    }
    if (rc) return rc;
    // half-width tail tiles (see the kernel): single-CTA / CTA-pair engines, no split-K, and a tile width whose half still
    // is a whole number of 32-column slabs per CTA
    static const char* env_tail = getenv("CAPDEC_GEMM_TAIL");   // "0" disables (bring-up / A-B)
    p.tail_ok = (mode <= 1 && p.splits == 1 && (bn_local / 2) % 32 == 0 && !(env_tail && env_tail[0] == '0')) ? 1 : 0;
    if (p.tail_ok) {
      if (!a.b_major) rc = make_map(&p.tmBh, pb, (uint64_t)K, (uint64_t)N, (uint64_t)a.ldb, kBlockK, b_rows / 2, eb.swz);
      else if (p.b_3d) rc = make_map_mn3d(&p.tmBh, pb, (uint64_t)N, (uint64_t)K, (uint64_t)a.ldb, b_slabs / 2, eb.swz);
      else p.tmBh = p.tmB;                      // 32-wide slab boxes: the same map serves both widths
      if (rc) return rc;
      p.idesc_h = (p.idesc & ~(0x3Fu << 17)) | ((uint32_t)((bn / 2) >> 3) << 17);
    }
  }
  rc = make_map(&p.tmC, a.C, (uint64_t)N, (uint64_t)M, (uint64_t)a.ldc, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B);  // per-warp {32 x 32} boxes
  if (rc) return rc;
  if (a.aux) {
    rc = make_map(&p.tmAux, a.aux, (uint64_t)N, (uint64_t)M, (uint64_t)a.ldc, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  if (a.mul_act) {
    rc = make_map(&p.tmMul, a.mul_in, (uint64_t)N, (uint64_t)M, (uint64_t)a.ldc, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }

  const int total_tiles = p.m_tiles * p.n_tiles * p.splits;
  // epilogue variant (see the kernel's kEpi): 1 plain, 2 gelu forward + derivative output, 3 multiply by a stored derivative
  int epi = 0;
  if (!p.lsu_store) {
    if (a.act == 0 && !a.aux && !a.mul_act && !a.colsum) epi = 1;
    else if (!split3 && a.act == 4 && a.aux && !a.mul_act && !a.colsum) epi = 2;
    else if (!split3 && a.mul_act == 4 && a.act == 0 && !a.aux) epi = 3;
  }
#define CAPDEC_GEMM_LAUNCH0(SPLIT, EPI, THREADS)                                                        \
  do {                                                                                                  \
    cudaLaunchConfig_t cfg0;                                                                            \
    memset(&cfg0, 0, sizeof(cfg0));                                                                     \
    cfg0.gridDim = dim3(grid); cfg0.blockDim = dim3(THREADS); cfg0.dynamicSmemBytes = smem_bytes; cfg0.stream = stream; \
    cudaLaunchAttribute at0;                                                                            \
    at0.id = cudaLaunchAttributeProgrammaticStreamSerialization;                                        \
    at0.val.programmaticStreamSerializationAllowed = 1;                                                 \
    cfg0.attrs = &at0; cfg0.numAttrs = gemm_pdl() ? 1 : 0;                                              \
    cudaError_t e0 = cudaLaunchKernelEx(&cfg0, gemm_tf32_kernel<0, SPLIT, EPI>, p);                     \
    if (e0 != cudaSuccess) return check_cuda(e0, "cudaLaunchKernelEx(gemm_tf32_kernel)");               \
  } while (0)
#define CAPDEC_GEMM_BY_EPI(MODE, CS)                                                                   \
  (epi == 1 ? launch_clustered<MODE, false, 1>(p, CS, total_tiles, smem_bytes, stream)                 \
   : epi == 2 ? launch_clustered<MODE, false, 2>(p, CS, total_tiles, smem_bytes, stream)               \
   : epi == 3 ? launch_clustered<MODE, false, 3>(p, CS, total_tiles, smem_bytes, stream)               \
              : launch_clustered<MODE, false, 0>(p, CS, total_tiles, smem_bytes, stream))
  if (mode == 0) {
    const int grid = total_tiles < num_sms() ? total_tiles : num_sms();
    if (split3) {
      if (epi == 1) CAPDEC_GEMM_LAUNCH0(true, 1, kSplitThreads); else CAPDEC_GEMM_LAUNCH0(true, 0, kSplitThreads);
    } else {
      if (epi == 1) CAPDEC_GEMM_LAUNCH0(false, 1, kThreads);
      else if (epi == 2) CAPDEC_GEMM_LAUNCH0(false, 2, kThreads);
      else if (epi == 3) CAPDEC_GEMM_LAUNCH0(false, 3, kThreads);
      else CAPDEC_GEMM_LAUNCH0(false, 0, kThreads);
    }
  } else if (mode == 1) {
    if (split3) rc = epi == 1 ? launch_clustered<1, true, 1>(p, 2, total_tiles, smem_bytes, stream)
                              : launch_clustered<1, true, 0>(p, 2, total_tiles, smem_bytes, stream);
    else rc = CAPDEC_GEMM_BY_EPI(1, 2);
    if (rc) return rc;
  } else if (mode == 2) {
    rc = CAPDEC_GEMM_BY_EPI(2, 4);
    if (rc) return rc;
  } else {
    rc = CAPDEC_GEMM_BY_EPI(3, 4);
    if (rc) return rc;
  }
#undef CAPDEC_GEMM_LAUNCH0
#undef CAPDEC_GEMM_BY_EPI
  g_launches.fetch_add(1);
  CAPDEC_LAUNCH_CHECK("gemm_tf32_kernel");
  return CAPDEC_OK;
}

// ---- measured plan selection ------------------------------------------------------------------------------------
// While tuning is enabled (capdec_gemm_autotune(1)), the first call for a problem signature times a short list of
// candidate plans on the caller's stream with the caller's operands (so device-side row / reduction limits are the
// real ones) and remembers the fastest; afterwards - and under CUDA-graph capture - the remembered plan is used.
// A tuning call launches the problem several times: accumulate outputs and fused column sums are garbage afterwards,
// so the caller runs tuning on a throw-away step (Trainer.autotune).
struct TuneKey {
  int v[18];
  bool operator<(const TuneKey& o) const { return memcmp(v, o.v, sizeof(v)) < 0; }
};
static std::map<TuneKey, GemmPlan> g_tuned;
static std::mutex g_tune_mu;
static std::atomic<int> g_tune_on{0};

static TuneKey tune_key(const GemmArgs& a, int block_n, int split_k) {
  TuneKey k;
  const int vals[18] = {a.M, a.N, a.K, a.a_major, a.b_major, a.accumulate, a.act, a.aux != nullptr, a.bias != nullptr,
                        a.mul_act, a.colsum != nullptr, a.precision, block_n, split_k, a.m_limit != nullptr,
                        a.k_limit != nullptr, (int)(a.lda % 32 == 0), (int)(a.ldb % 32 == 0)};
  memcpy(k.v, vals, sizeof(vals));
  return k;
}

static int tune_and_launch(const GemmArgs& a, int block_n, int split_k, const GemmPlan& base, cudaStream_t stream) {
  std::vector<GemmPlan> cands;
  cands.push_back(base);
  auto add = [&](int mode, int bn) {
    if (block_n && bn != block_n) return;
    if (bn == 192 && !plan_ok192(mode, a.b_major)) return;
    GemmPlan pl{mode, bn, a.accumulate ? split_k : 1};
    int bn_fixed = bn;
    const int m_live = (a.m_limit && t_m_hint > 0 && t_m_hint < a.M) ? t_m_hint : a.M;
    choose_tiling(m_live, a.N, (a.K + kBlockK - 1) / kBlockK, a.accumulate, plan_units(mode), plan_tile_m(mode), 1,
                  mode == 0 ? 64 : 128, true, bn_fixed, pl.splits);
    pl.bn = bn;
    for (const GemmPlan& c : cands) if (c.mode == pl.mode && c.bn == pl.bn && c.splits == pl.splits) return;
    cands.push_back(pl);
    if (a.accumulate && !split_k && pl.splits > 1) {   // neighbours of the heuristic split factor
      GemmPlan lo = pl, hi = pl;
      lo.splits = pl.splits - 1; hi.splits = pl.splits + 1;
      cands.push_back(lo);
      if ((a.K + kBlockK - 1) / kBlockK / hi.splits >= 4) cands.push_back(hi);
    }
  };
  if (base.mode == 0) {
    add(0, 256); add(0, 128); add(0, 64);
  } else {
    add(1, 256); add(1, 192); add(1, 128);
    if (!a.precision && (a.M > 2 * kBlockM || a.N > 256)) { add(3, 256); add(3, 192); }
  }
  cudaEvent_t e0, e1;
  if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) { cudaGetLastError(); return launch_plan(a, base, stream); }
  GemmPlan best = base;
  float best_ms = 1e30f;
  for (const GemmPlan& c : cands) {
    if (launch_plan(a, c, stream) != CAPDEC_OK) continue;   // warm-up (and legality check)
    const int reps = 4;
    cudaEventRecord(e0, stream);
    bool ok = true;
    for (int r = 0; r < reps && ok; ++r) ok = launch_plan(a, c, stream) == CAPDEC_OK;
    cudaEventRecord(e1, stream);
    if (!ok || cudaEventSynchronize(e1) != cudaSuccess) { cudaGetLastError(); continue; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best_ms) { best_ms = ms; best = c; }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  {
    std::lock_guard<std::mutex> lk(g_tune_mu);
    g_tuned[tune_key(a, block_n, split_k)] = best;
  }
  static const char* env_v = getenv("CAPDEC_GEMM_TUNE_VERBOSE");
  if (env_v && env_v[0] == '1')
    fprintf(stderr, "capdec gemm tune: M=%d N=%d K=%d maj=%d%d acc=%d mul=%d lim=%d%d -> mode %d bn %d splits %d (%.1f us; heuristic mode %d bn %d splits %d)\n",
            a.M, a.N, a.K, a.a_major, a.b_major, a.accumulate, a.mul_act, a.m_limit != nullptr, a.k_limit != nullptr, best.mode,
            best.bn, best.splits, best_ms * 250.f, base.mode, base.bn, base.splits);
  return launch_plan(a, best, stream);
}

}  // namespace capdec

// 1 = measure plans for problems not seen yet (see above), 0 = stop measuring (remembered plans stay in use),
// -1 = forget every remembered plan.  Returns the number of remembered plans.
extern "C" int capdec_gemm_autotune(int enable) {
  std::lock_guard<std::mutex> lk(g_tune_mu);
  if (enable < 0) g_tuned.clear();
  else g_tune_on.store(enable ? 1 : 0);
  return (int)g_tuned.size();
}

// The heuristic plan for a problem, without launching anything (host arithmetic only; usable without a GPU, where the SM
// count defaults to 148).  Returns engine | tile width << 8 | split-K << 20, or a negative error code.
extern "C" int capdec_gemm_plan_query(int M, int N, int K, int b_major, int accumulate, int block_n, int split_k,
                                      int row_limited) {
  CAPDEC_REQUIRE(M > 0 && N > 0 && K > 0, "gemm_plan_query: bad shape M=%d N=%d K=%d", M, N, K);
  CAPDEC_REQUIRE(block_n == 0 || block_n == 64 || block_n == 128 || block_n == 192 || block_n == 256, "gemm_plan_query: bad block_n");
  GemmArgs a;
  memset(&a, 0, sizeof(a));
  a.M = M; a.N = N; a.K = K; a.b_major = b_major ? 1 : 0; a.accumulate = accumulate ? 1 : 0;
  static const int32_t dummy_limit = 0;
  a.m_limit = row_limited ? &dummy_limit : nullptr;   // only its presence matters to the planner (never dereferenced)
  const GemmPlan pl = heuristic_plan(a, block_n, split_k, -1);
  return pl.mode | (pl.bn << 8) | (pl.splits << 20);
}

// common entry of capdec_gemm_tf32_ex / capdec_gemm_tf32_mul: validation, plan lookup / measurement, launch
static int gemm_entry(const float* A, int a_major, int64_t lda, const float* B, int b_major, int64_t ldb, float* C,
                      int64_t ldc, int M, int N, int K, const float* bias, int act, float* aux, int accumulate, int precision,
                      const float* a_lo, const float* b_lo, int block_n, int split_k, const int32_t* m_limit_dev,
                      const int32_t* k_limit_dev, const float* mul_in, int mul_act, float* colsum, cudaStream_t stream) {
  CAPDEC_REQUIRE(A && B && C, "gemm: null operand");
  CAPDEC_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: bad shape M=%d N=%d K=%d", M, N, K);
  CAPDEC_REQUIRE((lda % 4) == 0 && (ldb % 4) == 0 && (ldc % 4) == 0, "gemm: leading dims must be multiples of 4 (16 B TMA pitch): lda=%lld ldb=%lld ldc=%lld", (long long)lda, (long long)ldb, (long long)ldc);
  CAPDEC_REQUIRE(((uintptr_t)A % 16) == 0 && ((uintptr_t)B % 16) == 0 && ((uintptr_t)C % 16) == 0, "gemm: operands must be 16-byte aligned");
  CAPDEC_REQUIRE(lda >= (a_major ? M : K) && ldb >= (b_major ? N : K) && ldc >= N, "gemm: leading dim smaller than extent");
  CAPDEC_REQUIRE(precision == 0 || precision == 1, "gemm: precision must be 0 (1xTF32) or 1 (3xTF32, split in the pipeline)");
  (void)a_lo; (void)b_lo;   // kept in the signature for ABI stability: the hi/lo split happens inside the kernel
  CAPDEC_REQUIRE(block_n == 0 || block_n == 64 || block_n == 128 || block_n == 192 || block_n == 256, "gemm: block_n must be 0/64/128/192/256");
  CAPDEC_REQUIRE(act >= 0 && act <= 4, "gemm: bad act %d", act);
  CAPDEC_REQUIRE(!aux || ((uintptr_t)aux % 16) == 0, "gemm: aux must be 16-byte aligned");

  static std::atomic<bool> attr_set_dev[kMaxDevices];   // cudaFuncSetAttribute is per device
  std::atomic<bool>& attr_set = attr_set_dev[current_device()];
  if (!attr_set.load(std::memory_order_acquire)) {
    cudaError_t e = cudaSuccess;
    auto opt_in = [&](const void* fn) { if (e == cudaSuccess) e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit); };
#define CAPDEC_OPT_IN4(MODE) opt_in((const void*)gemm_tf32_kernel<MODE, false, 0>); opt_in((const void*)gemm_tf32_kernel<MODE, false, 1>); \
                             opt_in((const void*)gemm_tf32_kernel<MODE, false, 2>); opt_in((const void*)gemm_tf32_kernel<MODE, false, 3>)
    CAPDEC_OPT_IN4(0); CAPDEC_OPT_IN4(1); CAPDEC_OPT_IN4(2); CAPDEC_OPT_IN4(3);
#undef CAPDEC_OPT_IN4
    opt_in((const void*)gemm_tf32_kernel<0, true, 0>); opt_in((const void*)gemm_tf32_kernel<0, true, 1>);
    opt_in((const void*)gemm_tf32_kernel<1, true, 0>); opt_in((const void*)gemm_tf32_kernel<1, true, 1>);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(gemm_tf32_kernel)");
    attr_set.store(true, std::memory_order_release);
  }

  GemmArgs a;
  a.A = A; a.B = B; a.C = C; a.a_major = a_major ? 1 : 0; a.b_major = b_major ? 1 : 0; a.lda = lda; a.ldb = ldb; a.ldc = ldc;
  a.M = M; a.N = N; a.K = K; a.bias = bias; a.act = act; a.aux = aux; a.accumulate = accumulate ? 1 : 0; a.precision = precision;
  a.a_lo = a_lo; a.b_lo = b_lo; a.m_limit = m_limit_dev; a.k_limit = k_limit_dev;
  a.mul_in = mul_in; a.mul_act = mul_act; a.colsum = colsum;

  // ---- engine selection: 0 single CTA, 1 CTA pair, 2 quad sharing A (pairs side by side in N), 3 quad sharing B ----
  static const char* env_pair = getenv("CAPDEC_GEMM_PAIR");  // bring-up switch: 0 = cta_group::1 only, 1 = pairs only
  static const char* env_mode = getenv("CAPDEC_GEMM_MODE");  // bring-up switch: force engine 0..3 where legal
  const int forced = g_force_pair >= 0 ? g_force_pair : (env_mode ? atoi(env_mode) : (env_pair ? atoi(env_pair) : -1));
  GemmPlan pl = heuristic_plan(a, block_n, split_k, forced);
  if (forced < 0) {
    bool have = false;
    {
      std::lock_guard<std::mutex> lk(g_tune_mu);
      if (!g_tuned.empty() || g_tune_on.load()) {
        auto it = g_tuned.find(tune_key(a, block_n, split_k));
        if (it != g_tuned.end()) { pl = it->second; have = true; }
      }
    }
    if (!have && g_tune_on.load()) {
      cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
      if (cudaStreamIsCapturing(stream, &cs) == cudaSuccess && cs == cudaStreamCaptureStatusNone)
        return tune_and_launch(a, block_n, split_k, pl, stream);
      cudaGetLastError();
    }
  }
  return launch_plan(a, pl, stream);
}

extern "C" int capdec_gemm_tf32_ex(const float* A, int a_major, int64_t lda, const float* B, int b_major, int64_t ldb,
                                   float* C, int64_t ldc, int M, int N, int K, const float* bias, int act, float* aux,
                                   int accumulate, int precision, const float* a_lo, const float* b_lo, int block_n,
                                   int split_k, const int32_t* m_limit_dev, const int32_t* k_limit_dev,
                                   capdec_stream_t stream_) {
  return gemm_entry(A, a_major, lda, B, b_major, ldb, C, ldc, M, N, K, bias, act, aux, accumulate, precision, a_lo, b_lo,
                    block_n, split_k, m_limit_dev, k_limit_dev, nullptr, 0, nullptr, reinterpret_cast<cudaStream_t>(stream_));
}

// C = (A . B^T) * act'(mul_in), colsum[n] += sum_m C[m,n]: the dgrad GEMM that feeds an activation's backward, with the
// bias gradient of the layer below fused (HF:modeling_gpt2.py:238-243 c_fc -> gelu_new; train.py:106-118 tanh; :121 relu)
extern "C" int capdec_gemm_tf32_mul_ex(const float* A, int a_major, int64_t lda, const float* B, int b_major, int64_t ldb,
                                       float* C, int64_t ldc, int M, int N, int K, const float* mul_in, int mul_act,
                                       float* colsum, int block_n, const int32_t* m_limit_dev, int precision,
                                       capdec_stream_t stream_) {
  CAPDEC_REQUIRE(mul_in && mul_act >= 1 && mul_act <= 4, "gemm_mul: bad epilogue input");
  CAPDEC_REQUIRE(((uintptr_t)mul_in % 16) == 0, "gemm_mul: mul_in must be 16-byte aligned");
  return gemm_entry(A, a_major, lda, B, b_major, ldb, C, ldc, M, N, K, nullptr, 0, nullptr, 0, precision, nullptr, nullptr,
                    block_n, 1, m_limit_dev, nullptr, mul_in, mul_act, colsum, reinterpret_cast<cudaStream_t>(stream_));
}

extern "C" int capdec_gemm_tf32_mul(const float* A, int a_major, int64_t lda, const float* B, int b_major, int64_t ldb,
                                    float* C, int64_t ldc, int M, int N, int K, const float* mul_in, int mul_act,
                                    float* colsum, int block_n, const int32_t* m_limit_dev, capdec_stream_t stream_) {
  return capdec_gemm_tf32_mul_ex(A, a_major, lda, B, b_major, ldb, C, ldc, M, N, K, mul_in, mul_act, colsum, block_n,
                                 m_limit_dev, 0, stream_);
}
