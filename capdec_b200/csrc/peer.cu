// capdec_b200 — data-parallel optimizer step over NVLink peer memory (train.py:352-354 under N ranks).
//
// ONE kernel per step and rank replaces  ncclReduceScatter -> AdamW(1/N slice) -> ncclAllGather :
//   every rank owns the parameters [lo, lo + n) of the flat buffer.  Its kernel LOADS the gradients of that slice as every
//   rank computed them (summed in rank order, so the result does not depend on who owns the slice), runs the HF-AdamW
//   update with the moments it alone keeps, and STORES the new parameters into the parameter buffers of ALL ranks (NVLink
//   writes).  Where the gradients are read from is the caller's choice (g_slices): straight from the peers' gradient
//   buffers (NVLink reads: measured 365 GB/s per direction at N = 2, loads over the link are latency-bound), or - the
//   default of capdec_b200/trainer.py - from a local staging area into which every rank's COPY ENGINES pushed its share of
//   each gradient bucket while the backward pass was still running (capdec_copy_async under CUDA-graph capture = a memcpy
//   node; no SM is involved, so unlike an NCCL kernel the transfer takes nothing from the persistent GEMM CTAs).  Then the
//   only exposed NVLink traffic of a step is the parameter all-gather.  The cross-rank ordering (all gradients final before, all parameters
//   landed after) is two 16-byte NCCL all-reduces issued by the host code around the launch (capdec_b200/trainer.py); the
//   first of them is the global token count the update divides by, which the step needs anyway.
// The peer pointers come from CUDA IPC (one process per GPU): capdec_peer_export / capdec_peer_open below.
#include <string.h>

#include "../../include/capdec_b200.h"
#include "common.cuh"

namespace capdec {

constexpr int kMaxPeers = 8;   // one NVSwitch domain of this node

struct PeerSet {
  const float4* g[kMaxPeers];  // where the gradients of THIS rank's slice, as computed by rank 0..world-1, can be read
  float4* p[kMaxPeers];        // parameter buffers of rank 0..world-1 (element 0 = first trainable parameter)
};

// weak 128-bit load that never allocates in L1: peer lines are L1-cacheable but bypass the local L2, and the same
// addresses carry new gradients every step
__device__ __forceinline__ float4 ld_peer(const float4* p) {
  float4 r;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p)
               : "memory");
  return r;
}

template <int kWorld>
__global__ void __launch_bounds__(256) adamw_peer_kernel(const PeerSet ps, int rank, int64_t lo4, int64_t n4,
                                                        float4* __restrict__ m, float4* __restrict__ v,
                                                        const float* __restrict__ lr_dev, const float* __restrict__ t_dev,
                                                        float b1, float b2, float eps, float wd,
                                                        const float* __restrict__ denom_dev) {
  const float lr = *lr_dev, t = *t_dev;
  const float gs = denom_dev ? 1.0f / *denom_dev : 1.0f;
  const float bc1 = 1.0f - powf(b1, t), bc2 = 1.0f - powf(b2, t);
  const float step = lr * sqrtf(bc2) / bc1;
  const float decay = 1.0f - lr * wd;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = lo4 + i;
    float4 gr[kWorld];
#pragma unroll
    for (int r = 0; r < kWorld; ++r) gr[r] = ld_peer(ps.g[r] + i);   // kWorld independent loads in flight per thread
    float4 pp = ps.p[rank][e], mm = m[i], vv = v[i];
    float4 gg = gr[0];
#pragma unroll
    for (int r = 1; r < kWorld; ++r) { gg.x += gr[r].x; gg.y += gr[r].y; gg.z += gr[r].z; gg.w += gr[r].w; }
    gg.x *= gs; gg.y *= gs; gg.z *= gs; gg.w *= gs;
    mm.x = b1 * mm.x + (1.f - b1) * gg.x; mm.y = b1 * mm.y + (1.f - b1) * gg.y;
    mm.z = b1 * mm.z + (1.f - b1) * gg.z; mm.w = b1 * mm.w + (1.f - b1) * gg.w;
    vv.x = b2 * vv.x + (1.f - b2) * gg.x * gg.x; vv.y = b2 * vv.y + (1.f - b2) * gg.y * gg.y;
    vv.z = b2 * vv.z + (1.f - b2) * gg.z * gg.z; vv.w = b2 * vv.w + (1.f - b2) * gg.w * gg.w;
    pp.x = (pp.x - step * mm.x / (sqrtf(vv.x) + eps)) * decay; pp.y = (pp.y - step * mm.y / (sqrtf(vv.y) + eps)) * decay;
    pp.z = (pp.z - step * mm.z / (sqrtf(vv.z) + eps)) * decay; pp.w = (pp.w - step * mm.w / (sqrtf(vv.w) + eps)) * decay;
    m[i] = mm; v[i] = vv;
#pragma unroll
    for (int r = 0; r < kWorld; ++r) st_stream(ps.p[r] + e, pp);     // the all-gather: every replica gets the same bits
  }
  __threadfence_system();   // the parameter stores have left for the peers before this thread counts as finished
}

typedef CUresult (*GetRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
static GetRangeFn get_range_fn() {
  static GetRangeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<GetRangeFn>(sym);
    else
      cudaGetLastError();
  }
  return fn;
}

}  // namespace capdec

using namespace capdec;

extern "C" int capdec_peer_export(const void* ptr, void* handle64, int64_t* offset_out) {
  CAPDEC_REQUIRE(ptr && handle64 && offset_out, "peer_export: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
  GetRangeFn range = get_range_fn();
  CAPDEC_REQUIRE(range, "peer_export: cuMemGetAddressRange entry point unavailable");
  CUdeviceptr base = 0;
  size_t size = 0;
  if (range(&base, &size, (CUdeviceptr)(uintptr_t)ptr) != CUDA_SUCCESS) {
    set_last_error("peer_export: %p is not inside a device allocation", ptr);
    return CAPDEC_ERR_CUDA;
  }
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, reinterpret_cast<void*>((uintptr_t)base));
  if (e != cudaSuccess) return check_cuda(e, "cudaIpcGetMemHandle (allocations of an expandable-segments allocator cannot be exported)");
  memcpy(handle64, &h, 64);
  *offset_out = (int64_t)((uintptr_t)ptr - (uintptr_t)base);
  return CAPDEC_OK;
}

extern "C" int capdec_peer_open(const void* handle64, int64_t offset, void** ptr_out) {
  CAPDEC_REQUIRE(handle64 && ptr_out && offset >= 0, "peer_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* base = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return check_cuda(e, "cudaIpcOpenMemHandle");
  *ptr_out = static_cast<char*>(base) + offset;
  return CAPDEC_OK;
}

extern "C" int capdec_peer_close(void* ptr, int64_t offset) {
  CAPDEC_REQUIRE(ptr && offset >= 0, "peer_close: bad arguments");
  return check_cuda(cudaIpcCloseMemHandle(static_cast<char*>(ptr) - offset), "cudaIpcCloseMemHandle");
}

extern "C" int capdec_copy_async(void* dst, const void* src, int64_t bytes, capdec_stream_t stream_) {
  CAPDEC_REQUIRE(dst && src && bytes >= 0, "copy_async: bad arguments");
  if (bytes == 0) return CAPDEC_OK;
  return check_cuda(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, reinterpret_cast<cudaStream_t>(stream_)),
                    "cudaMemcpyAsync");
}

extern "C" int capdec_adamw_peer_step(void* const* g_slices, void* const* p_peers, int world, int rank, int64_t lo,
                                      int64_t n, float* m, float* v, const float* lr_dev, const float* t_dev,
                                      float beta1, float beta2, float eps, float weight_decay,
                                      const float* grad_denom_dev, capdec_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CAPDEC_REQUIRE(g_slices && p_peers && m && v && lr_dev && t_dev, "adamw_peer_step: null argument");
  CAPDEC_REQUIRE(world >= 2 && world <= kMaxPeers && rank >= 0 && rank < world, "adamw_peer_step: world=%d rank=%d (2..%d ranks)",
                 world, rank, kMaxPeers);
  CAPDEC_REQUIRE(lo >= 0 && n > 0 && lo % 4 == 0 && n % 4 == 0, "adamw_peer_step: lo and n must be multiples of 4");
  PeerSet ps;
  for (int r = 0; r < kMaxPeers; ++r) {
    const int s = r < world ? r : 0;
    CAPDEC_REQUIRE(g_slices[s] && p_peers[s], "adamw_peer_step: null peer pointer for rank %d", s);
    CAPDEC_REQUIRE(((uintptr_t)g_slices[s] & 15u) == 0 && ((uintptr_t)p_peers[s] & 15u) == 0, "adamw_peer_step: peer buffers must be 16-byte aligned");
    ps.g[r] = static_cast<const float4*>(g_slices[s]);
    ps.p[r] = static_cast<float4*>(p_peers[s]);
  }
  const int64_t n4 = n / 4;
  int64_t blocks = (n4 + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 8;
  const int grid = (int)(blocks < cap ? blocks : cap);
#define CAPDEC_PEER_LAUNCH(W)                                                                                      \
  adamw_peer_kernel<W><<<grid, 256, 0, stream>>>(ps, rank, lo / 4, n4, reinterpret_cast<float4*>(m),               \
                                                 reinterpret_cast<float4*>(v), lr_dev, t_dev, beta1, beta2, eps,   \
                                                 weight_decay, grad_denom_dev)
  switch (world) {
    case 2: CAPDEC_PEER_LAUNCH(2); break;
    case 3: CAPDEC_PEER_LAUNCH(3); break;
    case 4: CAPDEC_PEER_LAUNCH(4); break;
    case 5: CAPDEC_PEER_LAUNCH(5); break;
    case 6: CAPDEC_PEER_LAUNCH(6); break;
    case 7: CAPDEC_PEER_LAUNCH(7); break;
    default: CAPDEC_PEER_LAUNCH(8); break;
  }
#undef CAPDEC_PEER_LAUNCH
  g_launches.fetch_add(1);
  CAPDEC_LAUNCH_CHECK("adamw_peer_kernel");
  return CAPDEC_OK;
}
