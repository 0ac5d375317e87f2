// Data-parallel exchange over NVLink peer memory: the ranks of one node map each other's buffers (CUDA VMM /
// symmetric memory, set up by the host layer) and these kernels read and write them directly, so the gradient reduce-scatter
// and the weight all-gather are not separate collectives any more:
//
//   ltg_adam_peer     TF-Adam over this rank's row shard; the gradient of every row is the sum of the N ranks' rows, loaded
//                     straight from their gradient buffers (reduce-scatter fused into the consumer), and the updated bf16 row is
//                     stored into every rank's weight shadow (all-gather fused into the producer).
//   ltg_peer_push     copy a local block into the same slot of every rank's buffer (all-gather of small activations)
//   ltg_peer_reduce   out = sum over ranks of their buffer (small gradient buckets; out must not be a peer-visible buffer)
//   ltg_peer_allreduce_small  in-place all-reduce of <= 1024 floats in one single-CTA kernel (two internal barriers)
//   ltg_peer_barrier  all ranks have executed everything ordered before it on their stream
//
// ncu/bench showed why: at 2 GPUs the NCCL-collective version of the step spent 0.27 ms per step in eight latency-bound
// collectives (1.3 MB all-reduce = 40 us) on top of 0.56 ms of compute; a flag round trip over NVLink is ~2 us.
//
// Synchronisation: signal words live in a peer-mapped uint32 array `pads` ([slots][world] per rank). Barrier number e of a slot
// writes e into word [slot][my rank] of every rank and waits until its own [slot][r] >= e for all r; e comes from a local device
// counter, so a captured CUDA graph replays correctly. Each stream that issues barriers uses its own slot.
#include "ltg_common.cuh"
#include "../../include/ltgan.h"

namespace {

constexpr int PEER_MAX = 8;

struct PeerPtrs { void* p[PEER_MAX]; };

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer_f4(const float* p) {   // peer memory: bypass L1, system-coherent
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_peer_f(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}

// NVLS multicast (NVSwitch): one store to the multicast address lands in every rank's buffer (egress 1x instead of world-1 x),
// one ld_reduce returns the sum of every rank's value, added inside the switch (ingress 1x).
__device__ __forceinline__ float4 mm_ld_reduce_f4(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ void mm_st_b64(void* mc, uint2 v) {
  asm volatile("multimem.st.relaxed.sys.global.v2.f32 [%0], {%1,%2};" ::"l"(mc), "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)) : "memory");
}
__device__ __forceinline__ void mm_st_b128(void* mc, uint4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)),
               "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w)) : "memory");
}

constexpr unsigned long long PEER_SPIN_LIMIT_NS = 20ull * 1000ull * 1000ull * 1000ull;   // 20 s

// Executed by the first warp of ONE CTA. `e` is the barrier number (the same on every rank).
__device__ __forceinline__ void peer_barrier_warp(const PeerPtrs& pads, int rank, int world, int slot, uint32_t e) {
  const int lane = threadIdx.x & 31;
  __threadfence_system();
  if (lane < world) {
    st_release_sys(reinterpret_cast<uint32_t*>(pads.p[lane]) + slot * PEER_MAX + rank, e);
    const uint32_t* mine = reinterpret_cast<const uint32_t*>(pads.p[rank]) + slot * PEER_MAX + lane;
    // A peer that died (exception on its host, lost process) never arrives: after PEER_SPIN_LIMIT_NS the kernel traps, so the hang
    // surfaces as a CUDA error on this rank instead of a captured graph that spins forever.
    unsigned long long t0 = 0;
    uint32_t spins = 0;
    while ((int32_t)(ld_acquire_sys(mine) - e) < 0) {
      __nanosleep(20);
      if ((++spins & 0x3FFu) == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > PEER_SPIN_LIMIT_NS) __trap();
      }
    }
  }
  __syncwarp();
  __threadfence_system();
}

__global__ void peer_barrier_kernel(PeerPtrs pads, int rank, int world, int slot, uint32_t* epochs) {
  const uint32_t e = epochs[slot] + 1;
  __syncwarp();
  peer_barrier_warp(pads, rank, world, slot, e);
  if (threadIdx.x == 0) epochs[slot] = e;
}

__global__ void __launch_bounds__(1024)
peer_allreduce_small_kernel(PeerPtrs bufs, int64_t offset, int count, PeerPtrs pads, int rank, int world, int slot, uint32_t* epochs) {
  const uint32_t e = epochs[slot] + 1;
  __syncthreads();
  if (threadIdx.x < 32) peer_barrier_warp(pads, rank, world, slot, e);          // every rank's values are final
  __syncthreads();
  float s = 0.f;
  if ((int)threadIdx.x < count)
    for (int r = 0; r < world; ++r) s += ld_peer_f(reinterpret_cast<const float*>(bufs.p[r]) + offset + threadIdx.x);   // rank order: same sum everywhere
  __syncthreads();
  if (threadIdx.x < 32) peer_barrier_warp(pads, rank, world, slot, e + 1);      // every rank has read them
  __syncthreads();
  if ((int)threadIdx.x < count) reinterpret_cast<float*>(bufs.p[rank])[offset + threadIdx.x] = s;
  if (threadIdx.x == 0) epochs[slot] = e + 1;
}

__global__ void __launch_bounds__(256)
peer_reduce_kernel(PeerPtrs bufs, const float* bufs_mc, int64_t offset, int64_t n, int world, float* __restrict__ out) {
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bufs_mc != nullptr) {
      s = mm_ld_reduce_f4(bufs_mc + offset + 4 * i);   // one load, summed inside the NVSwitch
    } else {
      for (int r = 0; r < world; ++r) {
        const float4 o = ld_peer_f4(reinterpret_cast<const float*>(bufs.p[r]) + offset + 4 * i);
        s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w;
      }
    }
    *reinterpret_cast<float4*>(out + 4 * i) = s;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = (n4 << 2) + threadIdx.x;
    float s = 0.f;
    for (int r = 0; r < world; ++r) s += ld_peer_f(reinterpret_cast<const float*>(bufs.p[r]) + offset + i);
    out[i] = s;
  }
}

__global__ void __launch_bounds__(256)
peer_push_kernel(const uint4* __restrict__ src, int64_t n16, PeerPtrs dst, uint4* dst_mc, int64_t dst_off16, int world) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
    const uint4 v = __ldg(src + i);
    if (dst_mc != nullptr) {
      mm_st_b128(dst_mc + dst_off16 + i, v);
    } else {
      for (int r = 0; r < world; ++r) reinterpret_cast<uint4*>(dst.p[r])[dst_off16 + i] = v;
    }
  }
}

// Same update as adam_kernel (adam_kernels.cu); g = sum over ranks of their gradient rows, bf16 result stored to every rank.
__global__ void __launch_bounds__(256)
adam_peer_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, PeerPtrs grads, const float* grads_mc, PeerPtrs shadows,
                 __nv_bfloat16* shadows_mc, int64_t offset, int64_t n, int world, float lr_t, const float* __restrict__ scal, float b1,
                 float b2, float eps) {
  if (lr_t < 0.f) lr_t = scal[LTG_S_LR_T];
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = ld_stream_f4(p + 4 * i), mm = ld_stream_f4(m + 4 * i), vv = ld_stream_f4(v + 4 * i);
    float4 gg;
    if (grads_mc != nullptr) {
      gg = mm_ld_reduce_f4(grads_mc + offset + 4 * i);   // summed inside the NVSwitch; each element is reduced by its owner only
    } else {
      float4 g[PEER_MAX];
#pragma unroll
      for (int r = 0; r < PEER_MAX; ++r)
        if (r < world) g[r] = ld_peer_f4(reinterpret_cast<const float*>(grads.p[r]) + offset + 4 * i);
      gg = g[0];
#pragma unroll
      for (int r = 1; r < PEER_MAX; ++r)
        if (r < world) { gg.x += g[r].x; gg.y += g[r].y; gg.z += g[r].z; gg.w += g[r].w; }   // rank order on every rank
    }
    ltg_adam4(pp, mm, vv, gg, lr_t, b1, b2, eps);
    st_stream_f4(p + 4 * i, pp); st_stream_f4(m + 4 * i, mm); st_stream_f4(v + 4 * i, vv);
    uint2 s; s.x = pack_bf16x2(pp.x, pp.y); s.y = pack_bf16x2(pp.z, pp.w);
    if (shadows_mc != nullptr) {
      mm_st_b64(shadows_mc + offset + 4 * i, s);
    } else {
#pragma unroll
      for (int r = 0; r < PEER_MAX; ++r)
        if (r < world) *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(shadows.p[r]) + offset + 4 * i) = s;
    }
  }
}

// enc_adam_kernel (adam_kernels.cu) over this rank's item rows [item0, item0 + n_items): the compact gradient G already holds the
// global batch (activation exchange), so only the all-gather is fused in: the bf16 row goes to every rank's encoder shadow.
__global__ void __launch_bounds__(256)
enc_adam_peer_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, PeerPtrs shadows, __nv_bfloat16* shadows_mc,
                     int64_t offset, int n_items,
                     const int32_t* __restrict__ slot_of_item, const float* __restrict__ G, int world, float lr_t,
                     const float* __restrict__ scal, float b1, float b2, float eps) {
  constexpr int H4 = LTG_H / 4;
  if (lr_t < 0.f) lr_t = scal[LTG_S_LR_T];
  const int64_t n4 = (int64_t)n_items * H4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const int item = (int)(i / H4);
    const int c4 = (int)(i - (int64_t)item * H4);
    float4 pp = ld_stream_f4(p + 4 * i), mm = ld_stream_f4(m + 4 * i), vv = ld_stream_f4(v + 4 * i);
    const int slot = __ldg(slot_of_item + item);
    float4 gg = make_float4(0.f, 0.f, 0.f, 0.f);
    if (slot >= 0) gg = __ldg(reinterpret_cast<const float4*>(G + (size_t)slot * LTG_H) + c4);
    ltg_adam4(pp, mm, vv, gg, lr_t, b1, b2, eps);
    st_stream_f4(p + 4 * i, pp); st_stream_f4(m + 4 * i, mm); st_stream_f4(v + 4 * i, vv);
    uint2 s; s.x = pack_bf16x2(pp.x, pp.y); s.y = pack_bf16x2(pp.z, pp.w);
    if (shadows_mc != nullptr) {
      mm_st_b64(shadows_mc + offset + 4 * i, s);
    } else {
#pragma unroll
      for (int r = 0; r < PEER_MAX; ++r)
        if (r < world) *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(shadows.p[r]) + offset + 4 * i) = s;
    }
  }
}

int load_ptrs(PeerPtrs* dst, void* const* src, int world) {
  if (src == nullptr || world < 1 || world > PEER_MAX) { ltg_set_last_error("peer table: 1 <= world <= 8 pointers required", __FILE__, __LINE__); return LTG_ERR_ARG; }
  for (int r = 0; r < PEER_MAX; ++r) dst->p[r] = r < world ? src[r] : nullptr;
  for (int r = 0; r < world; ++r)
    if (src[r] == nullptr) { ltg_set_last_error("peer table holds a null pointer", __FILE__, __LINE__); return LTG_ERR_ARG; }
  return LTG_OK;
}

inline unsigned stream_grid(int64_t work_items) {
  int64_t b = (work_items + 255) / 256;
  const int64_t cap = 148 * 8;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

// Tables are HOST arrays of `world` device pointers (peer-mapped); they are passed to the kernels by value.
extern "C" int ltg_peer_barrier(void* const* pads, int rank, int world, int slot, uint32_t* epochs, void* stream) {
  PeerPtrs pp;
  int rc = load_ptrs(&pp, pads, world);
  if (rc) return rc;
  LTG_REQUIRE(epochs != nullptr && rank >= 0 && rank < world && slot >= 0 && slot < LTG_PEER_SLOTS);
  peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(pp, rank, world, slot, epochs);
  LTG_CHECK_LAUNCH();
  return LTG_OK;
}

extern "C" int ltg_peer_allreduce_small(void* const* bufs, int64_t offset, int count, void* const* pads, int rank, int world, int slot,
                                        uint32_t* epochs, void* stream) {
  PeerPtrs pb, pp;
  int rc = load_ptrs(&pb, bufs, world);
  if (rc) return rc;
  rc = load_ptrs(&pp, pads, world);
  if (rc) return rc;
  LTG_REQUIRE(epochs != nullptr && rank >= 0 && rank < world && slot >= 0 && slot < LTG_PEER_SLOTS && count >= 0 && count <= 1024 && offset >= 0);
  if (count == 0) return LTG_OK;
  peer_allreduce_small_kernel<<<1, ((count + 31) / 32) * 32, 0, (cudaStream_t)stream>>>(pb, offset, count, pp, rank, world, slot, epochs);
  LTG_CHECK_LAUNCH();
  return LTG_OK;
}

extern "C" int ltg_peer_reduce(void* const* bufs, const float* bufs_mc, int64_t offset, int64_t n, int world, float* out, void* stream) {
  PeerPtrs pb;
  int rc = load_ptrs(&pb, bufs, world);
  if (rc) return rc;
  LTG_REQUIRE(out != nullptr && offset >= 0 && offset % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0);
  for (int r = 0; r < world; ++r) LTG_REQUIRE((reinterpret_cast<uintptr_t>(bufs[r]) & 15) == 0);
  if (n <= 0) return LTG_OK;
  LTG_REQUIRE((reinterpret_cast<uintptr_t>(bufs_mc) & 15) == 0);
  peer_reduce_kernel<<<stream_grid(n >> 2), 256, 0, (cudaStream_t)stream>>>(pb, bufs_mc, offset, n, world, out);
  LTG_CHECK_LAUNCH();
  return LTG_OK;
}

extern "C" int ltg_peer_push(const void* src, int64_t bytes, void* const* dst, void* dst_mc, int64_t dst_offset_bytes, int world, void* stream) {
  PeerPtrs pd;
  int rc = load_ptrs(&pd, dst, world);
  if (rc) return rc;
  LTG_REQUIRE(src != nullptr && bytes >= 0 && bytes % 16 == 0 && dst_offset_bytes >= 0 && dst_offset_bytes % 16 == 0);
  LTG_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst_mc) & 15) == 0);
  for (int r = 0; r < world; ++r) LTG_REQUIRE((reinterpret_cast<uintptr_t>(dst[r]) & 15) == 0);
  if (bytes == 0) return LTG_OK;
  peer_push_kernel<<<stream_grid(bytes >> 4), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(src), bytes >> 4, pd,
                                                                              reinterpret_cast<uint4*>(dst_mc), dst_offset_bytes >> 4, world);
  LTG_CHECK_LAUNCH();
  return LTG_OK;
}

extern "C" int ltg_adam_peer(float* p, float* m, float* v, void* const* grads, const float* grads_mc, void* const* shadows_bf16,
                             void* shadows_mc, int64_t offset, int64_t n, int world, float lr_t, const float* scal, float beta1,
                             float beta2, float eps, void* stream) {
  PeerPtrs pg, ps;
  int rc = load_ptrs(&pg, grads, world);
  if (rc) return rc;
  rc = load_ptrs(&ps, shadows_bf16, world);
  if (rc) return rc;
  LTG_REQUIRE(p && m && v && offset >= 0 && offset % 4 == 0 && n % 4 == 0);
  LTG_REQUIRE(lr_t >= 0.f || scal != nullptr);
  LTG_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0);
  for (int r = 0; r < world; ++r)
    LTG_REQUIRE((reinterpret_cast<uintptr_t>(grads[r]) & 15) == 0 && (reinterpret_cast<uintptr_t>(shadows_bf16[r]) & 7) == 0);
  if (n <= 0) return LTG_OK;
  LTG_REQUIRE((reinterpret_cast<uintptr_t>(grads_mc) & 15) == 0 && (reinterpret_cast<uintptr_t>(shadows_mc) & 7) == 0);
  adam_peer_kernel<<<stream_grid(n >> 2), 256, 0, (cudaStream_t)stream>>>(p, m, v, pg, grads_mc, ps, reinterpret_cast<__nv_bfloat16*>(shadows_mc), offset, n,
                                                                          world, lr_t, scal, beta1, beta2, eps);
  LTG_CHECK_LAUNCH();
  return LTG_OK;
}

extern "C" int ltg_enc_adam_peer(float* p, float* m, float* v, void* const* shadows_bf16, void* shadows_mc, int64_t offset, int n_items,
                                 const int32_t* slot_of_item, const float* G, int world, float lr_t, const float* scal, float beta1,
                                 float beta2, float eps, void* stream) {
  PeerPtrs ps;
  int rc = load_ptrs(&ps, shadows_bf16, world);
  if (rc) return rc;
  LTG_REQUIRE(p && m && v && slot_of_item && G && offset >= 0 && offset % 4 == 0);
  LTG_REQUIRE(lr_t >= 0.f || scal != nullptr);
  for (int r = 0; r < world; ++r) LTG_REQUIRE((reinterpret_cast<uintptr_t>(shadows_bf16[r]) & 7) == 0);
  if (n_items <= 0) return LTG_OK;
  enc_adam_peer_kernel<<<stream_grid((int64_t)n_items * (LTG_H / 4)), 256, 0, (cudaStream_t)stream>>>(
      p, m, v, ps, reinterpret_cast<__nv_bfloat16*>(shadows_mc), offset, n_items, slot_of_item, G, world, lr_t, scal, beta1, beta2, eps);
  LTG_CHECK_LAUNCH();
  return LTG_OK;
}
