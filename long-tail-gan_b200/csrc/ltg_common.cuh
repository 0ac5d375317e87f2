// Shared device/host helpers for the Long-Tail-GAN sm_100a kernels.
// Everything here is header-only; each .cu includes it.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define LTG_OK 0
#define LTG_ERR_ARG -1
#define LTG_ERR_CUDA -2
#define LTG_ERR_DRIVER -3

#define LTG_CHECK_LAUNCH()                                   \
  do {                                                       \
    cudaError_t e__ = cudaGetLastError();                    \
    if (e__ != cudaSuccess) {                                \
      ltg_set_last_error(cudaGetErrorString(e__), __FILE__, __LINE__); \
      return LTG_ERR_CUDA;                                   \
    }                                                        \
  } while (0)

#define LTG_REQUIRE(cond)                                    \
  do {                                                       \
    if (!(cond)) {                                           \
      ltg_set_last_error("bad argument: " #cond, __FILE__, __LINE__); \
      return LTG_ERR_ARG;                                    \
    }                                                        \
  } while (0)

void ltg_set_last_error(const char* msg, const char* file, int line);

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL) -- OPTIONAL, off by default (LTG_PDL=1 turns it on). One GAN step is a chain of ~25 dependent
// kernels, most of them a few microseconds long, and each pays its own launch gap and prologue (barrier init, TMEM allocation, cluster
// sync, descriptor prefetch) after the previous kernel has drained. With LTG_PDL=1 the kernels of the step are launched with the
// programmatic-stream-serialization attribute (ltg_launch below; the edges survive CUDA-graph capture), every kernel
//   * calls pdl_trigger() first: its in-stream successor may be launched as soon as all of OUR blocks have started, and
//   * calls pdl_wait_cta() before its first global-memory access: it blocks until every prerequisite grid has completed and its
//     writes are visible -- so data dependencies (and write-after-read hazards) are exactly those of an ordinary launch;
// what overlaps is the successor's launch latency and prologue with this kernel's tail. Both instructions are no-ops for a kernel
// launched without the attribute, so a kernel that contains them is safe under any launch.
// MEASURED (round 2, B200, bench shape; all 58 GPU tests pass either way): early trigger 0.570 ms/step, implicit trigger at block exit
// (-DLTG_PDL_EARLY=0) 0.509 ms, PDL off 0.504 ms. The step graph is not one chain: blocks of the critical chain that are launched
// early sit on the SMs (a GEMM block holds 200 KB of shared memory while it waits) and keep the short-lived blocks of the HBM-bound
// side branches (Adam sweeps, weight-gradient GEMMs) from being scheduled, which lengthens the G update by 60 us; without the early
// trigger nothing is left to overlap. Hence off.
// RULE: a kernel may be passed to ltg_launch ONLY if it executes pdl_wait() before touching global memory.
// ---------------------------------------------------------------------------------------------
#ifndef LTG_PDL_EARLY
#define LTG_PDL_EARLY 1   // 1: explicit trigger at the top of every kernel; 0: implicit trigger when a block exits
#endif
__device__ __forceinline__ void pdl_trigger() {
#if LTG_PDL_EARLY
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// One thread waits for the prerequisite grids, the rest of the block parks at the barrier: a block that was launched early sits on its
// SM until its inputs exist, and 18 warps polling the dependency stole issue slots from the kernel still running there.
__device__ __forceinline__ void pdl_wait_cta() {
  if (threadIdx.x == 0) pdl_wait();
  __syncthreads();
}

bool ltg_pdl_enabled();   // runtime.cu: env LTG_PDL (default on)

template <typename... KArgs, typename... Args>
inline cudaError_t ltg_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = ltg_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10, stateless. key = (seed_lo, seed_hi); counter = (c0, c1, c2, c3).
// The numpy mirror lives in oracle/philox.py and tests/ check bit equality of the streams.
// ---------------------------------------------------------------------------------------------
struct Philox4 {
  uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ uint32_t ltg_mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                         uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = ltg_mulhi32(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = ltg_mulhi32(M1, c2), lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0;
    uint32_t n1 = lo1;
    uint32_t n2 = hi0 ^ c3 ^ k1;
    uint32_t n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  Philox4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
  return o;
}

// RNG stream ids (counter word c2). Shared with oracle/philox.py.
#define LTG_STREAM_ENC_DROPOUT 1u
#define LTG_STREAM_EPS 2u
#define LTG_STREAM_DISC_DROPOUT 3u   // + layer index (0: pop branch, 1: niche branch, 2: fc1)
#define LTG_STREAM_SAMPLE 8u
#define LTG_STREAM_PARTNER 9u

// One Bernoulli(keep) bit for element `idx` of stream `stream` at step `step`.
// Four consecutive idx share one Philox block (word idx&3).
__device__ __forceinline__ uint32_t ltg_rand_u32(uint64_t seed, uint32_t stream, uint32_t step, uint64_t idx) {
  uint64_t blk = idx >> 2;
  Philox4 r = philox4x32_10((uint32_t)blk, (uint32_t)(blk >> 32), stream, step, (uint32_t)seed, (uint32_t)(seed >> 32));
  uint32_t w = (uint32_t)(idx & 3);
  return w == 0 ? r.x : (w == 1 ? r.y : (w == 2 ? r.z : r.w));
}

// keep-threshold: element kept iff u32 < thr, thr = floor(keep * 2^32) clipped.
__host__ __device__ __forceinline__ uint32_t ltg_keep_threshold(float keep) {
  double t = (double)keep * 4294967296.0;
  if (t >= 4294967295.0) return 0xFFFFFFFFu;
  if (t <= 0.0) return 0u;
  return (uint32_t)t;
}

// ---------------------------------------------------------------------------------------------
// Cheap stateless dropout bits for the GEMM epilogues (discriminator.py:25,30,44 dropout layers), 16 bits per keep/drop
// decision (threshold floor(keep * 65536)). The key of a mask is a "lowbias32" hash of (seed, stream, step); the bits come from
// ltg_hash_quad below. Philox4x32-10 per element costs ~15 instructions inside an epilogue that is issue bound; this costs ~3.
// oracle/philox.py mirrors both functions bit for bit.
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t ltg_lowbias32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
__host__ __device__ __forceinline__ uint32_t ltg_hash_key(uint64_t seed, uint32_t stream, uint32_t step) {
  return ltg_lowbias32((uint32_t)seed ^ ltg_lowbias32((uint32_t)(seed >> 32) ^ ltg_lowbias32(stream * 0x9E3779B9u + step)));
}
// 64 decision bits for FOUR adjacent columns at once: group index g = (row * rng_ld + col) / 4 (col % 4 == 0), a Philox2x32
// with 5 rounds keyed by the hashed (seed, stream, step) key: one 32x32->64 multiply + one 3-input xor per round, i.e. ~3
// instructions per activation instead of ~6.5 for one lowbias32 per pair (ncu: the hash was half of the 21 instructions per
// activation in the discriminator epilogues). Column 4g+0 / +1 use the low / high half of .x, +2 / +3 those of .y.
// 5 rounds: keep-rate, row/column/lag-4/adjacent-key correlations of the masks are at the sampling-noise level
// (tests/test_oracle_cpu.py); 3 rounds are visibly correlated. oracle/philox.py mirrors it bit for bit.
__host__ __device__ __forceinline__ uint2 ltg_hash_quad(uint32_t key, uint64_t g) {
  uint32_t L = (uint32_t)g, R = (uint32_t)(g >> 32), k = key;
#pragma unroll
  for (int r = 0; r < 5; ++r) {
    const uint64_t p = (uint64_t)L * 0xD256D193ull;
    L = (uint32_t)(p >> 32) ^ k ^ R;
    R = (uint32_t)p;
    k += 0x9E3779B9u;
  }
  uint2 o; o.x = L; o.y = R;
  return o;
}
__host__ __device__ __forceinline__ uint32_t ltg_keep_threshold16(float keep) {
  int t = (int)((double)keep * 65536.0);
  return (uint32_t)(t < 0 ? 0 : (t > 65536 ? 65536 : t));
}

// uniform strictly inside (0,1): ((u32 >> 8) + 0.5) * 2^-24 (exact in fp32)
__device__ __forceinline__ float ltg_u01(uint32_t r) { return ((float)(r >> 8) + 0.5f) * (1.0f / 16777216.0f); }

// ---------------------------------------------------------------------------------------------
// small numeric helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t v) {
  __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&v);
  return __bfloat1622float2(h);
}
__device__ __forceinline__ float bf16_to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }

// 16-byte streaming load (read-once data: keep it out of L1)
__device__ __forceinline__ uint4 ld_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ld_stream_f4(const float* p) {
  float4 r;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
// Streaming accesses of the optimizer sweeps with an L2 evict-first policy: 0.6 GB of fp32 p/m/v passes through the 126 MB L2 every
// G step; marked evict-first it recycles its own lines instead of pushing out the bf16 weight shadows and activations that the
// latency-bound kernels of the same and the next step read (the shadows are STORED with evict-last by the same sweep).
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ float4 ld_stream_f4_hint(const float* p, uint64_t pol) {
  float4 r;
  asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ void st_stream_f4_hint(float* p, float4 v, uint64_t pol) {
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_b64_hint(void* p, uint2 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v2.u32 [%0], {%1,%2}, %3;" ::"l"(p), "r"(v.x), "r"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_stream_f4(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ---------------------------------------------------------------------------------------------
// TF1 Adam on one element (train.py:160-164): the ONE definition every optimizer kernel uses (adam / enc_adam, the
// peer variants, the GEMM epilogue). Explicit round-to-nearest intrinsics fix where the compiler may contract a multiply into an
// FMA, so two kernels given the same inputs produce bit-identical parameters (the data-parallel ranks must stay in lock step).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void ltg_adam1(float& p, float& m, float& v, float g, float lr_t, float b1, float b2, float eps) {
  m = __fmaf_rn(b1, m, __fmul_rn(1.f - b1, g));
  v = __fmaf_rn(b2, v, __fmul_rn(__fmul_rn(1.f - b2, g), g));
  p = __fsub_rn(p, __fdiv_rn(__fmul_rn(lr_t, m), __fadd_rn(sqrtf(v), eps)));
}
__device__ __forceinline__ void ltg_adam4(float4& p, float4& m, float4& v, const float4 g, float lr_t, float b1, float b2, float eps) {
  ltg_adam1(p.x, m.x, v.x, g.x, lr_t, b1, b2, eps); ltg_adam1(p.y, m.y, v.y, g.y, lr_t, b1, b2, eps);
  ltg_adam1(p.z, m.z, v.z, g.z, lr_t, b1, b2, eps); ltg_adam1(p.w, m.w, v.w, g.w, lr_t, b1, b2, eps);
}
