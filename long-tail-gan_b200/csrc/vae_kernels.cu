// MultiVAE-side kernels that are not GEMMs: CSR encoder gather, latent head, activation backward,
// catalog-softmax row statistics, d(loss)/d(logits).
#include "ltg_common.cuh"
#include "../../include/ltgan.h"

namespace {

constexpr int H = LTG_H;
constexpr int L = LTG_L;
constexpr int HV = H / 8;  // 75 16-byte vectors per bf16 weight row

// ---------------------------------------------------------------------------------------------
// a3: encoder. CTA (u, c) owns chunk c (ENC_CHUNK nonzeros) of user u's CSR row: it turns the chunk into
// (item, coef) pairs in shared memory (norm, Philox dropout bit per nonzero), then streams the surviving
// 1200-byte bf16 W_q0 rows. The kernel is a latency chain, not a bandwidth problem (ncu, round 2: 14 % occupancy, 0.45 TB/s,
// IPC 0.9 -- every CTA walks its rows in rounds of 8 loads per thread, one L2/HBM round trip per round, 12 rounds for a full
// chunk), so the chunk is spread over ENC_GROUPS groups of 80 threads: thread l < 75 of group g owns the 16-byte column slice l
// of the rows g, g+4, g+8, ... with 8 loads in flight, i.e. 32 rows per round trip and at most 4 rounds per chunk; the groups'
// partial sums meet in shared memory in a fixed order (deterministic). Users that fit one chunk (the common case) are finished in
// place; longer rows (heavy users, up to 2000 interactions) are spread over several CTAs that accumulate into an fp32 workspace
// row, and the last CTA to arrive applies bias + tanh and clears the workspace again (self-cleaning, no memset).
// Grid: 2-D (user, chunk) -- CTAs beyond a user's last chunk exit at once -- or, with a host-built work list (TrainData: one entry
// (chunk << 20 | user) per non-empty chunk, full chunks first), 1-D over exactly the chunks that exist.
// Restates MultiVAE.py:148 (l2_normalize), 149 (dropout), 152-155 (matmul + bias + tanh).
// ---------------------------------------------------------------------------------------------
constexpr int ENC_GROUPS = 4;
constexpr int ENC_GSZ = 80;     // threads per group; the first HV = 75 of them load
constexpr int ENC_THREADS = ENC_GROUPS * ENC_GSZ;
constexpr int ENC_CHUNK = 128;  // nonzeros per CTA (64 was tried: more atomics, no gain)
static_assert(ENC_GSZ >= HV && ENC_THREADS % 32 == 0 && ENC_THREADS >= ENC_CHUNK, "group size");

// PARTIAL (catalog-sharded layout): the CSR row holds only the interactions whose item lies in this rank's shard
// (local item ids, W = the shard's rows); the row norm comes from row_rnorm[] (it is over the user's WHOLE row), the dropout bit is
// keyed by the global item id (item + item_offset), and the fp32 partial pre-activation sum is ADDED into pre_ws[u] (zeroed by the
// caller) -- bias and tanh follow the cross-rank all-reduce (ltg_bias_tanh).
// 16-byte global -> shared copy that bypasses the register file (LDGSTS). A round of 8 row slices per thread written as register
// loads was serialised by ptxas into load -> use -> load with one or two requests in flight, whatever the source order (SASS check,
// round 2); an asynchronous copy has no result register to schedule around, so all 8 requests of a round are in flight at once.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

template <bool PARTIAL>
__global__ void __launch_bounds__(ENC_THREADS)
enc_gather_fwd_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices, const float* __restrict__ values,
                      int n_items, int64_t uid0, const uint4* __restrict__ W, const float* __restrict__ bias, float keep,
                      uint64_t seed, uint32_t step, const uint32_t* __restrict__ step_dev, __nv_bfloat16* __restrict__ h1,
                      int ld_h1, float* __restrict__ coef, float* __restrict__ pre_ws, int* __restrict__ counters,
                      const int32_t* __restrict__ slot_of_item, __nv_bfloat16* __restrict__ xc, int ld_xc,
                      const float* __restrict__ row_rnorm, int item_offset, const int32_t* __restrict__ work) {
  pdl_trigger();
  pdl_wait_cta();
  __shared__ int s_item[ENC_CHUNK];
  __shared__ float s_coef[ENC_CHUNK];
  __shared__ float s_red[ENC_THREADS / 32];
  __shared__ float4 s_acc[ENC_GROUPS - 1][HV][2];
  __shared__ int s_wcnt[ENC_CHUNK / 32];
  __shared__ int s_last;
  __shared__ uint4 s_stage[8 * ENC_GROUPS * HV];   // [slot of the round][group][column slice]: each thread reads back only its own slots
  int u = blockIdx.x, chunk = blockIdx.y;
  if (work != nullptr) { const int w = work[blockIdx.x]; u = w & 0xFFFFF; chunk = w >> 20; }
  const int tid = threadIdx.x;
  const int beg = indptr[u], end = indptr[u + 1];
  const int nchunks = max(1, (end - beg + ENC_CHUNK - 1) / ENC_CHUNK);
  if (chunk >= nchunks) return;
  if (step_dev != nullptr) step += *step_dev;

  // squared norm of the whole row (values == NULL: binary row)
  float ss = 0.f;
  if (values != nullptr) {
    for (int j = beg + tid; j < end; j += ENC_THREADS) { float v = values[j]; ss += v * v; }
    ss = warp_sum(ss);
    if ((tid & 31) == 0) s_red[tid >> 5] = ss;
    __syncthreads();
    ss = 0.f;
#pragma unroll
    for (int w = 0; w < ENC_THREADS / 32; ++w) ss += s_red[w];
    __syncthreads();
  } else {
    ss = (float)(end - beg);
  }
  const float rs = PARTIAL ? row_rnorm[u] : rsqrtf(fmaxf(ss, 1e-12f));
  const bool drop = keep > 0.f && keep < 1.f;
  const uint32_t thr = drop ? ltg_keep_threshold(keep) : 0xFFFFFFFFu;
  const float scale = drop ? rs / keep : rs;

  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;

  const int c0 = beg + chunk * ENC_CHUNK;
  const int cnt = min(ENC_CHUNK, end - c0);
  // thread j < cnt prepares nonzero j of the chunk (cnt <= ENC_CHUNK <= ENC_THREADS); the survivors of the dropout are compacted
  // in order (ballot + per-warp offsets: deterministic), so the load rounds below carry no empty slots
  int item = 0;
  float c = 0.f;
  if (tid < cnt) {
    item = indices[c0 + tid];
    const float val = values != nullptr ? values[c0 + tid] : 1.0f;
    c = val * scale;
    if (drop) {
      const uint32_t r = ltg_rand_u32(seed, LTG_STREAM_ENC_DROPOUT, step,
                                      (uint64_t)(uid0 + u) * (uint64_t)n_items + (uint64_t)(item + (PARTIAL ? item_offset : 0)));
      if (r >= thr) c = 0.f;
    }
    coef[c0 + tid] = c;
    // dense bf16 coefficient matrix over the batch's active items: the A operand of the encoder weight-gradient GEMM
    if (xc != nullptr) xc[(size_t)u * ld_xc + slot_of_item[item]] = __float2bfloat16(c);
  }
  const unsigned live = __ballot_sync(0xffffffffu, c != 0.f);
  if ((tid & 31) == 0 && tid < ENC_CHUNK) s_wcnt[tid >> 5] = __popc(live);
  __syncthreads();
  int nlive = 0;
#pragma unroll
  for (int w = 0; w < ENC_CHUNK / 32; ++w) {
    if (w == (tid >> 5) && c != 0.f) {
      const int pos = nlive + __popc(live & ((1u << (tid & 31)) - 1u));
      s_item[pos] = item;
      s_coef[pos] = c;
    }
    nlive += s_wcnt[w];
  }
  __syncthreads();
  const int grp = tid / ENC_GSZ, l = tid - grp * ENC_GSZ;
  if (l < HV) {
    constexpr int EU = 8;  // 16-byte copies in flight per thread
    uint4* st = s_stage + grp * HV + l;
    for (int j = grp; j < nlive; j += EU * ENC_GROUPS) {
#pragma unroll
      for (int q = 0; q < EU; ++q) {
        const int jj = j + q * ENC_GROUPS;
        if (jj < nlive) cp_async16(st + q * (ENC_GROUPS * HV), W + (size_t)s_item[jj] * HV + l);
      }
      cp_async_wait_all();
#pragma unroll
      for (int q = 0; q < EU; ++q) {
        const int jj = j + q * ENC_GROUPS;
        if (jj < nlive) {
          const float cf = s_coef[jj];
          const uint4 w = st[q * (ENC_GROUPS * HV)];
          float2 a = unpack_bf16x2(w.x), b = unpack_bf16x2(w.y), d = unpack_bf16x2(w.z), e = unpack_bf16x2(w.w);
          acc[0] = fmaf(cf, a.x, acc[0]); acc[1] = fmaf(cf, a.y, acc[1]);
          acc[2] = fmaf(cf, b.x, acc[2]); acc[3] = fmaf(cf, b.y, acc[3]);
          acc[4] = fmaf(cf, d.x, acc[4]); acc[5] = fmaf(cf, d.y, acc[5]);
          acc[6] = fmaf(cf, e.x, acc[6]); acc[7] = fmaf(cf, e.y, acc[7]);
        }
      }
    }
    if (grp > 0) {
      s_acc[grp - 1][l][0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
      s_acc[grp - 1][l][1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
  }
  __syncthreads();
  if (tid < HV) {   // group 0 collects the other groups' partial sums (fixed order)
#pragma unroll
    for (int g = 0; g < ENC_GROUPS - 1; ++g) {
      const float4 x = s_acc[g][tid][0], y = s_acc[g][tid][1];
      acc[0] += x.x; acc[1] += x.y; acc[2] += x.z; acc[3] += x.w;
      acc[4] += y.x; acc[5] += y.y; acc[6] += y.z; acc[7] += y.w;
    }
  }
  if constexpr (PARTIAL) {
    float* ws = pre_ws + (size_t)u * H;
    if (tid < HV && cnt > 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) atomicAdd(ws + tid * 8 + i, acc[i]);
    }
    return;
  }
  if (nchunks > 1) {
    // multi-CTA row: accumulate into the fp32 workspace, last arrival finishes
    float* ws = pre_ws + (size_t)u * H;
    if (tid < HV) {
#pragma unroll
      for (int i = 0; i < 8; ++i) atomicAdd(ws + tid * 8 + i, acc[i]);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(counters + u, 1) == nchunks - 1) ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (tid < HV) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { acc[i] = __ldcg(ws + tid * 8 + i); ws[tid * 8 + i] = 0.f; }
    }
    if (tid == 0) counters[u] = 0;
  }
  if (tid < HV) {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias) + tid * 2);
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias) + tid * 2 + 1);
    uint4 o;
    o.x = pack_bf16x2(tanhf(acc[0] + b0.x), tanhf(acc[1] + b0.y));
    o.y = pack_bf16x2(tanhf(acc[2] + b0.z), tanhf(acc[3] + b0.w));
    o.z = pack_bf16x2(tanhf(acc[4] + b1.x), tanhf(acc[5] + b1.y));
    o.w = pack_bf16x2(tanhf(acc[6] + b1.z), tanhf(acc[7] + b1.w));
    *reinterpret_cast<uint4*>(h1 + (size_t)u * ld_h1 + tid * 8) = o;
  }
}

// ---------------------------------------------------------------------------------------------
// a3/a4: latent head. MultiVAE.py:157-162 (mu/logvar split, std, KL) and 178-181 (reparameterise).
// ---------------------------------------------------------------------------------------------
__global__ void latent_fwd_kernel(const float* __restrict__ mulv, const float* __restrict__ eps, int B, int64_t uid0, float is_training,
                                  uint64_t seed, uint32_t step, const uint32_t* __restrict__ step_dev,
                                  __nv_bfloat16* __restrict__ z, int ld_z, float* __restrict__ zmu, float* __restrict__ scal) {
  __shared__ float s_red[8];
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  float kl = 0.f;
  if (idx < B * L) {
    const int u = idx / L, j = idx - u * L;
    const float mu = mulv[(size_t)u * 2 * L + j];
    const float lv = mulv[(size_t)u * 2 * L + L + j];
    const float ev = expf(lv);
    kl = 0.5f * (-lv + ev + mu * mu - 1.0f);
    float e = 0.f;
    if (is_training != 0.f) {
      if (eps != nullptr) {
        e = eps[idx];
      } else {
        if (step_dev != nullptr) step += *step_dev;
        const uint64_t g = (uint64_t)(uid0 + u) * (uint64_t)L + (uint64_t)j;
        Philox4 r = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), LTG_STREAM_EPS, step, (uint32_t)seed, (uint32_t)(seed >> 32));
        const float u1 = ltg_u01(r.x), u2 = ltg_u01(r.y);
        e = sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
      }
    }
    const float d = is_training * e * expf(0.5f * lv);
    zmu[idx] = d;
    z[(size_t)u * ld_z + j] = __float2bfloat16(mu + d);
  }
  kl = warp_sum(kl);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = kl;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_red[w];
    atomicAdd(scal + LTG_S_KL_SUM, t);
  }
}

// 2-D grid: blockIdx.x tiles the rows by COLSUM_ROWS, blockIdx.y tiles the columns by blockDim.x; every thread owns one
// column, sums it over the row tile (coalesced across the warp) and issues one atomic per column for the bias gradient.
constexpr int COLSUM_ROWS = 8;

__global__ void latent_bwd_kernel(const float* __restrict__ dz, const float* __restrict__ mulv, const float* __restrict__ zmu, int B,
                                  float inv_bg, float anneal, const float* __restrict__ scal, __nv_bfloat16* __restrict__ dmulv, int ld,
                                  float* __restrict__ db) {
  if (anneal < 0.f) anneal = scal[LTG_S_ANNEAL];
  const int r0 = blockIdx.x * COLSUM_ROWS;
  const int r1 = min(B, r0 + COLSUM_ROWS);
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c < 2 * L) {
    float cs = 0.f;
    const bool is_mu = c < L;
    const int j = is_mu ? c : c - L;
    for (int r = r0; r < r1; ++r) {
      const float g = dz[(size_t)r * L + j];
      float o;
      if (is_mu) {
        o = g + anneal * mulv[(size_t)r * 2 * L + j] * inv_bg;
      } else {
        const float lv = mulv[(size_t)r * 2 * L + L + j];
        o = g * zmu[(size_t)r * L + j] * 0.5f + anneal * 0.5f * (expf(lv) - 1.0f) * inv_bg;
      }
      dmulv[(size_t)r * ld + c] = __float2bfloat16(o);
      cs += o;
    }
    if (db != nullptr) atomicAdd(db + c, cs);
  }
}

__global__ void tanh_bwd_kernel(const float* __restrict__ dy, int ld_dy, int n_partials, int64_t partial_stride,
                                const __nv_bfloat16* __restrict__ y, int ld_y, int B, int N,
                                __nv_bfloat16* __restrict__ dxb, int ld_dxb, float* __restrict__ dxf, int ld_dxf, float* __restrict__ db) {
  pdl_trigger();
  pdl_wait_cta();
  constexpr int ROWS = 4;   // small row tile: this kernel sits on the critical path of the backward chain, parallelism first
  const int r0 = blockIdx.x * ROWS;
  const int r1 = min(B, r0 + ROWS);
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c < N) {
    float cs = 0.f;
    for (int r = r0; r < r1; ++r) {
      const float t = __bfloat162float(y[(size_t)r * ld_y + c]);
      const float* src = dy + (size_t)r * ld_dy + c;
      float d0 = src[0], d1 = 0.f, d2 = 0.f, d3 = 0.f;   // split-K partials: four independent load chains
      int sp = 1;
      for (; sp + 3 <= n_partials; sp += 3) {
        d1 += src[(size_t)sp * partial_stride]; d2 += src[(size_t)(sp + 1) * partial_stride]; d3 += src[(size_t)(sp + 2) * partial_stride];
      }
      for (; sp < n_partials; ++sp) d1 += src[(size_t)sp * partial_stride];
      const float d = (d0 + d1) + (d2 + d3);
      const float o = d * (1.0f - t * t);
      if (dxb != nullptr) dxb[(size_t)r * ld_dxb + c] = __float2bfloat16(o);
      if (dxf != nullptr) dxf[(size_t)r * ld_dxf + c] = o;
      cs += o;
    }
    if (db != nullptr) atomicAdd(db + c, cs);
  }
}

// Vector form (N % 4 == 0, 16-byte aligned rows): a thread owns four adjacent columns of TB4_ROWS rows and issues all of its
// TB4_ROWS x n_partials 16-byte loads before it consumes any of them. The scalar kernel above is latency bound (one 4-byte load chain
// per thread): 12 us alone and 30 us while the decoder weight-gradient GEMM keeps HBM busy next to it (timeline, round 2).
constexpr int TB4_ROWS = 4;
constexpr int TB4_MAXP = 8;

__global__ void __launch_bounds__(160)
tanh_bwd4_kernel(const float* __restrict__ dy, int ld_dy, int n_partials, int64_t partial_stride, const __nv_bfloat16* __restrict__ y, int ld_y,
                 int B, int N, __nv_bfloat16* __restrict__ dxb, int ld_dxb, float* __restrict__ dxf, int ld_dxf, float* __restrict__ db) {
  pdl_trigger();
  pdl_wait_cta();
  const int c = (blockIdx.y * blockDim.x + threadIdx.x) * 4;
  if (c >= N) return;
  const int r0 = blockIdx.x * TB4_ROWS;
  float4 part[TB4_ROWS][TB4_MAXP];
  uint2 yy[TB4_ROWS];
#pragma unroll
  for (int r = 0; r < TB4_ROWS; ++r) {
    const bool ok = r0 + r < B;
    const float* src = dy + (size_t)(ok ? r0 + r : 0) * ld_dy + c;
#pragma unroll
    for (int sp = 0; sp < TB4_MAXP; ++sp)
      part[r][sp] = (ok && sp < n_partials) ? ld_stream_f4(src + (size_t)sp * partial_stride) : make_float4(0.f, 0.f, 0.f, 0.f);
    yy[r] = ok ? *reinterpret_cast<const uint2*>(y + (size_t)(r0 + r) * ld_y + c) : make_uint2(0u, 0u);
  }
  float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int r = 0; r < TB4_ROWS; ++r) {
    if (r0 + r >= B) break;
    float4 d = part[r][0];
#pragma unroll
    for (int sp = 1; sp < TB4_MAXP; ++sp) { d.x += part[r][sp].x; d.y += part[r][sp].y; d.z += part[r][sp].z; d.w += part[r][sp].w; }
    const float2 t01 = unpack_bf16x2(yy[r].x), t23 = unpack_bf16x2(yy[r].y);
    float4 o;
    o.x = d.x * (1.0f - t01.x * t01.x); o.y = d.y * (1.0f - t01.y * t01.y);
    o.z = d.z * (1.0f - t23.x * t23.x); o.w = d.w * (1.0f - t23.y * t23.y);
    if (dxb != nullptr) {
      uint2 u; u.x = pack_bf16x2(o.x, o.y); u.y = pack_bf16x2(o.z, o.w);
      *reinterpret_cast<uint2*>(dxb + (size_t)(r0 + r) * ld_dxb + c) = u;
    }
    if (dxf != nullptr) *reinterpret_cast<float4*>(dxf + (size_t)(r0 + r) * ld_dxf + c) = o;
    cs.x += o.x; cs.y += o.y; cs.z += o.z; cs.w += o.w;
  }
  if (db != nullptr) { atomicAdd(db + c, cs.x); atomicAdd(db + c + 1, cs.y); atomicAdd(db + c + 2, cs.z); atomicAdd(db + c + 3, cs.w); }
}

// ---------------------------------------------------------------------------------------------
// a6: row statistics of the catalog softmax. One CTA per user (the interaction list of a heavy user is 20x the mean).
// ---------------------------------------------------------------------------------------------
constexpr int STATS_THREADS = 128;

__device__ __forceinline__ float block_sum_128(float v, float* s_red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  return s_red[0] + s_red[1] + s_red[2] + s_red[3];
}
__device__ __forceinline__ float block_max_128(float v, float* s_red) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  return fmaxf(fmaxf(s_red[0], s_red[1]), fmaxf(s_red[2], s_red[3]));
}

__global__ void __launch_bounds__(STATS_THREADS)
dec_row_stats_kernel(const float2* __restrict__ partial, int n_blocks, const __nv_bfloat16* __restrict__ logits, int ld, int B,
                     const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices, const float* __restrict__ values,
                     const int32_t* __restrict__ samp_ptr, const int32_t* __restrict__ samp_items,
                     const int32_t* __restrict__ samp_valid, float* __restrict__ lse_out, float* __restrict__ xw_out,
                     float* __restrict__ su_out, float* __restrict__ scal) {
  __shared__ float s_red[4];
  const int u = blockIdx.x;
  const int tid = threadIdx.x;
  float mx = -INFINITY;
  for (int b = tid; b < n_blocks; b += STATS_THREADS) mx = fmaxf(mx, partial[(size_t)b * B + u].x);
  mx = block_max_128(mx, s_red);
  float s = 0.f;
  for (int b = tid; b < n_blocks; b += STATS_THREADS) {
    const float2 p = partial[(size_t)b * B + u];
    s += p.y * __expf(p.x - mx);
  }
  s = block_sum_128(s, s_red);
  const float lse = mx + logf(s);
  const __nv_bfloat16* row = logits + (size_t)u * ld;
  float nll = 0.f, xw = 0.f;
  if (indptr != nullptr) {
    for (int j = indptr[u] + tid; j < indptr[u + 1]; j += STATS_THREADS) {
      const float v = values != nullptr ? values[j] : 1.0f;
      nll -= v * (__bfloat162float(row[indices[j]]) - lse);
      xw += v;
    }
    nll = block_sum_128(nll, s_red);
    xw = block_sum_128(xw, s_red);
  }
  float sp = 0.f;
  if (samp_ptr != nullptr) {
    for (int j = samp_ptr[u] + tid; j < samp_ptr[u + 1]; j += STATS_THREADS)
      if (samp_valid[j] > 0) sp += __expf(__bfloat162float(row[samp_items[j]]) - lse);
    sp = block_sum_128(sp, s_red);
  }
  if (tid == 0) {
    lse_out[u] = lse;
    if (xw_out != nullptr) xw_out[u] = xw;
    if (su_out != nullptr) su_out[u] = sp;
    if (indptr != nullptr) atomicAdd(scal + LTG_S_NLL_SUM, nll);
    if (samp_ptr != nullptr) atomicAdd(scal + LTG_S_SUM_P, sp);
  }
}

__global__ void dec_probs_kernel(const __nv_bfloat16* __restrict__ logits, int ld, const float* __restrict__ lse, int n_items,
                                 float* __restrict__ out, int ld_out) {
  const int u = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n_items) out[(size_t)u * ld_out + c] = __expf(__bfloat162float(logits[(size_t)u * ld + c]) - lse[u]);
}

// dense part of d g_loss / d logits: dl = pi * a_u, 8 bf16 per thread.
__global__ void dlogits_dense_kernel(const uint4* __restrict__ logits, int ld8, const float* __restrict__ lse, const float* __restrict__ xw,
                                     const float* __restrict__ s_u, int n_items, float inv_bg, float lam, const float* __restrict__ scal,
                                     uint4* __restrict__ dl) {
  const int u = blockIdx.y;
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= ld8) return;
  float ybar = 0.f;
  if (lam != 0.f) {
    const float cnt = scal[LTG_S_CNT];
    ybar = cnt > 0.f ? scal[LTG_S_SUM_Y] / cnt : 0.f;
  }
  const float a = xw[u] * inv_bg + lam * ybar * (s_u != nullptr ? s_u[u] : 0.f);
  const float l = lse[u];
  const uint4 x = ld_nc_v4(logits + (size_t)u * ld8 + v);
  const uint32_t xi[4] = {x.x, x.y, x.z, x.w};
  uint32_t o[4];
  const int c0 = v * 8;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float2 f = unpack_bf16x2(xi[q]);
    const float p0 = (c0 + 2 * q < n_items) ? __expf(f.x - l) * a : 0.f;
    const float p1 = (c0 + 2 * q + 1 < n_items) ? __expf(f.y - l) * a : 0.f;
    o[q] = pack_bf16x2(p0, p1);
  }
  dl[(size_t)u * ld8 + v] = make_uint4(o[0], o[1], o[2], o[3]);
}

// sparse fix-ups: -x_ui/Bg at the user's interactions, -lam*Ybar*pi_ui at the sampled (valid) items. One CTA per user
// (heavy users have up to 2000 interactions; a single warp made the kernel wait for them).
constexpr int SPARSE_THREADS = 128;
__global__ void __launch_bounds__(SPARSE_THREADS)
dlogits_sparse_kernel(const __nv_bfloat16* __restrict__ logits, int ld, const float* __restrict__ lse, int B, float inv_bg,
                      float lam, const float* __restrict__ scal, const int32_t* __restrict__ indptr,
                      const int32_t* __restrict__ indices, const float* __restrict__ values,
                      const int32_t* __restrict__ samp_ptr, const int32_t* __restrict__ samp_items,
                      const int32_t* __restrict__ samp_valid, __nv_bfloat16* __restrict__ dl) {
  const int u = blockIdx.x;
  const int tid = threadIdx.x;
  __nv_bfloat16* drow = dl + (size_t)u * ld;
  for (int j = indptr[u] + tid; j < indptr[u + 1]; j += SPARSE_THREADS) {
    const int i = indices[j];
    const float v = values != nullptr ? values[j] : 1.0f;
    drow[i] = __float2bfloat16(__bfloat162float(drow[i]) - v * inv_bg);
  }
  if (lam != 0.f && samp_ptr != nullptr) {
    __syncthreads();  // a sampled item may also be one of the user's interactions: order the two read-modify-writes
    const float cnt = scal[LTG_S_CNT];
    const float ybar = cnt > 0.f ? scal[LTG_S_SUM_Y] / cnt : 0.f;
    const float l = lse[u];
    const __nv_bfloat16* row = logits + (size_t)u * ld;
    for (int j = samp_ptr[u] + tid; j < samp_ptr[u + 1]; j += SPARSE_THREADS) {
      if (samp_valid[j] > 0) {
        const int i = samp_items[j];
        const float pi = __expf(__bfloat162float(row[i]) - l);
        drow[i] = __float2bfloat16(__bfloat162float(drow[i]) - lam * ybar * pi);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// a6 + a12 fused for the G step: one CTA per user does (1) the softmax row statistics (lse, NLL, sum of the sampled
// probabilities), (2) the dense part of d g_loss / d logits for its row, (3) the sparse fix-ups. Everything is row-local
// (the only cross-row quantity, Ybar = sum y / cnt, comes from the discriminator head that ran before), so the three
// launches + two grid-wide dependencies of the unfused path collapse into one.
// ---------------------------------------------------------------------------------------------
constexpr int ROWBWD_THREADS = 256;

__global__ void __launch_bounds__(ROWBWD_THREADS)
dec_row_bwd_kernel(const float2* __restrict__ partial, int n_blocks, const __nv_bfloat16* __restrict__ logits, int ld, int B, int n_items,
                   float inv_bg, float lam, const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                   const float* __restrict__ values, const int32_t* __restrict__ samp_ptr, const int32_t* __restrict__ samp_items,
                   const int32_t* __restrict__ samp_valid, float* __restrict__ lse_out, float* __restrict__ scal,
                   __nv_bfloat16* __restrict__ dl) {
  pdl_trigger();
  pdl_wait_cta();
  __shared__ float s_red[ROWBWD_THREADS / 32];
  const int u = blockIdx.x, tid = threadIdx.x;
  auto bsum = [&](float v) {
    v = warp_sum(v);
    __syncthreads();
    if ((tid & 31) == 0) s_red[tid >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < ROWBWD_THREADS / 32; ++w) t += s_red[w];
    return t;
  };
  auto bmax = [&](float v) {
    v = warp_max(v);
    __syncthreads();
    if ((tid & 31) == 0) s_red[tid >> 5] = v;
    __syncthreads();
    float t = s_red[0];
#pragma unroll
    for (int w = 1; w < ROWBWD_THREADS / 32; ++w) t = fmaxf(t, s_red[w]);
    return t;
  };
  // (1) statistics
  float mx = -INFINITY;
  for (int b = tid; b < n_blocks; b += ROWBWD_THREADS) mx = fmaxf(mx, partial[(size_t)b * B + u].x);
  mx = bmax(mx);
  float s = 0.f;
  for (int b = tid; b < n_blocks; b += ROWBWD_THREADS) {
    const float2 p = partial[(size_t)b * B + u];
    s += p.y * __expf(p.x - mx);
  }
  s = bsum(s);
  const float lse = mx + logf(s);
  const __nv_bfloat16* row = logits + (size_t)u * ld;
  __nv_bfloat16* drow = dl + (size_t)u * ld;
  float nll = 0.f, xw = 0.f, sp = 0.f;
  for (int j = indptr[u] + tid; j < indptr[u + 1]; j += ROWBWD_THREADS) {
    const float v = values != nullptr ? values[j] : 1.0f;
    nll -= v * (__bfloat162float(row[indices[j]]) - lse);
    xw += v;
  }
  const bool gan = lam != 0.f && samp_ptr != nullptr;
  if (samp_ptr != nullptr)
    for (int j = samp_ptr[u] + tid; j < samp_ptr[u + 1]; j += ROWBWD_THREADS)
      if (samp_valid[j] > 0) sp += __expf(__bfloat162float(row[samp_items[j]]) - lse);
  nll = bsum(nll); xw = bsum(xw); sp = bsum(sp);
  float ybar = 0.f;
  if (gan) {
    const float cnt = scal[LTG_S_CNT];
    ybar = cnt > 0.f ? scal[LTG_S_SUM_Y] / cnt : 0.f;
  }
  if (tid == 0) {
    lse_out[u] = lse;
    atomicAdd(scal + LTG_S_NLL_SUM, nll);
    if (samp_ptr != nullptr) atomicAdd(scal + LTG_S_SUM_P, sp);
  }
  // (2) dense: dl = pi * (xw/Bg + lam*Ybar*s_u)
  const float a = xw * inv_bg + (gan ? lam * ybar * sp : 0.f);
  const int ld8 = ld >> 3;
  const uint4* src = reinterpret_cast<const uint4*>(row);
  uint4* dst = reinterpret_cast<uint4*>(drow);
  for (int v0 = tid; v0 < ld8; v0 += 4 * ROWBWD_THREADS) {
    uint4 x[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int v = v0 + q * ROWBWD_THREADS;
      x[q] = v < ld8 ? ld_nc_v4(src + v) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int v = v0 + q * ROWBWD_THREADS;
      if (v >= ld8) continue;
      const uint32_t xi[4] = {x[q].x, x[q].y, x[q].z, x[q].w};
      uint32_t o[4];
      const int c0 = v * 8;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = unpack_bf16x2(xi[k]);
        const float p0 = (c0 + 2 * k < n_items) ? __expf(f.x - lse) * a : 0.f;
        const float p1 = (c0 + 2 * k + 1 < n_items) ? __expf(f.y - lse) * a : 0.f;
        o[k] = pack_bf16x2(p0, p1);
      }
      dst[v] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
  __syncthreads();
  // (3) sparse fix-ups
  for (int j = indptr[u] + tid; j < indptr[u + 1]; j += ROWBWD_THREADS) {
    const int i = indices[j];
    const float v = values != nullptr ? values[j] : 1.0f;
    drow[i] = __float2bfloat16(__bfloat162float(drow[i]) - v * inv_bg);
  }
  if (gan) {
    __syncthreads();
    for (int j = samp_ptr[u] + tid; j < samp_ptr[u + 1]; j += ROWBWD_THREADS) {
      if (samp_valid[j] > 0) {
        const int i = samp_items[j];
        const float pi = __expf(__bfloat162float(row[i]) - lse);
        drow[i] = __float2bfloat16(__bfloat162float(drow[i]) - lam * ybar * pi);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Data parallel encoder gradient. Rank r owns the item shard [row0, row0+R). For every interaction (u, i) of the GLOBAL
// batch whose item lies in the shard it rebuilds the forward's coefficient x_ui * rsqrt(|x_u|^2) * mask/keep -- the dropout
// bit is a stateless function of (seed, step, global uid, item), so no other rank has to send it -- and scatters it into the
// dense bf16 matrix Xc_glob[g_row, slot] that the shard's weight-gradient GEMM consumes. Only dh1pre travels (all-gather).
// ---------------------------------------------------------------------------------------------
__global__ void enc_coef_scatter_kernel(const int32_t* __restrict__ e_row, const int32_t* __restrict__ e_item, const int32_t* __restrict__ e_slot,
                                        const int64_t* __restrict__ row_uid, const float* __restrict__ row_rnorm, int n_entries, int n_items,
                                        float keep, uint64_t seed, uint32_t step, const uint32_t* __restrict__ step_dev,
                                        __nv_bfloat16* __restrict__ xc, int ld_xc) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_entries) return;
  if (step_dev != nullptr) step += *step_dev;
  const int r = e_row[e];
  const int item = e_item[e];
  const bool drop = keep > 0.f && keep < 1.f;
  float c = row_rnorm[r];
  if (drop) {
    const uint32_t rnd = ltg_rand_u32(seed, LTG_STREAM_ENC_DROPOUT, step, (uint64_t)row_uid[r] * (uint64_t)n_items + (uint64_t)item);
    c = rnd < ltg_keep_threshold(keep) ? c / keep : 0.f;
  }
  xc[(size_t)r * ld_xc + e_slot[e]] = __float2bfloat16(c);
}

// the same entries back to zero after the shard's weight-gradient GEMM has consumed the matrix (replaces a fill of the whole matrix)
__global__ void enc_coef_clear_kernel(const int32_t* __restrict__ e_row, const int32_t* __restrict__ e_slot, int n_entries,
                                      __nv_bfloat16* __restrict__ xc, int ld_xc) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n_entries) xc[(size_t)e_row[e] * ld_xc + e_slot[e]] = __float2bfloat16(0.f);
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" int ltg_enc_gather_fwd(const int32_t* indptr, const int32_t* indices, const float* values, int B, int n_items, int64_t uid0,
                                  const void* W_enc_bf16, const float* b_q0, float keep, uint64_t seed, uint32_t step,
                                  const uint32_t* step_dev, void* h1_bf16, int ld_h1, float* coef, int max_row_nnz, float* pre_ws,
                                  int32_t* counters, const int32_t* slot_of_item, void* xc_bf16, int ld_xc, const int32_t* work, int n_work,
                                  void* stream) {
  LTG_REQUIRE(indptr && indices && W_enc_bf16 && b_q0 && h1_bf16 && coef);
  LTG_REQUIRE(work == nullptr || (n_work >= B && B < (1 << 20)));
  LTG_REQUIRE(ld_h1 % 8 == 0 && ld_h1 >= H);
  LTG_REQUIRE(max_row_nnz <= ENC_CHUNK || (pre_ws != nullptr && counters != nullptr));
  LTG_REQUIRE(xc_bf16 == nullptr || slot_of_item != nullptr);
  if (B <= 0) return LTG_OK;
  const int chunks = max_row_nnz <= ENC_CHUNK ? 1 : (max_row_nnz + ENC_CHUNK - 1) / ENC_CHUNK;
  const dim3 grid = work != nullptr ? dim3(n_work) : dim3(B, chunks);
  ltg_launch(enc_gather_fwd_kernel<false>, dim3(grid), dim3(ENC_THREADS), 0, (cudaStream_t)stream, 
      indptr, indices, values, n_items, uid0, reinterpret_cast<const uint4*>(W_enc_bf16), b_q0, keep, seed, step, step_dev,
      reinterpret_cast<__nv_bfloat16*>(h1_bf16), ld_h1, coef, pre_ws, counters, slot_of_item, reinterpret_cast<__nv_bfloat16*>(xc_bf16), ld_xc,
      nullptr, 0, work);
  LTG_CHECK_LAUNCH();
  return LTG_OK;
}

extern "C" int ltg_enc_gather_partial(const int32_t* indptr, const int32_t* indices, int B, int n_items_global, int item_offset, int64_t uid0,
                                      const void* W_shard_bf16, const float* row_rnorm, float keep, uint64_t seed, uint32_t step,
                                      const uint32_t* step_dev, float* pre_sum, float* coef, int max_row_nnz, const int32_t* slot_of_item,
                                      void* xc_bf16, int ld_xc, void* stream) {
  LTG_REQUIRE(indptr && indices && W_shard_bf16 && row_rnorm && pre_sum && coef);
  LTG_REQUIRE(xc_bf16 == nullptr || slot_of_item != nullptr);
  if (B <= 0) return LTG_OK;
  const int chunks = max_row_nnz <= ENC_CHUNK ? 1 : (max_row_nnz + ENC_CHUNK - 1) / ENC_CHUNK;
  ltg_launch(enc_gather_fwd_kernel<true>, dim3(dim3(B, chunks)), dim3(ENC_THREADS), 0, (cudaStream_t)stream, 
      indptr, indices, nullptr, n_items_global, uid0, reinterpret_cast<const uint4*>(W_shard_bf16), nullptr, keep, seed, step, step_dev, nullptr, 0,
      coef, pre_sum, nullptr, slot_of_item, reinterpret_cast<__nv_bfloat16*>(xc_bf16), ld_xc, row_rnorm, item_offset, nullptr);
  LTG_CHECK_LAUNCH();
  return LTG_OK;
}

namespace {
__global__ void bias_tanh_kernel(const float* __restrict__ pre, int ld, const float* __restrict__ bias, int B, int N, __nv_bfloat16* __restrict__ out,
                                 int ld_out) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 4, r = blockIdx.y;
  if (c >= N || r >= B) return;
  const float4 x = *reinterpret_cast<const float4*>(pre + (size_t)r * ld + c);
  const float4 b = __ldg(reinterpret_cast<const float4*>(bias + c));
  uint2 o;
  o.x = pack_bf16x2(tanhf(x.x + b.x), tanhf(x.y + b.y)); o.y = pack_bf16x2(tanhf(x.z + b.z), tanhf(x.w + b.w));
  *reinterpret_cast<uint2*>(out + (size_t)r * ld_out + c) = o;
}
}  // namespace

extern "C" int ltg_bias_tanh(const float* pre, int ld, const float* bias, int B, int N, void* out_bf16, int ld_out, void* stream) {
  LTG_REQUIRE(pre && bias && out_bf16 && N % 4 == 0 && ld % 4 == 0 && ld_out % 4 == 0);
  if (B <= 0) return LTG_OK;
  bias_tanh_kernel<<<dim3((N / 4 + 127) / 128, B), 128, 0, (cudaStream_t)stream>>>(pre, ld, bias, B, N, reinterpret_cast<__nv_bfloat16*>(out_bf16), ld_out);
  LTG_CHECK_LAUNCH();
  return LTG_OK;
}

extern "C" int ltg_latent_fwd(const float* mulv, const float* eps, int B, int64_t uid0, float is_training, uint64_t seed, uint32_t step,
                              const uint32_t* step_dev, void* z_bf16, int ld_z, float* zmu, float* scal, void* stream) {
  LTG_REQUIRE(mulv && z_bf16 && zmu && scal);
  if (B <= 0) return LTG_OK;
  const int n = B * L;
  latent_fwd_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(mulv, eps, B, uid0, is_training, seed, step, step_dev,
                                                                        reinterpret_cast<__nv_bfloat16*>(z_bf16), ld_z, zmu, scal);
  LTG_CHECK_LAUNCH();
  return LTG_OK;
}

extern "C" int ltg_latent_bwd(const float* dz, const float* mulv, const float* zmu, int B, int B_global, float anneal, const float* scal,
                              void* dmulv_bf16, int ld, float* db_q1, void* stream) {
  LTG_REQUIRE(dz && mulv && zmu && dmulv_bf16);
  LTG_REQUIRE(anneal >= 0.f || scal != nullptr);
  if (B <= 0) return LTG_OK;
  latent_bwd_kernel<<<dim3((B + COLSUM_ROWS - 1) / COLSUM_ROWS, (2 * L + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      dz, mulv, zmu, B, 1.0f / (float)B_global, anneal, scal, reinterpret_cast<__nv_bfloat16*>(dmulv_bf16), ld, db_q1);
  LTG_CHECK_LAUNCH();
  return LTG_OK;
}

extern "C" int ltg_tanh_bwd(const float* dy, int ld_dy, int n_partials, int64_t partial_stride, const void* y_bf16, int ld_y, int B, int N,
                            void* dx_bf16, int ld_dxb, float* dx_f32, int ld_dxf, float* dbias, void* stream) {
  LTG_REQUIRE(dy && y_bf16 && n_partials >= 1);
  if (B <= 0) return LTG_OK;
  const bool vec = N % 4 == 0 && n_partials <= TB4_MAXP && ld_dy % 4 == 0 && partial_stride % 4 == 0 && ld_y % 4 == 0 &&
                   (reinterpret_cast<uintptr_t>(dy) & 15) == 0 && (reinterpret_cast<uintptr_t>(y_bf16) & 7) == 0 &&
                   (dx_bf16 == nullptr || (ld_dxb % 4 == 0 && (reinterpret_cast<uintptr_t>(dx_bf16) & 7) == 0)) &&
                   (dx_f32 == nullptr || (ld_dxf % 4 == 0 && (reinterpret_cast<uintptr_t>(dx_f32) & 15) == 0));
  if (vec) {
    const int n4 = N / 4;
    ltg_launch(tanh_bwd4_kernel, dim3(dim3((B + TB4_ROWS - 1) / TB4_ROWS, (n4 + 159) / 160)), dim3(160), 0, (cudaStream_t)stream, 
        dy, ld_dy, n_partials, partial_stride, reinterpret_cast<const __nv_bfloat16*>(y_bf16), ld_y, B, N,
        reinterpret_cast<__nv_bfloat16*>(dx_bf16), ld_dxb, dx_f32, ld_dxf, dbias);
    LTG_CHECK_LAUNCH();
    return LTG_OK;
  }
  ltg_launch(tanh_bwd_kernel, dim3(dim3((B + 3) / 4, (N + 127) / 128)), dim3(128), 0, (cudaStream_t)stream, 
      dy, ld_dy, n_partials, partial_stride, reinterpret_cast<const __nv_bfloat16*>(y_bf16), ld_y, B, N, reinterpret_cast<__nv_bfloat16*>(dx_bf16), ld_dxb, dx_f32,
      ld_dxf, dbias);
  LTG_CHECK_LAUNCH();
  return LTG_OK;
}

extern "C" int ltg_dec_row_stats(const float* partial, int n_blocks, const void* logits_bf16, int ld_logits, int B,
                                 const int32_t* indptr, const int32_t* indices, const float* values,
                                 const int32_t* samp_ptr, const int32_t* samp_items, const int32_t* samp_valid,
                                 float* lse, float* xw, float* s_u, float* scal, void* stream) {
  LTG_REQUIRE(partial && lse && scal);
  LTG_REQUIRE((indptr == nullptr && samp_ptr == nullptr) || logits_bf16 != nullptr);
  if (B <= 0) return LTG_OK;
  dec_row_stats_kernel<<<B, STATS_THREADS, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float2*>(partial), n_blocks, reinterpret_cast<const __nv_bfloat16*>(logits_bf16), ld_logits, B, indptr, indices,
      values, samp_ptr, samp_items, samp_valid, lse, xw, s_u, scal);
  LTG_CHECK_LAUNCH();
  return LTG_OK;
}

extern "C" int ltg_dec_probs(const void* logits_bf16, int ld_logits, const float* lse, int B, int n_items, float* out, int ld_out, void* stream) {
  LTG_REQUIRE(logits_bf16 && lse && out);
  if (B <= 0) return LTG_OK;
  dim3 grid((n_items + 255) / 256, B);
  dec_probs_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(logits_bf16), ld_logits, lse, n_items, out, ld_out);
  LTG_CHECK_LAUNCH();
  return LTG_OK;
}

extern "C" int ltg_dec_dlogits(const void* logits_bf16, int ld, const float* lse, const float* xw, const float* s_u, int B, int n_items,
                               int B_global, float lam, const float* scal,
                               const int32_t* indptr, const int32_t* indices, const float* values,
                               const int32_t* samp_ptr, const int32_t* samp_items, const int32_t* samp_valid,
                               void* dl_bf16, void* stream) {
  LTG_REQUIRE(logits_bf16 && lse && xw && dl_bf16 && indptr && indices);
  LTG_REQUIRE(ld % 8 == 0 && ld >= n_items);
  LTG_REQUIRE(lam == 0.f || scal != nullptr);
  if (B <= 0) return LTG_OK;
  const int ld8 = ld / 8;
  dim3 grid((ld8 + 255) / 256, B);
  const float inv_bg = 1.0f / (float)B_global;
  dlogits_dense_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(logits_bf16), ld8, lse, xw, s_u, n_items, inv_bg,
                                                                lam, scal, reinterpret_cast<uint4*>(dl_bf16));
  LTG_CHECK_LAUNCH();
  dlogits_sparse_kernel<<<B, SPARSE_THREADS, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(logits_bf16), ld, lse, B, inv_bg, lam, scal, indptr, indices, values, samp_ptr, samp_items,
      samp_valid, reinterpret_cast<__nv_bfloat16*>(dl_bf16));
  LTG_CHECK_LAUNCH();
  return LTG_OK;
}

extern "C" int ltg_dec_row_bwd(const float* partial, int n_blocks, const void* logits_bf16, int ld, int B, int n_items, int B_global, float lam,
                               const int32_t* indptr, const int32_t* indices, const float* values,
                               const int32_t* samp_ptr, const int32_t* samp_items, const int32_t* samp_valid,
                               float* lse, float* scal, void* dl_bf16, void* stream) {
  LTG_REQUIRE(partial && logits_bf16 && indptr && indices && lse && scal && dl_bf16);
  LTG_REQUIRE(ld % 8 == 0 && ld >= n_items);
  if (B <= 0) return LTG_OK;
  ltg_launch(dec_row_bwd_kernel, dim3(B), dim3(ROWBWD_THREADS), 0, (cudaStream_t)stream, 
      reinterpret_cast<const float2*>(partial), n_blocks, reinterpret_cast<const __nv_bfloat16*>(logits_bf16), ld, B, n_items,
      1.0f / (float)B_global, lam, indptr, indices, values, samp_ptr, samp_items, samp_valid, lse, scal, reinterpret_cast<__nv_bfloat16*>(dl_bf16));
  LTG_CHECK_LAUNCH();
  return LTG_OK;
}

extern "C" int ltg_enc_coef_scatter(const int32_t* e_row, const int32_t* e_item, const int32_t* e_slot, const int64_t* row_uid,
                                    const float* row_rnorm, int n_entries, int n_items, float keep, uint64_t seed, uint32_t step,
                                    const uint32_t* step_dev, void* xc_bf16, int ld_xc, void* stream) {
  LTG_REQUIRE(e_row && e_item && e_slot && row_uid && row_rnorm && xc_bf16);
  if (n_entries <= 0) return LTG_OK;
  enc_coef_scatter_kernel<<<(n_entries + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      e_row, e_item, e_slot, row_uid, row_rnorm, n_entries, n_items, keep, seed, step, step_dev, reinterpret_cast<__nv_bfloat16*>(xc_bf16), ld_xc);
  LTG_CHECK_LAUNCH();
  return LTG_OK;
}

extern "C" int ltg_enc_coef_clear(const int32_t* e_row, const int32_t* e_slot, int n_entries, void* xc_bf16, int ld_xc, void* stream) {
  LTG_REQUIRE(e_row && e_slot && xc_bf16);
  if (n_entries <= 0) return LTG_OK;
  enc_coef_clear_kernel<<<(n_entries + 255) / 256, 256, 0, (cudaStream_t)stream>>>(e_row, e_slot, n_entries, reinterpret_cast<__nv_bfloat16*>(xc_bf16), ld_xc);
  LTG_CHECK_LAUNCH();
  return LTG_OK;
}
