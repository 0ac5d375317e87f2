// Instantiations of the tcgen05 GEMM: the generic entry point and the decoder forward with the
// softmax-statistics epilogue.
#include "gemm_sm100.cuh"
#include "../../include/ltgan.h"

using namespace ltg;

template <int BN>
static int dispatch_major(const __nv_bfloat16* A, int lda, int a_mn, const __nv_bfloat16* B, int ldb, int b_mn, int M, int N, int K,
                          int splits, const EpiStore::Params& ep, cudaStream_t st) {
  if (!a_mn && !b_mn) return launch_gemm<BN, false, false, EpiStore>(A, lda, B, ldb, M, N, K, splits, ep, st);
  if (!a_mn && b_mn) return launch_gemm<BN, false, true, EpiStore>(A, lda, B, ldb, M, N, K, splits, ep, st);
  if (a_mn && !b_mn) return launch_gemm<BN, true, false, EpiStore>(A, lda, B, ldb, M, N, K, splits, ep, st);
  return launch_gemm<BN, true, true, EpiStore>(A, lda, B, ldb, M, N, K, splits, ep, st);
}

extern "C" int ltg_gemm_bf16(const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, int M, int N, int K, int splits, int bn,
                             float* out_f32, int ld_f32, void* out_bf16, int ld_bf16, const float* bias, int act, float alpha, int atomic,
                             float keep, uint64_t seed, uint32_t rng_stream, uint32_t rng_step, const uint32_t* rng_step_dev, int rng_ld,
                             int aux_col, float* aux_out, const void* dact_src, int dact_ld, float dact_keep, int64_t split_stride, void* stream) {
  LTG_REQUIRE(A != nullptr && B != nullptr);
  LTG_REQUIRE(out_f32 != nullptr || out_bf16 != nullptr);
  LTG_REQUIRE(!atomic || (out_f32 != nullptr && out_bf16 == nullptr && act == 0 && bias == nullptr));
  LTG_REQUIRE(splits <= 1 || atomic || (split_stride > 0 && out_f32 != nullptr && out_bf16 == nullptr && act == 0 && bias == nullptr));
  LTG_REQUIRE(aux_col < 0 || aux_out != nullptr);
  LTG_REQUIRE(dact_src == nullptr || dact_ld >= N);
  const bool drop = keep > 0.f && keep < 1.f;
  LTG_REQUIRE(!drop || (rng_ld % 4 == 0 && rng_ld >= N));
  EpiStore::Params ep;
  ep.out_f32 = out_f32; ep.ld_f32 = ld_f32;
  ep.out_bf16 = reinterpret_cast<__nv_bfloat16*>(out_bf16); ep.ld_bf16 = ld_bf16;
  ep.bias = bias; ep.act = act; ep.atomic = atomic; ep.alpha = alpha; ep.split_stride = splits > 1 ? split_stride : 0;
  ep.keep = keep; ep.seed = seed; ep.rng_stream = rng_stream; ep.rng_step = rng_step; ep.rng_step_dev = rng_step_dev; ep.rng_ld = rng_ld;
  ep.aux_col = aux_col; ep.aux_out = aux_out;
  ep.dact_src = reinterpret_cast<const __nv_bfloat16*>(dact_src); ep.dact_ld = dact_ld; ep.dact_keep = dact_keep;
  const __nv_bfloat16* a = reinterpret_cast<const __nv_bfloat16*>(A);
  const __nv_bfloat16* b = reinterpret_cast<const __nv_bfloat16*>(B);
  cudaStream_t st = (cudaStream_t)stream;
  switch (bn) {
    case 64: return dispatch_major<64>(a, lda, a_mn, b, ldb, b_mn, M, N, K, splits, ep, st);
    case 128: return dispatch_major<128>(a, lda, a_mn, b, ldb, b_mn, M, N, K, splits, ep, st);
    case 192: return dispatch_major<192>(a, lda, a_mn, b, ldb, b_mn, M, N, K, splits, ep, st);
    case 256: return dispatch_major<256>(a, lda, a_mn, b, ldb, b_mn, M, N, K, splits, ep, st);
    default:
      ltg_set_last_error("bn must be 64, 128, 192 or 256", __FILE__, __LINE__);
      return LTG_ERR_ARG;
  }
}

// Tile shape of the decoder forward. The kernel is epilogue-bound (ncu, round 2), so a tile costs about its width, and at batch 500
// the catalog gives only a few tiles per SM: what matters is how many waves of (cluster) tiles the grid needs. 256-wide tiles in
// clusters of four -- the round-1 choice -- put 79 cluster tiles on 33 clusters (4-CTA clusters reach 132 of the 148 SMs): 3 waves of
// 256 columns, the last one a third full; 192-wide tiles in CTA pairs are 210 cluster tiles on 74 pairs: 3 waves of 192.
// cost = waves x (BN + per-tile overhead); ties go to the wider tile.
static void logits_tile_shape(int B, int n_items, int* bn_out, int* cm_out) {
  const int m_blocks = (B + GEMM_BM - 1) / GEMM_BM;
  const int sms = ltg_num_sms();
  static const int cand[][2] = {{256, 4}, {256, 2}, {192, 2}, {128, 2}, {256, 1}, {192, 1}, {128, 1}};
  long best = -1;
  for (const auto& c : cand) {
    const int bn = c[0], cm = c[1];
    if (cm > 1 && m_blocks < cm) continue;            // no row blocks to share the B tile with
    if (cm == 1 && m_blocks > 1) continue;            // (clusters save L2 -> SM operand traffic whenever there is something to share)
    if (cm == 4 && m_blocks != 3 && m_blocks != 4) continue;
    const int clusters = cm == 4 ? sms * 33 / 148 : sms / cm;
    const long tiles = (long)((m_blocks + cm - 1) / cm) * ((n_items + bn - 1) / bn);
    const long cost = ((tiles + clusters - 1) / clusters) * (bn + 48);
    if (best < 0 || cost < best) { best = cost; *bn_out = bn; *cm_out = cm; }
  }
}

extern "C" int ltg_dec_logits_nblk(int B, int n_items) {
  if (B <= 0 || n_items <= 0) return 0;
  int bn = 256, cm = 1;
  logits_tile_shape(B, n_items, &bn, &cm);
  return 4 * ((n_items + bn - 1) / bn);
}

template <int BN>
static int launch_logits(int cm, const __nv_bfloat16* A, int lda, const __nv_bfloat16* Bm, int B, int n_items, const EpiLogitsStats::Params& ep,
                         cudaStream_t stream) {
  GemmShape s;
  s.M = B; s.N = n_items; s.K = LTG_H;
  s.m_blocks = (B + GEMM_BM - 1) / GEMM_BM;
  s.n_blocks = (n_items + BN - 1) / BN;
  s.k_blocks = (LTG_H + GEMM_BK - 1) / GEMM_BK;
  s.kb_per_split = s.k_blocks;
  s.splits = 1;
  if (cm == 4) return launch_gemm_cm<BN, false, false, 4, EpiLogitsStats>(A, lda, Bm, LTG_H, s, ep, stream);
  if (cm == 2) return launch_gemm_cm<BN, false, false, 2, EpiLogitsStats>(A, lda, Bm, LTG_H, s, ep, stream);
  return launch_gemm_cm<BN, false, false, 1, EpiLogitsStats>(A, lda, Bm, LTG_H, s, ep, stream);
}

extern "C" int ltg_dec_logits_fwd(const void* h2_bf16, int ld_h2, const void* WdT_bf16, const float* b_dec, int B, int n_items,
                                  void* logits_bf16, int ld_logits, float* partial, void* stream) {
  LTG_REQUIRE(h2_bf16 != nullptr && WdT_bf16 != nullptr && b_dec != nullptr && (partial != nullptr || logits_bf16 != nullptr));
  LTG_REQUIRE(logits_bf16 == nullptr || ld_logits >= n_items);
  if (B <= 0 || n_items <= 0) return LTG_OK;
  EpiLogitsStats::Params ep;
  ep.logits = reinterpret_cast<__nv_bfloat16*>(logits_bf16);
  ep.ld = ld_logits;
  ep.bias = b_dec;
  ep.partial = reinterpret_cast<float2*>(partial);
  int bn = 256, cm = 1;
  logits_tile_shape(B, n_items, &bn, &cm);
  const __nv_bfloat16* A = reinterpret_cast<const __nv_bfloat16*>(h2_bf16);
  const __nv_bfloat16* W = reinterpret_cast<const __nv_bfloat16*>(WdT_bf16);
  if (bn == 256) return launch_logits<256>(cm, A, ld_h2, W, B, n_items, ep, (cudaStream_t)stream);
  if (bn == 192) return launch_logits<192>(cm, A, ld_h2, W, B, n_items, ep, (cudaStream_t)stream);
  return launch_logits<128>(cm, A, ld_h2, W, B, n_items, ep, (cudaStream_t)stream);
}

// dW = A^T-layout GEMM (A stored [K][M], B stored [K][N]) with the Adam step fused into the epilogue; see EpiAdam.
extern "C" int ltg_wgrad_adam(const void* A, int lda, const void* B, int ldb, int M, int N, int K, float* p, float* m, float* v,
                              void* shadow_bf16, int ld, int n_cols, int aux_col, float* aux_out, float lr_t, const float* scal,
                              float beta1, float beta2, float eps, void* stream) {
  LTG_REQUIRE(A && B && p && m && v && shadow_bf16);
  LTG_REQUIRE(ld % 4 == 0 && n_cols % 4 == 0 && n_cols <= ld && n_cols <= N && (aux_col < 0 || aux_out != nullptr));
  LTG_REQUIRE(lr_t >= 0.f || scal != nullptr);
  LTG_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(shadow_bf16) & 7) == 0);
  EpiAdam::Params ep;
  ep.p = p; ep.m = m; ep.v = v; ep.shadow = reinterpret_cast<__nv_bfloat16*>(shadow_bf16); ep.ld = ld; ep.n_cols = n_cols;
  ep.aux_col = aux_col; ep.aux_out = aux_out; ep.lr_t = lr_t; ep.scal = scal; ep.b1 = beta1; ep.b2 = beta2; ep.eps = eps;
  return launch_gemm<128, true, true, EpiAdam>(reinterpret_cast<const __nv_bfloat16*>(A), lda, reinterpret_cast<const __nv_bfloat16*>(B), ldb,
                                               M, N, K, 1, ep, (cudaStream_t)stream);
}
