// Runtime glue of libltgan.so: error reporting, driver entry points, device step state.
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include "gemm_sm100.cuh"
#include "../../include/ltgan.h"

static thread_local char g_err[512] = "";

void ltg_set_last_error(const char* msg, const char* file, int line) {
  snprintf(g_err, sizeof(g_err), "%s (%s:%d)", msg, file, line);
}

bool ltg_pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("LTG_PDL");
    v = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return v != 0;
}

namespace ltg {

PFN_encodeTiled ltg_get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (fn != nullptr) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || p == nullptr) {
    ltg_set_last_error("cuTensorMapEncodeTiled not available from the CUDA driver", __FILE__, __LINE__);
    return nullptr;
  }
  fn = reinterpret_cast<PFN_encodeTiled>(p);
  return fn;
}

int ltg_gemm_cluster_override() {
  static int v = -2;
  if (v == -2) {
    const char* e = getenv("LTG_GEMM_CM");
    v = e != nullptr ? atoi(e) : -1;
  }
  return v;
}

int ltg_num_sms() {
  static int n = 0;
  if (n > 0) return n;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  return n;
}

}  // namespace ltg

extern "C" const char* ltg_last_error(void) { return g_err; }
extern "C" int ltg_version(void) { return 100; }

extern "C" int ltg_init(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { ltg_set_last_error(cudaGetErrorString(e), __FILE__, __LINE__); return LTG_ERR_CUDA; }
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10) { ltg_set_last_error("libltgan requires an sm_100a (B200) device", __FILE__, __LINE__); return LTG_ERR_CUDA; }
  if (ltg::ltg_get_encode_tiled() == nullptr) return LTG_ERR_DRIVER;
  ltg::ltg_num_sms();
  return LTG_OK;
}

// words[0] rng step, words[1] Adam t, words[2] G-update count.
// Also clears one caller-chosen buffer (the phase's atomically accumulated gradients / counters): one memset launch less at the
// head of every phase's critical chain.
__global__ void step_advance_kernel(uint32_t* words, float* scal, int kind, float lr, double beta1, double beta2,
                                    float anneal_cap, float total_anneal_steps, uint32_t* zero_buf, int64_t zero_words,
                                    uint32_t* step_snapshot) {
  pdl_trigger();
  pdl_wait_cta();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < zero_words; i += (int64_t)gridDim.x * blockDim.x) zero_buf[i] = 0u;
  if (blockIdx.x != 0) return;
  if (threadIdx.x < 8) scal[threadIdx.x] = 0.f;   // per-step accumulators (KL, NLL, sum p, sum y, cnt, d_loss)
  if (threadIdx.x != 0) return;
  words[0] += 1;
  // a phase whose kernels may run beside another phase's (engine.run_step: the G forward beside the D update) reads its rng step
  // from its own word instead of the live counter
  if (step_snapshot != nullptr) *step_snapshot = words[0];
  if (kind >= 1) {
    const uint32_t t = ++words[1];
    // TF1 AdamOptimizer: lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t)   [ext, SURVEY F6]
    const double b1t = pow(beta1, (double)t), b2t = pow(beta2, (double)t);
    scal[LTG_S_LR_T] = (float)((double)lr * sqrt(1.0 - b2t) / (1.0 - b1t));
  }
  if (kind == 2) {
    // train.py:319-324: anneal from the pre-increment update_count
    const uint32_t c = words[2];
    float a = anneal_cap;
    if (total_anneal_steps > 0.f) a = fminf(anneal_cap, (float)c / total_anneal_steps);
    scal[LTG_S_ANNEAL] = a;
    words[2] = c + 1;
  }
}

// The three advances of one A -> D -> G step (engine.run_step) in ONE launch: state after it == three ltg_step_advance launches with
// kind 0, 1, 2 in this order (each phase gets its own scalar row, cleared buffer and rng-step snapshot). Two launches less on the
// critical chain of the step.
struct StepPhase { float* scal; uint32_t* zero_buf; int64_t zero_words; uint32_t* snapshot; };
__global__ void step_advance3_kernel(uint32_t* words, StepPhase pa, StepPhase pd, StepPhase pg, float lr, double beta1, double beta2,
                                     float anneal_cap, float total_anneal_steps) {
  pdl_trigger();
  pdl_wait_cta();
  const StepPhase ph[3] = {pa, pd, pg};
#pragma unroll
  for (int q = 0; q < 3; ++q)
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ph[q].zero_words; i += (int64_t)gridDim.x * blockDim.x)
      ph[q].zero_buf[i] = 0u;
  if (blockIdx.x != 0) return;
  if (threadIdx.x < 24) ph[threadIdx.x >> 3].scal[threadIdx.x & 7] = 0.f;
  if (threadIdx.x != 0) return;
  uint32_t step = words[0], t = words[1];
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    ++step;
    if (ph[q].snapshot != nullptr) *ph[q].snapshot = step;
    if (q >= 1) {
      ++t;
      const double b1t = pow(beta1, (double)t), b2t = pow(beta2, (double)t);
      ph[q].scal[LTG_S_LR_T] = (float)((double)lr * sqrt(1.0 - b2t) / (1.0 - b1t));
    }
  }
  words[0] = step; words[1] = t;
  const uint32_t c = words[2];   // train.py:319-324: anneal from the pre-increment update_count
  float a = anneal_cap;
  if (total_anneal_steps > 0.f) a = fminf(anneal_cap, (float)c / total_anneal_steps);
  pg.scal[LTG_S_ANNEAL] = a;
  words[2] = c + 1;
}

extern "C" int ltg_step_advance3(uint32_t* words, float* scal_a, float* scal_d, float* scal_g, float lr, float beta1, float beta2,
                                 float anneal_cap, float total_anneal_steps, void* zero_a, int64_t zero_a_words, void* zero_d,
                                 int64_t zero_d_words, void* zero_g, int64_t zero_g_words, uint32_t* snap_a, uint32_t* snap_d,
                                 uint32_t* snap_g, void* stream) {
  LTG_REQUIRE(words != nullptr && scal_a != nullptr && scal_d != nullptr && scal_g != nullptr);
  LTG_REQUIRE(scal_a != scal_d && scal_a != scal_g && scal_d != scal_g);
  LTG_REQUIRE(zero_a_words >= 0 && zero_d_words >= 0 && zero_g_words >= 0);
  LTG_REQUIRE((zero_a_words == 0 || zero_a != nullptr) && (zero_d_words == 0 || zero_d != nullptr) && (zero_g_words == 0 || zero_g != nullptr));
  LTG_REQUIRE(((reinterpret_cast<uintptr_t>(zero_a) | reinterpret_cast<uintptr_t>(zero_d) | reinterpret_cast<uintptr_t>(zero_g)) & 3) == 0);
  StepPhase pa = {scal_a, reinterpret_cast<uint32_t*>(zero_a), zero_a_words, snap_a};
  StepPhase pd = {scal_d, reinterpret_cast<uint32_t*>(zero_d), zero_d_words, snap_d};
  StepPhase pg = {scal_g, reinterpret_cast<uint32_t*>(zero_g), zero_g_words, snap_g};
  int64_t mx = zero_a_words > zero_d_words ? zero_a_words : zero_d_words;
  if (zero_g_words > mx) mx = zero_g_words;
  int64_t blocks = (mx + 1023) / 1024;
  if (blocks < 1) blocks = 1;
  if (blocks > 148) blocks = 148;
  ltg_launch(step_advance3_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, words, pa, pd, pg, lr, (double)beta1, (double)beta2, anneal_cap,
                                                                           total_anneal_steps);
  LTG_CHECK_LAUNCH();
  return LTG_OK;
}

extern "C" int ltg_step_advance(uint32_t* words, float* scal, int kind, float lr, float beta1, float beta2,
                                float anneal_cap, float total_anneal_steps, void* zero_buf, int64_t zero_words, uint32_t* step_snapshot,
                                void* stream) {
  LTG_REQUIRE(words != nullptr && scal != nullptr && zero_words >= 0 && (zero_words == 0 || zero_buf != nullptr));
  LTG_REQUIRE((reinterpret_cast<uintptr_t>(zero_buf) & 3) == 0);
  int64_t blocks = (zero_words + 1023) / 1024;   // 256 threads x 4 words each
  if (blocks < 1) blocks = 1;
  if (blocks > 148) blocks = 148;
  ltg_launch(step_advance_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, words, scal, kind, lr, (double)beta1, (double)beta2, anneal_cap,
                                                                          total_anneal_steps, reinterpret_cast<uint32_t*>(zero_buf), zero_words, step_snapshot);
  LTG_CHECK_LAUNCH();
  return LTG_OK;
}
