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

extern "C" int ltg_step_advance(uint32_t* words, float* scal, int kind, float lr, float beta1, float beta2,
                                float anneal_cap, float total_anneal_steps, void* zero_buf, int64_t zero_words, uint32_t* step_snapshot,
                                void* stream) {
  LTG_REQUIRE(words != nullptr && scal != nullptr && zero_words >= 0 && (zero_words == 0 || zero_buf != nullptr));
  LTG_REQUIRE((reinterpret_cast<uintptr_t>(zero_buf) & 3) == 0);
  int64_t blocks = (zero_words + 1023) / 1024;   // 256 threads x 4 words each
  if (blocks < 1) blocks = 1;
  if (blocks > 148) blocks = 148;
  step_advance_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(words, scal, kind, lr, (double)beta1, (double)beta2, anneal_cap,
                                                                          total_anneal_steps, reinterpret_cast<uint32_t*>(zero_buf), zero_words, step_snapshot);
  LTG_CHECK_LAUNCH();
  return LTG_OK;
}
