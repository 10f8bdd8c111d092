// lib.cu — library state: init, error string, launch counter.
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>
#include <mutex>
#include "common.cuh"

namespace isb {

static thread_local char t_err[512] = "";
std::atomic<uint64_t> g_launches{0};
static std::atomic<bool> g_init{false};
static int g_num_sms = 148;
static tensormap_encode_fn g_encode = nullptr;
static std::mutex g_mu;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
}
bool is_initialised() { return g_init.load(); }
// ISB_PDL: 0 = off, 1 = every kernel, 2 (default) = conv_tc only.  Measured on B200 (profiles/): making every
// kernel a programmatic dependent slows the memory-bound kernels down (their early CTAs squat on SMs while
// the predecessor drains); the conv kernel alone profits — it streams weight tiles during the wait.
static int pdl_mode() {
  static const int mode = [] {
    const char* e = getenv("ISB_PDL");
    return e ? atoi(e) : 2;
  }();
  return mode;
}
thread_local int t_pdl_family = 4;
bool pdl_enabled_family(int family) {
  static const int mask = [] {
    const char* e = getenv("ISB_PDL_MASK");
    return e ? atoi(e) : 0;
  }();
  return ((mask >> family) & 1) != 0;
}
bool pdl_enabled() { return pdl_mode() == 1; }
bool pdl_enabled_conv() { return pdl_mode() >= 1; }
int num_sms() { return g_num_sms; }
tensormap_encode_fn get_tensormap_encode() { return g_encode; }

int conv_tc_init();     // conv_tc.cu: raise dynamic smem limit
void conv_tc_set_trace(void* ptr);
int decode_init();      // decode.cu
int mc_init();          // mc.cu: case table -> constant memory
int attention_flash_init();   // attention_flash.cu

// profiling only: a chain of n dependent, (almost) empty kernels — the floor of one dependent launch in a stream /
// graph, with and without programmatic dependent launch (tools/launch_floor.py)
__global__ void debug_chain_kernel(float* p) {
  pdl_wait();
  pdl_trigger();
  if (p != nullptr && threadIdx.x == 0 && blockIdx.x == 0) p[0] += 1.0f;
}
int debug_launch_chain(int n, int ctas, int threads, int pdl, cudaStream_t st) {
  for (int i = 0; i < n; ++i) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas);
    cfg.blockDim = dim3(threads);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    ISB_CUDA(cudaLaunchKernelEx(&cfg, debug_chain_kernel, static_cast<float*>(nullptr)));
  }
  return ISB_OK;
}

}  // namespace isb

extern "C" {

int isb_abi_version(void) { return ISB_ABI_VERSION; }

const char* isb_last_error(void) { return isb::t_err; }

uint64_t isb_launch_count(void) { return isb::g_launches.load(); }

void isb_debug_set_trace(void* device_buffer) { isb::conv_tc_set_trace(device_buffer); }

int isb_debug_launch_chain(int n, int ctas, int threads, int pdl, isb_stream_t stream) {
  return isb::debug_launch_chain(n, ctas, threads, pdl, isb::as_stream(stream));
}

int isb_init(int device) {
  std::lock_guard<std::mutex> lk(isb::g_mu);
  ISB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  ISB_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    isb::set_error("isb_init: device %d is sm_%d%d; this library is built for sm_100a (B200) only",
                   device, prop.major, prop.minor);
    return ISB_ERR_CUDA;
  }
  isb::g_num_sms = prop.multiProcessorCount;
  if (!isb::g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    ISB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (q != cudaDriverEntryPointSuccess || fn == nullptr) {
      isb::set_error("isb_init: cuTensorMapEncodeTiled not available from the driver");
      return ISB_ERR_CUDA;
    }
    isb::g_encode = reinterpret_cast<isb::tensormap_encode_fn>(fn);
  }
  int rc = isb::conv_tc_init();
  if (rc) return rc;
  rc = isb::mc_init();
  if (rc) return rc;
  rc = isb::decode_init();
  if (rc) return rc;
  rc = isb::attention_flash_init();
  if (rc) return rc;
  isb::g_init.store(true);
  return ISB_OK;
}

}  // extern "C"
