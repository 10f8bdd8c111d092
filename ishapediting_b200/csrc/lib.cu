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
bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("ISB_PDL");
    return e && e[0] == '1';   // measured neutral under CUDA-graph replay (profiles/): opt-in
  }();
  return on;
}
int num_sms() { return g_num_sms; }
tensormap_encode_fn get_tensormap_encode() { return g_encode; }

int conv_tc_init();     // conv_tc.cu: raise dynamic smem limit
int decode_init();      // decode.cu

}  // namespace isb

extern "C" {

int isb_abi_version(void) { return ISB_ABI_VERSION; }

const char* isb_last_error(void) { return isb::t_err; }

uint64_t isb_launch_count(void) { return isb::g_launches.load(); }

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
  rc = isb::decode_init();
  if (rc) return rc;
  isb::g_init.store(true);
  return ISB_OK;
}

}  // extern "C"
