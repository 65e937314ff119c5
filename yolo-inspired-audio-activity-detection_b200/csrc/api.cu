// Library state: error string, device check, driver entry points.
#include "common.cuh"
#include <cuda.h>
#include <mutex>
#include <string.h>
#include <stdlib.h>

namespace yad {

static thread_local char g_err[512] = "";
static void* g_encode = nullptr;
static int g_sm_count = 0;
static int g_inited_device = -1;
static std::mutex g_mu;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void* tensor_map_encode_fn() { return g_encode; }
int sm_count() { return g_sm_count; }
int pdl_mode() {
  static const int mode = [] {
    const char* e = getenv("YAD_PDL");
    return e ? atoi(e) : 1;
  }();
  return mode;
}

int init_conv_tc_attrs();   // conv_tc.cu
int init_conv_flat_attrs(); // conv_flat.cu
int init_frontend_attrs();  // frontend.cu
int init_conv_stem_tc_attrs();  // conv_stem_tc.cu
int init_conv_stem_fused_attrs();  // conv_stem_fused.cu
int init_conv_tf32_attrs();  // conv_tf32.cu
int init_neck_fused_attrs(); // neck_fused.cu

}  // namespace yad

extern "C" {

int yad_version(void) { return 100; }

const char* yad_last_error(void) { return yad::g_err; }

int yad_init(int device) {
  std::lock_guard<std::mutex> lk(yad::g_mu);
  if (yad::g_inited_device == device) return YAD_OK;
  int n = 0;
  YAD_CUDA(cudaGetDeviceCount(&n));
  YAD_CHECK_ARG(device >= 0 && device < n, "yad_init: device %d out of range (%d visible)", device, n);
  cudaDeviceProp prop;
  YAD_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    yad::set_error("yad_init: device %d is sm_%d%d; this library contains sm_100a code only (no fallback)",
                   device, prop.major, prop.minor);
    return YAD_ERR_ARCH;
  }
  YAD_CUDA(cudaSetDevice(device));
  yad::g_sm_count = prop.multiProcessorCount;
  cudaDriverEntryPointQueryResult q;
  void* fn = nullptr;
  YAD_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (q != cudaDriverEntryPointSuccess || fn == nullptr) {
    yad::set_error("yad_init: cuTensorMapEncodeTiled not available from the driver");
    return YAD_ERR_CUDA;
  }
  yad::g_encode = fn;
  int rc = yad::init_conv_tc_attrs();
  if (rc) return rc;
  rc = yad::init_conv_flat_attrs();
  if (rc) return rc;
  rc = yad::init_conv_stem_tc_attrs();
  if (rc) return rc;
  rc = yad::init_conv_stem_fused_attrs();
  if (rc) return rc;
  rc = yad::init_frontend_attrs();
  if (rc) return rc;
  rc = yad::init_conv_tf32_attrs();
  if (rc) return rc;
  rc = yad::init_neck_fused_attrs();
  if (rc) return rc;
  yad::g_inited_device = device;
  return YAD_OK;
}

}  // extern "C"
