// vi_api.cu - library-level entry points: init, version, thread-local error string.
#include "vi_common.cuh"

#include <stdarg.h>
#include <stdlib.h>

static thread_local char g_err[512] = "";
static int g_num_sms = 0;

void vi_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int vi_num_sms() {
  if (g_num_sms == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      g_num_sms = n;
    else
      g_num_sms = 148;
  }
  return g_num_sms;
}

bool vi_pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("VI_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}

int vi_attn_init();

extern "C" int vi_version(void) { return 100; }   // 0.1.0

extern "C" const char* vi_last_error(void) { return g_err; }

extern "C" int vi_init(int device) {
  VI_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  VI_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    vi_set_error("libvlnimagine is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
    return VI_ERR_UNSUPPORTED;
  }
  g_num_sms = prop.multiProcessorCount;
  if (int rc = vi_attn_init()) return rc;
  return VI_OK;
}
