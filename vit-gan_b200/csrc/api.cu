// api.cu -- C-ABI glue: error state, device queries, vg_gemm dispatch.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace vg {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    (void)cudaGetLastError();
    return VG_ERR_LAUNCH;
  }
  return VG_OK;
}

bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("VG_PDL"); return !(e && e[0] == '0'); }();   // read once, thread-safe
  return on;
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace vg

using namespace vg;

extern "C" int vg_version(void) { return VG_ABI_VERSION; }
extern "C" const char* vg_last_error(void) { return g_err; }

extern "C" int vg_device_is_sm100(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10;
}

extern "C" int vg_gemm(const vg_gemm_args* args, void* stream) {
  VG_REQUIRE(args != nullptr, VG_ERR_ARG, "vg_gemm: NULL args");
  const vg_gemm_args& a = *args;
  VG_REQUIRE(a.A && a.B && a.C, VG_ERR_ARG, "vg_gemm: NULL operand");
  cudaStream_t st = as_stream(stream);
  if (a.path == VG_GEMM_SIMT) {
    VG_REQUIRE(!a.a_rowsum && !a.ln_gamma, VG_ERR_UNSUPPORTED, "vg_gemm(simt): a_rowsum / fused LayerNorm are tcgen05-path features");
    return gemm_simt_launch(a, st);
  }
  const char* why = "";
  const bool ok = gemm_tc_supported(a, &why);
  if (a.path == VG_GEMM_TCGEN05) {
    VG_REQUIRE(ok, VG_ERR_UNSUPPORTED, "vg_gemm(tcgen05): unsupported problem: %s", why);
    return gemm_tc_launch(a, st);
  }
  // AUTO: tensor cores for bf16 problems the tcgen05 kernel accepts, CUDA cores otherwise (fp32 parity path,
  // tiny heads).  Both are this library's own sm_100a kernels -- there is no CPU or vendor-library fallback.
  VG_REQUIRE(ok || (!a.a_rowsum && !a.ln_gamma), VG_ERR_UNSUPPORTED, "vg_gemm: a_rowsum / fused LayerNorm unsupported here: %s", why);
  return ok ? gemm_tc_launch(a, st) : gemm_simt_launch(a, st);
}

namespace vg { void attention_mt_set_trace(unsigned long long* p); }
extern "C" int vg_attention_set_trace(void* buffer) {
  vg::attention_mt_set_trace(static_cast<unsigned long long*>(buffer));
  return VG_OK;
}
extern "C" int vg_gemm_set_trace(void* buffer) {
  gemm_tc_set_trace(static_cast<unsigned long long*>(buffer));
  return VG_OK;
}
