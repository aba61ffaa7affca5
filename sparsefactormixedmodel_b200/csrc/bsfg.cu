// libbsfg_b200: single translation unit (kernels are header-defined).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include "common.cuh"
#include "gemm_dmma.cuh"
#include "kernels.cuh"
#include "chol_blocked.cuh"
#include <algorithm>
#include <map>
#include <cmath>
#include <functional>

namespace bsfg {
static thread_local char g_err[1024] = "";
void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
const char* get_last_error() { return g_err; }
}  // namespace bsfg

extern "C" int bsfg_version(void) { return 100; }
extern "C" const char* bsfg_last_error(void) { return bsfg::get_last_error(); }
extern "C" int bsfg_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

#include "gemm_host.inl"
#include "stateless.inl"
#include "sweep.inl"
#include "init_device.inl"
