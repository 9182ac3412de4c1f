// api.cu -- error reporting, device cache and parameter-layout queries of libgcbf_b200.so.
#include <atomic>
#include <mutex>
#include <stdarg.h>

#include "common.cuh"

namespace gcbf {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};
static std::once_flag g_dev_once;
static int g_sm_count = 148;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int32_t check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return 0;
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int32_t)e;
}

void count_launch(int64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count() {
    std::call_once(g_dev_once, [] {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            g_sm_count = n;
    });
    return g_sm_count;
}

}  // namespace gcbf

extern "C" __attribute__((visibility("default"))) const char* gcbf_last_error_string(void) { return gcbf::g_err; }
extern "C" __attribute__((visibility("default"))) int32_t gcbf_version(void) { return 100; }
extern "C" __attribute__((visibility("default"))) int64_t gcbf_launch_count(void) { return gcbf::g_launches.load(); }

extern "C" __attribute__((visibility("default"))) int32_t gcbf_param_count(int32_t edge_dim, int32_t out_dim) {
    if (edge_dim < 1 || edge_dim > 6 || out_dim < 1 || out_dim > 4) return -1;
    return gcbf::make_layout(edge_dim, out_dim).total;
}

extern "C" __attribute__((visibility("default"))) int32_t gcbf_param_offsets(int32_t edge_dim, int32_t out_dim, int32_t* off) {
    if (edge_dim < 1 || edge_dim > 6 || out_dim < 1 || out_dim > 4 || !off) {
        gcbf::set_error("gcbf_param_offsets: bad argument");
        return -1;
    }
    gcbf::ParamLayout L = gcbf::make_layout(edge_dim, out_dim);
    for (int i = 0; i < 12; ++i) {
        off[2 * i] = L.w[i];
        off[2 * i + 1] = L.b[i];
    }
    return 0;
}
