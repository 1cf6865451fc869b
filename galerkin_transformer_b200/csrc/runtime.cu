// Library-level plumbing: error strings, device selection, launch accounting.
#include <atomic>
#include <cstdarg>

#include "common.cuh"

namespace gb200 {

static thread_local char g_err[512] = "";
static thread_local int g_device = -1;
static std::atomic<unsigned long long> g_launches{0};
static const unsigned long long* g_rng_offset = nullptr;

const unsigned long long* rng_offset_ptr() { return g_rng_offset; }

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void use_device(int device) {
    if (device >= 0 && device != g_device) {
        cudaSetDevice(device);
        g_device = device;
    }
}

int check_launch(const char* what, int nkernels) {
    g_launches.fetch_add((unsigned long long)nkernels, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
        return GB200_ERR_CUDA;
    }
    return GB200_OK;
}

}  // namespace gb200

extern "C" int gb200_version(void) { return 101; }
extern "C" int gb200_set_rng_offset_ptr(const unsigned long long* device_counter) {
    gb200::g_rng_offset = device_counter;
    return GB200_OK;
}
extern "C" const char* gb200_last_error(void) { return gb200::g_err; }
extern "C" unsigned long long gb200_launch_count(void) { return gb200::g_launches.load(); }
