// api_core.cu -- lifecycle, error reporting, scratch/pinned arenas.
#include <cstring>

#include "common.cuh"

namespace kzg {

static thread_local char g_err[512] = "";

Context& ctx() {
    static Context c;
    return c;
}

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    return KZGB_ERR_CUDA;
}

void* scratch_reserve(size_t bytes) {
    Context& c = ctx();
    if (bytes <= c.scratch_bytes) return c.scratch;
    if (c.scratch) {
        cudaStreamSynchronize(c.stream);
        cudaFree(c.scratch);
        c.scratch = nullptr;
        c.scratch_bytes = 0;
    }
    size_t want = bytes + (bytes >> 2);
    cudaError_t e = cudaMalloc(&c.scratch, want);
    if (e != cudaSuccess) {
        cuda_fail(e, "cudaMalloc(scratch)", __FILE__, __LINE__);
        return nullptr;
    }
    c.scratch_bytes = want;
    return c.scratch;
}

void* pinned_reserve(size_t bytes) {
    Context& c = ctx();
    if (bytes <= c.pinned_bytes) return c.pinned;
    if (c.pinned) {
        cudaStreamSynchronize(c.stream);
        cudaFreeHost(c.pinned);
        c.pinned = nullptr;
        c.pinned_bytes = 0;
    }
    cudaError_t e = cudaMallocHost(&c.pinned, bytes);
    if (e != cudaSuccess) {
        cuda_fail(e, "cudaMallocHost(staging)", __FILE__, __LINE__);
        return nullptr;
    }
    c.pinned_bytes = bytes;
    return c.pinned;
}

}  // namespace kzg

using namespace kzg;

extern "C" {

int kzgb_init(int device) {
    Context& c = ctx();
    std::lock_guard<std::recursive_mutex> g(c.mu);
    if (c.initialised) {
        if (c.device == device) return KZGB_OK;
        set_error("kzgb_init: already initialised on device %d", c.device);
        return KZGB_ERR_INVALID_ARG;
    }
    int count = 0;
    KZG_CUDA(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) {
        set_error("kzgb_init: device %d out of range (%d visible)", device, count);
        return KZGB_ERR_INVALID_ARG;
    }
    KZG_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    KZG_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("kzgb_init: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                  prop.minor);
        return KZGB_ERR_INVALID_ARG;
    }
    c.sm_count = prop.multiProcessorCount;
    KZG_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    c.device = device;
    c.initialised = true;
    c.launches = 0;
    return KZGB_OK;
}

int kzgb_shutdown(void) {
    Context& c = ctx();
    std::lock_guard<std::recursive_mutex> g(c.mu);
    if (!c.initialised) return KZGB_OK;
    cudaStreamSynchronize(c.stream);
    if (c.scratch) cudaFree(c.scratch);
    if (c.pinned) cudaFreeHost(c.pinned);
    cudaStreamDestroy(c.stream);
    c.scratch = nullptr;
    c.scratch_bytes = 0;
    c.pinned = nullptr;
    c.pinned_bytes = 0;
    c.stream = nullptr;
    c.initialised = false;
    return KZGB_OK;
}

const char* kzgb_last_error_message(void) { return g_err; }

uint64_t kzgb_launch_count(void) { return ctx().launches; }

}  // extern "C"
