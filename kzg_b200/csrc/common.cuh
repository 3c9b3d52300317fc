// common.cuh -- library-internal plumbing: status codes, error text, launch counting.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include <mutex>

#include "../../include/kzg_b200.h"

namespace kzg {

struct Context {
    bool initialised = false;
    int device = -1;
    int sm_count = 0;
    cudaStream_t stream = nullptr;  // library-owned stream for host-pointer entry points
    cudaStream_t last_stream = nullptr;  // stream of the previous API call (arenas are shared: see use_stream)
    bool have_last = false;
    std::recursive_mutex mu;
    uint64_t launches = 0;
    // pinned staging (host-pointer entry points)
    void* pinned = nullptr;
    size_t pinned_bytes = 0;
    // device scratch arena (grown on demand, reused across calls)
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
    // device staging arena: device copies of host-pointer inputs (outer level; scratch is the inner level)
    void* staging = nullptr;
    size_t staging_bytes = 0;
};

Context& ctx();
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
// Every entry point runs its work on ONE stream.  The scratch/staging arenas and per-SRS work areas are shared
// between calls, so when consecutive calls use different streams the earlier stream is drained first.
cudaStream_t use_stream(cudaStream_t s);
void* scratch_reserve(size_t bytes);   // returns device pointer or nullptr (error set)
void* staging_reserve(size_t bytes);   // device staging for host-pointer entry points
void* pinned_reserve(size_t bytes);    // returns pinned host pointer or nullptr (error set)

#define KZG_CUDA(expr)                                                              \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess) return ::kzg::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)

#define KZG_REQUIRE_INIT()                                                          \
    do {                                                                            \
        if (!::kzg::ctx().initialised) {                                            \
            ::kzg::set_error("kzgb_init() has not been called");                    \
            return KZGB_ERR_NOT_INITIALISED;                                        \
        }                                                                           \
    } while (0)

// count + check a kernel launch
#define KZG_LAUNCHED()                                                              \
    do {                                                                            \
        ::kzg::ctx().launches++;                                                    \
        ::kzg::trace_mark(__FILE__, __LINE__);                                      \
        KZG_CUDA(cudaGetLastError());                                               \
    } while (0)

// ---- stage timer: CUDA events around each pipeline stage, on the launching stream ----
// (bench.py's roofline needs the dominant kernel's duration measured live, outside any profiler)
enum Stage {
    ST_MSM_DIGITS = 0, ST_MSM_SORT, ST_MSM_ACCUMULATE, ST_MSM_FIXUP, ST_MSM_BUCKET_REDUCE, ST_MSM_FINISH,
    ST_NTT, ST_SCAN, ST_DIV_OMEGA, ST_INTERP, ST_BATCH_DIV, ST_COUNT
};
void stage_begin(int stage, cudaStream_t stream);
// per-launch trace (diagnostics): when enabled, every KZG_LAUNCHED() records an event on the call's stream, so the
// warm, in-stream duration of each kernel of a pipeline can be read back (ncu launch lists are cold and serialised)
void trace_mark(const char* file, int line);
void stage_end(int stage, cudaStream_t stream);
struct StageScope {
    int stage; cudaStream_t stream;
    StageScope(int s, cudaStream_t st) : stage(s), stream(st) { stage_begin(s, st); }
    ~StageScope() { stage_end(stage, stream); }
};

static inline unsigned div_up(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

}  // namespace kzg
