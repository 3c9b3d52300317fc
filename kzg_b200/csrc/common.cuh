// common.cuh -- library-internal plumbing: status codes, error text, launch counting.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include <atomic>
#include <condition_variable>
#include <mutex>

#include "../../include/kzg_b200.h"

namespace kzg {

// One in-flight API call.  A lane owns a stream (host-pointer entry points run on it) and the three arenas a call may
// use; a thread holds exactly one lane for the duration of a public call (nested calls share it), so calls from
// different host threads overlap on the device instead of queueing behind one process-wide mutex.  Lane 0 is always
// tried first: a single-threaded caller only ever touches one set of arenas.
struct Lane {
    cudaStream_t stream = nullptr;       // library-owned stream for host-pointer entry points
    cudaStream_t last_stream = nullptr;  // stream of the previous call that used this lane's arenas (see use_stream)
    bool have_last = false;
    bool busy = false;
    void* pinned = nullptr;   // pinned staging (host-pointer entry points)
    size_t pinned_bytes = 0;
    void* scratch = nullptr;  // device scratch arena (grown on demand, reused across calls)
    size_t scratch_bytes = 0;
    void* staging = nullptr;  // device copies of host-pointer inputs (outer level; scratch is the inner level)
    size_t staging_bytes = 0;
    // fork/join inside one call (independent sub-chains of a pipeline run beside each other, see batched.cu)
    cudaStream_t side = nullptr;
    cudaEvent_t fork_ev = nullptr, join_ev = nullptr;
    // relay ring of upload_async (pageable host inputs): one event per pinned slice, created on first use
    cudaEvent_t relay_ev[16] = {nullptr};
};
constexpr int KZG_MAX_LANES = 4;


struct Context {
    bool initialised = false;
    int device = -1;
    int sm_count = 0;
    std::mutex mu;                 // guards lanes[].busy and init/shutdown
    std::condition_variable cv;    // a lane was released
    Lane lanes[KZG_MAX_LANES];
    std::atomic<uint64_t> launches{0};
    // per-device caches shared by every lane (built once, then read-only): NTT twiddle tables and kernel attributes
    std::mutex cache_mu;
    void* ntt_tw[33] = {nullptr};
    bool ntt_attr_set = false;
    bool interp_attr_set = false;
};

Context& ctx();
Lane& lane();   // the calling thread's lane (valid inside an ApiScope)
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

// RAII guard of every public entry point: selects the library's device on the calling thread (entry points may be
// called from any host thread -- the reference's provers are Send + Sync, coeff_form.rs:37-41 -- and CUDA's current
// device is per thread), acquires a lane, and on exit restores the caller's device.  Nested entry points reuse the
// outer scope's lane.
struct ApiScope {
    int prev_device = -1;
    bool outer = false;
    cudaError_t err = cudaSuccess;
    ApiScope();
    ~ApiScope();
};

// Every entry point runs its work on ONE stream.  A lane's arenas (and an SRS handle's work areas, see msm.cuh) are
// reused between calls, so when consecutive users run on different streams the earlier stream is drained first.
cudaStream_t use_stream(cudaStream_t s);
int lane_side_stream(cudaStream_t* side, cudaEvent_t* fork_ev, cudaEvent_t* join_ev);   // created on first use
void* scratch_reserve(size_t bytes);   // returns device pointer or nullptr (error set)
void* staging_reserve(size_t bytes);   // device staging for host-pointer entry points
void* pinned_reserve(size_t bytes);    // returns pinned host pointer or nullptr (error set)
// Host -> device copy of a (possibly pageable) caller buffer, ordered on `st`.  Pinned or registered sources go
// straight to cudaMemcpyAsync; large pageable ones (what a Rust Vec<Fr> is) are relayed through the lane's pinned ring
// by a few host threads, so the PCIe transfer of one slice overlaps the host memcpy of the next ones.  Like a pageable
// cudaMemcpyAsync it returns once the source has been read.
cudaError_t upload_async(void* dst, const void* src, size_t bytes, cudaStream_t st);
// Device -> host: plain cudaMemcpyAsync for pinned destinations (synchronise `st` before reading), the threaded relay
// for large pageable ones (complete on return).
cudaError_t download_async(void* dst, const void* src, size_t bytes, cudaStream_t st);

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

// first statement (after KZG_REQUIRE_INIT) of every public entry point that touches the device
#define KZG_ENTER()                                                                 \
    ::kzg::ApiScope _api_scope;                                                     \
    if (_api_scope.err != cudaSuccess)                                              \
        return ::kzg::cuda_fail(_api_scope.err, "cudaSetDevice(library device)", __FILE__, __LINE__)

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
