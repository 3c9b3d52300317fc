// api_core.cu -- lifecycle, error reporting, scratch/pinned arenas.
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "common.cuh"
#include "msm.cuh"
#include "ntt.cuh"

namespace kzg {

static thread_local char g_err[512] = "";

Context& ctx() {
    static Context c;
    return c;
}

static thread_local Lane* t_lane = nullptr;
static thread_local int t_depth = 0;

Lane& lane() {
    // outside an ApiScope (library-internal misuse) fall back to lane 0 rather than dereferencing null
    return t_lane ? *t_lane : ctx().lanes[0];
}

ApiScope::ApiScope() {
    Context& c = ctx();
    if (t_depth++ > 0) return;   // nested entry point: the outer scope holds the lane and the device
    outer = true;
    if (cudaGetDevice(&prev_device) != cudaSuccess) prev_device = -1;
    if (prev_device != c.device) err = cudaSetDevice(c.device);
    std::unique_lock<std::mutex> g(c.mu);
    for (;;) {
        for (int i = 0; i < KZG_MAX_LANES; i++) {
            if (!c.lanes[i].busy) {
                c.lanes[i].busy = true;
                t_lane = &c.lanes[i];
                break;
            }
        }
        if (t_lane) break;
        c.cv.wait(g);
    }
    if (!t_lane->stream && err == cudaSuccess) err = cudaStreamCreateWithFlags(&t_lane->stream, cudaStreamNonBlocking);
}

ApiScope::~ApiScope() {
    if (--t_depth > 0 || !outer) return;
    Context& c = ctx();
    {
        std::lock_guard<std::mutex> g(c.mu);
        if (t_lane) t_lane->busy = false;
        t_lane = nullptr;
    }
    c.cv.notify_one();
    if (prev_device >= 0 && prev_device != c.device) cudaSetDevice(prev_device);
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

cudaStream_t use_stream(cudaStream_t s) {
    Lane& l = lane();
    if (l.have_last && l.last_stream != s) cudaStreamSynchronize(l.last_stream);
    l.last_stream = s;
    l.have_last = true;
    return s;
}

int lane_side_stream(cudaStream_t* side, cudaEvent_t* fork_ev, cudaEvent_t* join_ev) {
    Lane& l = lane();
    if (!l.side) {
        KZG_CUDA(cudaStreamCreateWithFlags(&l.side, cudaStreamNonBlocking));
        KZG_CUDA(cudaEventCreateWithFlags(&l.fork_ev, cudaEventDisableTiming));
        KZG_CUDA(cudaEventCreateWithFlags(&l.join_ev, cudaEventDisableTiming));
    }
    *side = l.side;
    *fork_ev = l.fork_ev;
    *join_ev = l.join_ev;
    return KZGB_OK;
}

// Arena growth frees the old block: everything that may still be reading it (the lane's last stream) is drained first.
static void lane_drain(Lane& l) {
    if (l.have_last) cudaStreamSynchronize(l.last_stream);
    if (l.stream) cudaStreamSynchronize(l.stream);
}

void* scratch_reserve(size_t bytes) {
    Lane& c = lane();
    if (bytes <= c.scratch_bytes) return c.scratch;
    if (c.scratch) {
        lane_drain(c);
        cudaFree(c.scratch);
        c.scratch = nullptr;
        c.scratch_bytes = 0;
    }
    // never smaller than 4 MiB: growing the arena synchronises and goes through cudaFree/cudaMalloc
    // (measured: up to 0.5 s with several processes on one box), so tiny requests must not trigger it repeatedly
    size_t want = bytes + (bytes >> 2);
    if (want < ((size_t)4 << 20)) want = (size_t)4 << 20;
    cudaError_t e = cudaMalloc(&c.scratch, want);
    if (e != cudaSuccess) {
        cuda_fail(e, "cudaMalloc(scratch)", __FILE__, __LINE__);
        return nullptr;
    }
    c.scratch_bytes = want;
    return c.scratch;
}

void* staging_reserve(size_t bytes) {
    Lane& c = lane();
    if (bytes <= c.staging_bytes) return c.staging;
    if (c.staging) {
        lane_drain(c);
        cudaFree(c.staging);
        c.staging = nullptr;
        c.staging_bytes = 0;
    }
    cudaError_t e = cudaMalloc(&c.staging, bytes);
    if (e != cudaSuccess) {
        cuda_fail(e, "cudaMalloc(staging)", __FILE__, __LINE__);
        return nullptr;
    }
    c.staging_bytes = bytes;
    return c.staging;
}

void* pinned_reserve(size_t bytes) {
    Lane& c = lane();
    if (bytes <= c.pinned_bytes) return c.pinned;
    if (c.pinned) {
        lane_drain(c);
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

// ---- host -> device upload with a pinned relay for pageable sources ----
namespace {
constexpr size_t RELAY_SLICE_MAX = size_t(2) << 20;   // bytes per pinned ring slot
constexpr int RELAY_MAX_THREADS = 8;              // 2 slices per thread
constexpr size_t RELAY_MIN_BYTES = size_t(4) << 20;

int relay_threads() {
    static int n = [] {
        const char* e = getenv("KZGB_UPLOAD_THREADS");   // 0 = leave pageable copies to the driver
        int v = e ? atoi(e) : 4;
        return v < 0 ? 0 : (v > RELAY_MAX_THREADS ? RELAY_MAX_THREADS : v);
    }();
    return n;
}

// bytes per relayed slice (shorter slices pay more per-copy overhead than they save in pipeline fill)
size_t relay_slice() {
    static size_t n = [] {
        const char* e = getenv("KZGB_UPLOAD_SLICE_KB");
        size_t v = e ? (size_t)atol(e) << 10 : size_t(1024) << 10;   // 128 KiB .. 2 MiB measured: profiles/r02s_pageable_upload_relay.log
        return v < 4096 ? size_t(4096) : (v > RELAY_SLICE_MAX ? RELAY_SLICE_MAX : v);
    }();
    return n;
}

bool host_is_pageable(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}
}  // namespace

cudaError_t upload_async(void* dst, const void* src, size_t bytes, cudaStream_t st) {
    const int T = relay_threads();
    if (bytes < RELAY_MIN_BYTES || T == 0 || !host_is_pageable(src))
        return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st);
    Lane& l = lane();
    char* ring = (char*)pinned_reserve(size_t(2 * RELAY_MAX_THREADS) * RELAY_SLICE_MAX);
    if (!ring) return cudaErrorMemoryAllocation;
    for (int i = 0; i < 2 * T; i++)
        if (!l.relay_ev[i]) {
            cudaError_t e = cudaEventCreateWithFlags(&l.relay_ev[i], cudaEventDisableTiming);
            if (e != cudaSuccess) return e;
        }
    const size_t RELAY_SLICE = relay_slice();
    const size_t slices = (bytes + RELAY_SLICE - 1) / RELAY_SLICE;
    const int device = ctx().device;
    cudaError_t errs[RELAY_MAX_THREADS];
    // worker t relays slices t, t+T, t+2T, ... through its own two pinned slots
    auto work = [&](int t) {
        cudaError_t e = cudaSetDevice(device);
        size_t j = 0;
        for (size_t k = t; k < slices && e == cudaSuccess; k += T, j++) {
            const int slot = 2 * t + int(j & 1);
            char* buf = ring + size_t(slot) * RELAY_SLICE;
            const size_t off = k * RELAY_SLICE, len = bytes - off < RELAY_SLICE ? bytes - off : RELAY_SLICE;
            e = cudaEventSynchronize(l.relay_ev[slot]);   // the slot's previous transfer (this call's or an earlier one's) has left
            if (e != cudaSuccess) break;
            memcpy(buf, (const char*)src + off, len);
            e = cudaMemcpyAsync((char*)dst + off, buf, len, cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) e = cudaEventRecord(l.relay_ev[slot], st);
        }
        errs[t] = e;
    };
    // a thread that cannot be started (resource limits) leaves its share to the calling thread: never unwind through the C ABI
    std::thread th[RELAY_MAX_THREADS];
    bool started[RELAY_MAX_THREADS] = {false};
    for (int t = 1; t < T; t++) {
        try {
            th[t] = std::thread(work, t);
            started[t] = true;
        } catch (...) {
            started[t] = false;
        }
    }
    work(0);
    for (int t = 1; t < T; t++) {
        if (started[t]) th[t].join();
        else work(t);
    }
    for (int t = 0; t < T; t++)
        if (errs[t] != cudaSuccess) return errs[t];
    return cudaSuccess;   // the source has been read; the last slices may still be in flight on `st`
}

// Device -> host counterpart: pinned destinations get a plain cudaMemcpyAsync (the caller synchronises `st` as before);
// large pageable ones are filled through the same ring, each worker keeping one slice in flight while it copies the
// previous one out.  On return from the relayed path the destination is complete.
cudaError_t download_async(void* dst, const void* src, size_t bytes, cudaStream_t st) {
    const int T = relay_threads();
    if (bytes < RELAY_MIN_BYTES || T == 0 || !host_is_pageable(dst))
        return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st);
    Lane& l = lane();
    char* ring = (char*)pinned_reserve(size_t(2 * RELAY_MAX_THREADS) * RELAY_SLICE_MAX);
    if (!ring) return cudaErrorMemoryAllocation;
    for (int i = 0; i < 2 * T; i++)
        if (!l.relay_ev[i]) {
            cudaError_t e = cudaEventCreateWithFlags(&l.relay_ev[i], cudaEventDisableTiming);
            if (e != cudaSuccess) return e;
        }
    const size_t RELAY_SLICE = relay_slice();
    const size_t slices = (bytes + RELAY_SLICE - 1) / RELAY_SLICE;
    const int device = ctx().device;
    cudaError_t errs[RELAY_MAX_THREADS];
    auto work = [&](int t) {
        cudaError_t e = cudaSetDevice(device);
        auto span = [&](size_t k, size_t* off, size_t* len) {
            *off = k * RELAY_SLICE;
            *len = bytes - *off < RELAY_SLICE ? bytes - *off : RELAY_SLICE;
        };
        auto fetch = [&](size_t k, int slot) {   // device -> ring slot, ordered on st
            size_t off, len;
            span(k, &off, &len);
            cudaError_t f = cudaEventSynchronize(l.relay_ev[slot]);   // an earlier upload may still be reading the slot
            if (f == cudaSuccess) f = cudaMemcpyAsync(ring + size_t(slot) * RELAY_SLICE, (const char*)src + off, len, cudaMemcpyDeviceToHost, st);
            if (f == cudaSuccess) f = cudaEventRecord(l.relay_ev[slot], st);
            return f;
        };
        size_t j = 0;
        if (e == cudaSuccess && size_t(t) < slices) e = fetch(t, 2 * t);
        for (size_t k = t; k < slices && e == cudaSuccess; k += T, j++) {
            const int slot = 2 * t + int(j & 1);
            if (k + T < slices) e = fetch(k + T, 2 * t + int((j + 1) & 1));
            if (e != cudaSuccess) break;
            e = cudaEventSynchronize(l.relay_ev[slot]);
            if (e != cudaSuccess) break;
            size_t off, len;
            span(k, &off, &len);
            memcpy((char*)dst + off, ring + size_t(slot) * RELAY_SLICE, len);
        }
        errs[t] = e;
    };
    // a thread that cannot be started (resource limits) leaves its share to the calling thread: never unwind through the C ABI
    std::thread th[RELAY_MAX_THREADS];
    bool started[RELAY_MAX_THREADS] = {false};
    for (int t = 1; t < T; t++) {
        try {
            th[t] = std::thread(work, t);
            started[t] = true;
        } catch (...) {
            started[t] = false;
        }
    }
    work(0);
    for (int t = 1; t < T; t++) {
        if (started[t]) th[t].join();
        else work(t);
    }
    for (int t = 0; t < T; t++)
        if (errs[t] != cudaSuccess) return errs[t];
    return cudaSuccess;
}

// ---- stage timer ----
// (diagnostics; one mutex of their own: they are shared by all lanes)
struct StageRec { int stage; cudaEvent_t a, b; };
static std::mutex g_prof_mu;
static bool g_prof = false;
static std::vector<StageRec> g_recs;
static std::vector<cudaEvent_t> g_free_events;
static thread_local cudaEvent_t t_open[ST_COUNT];

static cudaEvent_t get_event() {
    if (!g_free_events.empty()) { cudaEvent_t e = g_free_events.back(); g_free_events.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
void stage_begin(int stage, cudaStream_t stream) {
    if (!g_prof) return;
    std::lock_guard<std::mutex> g(g_prof_mu);
    cudaEvent_t e = get_event();
    cudaEventRecord(e, stream);
    t_open[stage] = e;
}
void stage_end(int stage, cudaStream_t stream) {
    if (!g_prof || !t_open[stage]) return;
    std::lock_guard<std::mutex> g(g_prof_mu);
    cudaEvent_t e = get_event();
    cudaEventRecord(e, stream);
    g_recs.push_back(StageRec{stage, t_open[stage], e});
    t_open[stage] = nullptr;
}

// ---- per-launch trace ----
struct TraceRec { const char* file; int line; cudaEvent_t e; };
static bool g_trace = false;
static std::vector<TraceRec> g_trace_recs;
void trace_mark(const char* file, int line) {
    if (!g_trace) return;
    std::lock_guard<std::mutex> g(g_prof_mu);
    Lane& l = lane();
    cudaEvent_t e = get_event();
    cudaEventRecord(e, l.have_last ? l.last_stream : l.stream);
    g_trace_recs.push_back(TraceRec{file, line, e});
}

// ---- live SRS handles (kzgb_shutdown releases their device memory; see msm.cuh) ----
static std::mutex g_srs_mu;
static std::vector<kzgb_srs*> g_live_srs;
void srs_register(kzgb_srs* s) {
    std::lock_guard<std::mutex> g(g_srs_mu);
    g_live_srs.push_back(s);
}
void srs_unregister(kzgb_srs* s) {
    std::lock_guard<std::mutex> g(g_srs_mu);
    for (size_t i = 0; i < g_live_srs.size(); i++)
        if (g_live_srs[i] == s) { g_live_srs[i] = g_live_srs.back(); g_live_srs.pop_back(); break; }
}

}  // namespace kzg

using namespace kzg;

// on = 1: start tracing; on = 0: stop
extern "C" KZGB_API int kzgb_trace_enable(int on) {
    KZG_REQUIRE_INIT();
    std::lock_guard<std::mutex> g(g_prof_mu);
    g_trace = on != 0;
    return KZGB_OK;
}
// Synchronises, then writes up to cap records: lines[i] = source line of the i-th KZG_LAUNCHED(), ms[i] = time since
// the previous mark (0 for the first); returns the number of records (negative on error) and clears the trace.
extern "C" KZGB_API int kzgb_trace_read(int* lines, float* ms, int cap) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    KZG_CUDA(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> g(g_prof_mu);
    int n = 0;
    for (size_t i = 0; i < g_trace_recs.size(); i++) {
        if (n < cap) {
            float t = 0;
            if (i > 0) cudaEventElapsedTime(&t, g_trace_recs[i - 1].e, g_trace_recs[i].e);
            lines[n] = g_trace_recs[i].line;
            ms[n] = t;
            n++;
        }
        g_free_events.push_back(g_trace_recs[i].e);
    }
    g_trace_recs.clear();
    return n;
}

extern "C" KZGB_API int kzgb_profile_enable(int on) {
    std::lock_guard<std::mutex> g(g_prof_mu);
    g_prof = on != 0;
    return KZGB_OK;
}
// Synchronises the device, adds every recorded stage interval into sum_ms[stage] / counts[stage]
// (arrays of at least kzgb_profile_stage_count() entries, zeroed by the caller) and clears the records.
extern "C" KZGB_API int kzgb_profile_read(double* sum_ms, uint64_t* counts) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    KZG_CUDA(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> g(g_prof_mu);
    for (auto& r : g_recs) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            if (sum_ms) sum_ms[r.stage] += ms;
            if (counts) counts[r.stage] += 1;
        }
        g_free_events.push_back(r.a);
        g_free_events.push_back(r.b);
    }
    g_recs.clear();
    return KZGB_OK;
}
extern "C" KZGB_API int kzgb_profile_stage_count(void) { return ST_COUNT; }
extern "C" KZGB_API const char* kzgb_profile_stage_name(int s) {
    static const char* names[ST_COUNT] = {"msm_digits", "msm_sort", "msm_accumulate", "msm_fixup", "msm_bucket_reduce",
                                          "msm_finish", "ntt", "scan", "div_by_omega", "interpolate", "batch_div"};
    return (s >= 0 && s < ST_COUNT) ? names[s] : "";
}

extern "C" {

int kzgb_init(int device) {
    Context& c = ctx();
    std::lock_guard<std::mutex> g(c.mu);
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
    cudaDeviceProp prop;
    KZG_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("kzgb_init: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                  prop.minor);
        return KZGB_ERR_INVALID_ARG;
    }
    // The device is selected per call and per thread by ApiScope (CUDA's current device is thread-local), and the
    // caller's own current device is left as it was.
    c.sm_count = prop.multiProcessorCount;
    c.device = device;
    c.initialised = true;
    c.launches = 0;
    return KZGB_OK;
}

int kzgb_shutdown(void) {
    Context& c = ctx();
    if (!c.initialised) return KZGB_OK;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(c.device);
    cudaDeviceSynchronize();
    comm_shutdown();
    poly_cache_shutdown();
    {   // live SRS handles: release their device memory now; kzgb_srs_destroy then only frees the host struct
        std::lock_guard<std::mutex> g(g_srs_mu);
        for (kzgb_srs* s : g_live_srs) srs_release_device(s);
        g_live_srs.clear();
    }
    {
        std::lock_guard<std::mutex> g(c.cache_mu);
        for (int i = 0; i < 33; i++) {
            if (c.ntt_tw[i]) cudaFree(c.ntt_tw[i]);
            c.ntt_tw[i] = nullptr;
        }
        c.ntt_attr_set = false;
        c.interp_attr_set = false;   // cudaFuncSetAttribute is per device: a re-init elsewhere must set it again
    }
    std::lock_guard<std::mutex> g(c.mu);
    for (Lane& l : c.lanes) {
        if (l.scratch) cudaFree(l.scratch);
        if (l.staging) cudaFree(l.staging);
        if (l.pinned) cudaFreeHost(l.pinned);
        if (l.stream) cudaStreamDestroy(l.stream);
        if (l.side) cudaStreamDestroy(l.side);
        if (l.fork_ev) cudaEventDestroy(l.fork_ev);
        if (l.join_ev) cudaEventDestroy(l.join_ev);
        for (cudaEvent_t e : l.relay_ev)
            if (e) cudaEventDestroy(e);
        l = Lane();
    }
    c.initialised = false;
    if (prev >= 0 && prev != c.device) cudaSetDevice(prev);
    c.device = -1;
    return KZGB_OK;
}

const char* kzgb_last_error_message(void) { return g_err; }

uint64_t kzgb_launch_count(void) { return ctx().launches.load(); }

}  // extern "C"
