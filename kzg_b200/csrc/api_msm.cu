// api_msm.cu -- C ABI: SRS handles and the G1 MSM entry points.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <vector>

#include "common.cuh"
#include "ec.cuh"
#include "evalform.cuh"
#include "msm.cuh"
#include "scan.cuh"

using namespace kzg;


static cudaStream_t pick_stream(void* stream) { return use_stream((cudaStream_t)stream); }

extern "C" {

KZGB_API int kzgb_srs_create_g1(const kzgb_p1* points, size_t n, int window_bits, kzgb_srs** out) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    use_stream(lane().stream);
    if (!points || !out) { set_error("kzgb_srs_create_g1: null argument"); return KZGB_ERR_INVALID_ARG; }
    kzgb_srs* s = nullptr;
    int st = srs_alloc(&s, n, window_bits);
    if (st) return st;
    cudaStream_t stream = lane().stream;
    void* d_jac = nullptr;
    cudaError_t e = cudaMalloc(&d_jac, n * sizeof(kzgb_p1));
    if (e != cudaSuccess) { srs_free(s); return cuda_fail(e, "cudaMalloc(jacobian staging)", __FILE__, __LINE__); }
    st = KZGB_OK;
    do {
        if ((e = cudaMemcpyAsync(d_jac, points, n * sizeof(kzgb_p1), cudaMemcpyHostToDevice, stream)) != cudaSuccess) {
            st = cuda_fail(e, "H2D srs", __FILE__, __LINE__); break;
        }
        if ((st = jac_to_affine_dev(d_jac, s->table, n, stream))) break;
        if ((st = srs_build_table(s, stream))) break;
        if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) { st = cuda_fail(e, "sync", __FILE__, __LINE__); break; }
    } while (0);
    cudaFree(d_jac);
    if (st) { srs_free(s); return st; }
    *out = s;
    return KZGB_OK;
}

KZGB_API int kzgb_srs_create_g1_affine(const kzgb_p1_affine* points, size_t n, int window_bits, int on_device,
                                       kzgb_srs** out) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    use_stream(lane().stream);
    if (!points || !out) { set_error("kzgb_srs_create_g1_affine: null argument"); return KZGB_ERR_INVALID_ARG; }
    kzgb_srs* s = nullptr;
    int st = srs_alloc(&s, n, window_bits);
    if (st) return st;
    cudaStream_t stream = lane().stream;
    cudaError_t e = cudaMemcpyAsync(s->table, points, n * sizeof(kzgb_p1_affine),
                                    on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) { srs_free(s); return cuda_fail(e, "copy srs", __FILE__, __LINE__); }
    st = srs_build_table(s, stream);
    if (!st && (e = cudaStreamSynchronize(stream)) != cudaSuccess) st = cuda_fail(e, "sync", __FILE__, __LINE__);
    if (st) { srs_free(s); return st; }
    *out = s;
    return KZGB_OK;
}

KZGB_API int kzgb_srs_destroy(kzgb_srs* s) {
    if (!s) return KZGB_OK;
    // after kzgb_shutdown the handle's device memory is already gone (kzgb_shutdown released it): free the host side
    if (s->dead || !ctx().initialised) { srs_free(s); return KZGB_OK; }
    KZG_ENTER();
    {   // wait for the calls that still use this handle (other threads, or device work it has in flight)
        KZG_LOCK_SRS(s);
        if (s->have_last) cudaStreamSynchronize(s->last_stream);
    }
    srs_free(s);
    return KZGB_OK;
}

KZGB_API size_t kzgb_srs_size(const kzgb_srs* s) { return s ? s->n : 0; }
KZGB_API int kzgb_srs_window_bits(const kzgb_srs* s) { return s ? s->c : 0; }
KZGB_API size_t kzgb_srs_device_bytes(const kzgb_srs* s) {
    return s ? (size_t)s->nwin * s->n * sizeof(kzgb_p1_affine) + s->work_bytes + s->pipe_work_bytes[0] + s->pipe_work_bytes[1] : 0;
}
KZGB_API const void* kzgb_srs_device_points(const kzgb_srs* s) { return s ? s->table : nullptr; }

KZGB_API int kzgb_srs_read_affine(const kzgb_srs* s, size_t start, size_t count, kzgb_p1_affine* out) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    use_stream(lane().stream);
    if (!s || !out || start + count > s->n || s->dead) { set_error("kzgb_srs_read_affine: bad range"); return KZGB_ERR_INVALID_ARG; }
    KZG_LOCK_SRS(s);
    KZG_CUDA(cudaStreamSynchronize(lane().stream));
    KZG_CUDA(cudaMemcpy(out, (const kzgb_p1_affine*)s->table + start, count * sizeof(kzgb_p1_affine),
                        cudaMemcpyDeviceToHost));
    return KZGB_OK;
}

// out_mode: 0 affine (96 B), 1 Jacobian (144 B), 2 raw XYZZ partial (192 B, for multi-GPU combine)
KZGB_API int kzgb_msm_g1_dev(kzgb_srs* s, const void* d_scalars, size_t m, void* d_out, int out_mode, void* stream) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    if (!s || (!d_scalars && m) || !d_out || out_mode < 0 || out_mode > 2) {
        set_error("kzgb_msm_g1_dev: bad argument");
        return KZGB_ERR_INVALID_ARG;
    }
    KZG_LOCK_SRS(s);
    cudaStream_t st = pick_stream(stream);
    const XYZZ* res = nullptr;
    int rc = msm_run(s, d_scalars, m, st, &res);
    if (rc) return rc;
    return msm_finish(res, d_out, out_mode, st);
}

static int msm_host(kzgb_srs* s, const kzgb_fr* scalars, size_t m, void* out, int mode, size_t out_bytes) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    use_stream(lane().stream);
    if (!s || (!scalars && m) || !out) { set_error("kzgb_msm_g1: null argument"); return KZGB_ERR_INVALID_ARG; }
    if (m > s->n) { set_error("kzgb_msm_g1: %zu scalars > %zu SRS points", m, s->n); return KZGB_ERR_INVALID_ARG; }
    KZG_LOCK_SRS(s);
    size_t bytes = m * sizeof(kzgb_fr);
    char* d = (char*)staging_reserve(bytes + 256);
    if (!d) return KZGB_ERR_CUDA;
    if (m) KZG_CUDA(upload_async(d + 256, scalars, bytes, lane().stream));
    const XYZZ* res = nullptr;
    int rc = msm_run(s, d + 256, m, lane().stream, &res);
    if (rc) return rc;
    rc = msm_finish(res, d, mode, lane().stream);
    if (rc) return rc;
    KZG_CUDA(cudaMemcpyAsync(out, d, out_bytes, cudaMemcpyDeviceToHost, lane().stream));
    KZG_CUDA(cudaStreamSynchronize(lane().stream));
    return KZGB_OK;
}

KZGB_API int kzgb_msm_g1(kzgb_srs* s, const kzgb_fr* scalars, size_t m, kzgb_p1_affine* out) {
    return msm_host(s, scalars, m, out, 0, sizeof(kzgb_p1_affine));
}
KZGB_API int kzgb_msm_g1_jacobian(kzgb_srs* s, const kzgb_fr* scalars, size_t m, kzgb_p1* out) {
    return msm_host(s, scalars, m, out, 1, sizeof(kzgb_p1));
}

// ---- pipelined batch of MSMs against one SRS (commit to many polynomials) -------------------------------------
// A single MSM ends in a latency-bound tail (fix-up levels, bucket reduction, affine conversion: ~1.4 ms of ~7 ms at
// 2^20 on small grids) and, through the host-pointer entry point, starts with a 32 MiB host-to-device copy.  In a
// batch, consecutive MSMs alternate between two work areas:
//   * ONE low-priority "heavy" stream runs digits -> sort -> accumulation of every item in order (two accumulations
//     sharing the machine would only fight for the same pipe),
//   * the tail of item i runs on the high-priority stream of its slot, so its few blocks slot in while the heavy
//     stream is already accumulating item i+1,
//   * a copy stream moves item i+1's scalars while item i computes.
// Events hand the slot (staging + work area) back once the tail of item i is done; item i+2 waits for it.
// Results are identical to `count` separate calls.
namespace {
int pipe_reserve_results(MsmPipe& g_pipe, size_t count);   // defined with the many-witness helpers below
size_t out_size(int mode) { return mode == 0 ? sizeof(kzgb_p1_affine) : (mode == 1 ? sizeof(kzgb_p1) : 192); }

int pipe_drain(MsmPipe& g_pipe, int rc) {
    cudaStream_t all[4] = {g_pipe.copy, g_pipe.heavy, g_pipe.tail_s[0], g_pipe.tail_s[1]};
    for (cudaStream_t st : all) {
        cudaError_t e = cudaStreamSynchronize(st);
        if (rc == KZGB_OK && e != cudaSuccess) rc = cuda_fail(e, "sync(batch)", __FILE__, __LINE__);
    }
    return rc;
}

// where: bit 0 = scalars[i] are device pointers (else host), bit 1 = `out` is a device array (else host); results are
// out_size(mode) bytes apart.  `after`: a stream whose already-enqueued work the batch must wait for and which
// receives the completion (device output), or nullptr.
// KZGB_PIPE_DEBUG=1: per-item timeline on stderr (host enqueue time, GPU time of the hand-off and of completion)
struct PipeDbg { double host_ms; cudaEvent_t acc_done, done; };
static std::vector<PipeDbg> g_dbg;
static cudaEvent_t g_dbg_start = nullptr;
static bool pipe_debug() { static int v = -1; if (v < 0) { const char* e = getenv("KZGB_PIPE_DEBUG"); v = (e && *e == '1') ? 1 : 0; } return v == 1; }
static double now_ms() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }

int msm_batch_run(kzgb_srs* s, MsmPipe& g_pipe, const void* const* scalars, const size_t* ms, size_t count, void* out,
                  int mode, int where, cudaStream_t after) {
    const bool sc_dev = (where & 1) != 0, out_dev = (where & 2) != 0;
    const size_t osz = out_size(mode);
    const bool dbg = pipe_debug();
    double h0 = now_ms();
    if (dbg) {
        g_dbg.clear();
        cudaEventCreate(&g_dbg_start);
        cudaEventRecord(g_dbg_start, g_pipe.heavy);
    }
    if (!out_dev && count) {
        int rc = pipe_reserve_results(g_pipe, count);
        if (rc) return rc;
    }
    for (size_t i = 0; i < count; i++) {
        const int k = (int)(i & 1);
        cudaStream_t ts_ = g_pipe.tail_s[k];
        const size_t m = ms[i];
        if (m > s->n) {
            set_error("kzgb_msm_g1_batch: item %zu has %zu scalars > %zu SRS points", i, m, s->n);
            return KZGB_ERR_INVALID_ARG;
        }
        const void* d_sc = scalars[i];
        if (!sc_dev) {
            const size_t bytes = m * sizeof(kzgb_fr);
            if (g_pipe.stage_bytes[k] < bytes) {
                int rc = pipe_drain(g_pipe, KZGB_OK);
                if (rc) return rc;
                if (g_pipe.d_stage[k]) cudaFree(g_pipe.d_stage[k]);
                g_pipe.d_stage[k] = nullptr;
                g_pipe.stage_bytes[k] = 0;
                KZG_CUDA(cudaMalloc(&g_pipe.d_stage[k], bytes));
                g_pipe.stage_bytes[k] = bytes;
            }
            // the staging buffer of this slot is free once the tail of item i-2 has finished
            if (g_pipe.used[k]) KZG_CUDA(cudaStreamWaitEvent(g_pipe.copy, g_pipe.done[k], 0));
            if (m) KZG_CUDA(upload_async(g_pipe.d_stage[k], scalars[i], bytes, g_pipe.copy));
            KZG_CUDA(cudaEventRecord(g_pipe.staged[k], g_pipe.copy));
            KZG_CUDA(cudaStreamWaitEvent(g_pipe.heavy, g_pipe.staged[k], 0));
            d_sc = g_pipe.d_stage[k];
        }
        // ... and so is its work area
        if (g_pipe.used[k]) KZG_CUDA(cudaStreamWaitEvent(g_pipe.heavy, g_pipe.done[k], 0));
        const XYZZ* res = nullptr;
        int rc = msm_run(s, d_sc, m, g_pipe.heavy, &res, k, ts_, g_pipe.handoff[k]);
        if (rc) return rc;
        PipeDbg d{};
        if (dbg) {
            cudaEventCreate(&d.acc_done);
            cudaEventRecord(d.acc_done, g_pipe.heavy);
        }
        void* d_o = out_dev ? (void*)((char*)out + i * osz) : (void*)((char*)g_pipe.d_out + i * 192);
        rc = msm_finish(res, d_o, mode, ts_);
        if (rc) return rc;
        if (!out_dev)
            KZG_CUDA(cudaMemcpyAsync((char*)g_pipe.h_out + i * 192, d_o, osz, cudaMemcpyDeviceToHost, ts_));
        KZG_CUDA(cudaEventRecord(g_pipe.done[k], ts_));
        g_pipe.used[k] = true;
        if (dbg) {
            cudaEventCreate(&d.done);
            cudaEventRecord(d.done, ts_);
            d.host_ms = now_ms() - h0;
            g_dbg.push_back(d);
        }
    }
    // the batch is complete when both tails are; the device variant hands completion back to `after`
    if (out_dev && after)
        for (int k = 0; k < 2 && (size_t)k < count; k++) KZG_CUDA(cudaStreamWaitEvent(after, g_pipe.done[k], 0));
    return KZGB_OK;
}

int msm_batch(kzgb_srs* s, const void* const* scalars, const size_t* ms, size_t count, void* out, int mode, int where,
              cudaStream_t after) {
    const bool out_dev = (where & 2) != 0;
    MsmPipe* pp = nullptr;
    int rc = msm_pipe_get(s, &pp);
    if (rc) return rc;
    MsmPipe& g_pipe = *pp;
    // the handle's main work area may still be in use by an earlier single call on another stream
    if (s->have_last) { KZG_CUDA(cudaStreamSynchronize(s->last_stream)); s->have_last = false; }
    if (after) {  // the heavy stream starts after the work already enqueued on the caller's stream
        cudaEvent_t start_ev;
        KZG_CUDA(cudaEventCreateWithFlags(&start_ev, cudaEventDisableTiming));
        KZG_CUDA(cudaEventRecord(start_ev, after));
        KZG_CUDA(cudaStreamWaitEvent(g_pipe.heavy, start_ev, 0));
        KZG_CUDA(cudaStreamWaitEvent(g_pipe.copy, start_ev, 0));
        cudaEventDestroy(start_ev);
    }
    rc = msm_batch_run(s, g_pipe, scalars, ms, count, out, mode, where, after);
    // errors, host output and host scalars (the caller may reuse its buffers on return): wait here
    if (rc != KZGB_OK || !(out_dev && after) || !(where & 1) || pipe_debug()) rc = pipe_drain(g_pipe, rc);
    if (pipe_debug() && g_dbg_start) {
        fprintf(stderr, "[kzgb pipe] %zu items: host enqueue done at / accumulate done at / item done at (ms)\n", g_dbg.size());
        for (size_t i = 0; i < g_dbg.size(); i++) {
            float a = 0, b = 0;
            cudaEventElapsedTime(&a, g_dbg_start, g_dbg[i].acc_done);
            cudaEventElapsedTime(&b, g_dbg_start, g_dbg[i].done);
            fprintf(stderr, "[kzgb pipe] %2zu  %8.3f  %8.3f  %8.3f\n", i, g_dbg[i].host_ms, a, b);
            cudaEventDestroy(g_dbg[i].acc_done);
            cudaEventDestroy(g_dbg[i].done);
        }
        cudaEventDestroy(g_dbg_start);
        g_dbg_start = nullptr;
        g_dbg.clear();
    }
    if (rc == KZGB_OK && !out_dev)
        for (size_t i = 0; i < count; i++) memcpy((char*)out + i * out_size(mode), (char*)g_pipe.h_out + i * 192, out_size(mode));
    return rc;
}
}  // namespace

extern "C++" {
namespace kzg {
// the pipelined batch for other translation units (comm.cu: sharded batches); the caller holds the handle's lock
int msm_batch_entry(kzgb_srs* s, const void* const* scalars, const size_t* ms, size_t count, void* out, int mode, int where,
                    cudaStream_t after) {
    return msm_batch(s, scalars, ms, count, out, mode, where, after);
}
}  // namespace kzg
}  // extern "C++"

// `count` commits against one SRS: scalars[i] points at ms[i] host scalars, out[i] receives the affine result.
// Same results as `count` calls of kzgb_msm_g1 (coeff_form.rs:59-64 once per polynomial), pipelined as described above.
KZGB_API int kzgb_msm_g1_batch(kzgb_srs* s, const kzgb_fr* const* scalars, const size_t* ms, size_t count,
                               kzgb_p1_affine* out) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    use_stream(lane().stream);
    if (!s || (count && (!scalars || !ms || !out))) { set_error("kzgb_msm_g1_batch: null argument"); return KZGB_ERR_INVALID_ARG; }
    for (size_t i = 0; i < count; i++)
        if (ms[i] && !scalars[i]) { set_error("kzgb_msm_g1_batch: item %zu is null", i); return KZGB_ERR_INVALID_ARG; }
    KZG_LOCK_SRS(s);
    KZG_CUDA(cudaStreamSynchronize(lane().stream));
    return msm_batch(s, (const void* const*)scalars, ms, count, out, 0, 0, nullptr);
}

// Device-resident variant: d_scalars[i] are device pointers (the array itself is a host array), d_out a device
// array of `count` results in out_mode layout (0 affine 96 B, 1 Jacobian 144 B, 2 raw XYZZ 192 B).  Ordered after
// the work already enqueued on `stream`; work enqueued on `stream` afterwards sees the results.
KZGB_API int kzgb_msm_g1_batch_dev(kzgb_srs* s, const void* const* d_scalars, const size_t* ms, size_t count, void* d_out,
                                   int out_mode, void* stream) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    if (!s || out_mode < 0 || out_mode > 2 || (count && (!d_scalars || !ms || !d_out))) {
        set_error("kzgb_msm_g1_batch_dev: bad argument");
        return KZGB_ERR_INVALID_ARG;
    }
    for (size_t i = 0; i < count; i++)
        if (ms[i] && !d_scalars[i]) { set_error("kzgb_msm_g1_batch_dev: item %zu is null", i); return KZGB_ERR_INVALID_ARG; }
    KZG_LOCK_SRS(s);
    cudaStream_t st = pick_stream(stream);
    return msm_batch(s, d_scalars, ms, count, d_out, out_mode, 3, st);
}

// Host scalars in, device results out (the multi-GPU commit path from host buffers: the K partial sums stay on the
// device for the all-gather).  Returns when the batch is complete (the host buffers may be reused).
KZGB_API int kzgb_msm_g1_batch_h2d(kzgb_srs* s, const kzgb_fr* const* scalars, const size_t* ms, size_t count, void* d_out,
                                   int out_mode, void* stream) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    if (!s || out_mode < 0 || out_mode > 2 || (count && (!scalars || !ms || !d_out))) {
        set_error("kzgb_msm_g1_batch_h2d: bad argument");
        return KZGB_ERR_INVALID_ARG;
    }
    for (size_t i = 0; i < count; i++)
        if (ms[i] && !scalars[i]) { set_error("kzgb_msm_g1_batch_h2d: item %zu is null", i); return KZGB_ERR_INVALID_ARG; }
    KZG_LOCK_SRS(s);
    cudaStream_t st = pick_stream(stream);
    return msm_batch(s, (const void* const*)scalars, ms, count, d_out, out_mode, 2, st);
}

// ---- shared pieces of the "many witnesses" entry points (host results; scalars produced on the heavy stream) ----
namespace {
int pipe_reserve_results(MsmPipe& g_pipe, size_t count) {
    if (g_pipe.out_bytes >= count * 192) return KZGB_OK;
    if (count < 256) count = 256;   // pinned allocations are slow (up to 0.1 s): never regrow for a handful of results
    if (g_pipe.d_out) { cudaFree(g_pipe.d_out); cudaFreeHost(g_pipe.h_out); g_pipe.d_out = g_pipe.h_out = nullptr; g_pipe.out_bytes = 0; }
    KZG_CUDA(cudaMalloc(&g_pipe.d_out, count * 192));
    KZG_CUDA(cudaMallocHost(&g_pipe.h_out, count * 192));
    g_pipe.out_bytes = count * 192;
    return KZGB_OK;
}
int pipe_reserve_stage(MsmPipe& g_pipe, size_t bytes) {   // both slots; drains the pipeline when it has to grow
    for (int k = 0; k < 2; k++) {
        if (g_pipe.stage_bytes[k] >= bytes) continue;
        int rc = pipe_drain(g_pipe, KZGB_OK);
        if (rc) return rc;
        if (g_pipe.d_stage[k]) cudaFree(g_pipe.d_stage[k]);
        g_pipe.d_stage[k] = nullptr;
        g_pipe.stage_bytes[k] = 0;
        KZG_CUDA(cudaMalloc(&g_pipe.d_stage[k], bytes));
        g_pipe.stage_bytes[k] = bytes;
    }
    return KZGB_OK;
}
// MSM of the m scalars in the slot's stage buffer -> affine result i (device + pinned mirror); releases the slot
int pipe_msm_item(kzgb_srs* s, MsmPipe& g_pipe, int k, size_t m, size_t i) {
    cudaStream_t ts_ = g_pipe.tail_s[k];
    const XYZZ* res = nullptr;
    int rc = msm_run(s, g_pipe.d_stage[k], m, g_pipe.heavy, &res, k, ts_, g_pipe.handoff[k]);
    if (rc) return rc;
    void* d_o = (char*)g_pipe.d_out + i * 192;
    rc = msm_finish(res, d_o, 0, ts_);
    if (rc) return rc;
    KZG_CUDA(cudaMemcpyAsync((char*)g_pipe.h_out + i * 192, d_o, sizeof(kzgb_p1_affine), cudaMemcpyDeviceToHost, ts_));
    KZG_CUDA(cudaEventRecord(g_pipe.done[k], ts_));
    g_pipe.used[k] = true;
    return KZGB_OK;
}
}  // namespace

// create_witness (coeff_form.rs:66-81) of ONE polynomial at m points in one pipelined call (multi-opening workloads:
// SURVEY 8 f2).  Item i = scan (p - y_i) / (X - x_i) on the heavy stream into the slot's quotient buffer, then the MSM of
// the n-1 quotient coefficients; the tail of witness i overlaps the scan + accumulation of witness i+1, and p is
// uploaded once.  status[i] = 0, or 1 (KZGB_POINT_NOT_ON_POLYNOMIAL: p(x_i) != y_i, coeff_form.rs:72-75; out[i] is
// then unspecified, as the reference returns Err).  Same bytes as m calls of kzgb_witness_coeff.
namespace {
__global__ void batch_rem_flag_kernel(const Fr* rem, int* flag) {
    if (threadIdx.x == 0 && blockIdx.x == 0) *flag = rem->is_zero() ? 0 : 1;
}
}  // namespace

KZGB_API int kzgb_witness_coeff_many(kzgb_srs* s, const kzgb_fr* p, size_t n, const kzgb_fr* xs, const kzgb_fr* ys, size_t m,
                                     kzgb_p1_affine* out, int* status) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    use_stream(lane().stream);
    if (!s || !p || n == 0 || (m && (!xs || !ys || !out || !status))) {
        set_error("kzgb_witness_coeff_many: bad argument");
        return KZGB_ERR_INVALID_ARG;
    }
    const size_t qn = n > 1 ? n - 1 : 1;   // psi has n-1 coefficients (n == 1: one zero coefficient)
    if (qn > s->n) {
        set_error("kzgb_witness_coeff_many: quotient has %zu coefficients but the SRS holds %zu points", qn, s->n);
        return KZGB_ERR_INVALID_ARG;
    }
    if (m == 0) return KZGB_OK;
    KZG_LOCK_SRS(s);
    KZG_CUDA(cudaStreamSynchronize(lane().stream));
    MsmPipe* pp = nullptr;
    int rc = msm_pipe_get(s, &pp);
    if (rc) return rc;
    MsmPipe& g_pipe = *pp;
    if (s->have_last) { KZG_CUDA(cudaStreamSynchronize(s->last_stream)); s->have_last = false; }
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    // staging arena: [p | (x_i, y_i) pairs | remainders | flags]; scan workspace in the scratch arena
    const size_t pb = al(n * 32), xyb = al(m * 64), rb = al(m * 32), fb = al(m * sizeof(int));
    char* buf = (char*)staging_reserve(pb + xyb + rb + fb);
    if (!buf) return KZGB_ERR_CUDA;
    char* ws = (char*)scratch_reserve(al(scan_workspace_bytes(n)));
    if (!ws) return KZGB_ERR_CUDA;
    Fr* d_p = (Fr*)buf;
    Fr* d_xy = (Fr*)(buf + pb);
    Fr* d_rem = (Fr*)(buf + pb + xyb);
    int* d_flag = (int*)(buf + pb + xyb + rb);
    if ((rc = pipe_reserve_results(g_pipe, m)) || (rc = pipe_reserve_stage(g_pipe, qn * 32))) return rc;
    cudaStream_t hs = g_pipe.heavy;
    KZG_CUDA(upload_async(d_p, p, n * 32, hs));
    // interleave (x_i, y_i) on the device side: two strided copies
    KZG_CUDA(cudaMemcpy2DAsync(d_xy, 64, xs, 32, 32, m, cudaMemcpyHostToDevice, hs));
    KZG_CUDA(cudaMemcpy2DAsync(d_xy + 1, 64, ys, 32, 32, m, cudaMemcpyHostToDevice, hs));
    for (size_t i = 0; i < m && rc == KZGB_OK; i++) {
        const int k = (int)(i & 1);
        if (g_pipe.used[k]) KZG_CUDA(cudaStreamWaitEvent(hs, g_pipe.done[k], 0));   // quotient buffer + work area free
        Fr* d_q = (Fr*)g_pipe.d_stage[k];
        rc = scan_divide_linear(d_p, n, d_xy + 2 * i, d_xy + 2 * i + 1, d_q, d_rem + i, ws, hs);
        if (rc) break;
        batch_rem_flag_kernel<<<1, 32, 0, hs>>>(d_rem + i, d_flag + i);
        KZG_LAUNCHED();
        if (n == 1) KZG_CUDA(cudaMemsetAsync(d_q, 0, 32, hs));
        rc = pipe_msm_item(s, g_pipe, k, qn, i);
    }
    rc = pipe_drain(g_pipe, rc);
    if (rc) return rc;
    KZG_CUDA(cudaMemcpy(status, d_flag, m * sizeof(int), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < m; i++) {
        status[i] = status[i] ? KZGB_POINT_NOT_ON_POLYNOMIAL : KZGB_OK;
        memcpy(out + i, (char*)g_pipe.h_out + i * 192, sizeof(kzgb_p1_affine));
    }
    return KZGB_OK;
}

// KZGProverEvalForm::create_witness (eval_form.rs:124-140) of one evaluation vector at m indices in one pipelined
// call: evals uploaded once; item j = div_by_omega_i (eval_form.rs:58-84) at indices[j] into the slot's quotient
// buffer, then the MSM over the Lagrange-basis SRS.  Same bytes as m calls of kzgb_witness_eval (which is infallible).
KZGB_API int kzgb_witness_eval_many(kzgb_srs* lagrange_srs, const kzgb_fr* evals, uint32_t log_d, const size_t* indices,
                                    size_t m, kzgb_p1_affine* out) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    use_stream(lane().stream);
    if (log_d >= 32) return KZGB_DEGREE_TOO_LARGE;
    const size_t d = (size_t)1 << log_d;
    if (!lagrange_srs || !evals || (m && (!indices || !out)) || d > lagrange_srs->n) {
        set_error("kzgb_witness_eval_many: bad argument");
        return KZGB_ERR_INVALID_ARG;
    }
    for (size_t j = 0; j < m; j++)
        if (indices[j] >= d) { set_error("kzgb_witness_eval_many: index %zu out of bounds", indices[j]); return KZGB_ERR_INVALID_ARG; }
    if (m == 0) return KZGB_OK;
    KZG_LOCK_SRS(lagrange_srs);
    KZG_CUDA(cudaStreamSynchronize(lane().stream));
    MsmPipe* pp = nullptr;
    int rc = msm_pipe_get(lagrange_srs, &pp);
    if (rc) return rc;
    MsmPipe& g_pipe = *pp;
    if (lagrange_srs->have_last) { KZG_CUDA(cudaStreamSynchronize(lagrange_srs->last_stream)); lagrange_srs->have_last = false; }
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    char* d_evals = (char*)staging_reserve(al(d * 32));
    if (!d_evals) return KZGB_ERR_CUDA;
    char* ws = (char*)scratch_reserve(al(div_by_omega_workspace_bytes(log_d)) + 256);
    if (!ws) return KZGB_ERR_CUDA;
    if ((rc = pipe_reserve_results(g_pipe, m)) || (rc = pipe_reserve_stage(g_pipe, d * 32))) return rc;
    cudaStream_t hs = g_pipe.heavy;
    KZG_CUDA(upload_async(d_evals, evals, d * 32, hs));
    for (size_t j = 0; j < m && rc == KZGB_OK; j++) {
        const int k = (int)(j & 1);
        if (g_pipe.used[k]) KZG_CUDA(cudaStreamWaitEvent(hs, g_pipe.done[k], 0));
        if (log_d == 0) KZG_CUDA(cudaMemsetAsync(g_pipe.d_stage[k], 0, 32, hs));
        else rc = div_by_omega_run((const Fr*)d_evals, log_d, (uint32_t)indices[j], (Fr*)g_pipe.d_stage[k], ws, hs);
        if (rc) break;
        rc = pipe_msm_item(lagrange_srs, g_pipe, k, d, j);
    }
    rc = pipe_drain(g_pipe, rc);
    if (rc) return rc;
    for (size_t j = 0; j < m; j++) memcpy(out + j, (char*)g_pipe.h_out + j * 192, sizeof(kzgb_p1_affine));
    return KZGB_OK;
}

// Multi-GPU combine: `count` raw XYZZ partial sums (192 B each, device) -> one point.
KZGB_API int kzgb_g1_combine_dev(const void* d_partials, uint32_t count, void* d_out, int out_mode, void* stream) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    if (!d_partials || !d_out || count == 0) { set_error("kzgb_g1_combine_dev: bad argument"); return KZGB_ERR_INVALID_ARG; }
    cudaStream_t st = pick_stream(stream);
    char* d = (char*)scratch_reserve(256);
    if (!d) return KZGB_ERR_CUDA;
    int rc = msm_combine(d_partials, count, d, st);
    if (rc) return rc;
    return msm_finish((const XYZZ*)d, d_out, out_mode, st);
}


// Batched combine (K polynomials at once): d_partials is the all-gathered [count ranks][items][192 B] array;
// d_out receives `items` results in out_mode layout.  One block per item, so the K affine conversions run side by side.
KZGB_API int kzgb_g1_combine_batch_dev(const void* d_partials, uint32_t count, uint32_t items, void* d_out, int out_mode,
                                       void* stream) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    if (!d_partials || !d_out || count == 0 || out_mode < 0 || out_mode > 2) {
        set_error("kzgb_g1_combine_batch_dev: bad argument");
        return KZGB_ERR_INVALID_ARG;
    }
    if (items == 0) return KZGB_OK;
    cudaStream_t st = pick_stream(stream);
    char* d = (char*)scratch_reserve((size_t)items * 192);
    if (!d) return KZGB_ERR_CUDA;
    int rc = msm_combine(d_partials, count, d, st, items);
    if (rc) return rc;
    return msm_finish((const XYZZ*)d, d_out, out_mode, st, items);
}

}  // extern "C"
