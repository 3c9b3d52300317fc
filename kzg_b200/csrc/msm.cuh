// msm.cuh -- SRS handle and MSM driver interface (internal).
#pragma once
#include <cstddef>
#include <cuda_runtime.h>
#include <mutex>

#include "../../include/kzg_b200.h"

namespace kzg {
struct Affine;
struct XYZZ;
}  // namespace kzg

// Streams, events and staging of the pipelined batch entry points (kzgb_msm_g1_batch*, kzgb_witness_*_many); one per
// SRS handle, created on first use.
struct MsmPipe {
    cudaStream_t heavy = nullptr, copy = nullptr, tail_s[2] = {nullptr, nullptr};
    cudaEvent_t handoff[2] = {nullptr, nullptr}, done[2] = {nullptr, nullptr}, staged[2] = {nullptr, nullptr};
    void* d_stage[2] = {nullptr, nullptr};
    size_t stage_bytes[2] = {0, 0};
    void* d_out = nullptr;      // count x 192 B device results (host variant)
    size_t out_bytes = 0;
    void* h_out = nullptr;      // pinned mirror of d_out
    bool used[2] = {false, false};  // done[k] has been recorded at least once
};

// Device-resident SRS: `table[w*n + i] = 2^(c*w) * P_i` in affine Montgomery form
// (blst_p1_affine layout).  Row 0 is the SRS itself.
// Concurrency: `mu` serialises the calls that use THIS handle's work areas; calls on different handles (and calls that
// take no handle) run concurrently, each on its own lane (common.cuh).  A work area is reused by the next call, so
// when that call runs on another stream the previous stream is drained first (`last_stream`).
struct kzgb_srs {
    size_t n = 0;
    int c = 0;
    int nwin = 0;
    kzg::Affine* table = nullptr;
    void* work = nullptr;  // per-call MSM scratch, grown on demand
    size_t work_bytes = 0;
    // two more work areas for the pipelined batch entry points (kzgb_msm_g1_batch*): consecutive MSMs alternate
    // between them so that one MSM's latency-bound tail overlaps the next one's copy, sort and accumulation
    void* pipe_work[2] = {nullptr, nullptr};
    size_t pipe_work_bytes[2] = {0, 0};
    MsmPipe* pipe = nullptr;
    mutable std::recursive_mutex mu;
    cudaStream_t last_stream = nullptr;
    bool have_last = false;
    bool dead = false;   // kzgb_shutdown released the device memory; only kzgb_srs_destroy is valid now
};

#define KZG_LOCK_SRS(s) std::lock_guard<std::recursive_mutex> _srs_guard((s)->mu)

namespace kzg {
int msm_choose_window(size_t n);
int srs_alloc(kzgb_srs** out, size_t n, int c);
void srs_free(kzgb_srs* s);
int srs_build_table(kzgb_srs* s, cudaStream_t stream);
// slot < 0: the handle's main work area; 0/1: a pipeline work area.  tail_stream != nullptr: everything after the
// accumulation (fix-up, bucket reduction; the caller's msm_finish) runs on tail_stream, ordered after `stream` through
// `handoff` (an event owned by the caller).
int msm_run(kzgb_srs* s, const void* d_scalars, size_t m, cudaStream_t stream, const XYZZ** d_result_xyzz, int slot = -1,
            cudaStream_t tail_stream = nullptr, cudaEvent_t handoff = nullptr);
int msm_finish(const XYZZ* d_xyzz, void* d_out, int mode, cudaStream_t stream, uint32_t items = 1);
// items > 1: partial of (rank g, item b) at index g*items + b; out[b] = sum over g
int msm_combine(const void* d_partials_xyzz, uint32_t count, void* d_out_xyzz, cudaStream_t stream, uint32_t items = 1);
// where: bit 0 = scalars[i] are device pointers (else host), bit 1 = `out` is a device array (else host)
int msm_batch_entry(kzgb_srs* s, const void* const* scalars, const size_t* ms, size_t count, void* out, int mode, int where,
                    cudaStream_t after);
int msm_pipe_get(kzgb_srs* s, MsmPipe** out);   // creates the handle's batch pipeline on first use
void srs_register(kzgb_srs* s);                 // live-handle list (api_core.cu): kzgb_shutdown releases device memory
void srs_unregister(kzgb_srs* s);
void srs_release_device(kzgb_srs* s);           // frees everything on the device, marks the handle dead
void comm_shutdown();                           // multi-GPU communicator (comm.cu)
void poly_cache_shutdown();                     // parked kzgb_poly buffers (api_handle.cu)
int jac_to_affine_dev(const void* d_jac, void* d_aff, size_t n, cudaStream_t stream);
}  // namespace kzg
