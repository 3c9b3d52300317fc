// msm.cuh -- SRS handle and MSM driver interface (internal).
#pragma once
#include <cstddef>
#include <cuda_runtime.h>

#include "../../include/kzg_b200.h"

namespace kzg {
struct Affine;
struct XYZZ;
}  // namespace kzg

// Device-resident SRS: `table[w*n + i] = 2^(c*w) * P_i` in affine Montgomery form
// (blst_p1_affine layout).  Row 0 is the SRS itself.
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
};

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
void msm_pipe_shutdown();
int jac_to_affine_dev(const void* d_jac, void* d_aff, size_t n, cudaStream_t stream);
}  // namespace kzg
