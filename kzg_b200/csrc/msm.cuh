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
};

namespace kzg {
int msm_choose_window(size_t n);
int srs_alloc(kzgb_srs** out, size_t n, int c);
void srs_free(kzgb_srs* s);
int srs_build_table(kzgb_srs* s, cudaStream_t stream);
int msm_run(kzgb_srs* s, const void* d_scalars, size_t m, cudaStream_t stream, const XYZZ** d_result_xyzz);
int msm_finish(const XYZZ* d_xyzz, void* d_out, int mode, cudaStream_t stream);
int msm_combine(const void* d_partials_xyzz, uint32_t count, void* d_out_xyzz, cudaStream_t stream);
int jac_to_affine_dev(const void* d_jac, void* d_aff, size_t n, cudaStream_t stream);
}  // namespace kzg
