// api_msm.cu -- C ABI: SRS handles and the G1 MSM entry points.
#include "common.cuh"
#include "ec.cuh"
#include "msm.cuh"

using namespace kzg;

#define KZG_LOCK() std::lock_guard<std::recursive_mutex> _g(ctx().mu)

static cudaStream_t pick_stream(void* stream) { return use_stream((cudaStream_t)stream); }

extern "C" {

KZGB_API int kzgb_srs_create_g1(const kzgb_p1* points, size_t n, int window_bits, kzgb_srs** out) {
    KZG_REQUIRE_INIT();
    KZG_LOCK();
    use_stream(ctx().stream);
    if (!points || !out) { set_error("kzgb_srs_create_g1: null argument"); return KZGB_ERR_INVALID_ARG; }
    kzgb_srs* s = nullptr;
    int st = srs_alloc(&s, n, window_bits);
    if (st) return st;
    cudaStream_t stream = ctx().stream;
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
    KZG_LOCK();
    use_stream(ctx().stream);
    if (!points || !out) { set_error("kzgb_srs_create_g1_affine: null argument"); return KZGB_ERR_INVALID_ARG; }
    kzgb_srs* s = nullptr;
    int st = srs_alloc(&s, n, window_bits);
    if (st) return st;
    cudaStream_t stream = ctx().stream;
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
    KZG_REQUIRE_INIT();
    KZG_LOCK();
    use_stream(ctx().stream);
    cudaDeviceSynchronize();
    srs_free(s);
    return KZGB_OK;
}

KZGB_API size_t kzgb_srs_size(const kzgb_srs* s) { return s ? s->n : 0; }
KZGB_API int kzgb_srs_window_bits(const kzgb_srs* s) { return s ? s->c : 0; }
KZGB_API size_t kzgb_srs_device_bytes(const kzgb_srs* s) {
    return s ? (size_t)s->nwin * s->n * sizeof(kzgb_p1_affine) + s->work_bytes : 0;
}
KZGB_API const void* kzgb_srs_device_points(const kzgb_srs* s) { return s ? s->table : nullptr; }

KZGB_API int kzgb_srs_read_affine(const kzgb_srs* s, size_t start, size_t count, kzgb_p1_affine* out) {
    KZG_REQUIRE_INIT();
    KZG_LOCK();
    use_stream(ctx().stream);
    if (!s || !out || start + count > s->n) { set_error("kzgb_srs_read_affine: bad range"); return KZGB_ERR_INVALID_ARG; }
    KZG_CUDA(cudaStreamSynchronize(ctx().stream));
    KZG_CUDA(cudaMemcpy(out, (const kzgb_p1_affine*)s->table + start, count * sizeof(kzgb_p1_affine),
                        cudaMemcpyDeviceToHost));
    return KZGB_OK;
}

// out_mode: 0 affine (96 B), 1 Jacobian (144 B), 2 raw XYZZ partial (192 B, for multi-GPU combine)
KZGB_API int kzgb_msm_g1_dev(kzgb_srs* s, const void* d_scalars, size_t m, void* d_out, int out_mode, void* stream) {
    KZG_REQUIRE_INIT();
    KZG_LOCK();
    if (!s || (!d_scalars && m) || !d_out || out_mode < 0 || out_mode > 2) {
        set_error("kzgb_msm_g1_dev: bad argument");
        return KZGB_ERR_INVALID_ARG;
    }
    cudaStream_t st = pick_stream(stream);
    const XYZZ* res = nullptr;
    int rc = msm_run(s, d_scalars, m, st, &res);
    if (rc) return rc;
    return msm_finish(res, d_out, out_mode, st);
}

static int msm_host(kzgb_srs* s, const kzgb_fr* scalars, size_t m, void* out, int mode, size_t out_bytes) {
    KZG_REQUIRE_INIT();
    KZG_LOCK();
    use_stream(ctx().stream);
    if (!s || (!scalars && m) || !out) { set_error("kzgb_msm_g1: null argument"); return KZGB_ERR_INVALID_ARG; }
    if (m > s->n) { set_error("kzgb_msm_g1: %zu scalars > %zu SRS points", m, s->n); return KZGB_ERR_INVALID_ARG; }
    Context& c = ctx();
    size_t bytes = m * sizeof(kzgb_fr);
    char* d = (char*)staging_reserve(bytes + 256);
    if (!d) return KZGB_ERR_CUDA;
    if (m) KZG_CUDA(cudaMemcpyAsync(d + 256, scalars, bytes, cudaMemcpyHostToDevice, c.stream));
    const XYZZ* res = nullptr;
    int rc = msm_run(s, d + 256, m, c.stream, &res);
    if (rc) return rc;
    rc = msm_finish(res, d, mode, c.stream);
    if (rc) return rc;
    KZG_CUDA(cudaMemcpyAsync(out, d, out_bytes, cudaMemcpyDeviceToHost, c.stream));
    KZG_CUDA(cudaStreamSynchronize(c.stream));
    return KZGB_OK;
}

KZGB_API int kzgb_msm_g1(kzgb_srs* s, const kzgb_fr* scalars, size_t m, kzgb_p1_affine* out) {
    return msm_host(s, scalars, m, out, 0, sizeof(kzgb_p1_affine));
}
KZGB_API int kzgb_msm_g1_jacobian(kzgb_srs* s, const kzgb_fr* scalars, size_t m, kzgb_p1* out) {
    return msm_host(s, scalars, m, out, 1, sizeof(kzgb_p1));
}

// Multi-GPU combine: `count` raw XYZZ partial sums (192 B each, device) -> one point.
KZGB_API int kzgb_g1_combine_dev(const void* d_partials, uint32_t count, void* d_out, int out_mode, void* stream) {
    KZG_REQUIRE_INIT();
    KZG_LOCK();
    if (!d_partials || !d_out || count == 0) { set_error("kzgb_g1_combine_dev: bad argument"); return KZGB_ERR_INVALID_ARG; }
    cudaStream_t st = pick_stream(stream);
    char* d = (char*)scratch_reserve(256);
    if (!d) return KZGB_ERR_CUDA;
    int rc = msm_combine(d_partials, count, d, st);
    if (rc) return rc;
    return msm_finish((const XYZZ*)d, d_out, out_mode, st);
}

}  // extern "C"
