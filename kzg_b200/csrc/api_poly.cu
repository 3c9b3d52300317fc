// api_poly.cu -- C ABI: quotient division and the witness entry points.
#include "common.cuh"
#include "ec.cuh"
#include "evalform.cuh"
#include "msm.cuh"
#include "scan.cuh"

using namespace kzg;

static size_t al256(size_t v) { return (v + 255) & ~(size_t)255; }

__global__ void fr_is_zero_kernel(const Fr* v, int* flag) {
    if (threadIdx.x == 0 && blockIdx.x == 0) *flag = v->is_zero() ? 0 : 1;
}

extern "C" {

// q = (p - y) div (X - x), rem = p(x) - y.   y may be NULL (plain division).  n >= 1 coefficients, low -> high.
KZGB_API int kzgb_divide_linear(const kzgb_fr* p, size_t n, const kzgb_fr* x, const kzgb_fr* y, kzgb_fr* q_out,
                                kzgb_fr* rem_out) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    use_stream(lane().stream);
    if (!p || !x || n == 0 || (!q_out && n > 1) || !rem_out) {
        set_error("kzgb_divide_linear: bad argument");
        return KZGB_ERR_INVALID_ARG;
    }
    size_t pb = al256(n * 32), ws = al256(scan_workspace_bytes(n));
    char* d = (char*)scratch_reserve(2 * pb + ws + 1024);
    if (!d) return KZGB_ERR_CUDA;
    Fr* d_p = (Fr*)d;
    Fr* d_q = (Fr*)(d + pb);
    char* d_ws = d + 2 * pb;
    Fr* d_x = (Fr*)(d + 2 * pb + ws);
    Fr* d_y = d_x + 1;
    Fr* d_rem = d_x + 2;
    KZG_CUDA(upload_async(d_p, p, n * 32, lane().stream));
    KZG_CUDA(cudaMemcpyAsync(d_x, x, 32, cudaMemcpyHostToDevice, lane().stream));
    if (y) KZG_CUDA(cudaMemcpyAsync(d_y, y, 32, cudaMemcpyHostToDevice, lane().stream));
    int st = scan_divide_linear(d_p, n, d_x, y ? d_y : nullptr, d_q, d_rem, d_ws, lane().stream);
    if (st) return st;
    if (n > 1) KZG_CUDA(download_async(q_out, d_q, (n - 1) * 32, lane().stream));
    KZG_CUDA(cudaMemcpyAsync(rem_out, d_rem, 32, cudaMemcpyDeviceToHost, lane().stream));
    KZG_CUDA(cudaStreamSynchronize(lane().stream));
    return KZGB_OK;
}

// Device variant: d_xy = {x, y} (two Fr); d_q: n-1 elements; d_rem: 1 element.
KZGB_API int kzgb_divide_linear_dev(const void* d_p, size_t n, const void* d_xy, int use_y, void* d_q, void* d_rem,
                                    void* stream) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    if (!d_p || !d_xy || n == 0 || !d_rem || (!d_q && n > 1)) {
        set_error("kzgb_divide_linear_dev: bad argument");
        return KZGB_ERR_INVALID_ARG;
    }
    cudaStream_t st = use_stream((cudaStream_t)stream);
    char* ws = (char*)scratch_reserve(al256(scan_workspace_bytes(n)));
    if (!ws) return KZGB_ERR_CUDA;
    const Fr* xy = (const Fr*)d_xy;
    return scan_divide_linear((const Fr*)d_p, n, xy, use_y ? xy + 1 : nullptr, (Fr*)d_q, (Fr*)d_rem, ws, st);
}

// create_witness (coeff_form.rs:66-81) with device-resident polynomial.  d_p: n = degree+1 coefficients.
// d_xy = {x, y}.  d_work: n Fr of scratch for psi.  d_out: 96 B affine.  Synchronises `stream`.
KZGB_API int kzgb_witness_coeff_dev(kzgb_srs* srs, const void* d_p, size_t n, const void* d_xy, void* d_work,
                                    void* d_out, void* stream) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    if (!srs || !d_p || !d_xy || !d_work || !d_out || n == 0) {
        set_error("kzgb_witness_coeff_dev: bad argument");
        return KZGB_ERR_INVALID_ARG;
    }
    if (n - (n > 1) > srs->n) {
        set_error("kzgb_witness_coeff_dev: quotient has %zu coefficients but the SRS holds %zu points", n - 1, srs->n);
        return KZGB_ERR_INVALID_ARG;
    }
    KZG_LOCK_SRS(srs);
    cudaStream_t st = use_stream((cudaStream_t)stream);
    size_t wsb = al256(scan_workspace_bytes(n));
    char* ws = (char*)scratch_reserve(wsb + 512);
    if (!ws) return KZGB_ERR_CUDA;
    Fr* d_rem = (Fr*)(ws + wsb);
    int* d_flag = (int*)(ws + wsb + 64);
    const Fr* xy = (const Fr*)d_xy;
    int rc = scan_divide_linear((const Fr*)d_p, n, xy, xy + 1, (Fr*)d_work, d_rem, ws, st);
    if (rc) return rc;
    fr_is_zero_kernel<<<1, 32, 0, st>>>(d_rem, d_flag);
    KZG_LAUNCHED();
    // psi has n-1 coefficients (n == 1: the zero polynomial, one coefficient equal to zero)
    const XYZZ* res = nullptr;
    if (n == 1) KZG_CUDA(cudaMemsetAsync(d_work, 0, 32, st));
    rc = msm_run(srs, d_work, n > 1 ? n - 1 : 1, st, &res);
    if (rc) return rc;
    rc = msm_finish(res, d_out, 0, st);
    if (rc) return rc;
    int flag = 0;
    KZG_CUDA(cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    KZG_CUDA(cudaStreamSynchronize(st));
    return flag ? KZGB_POINT_NOT_ON_POLYNOMIAL : KZGB_OK;
}

// create_witness with host buffers (the drop-in call): p = polynomial.slice_coeffs() (n = degree+1).
KZGB_API int kzgb_witness_coeff(kzgb_srs* srs, const kzgb_fr* p, size_t n, const kzgb_fr* x, const kzgb_fr* y,
                                kzgb_p1_affine* out) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    use_stream(lane().stream);
    if (!srs || !p || !x || !y || !out || n == 0) {
        set_error("kzgb_witness_coeff: bad argument");
        return KZGB_ERR_INVALID_ARG;
    }
    KZG_LOCK_SRS(srs);
    // arena: [p | psi | xy | out]; the scan workspace lives in ctx scratch, so use a dedicated allocation here
    size_t pb = al256(n * 32);
    char* buf = (char*)staging_reserve(2 * pb + 512);
    if (!buf) return KZGB_ERR_CUDA;
    Fr* d_p = (Fr*)buf;
    Fr* d_psi = (Fr*)(buf + pb);
    Fr* d_xy = (Fr*)(buf + 2 * pb);
    char* d_out = buf + 2 * pb + 128;
    int rc = KZGB_OK;
    cudaError_t e;
    do {
        if ((e = upload_async(d_p, p, n * 32, lane().stream)) != cudaSuccess ||
            (e = cudaMemcpyAsync(d_xy, x, 32, cudaMemcpyHostToDevice, lane().stream)) != cudaSuccess ||
            (e = cudaMemcpyAsync(d_xy + 1, y, 32, cudaMemcpyHostToDevice, lane().stream)) != cudaSuccess) {
            rc = cuda_fail(e, "H2D witness inputs", __FILE__, __LINE__);
            break;
        }
        rc = kzgb_witness_coeff_dev(srs, d_p, n, d_xy, d_psi, d_out, lane().stream);
        if (rc < 0) break;
        if ((e = cudaMemcpy(out, d_out, sizeof(kzgb_p1_affine), cudaMemcpyDeviceToHost)) != cudaSuccess)
            rc = cuda_fail(e, "D2H witness", __FILE__, __LINE__);
    } while (0);
    return rc;
}

// div_by_omega_i(evals - evals[i], i) (eval_form.rs:58-84,125-131) on host vectors; d = 2^log_d.
KZGB_API int kzgb_div_by_omega_i(const kzgb_fr* evals, uint32_t log_d, size_t i, kzgb_fr* q_out) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    use_stream(lane().stream);
    if (log_d >= 32) return KZGB_DEGREE_TOO_LARGE;
    size_t d = (size_t)1 << log_d;
    if (!evals || !q_out || i >= d) { set_error("kzgb_div_by_omega_i: bad argument"); return KZGB_ERR_INVALID_ARG; }
    if (log_d == 0) { memset(q_out, 0, 32); return KZGB_OK; }  // empty sum (eval_form.rs:63-76)
    size_t pb = al256(d * 32);
    char* buf = (char*)scratch_reserve(2 * pb + al256(div_by_omega_workspace_bytes(log_d)));
    if (!buf) return KZGB_ERR_CUDA;
    KZG_CUDA(upload_async(buf, evals, d * 32, lane().stream));
    int rc = div_by_omega_run((const Fr*)buf, log_d, (uint32_t)i, (Fr*)(buf + pb), buf + 2 * pb, lane().stream);
    if (rc) return rc;
    KZG_CUDA(download_async(q_out, buf + pb, d * 32, lane().stream));
    KZG_CUDA(cudaStreamSynchronize(lane().stream));
    return KZGB_OK;
}

// KZGProverEvalForm::create_witness (eval_form.rs:124-140), device-resident evals.  d_work: d Fr scratch.
KZGB_API int kzgb_witness_eval_dev(kzgb_srs* lagrange_srs, const void* d_evals, uint32_t log_d, size_t i, void* d_work,
                                   void* d_out, void* stream) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    if (log_d >= 32) return KZGB_DEGREE_TOO_LARGE;
    size_t d = (size_t)1 << log_d;
    if (!lagrange_srs || !d_evals || !d_work || !d_out || i >= d || d > lagrange_srs->n) {
        set_error("kzgb_witness_eval_dev: bad argument");
        return KZGB_ERR_INVALID_ARG;
    }
    KZG_LOCK_SRS(lagrange_srs);
    cudaStream_t st = use_stream((cudaStream_t)stream);
    if (log_d == 0) {
        KZG_CUDA(cudaMemsetAsync(d_work, 0, 32, st));
    } else {
        char* ws = (char*)scratch_reserve(al256(div_by_omega_workspace_bytes(log_d)));
        if (!ws) return KZGB_ERR_CUDA;
        int rc = div_by_omega_run((const Fr*)d_evals, log_d, (uint32_t)i, (Fr*)d_work, ws, st);
        if (rc) return rc;
    }
    const XYZZ* res = nullptr;
    int rc = msm_run(lagrange_srs, d_work, d, st, &res);
    if (rc) return rc;
    return msm_finish(res, d_out, 0, st);
}

KZGB_API int kzgb_witness_eval(kzgb_srs* lagrange_srs, const kzgb_fr* evals, uint32_t log_d, size_t i,
                               kzgb_p1_affine* out) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    use_stream(lane().stream);
    if (log_d >= 32) return KZGB_DEGREE_TOO_LARGE;
    size_t d = (size_t)1 << log_d;
    if (!lagrange_srs || !evals || !out || i >= d) { set_error("kzgb_witness_eval: bad argument"); return KZGB_ERR_INVALID_ARG; }
    KZG_LOCK_SRS(lagrange_srs);
    size_t pb = al256(d * 32);
    char* buf = (char*)staging_reserve(2 * pb + 256);
    if (!buf) return KZGB_ERR_CUDA;
    int rc;
    cudaError_t e = upload_async(buf, evals, d * 32, lane().stream);
    if (e != cudaSuccess) rc = cuda_fail(e, "H2D evals", __FILE__, __LINE__);
    else rc = kzgb_witness_eval_dev(lagrange_srs, buf, log_d, i, buf + pb, buf + 2 * pb, lane().stream);
    if (rc == KZGB_OK) {
        e = cudaMemcpyAsync(out, buf + 2 * pb, sizeof(kzgb_p1_affine), cudaMemcpyDeviceToHost, lane().stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(lane().stream);
        if (e != cudaSuccess) rc = cuda_fail(e, "D2H witness", __FILE__, __LINE__);
    }
    cudaStreamSynchronize(lane().stream);
    return rc;
}

}  // extern "C"
