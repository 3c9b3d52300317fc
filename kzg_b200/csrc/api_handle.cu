// api_handle.cu -- device-resident polynomial / evaluation-vector handles.
//
// The reference's provers borrow `&Polynomial` (polynomial.rs:24-27) / `&EvaluationDomain` (ft.rs:17-25) and a prover
// typically issues commit -> create_witness -> create_witness_batched on the SAME polynomial.  Through the
// host-pointer entry points every one of those calls re-uploads the 32 MiB coefficient vector; a kzgb_poly keeps it
// in HBM (one H2D at upload) together with the quotient scratch the witness paths need, so the per-call traffic is
// the point (x, y) in and 96 bytes out.
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "ec.cuh"
#include "msm.cuh"
#include "ntt.cuh"

struct kzgb_poly {
    size_t n = 0;
    kzg::Fr* d_data = nullptr;
    kzg::Fr* d_work = nullptr;   // n + 1 elements of quotient scratch, allocated on first witness
    size_t data_cap = 0, work_cap = 0;
    std::recursive_mutex mu;
};

using namespace kzg;

// Handles come and go with the caller's polynomials; cudaMalloc / cudaFree cost anything from 10 us to 100 ms on a busy
// box (measured: a commit + 8 witnesses sequence took 65 or 184 ms depending on them), so released buffers are parked
// in a small cache and handed to the next handle that fits.
namespace {
struct Parked { void* p; size_t bytes; };
std::mutex g_cache_mu;
std::vector<Parked> g_cache;
constexpr size_t POLY_CACHE_ENTRIES = 8;

cudaError_t poly_alloc(void** out, size_t bytes, size_t* capacity) {
    {
        std::lock_guard<std::mutex> g(g_cache_mu);
        int best = -1;
        for (size_t i = 0; i < g_cache.size(); i++)
            if (g_cache[i].bytes >= bytes && g_cache[i].bytes <= 2 * bytes + 4096 && (best < 0 || g_cache[i].bytes < g_cache[best].bytes))
                best = (int)i;
        if (best >= 0) {
            *out = g_cache[best].p;
            *capacity = g_cache[best].bytes;
            g_cache[best] = g_cache.back();
            g_cache.pop_back();
            return cudaSuccess;
        }
    }
    *capacity = bytes;
    return cudaMalloc(out, bytes);
}
void poly_release(void* p, size_t capacity) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> g(g_cache_mu);
        if (g_cache.size() < POLY_CACHE_ENTRIES) { g_cache.push_back(Parked{p, capacity}); return; }
    }
    cudaFree(p);
}
}  // namespace

namespace kzg {
void poly_cache_shutdown() {
    std::lock_guard<std::mutex> g(g_cache_mu);
    for (auto& e : g_cache) cudaFree(e.p);
    g_cache.clear();
}
}  // namespace kzg

#define KZG_LOCK_POLY(p) std::lock_guard<std::recursive_mutex> _poly_guard((p)->mu)

static int poly_work(kzgb_poly* p) {
    if (!p->d_work) KZG_CUDA(poly_alloc((void**)&p->d_work, (p->n + 1) * sizeof(Fr), &p->work_cap));
    return KZGB_OK;
}

extern "C" {

KZGB_API int kzgb_poly_upload(const kzgb_fr* coeffs, size_t n, kzgb_poly** out) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    if (!coeffs || !out || n == 0) { set_error("kzgb_poly_upload: bad argument"); return KZGB_ERR_INVALID_ARG; }
    kzgb_poly* p = new kzgb_poly();
    p->n = n;
    cudaError_t e = poly_alloc((void**)&p->d_data, n * sizeof(Fr), &p->data_cap);
    if (e == cudaSuccess) e = upload_async(p->d_data, coeffs, n * sizeof(Fr), lane().stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(lane().stream);
    if (e != cudaSuccess) {
        if (p->d_data) cudaFree(p->d_data);
        delete p;
        return cuda_fail(e, "kzgb_poly_upload", __FILE__, __LINE__);
    }
    *out = p;
    return KZGB_OK;
}

KZGB_API int kzgb_poly_free(kzgb_poly* p) {
    if (!p) return KZGB_OK;
    if (ctx().initialised) {   // after kzgb_shutdown the context's memory is gone with the device reset of the caller
        KZG_ENTER();
        // every entry point that uses the handle drains its stream before returning, so nothing is in flight on it
        { KZG_LOCK_POLY(p); }
        poly_release(p->d_data, p->data_cap);
        poly_release(p->d_work, p->work_cap);
    }
    delete p;
    return KZGB_OK;
}

KZGB_API size_t kzgb_poly_len(const kzgb_poly* p) { return p ? p->n : 0; }
KZGB_API const void* kzgb_poly_device_ptr(const kzgb_poly* p) { return p ? p->d_data : nullptr; }

KZGB_API int kzgb_poly_download(kzgb_poly* p, kzgb_fr* out) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    if (!p || !out) { set_error("kzgb_poly_download: null argument"); return KZGB_ERR_INVALID_ARG; }
    KZG_LOCK_POLY(p);
    cudaStream_t st = use_stream(lane().stream);
    KZG_CUDA(download_async(out, p->d_data, p->n * sizeof(Fr), st));
    KZG_CUDA(cudaStreamSynchronize(st));
    return KZGB_OK;
}

// KZGProver::commit (coeff_form.rs:59-64) / KZGProverEvalForm::commit (eval_form.rs:114-122) of the first m
// coefficients (m = polynomial.num_coeffs() <= len).
KZGB_API int kzgb_poly_commit(kzgb_srs* srs, kzgb_poly* p, size_t m, kzgb_p1_affine* out) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    if (!srs || !p || !out || m > p->n) { set_error("kzgb_poly_commit: bad argument"); return KZGB_ERR_INVALID_ARG; }
    KZG_LOCK_POLY(p);
    cudaStream_t st = use_stream(lane().stream);
    char* d_out = (char*)staging_reserve(256);
    if (!d_out) return KZGB_ERR_CUDA;
    int rc = kzgb_msm_g1_dev(srs, p->d_data, m, d_out, 0, st);
    if (rc) return rc;
    KZG_CUDA(cudaMemcpyAsync(out, d_out, sizeof(kzgb_p1_affine), cudaMemcpyDeviceToHost, st));
    KZG_CUDA(cudaStreamSynchronize(st));
    return KZGB_OK;
}

// KZGProver::create_witness (coeff_form.rs:66-81) on the first n coefficients
KZGB_API int kzgb_poly_witness(kzgb_srs* srs, kzgb_poly* p, size_t n, const kzgb_fr* x, const kzgb_fr* y, kzgb_p1_affine* out) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    if (!srs || !p || !x || !y || !out || n == 0 || n > p->n) { set_error("kzgb_poly_witness: bad argument"); return KZGB_ERR_INVALID_ARG; }
    KZG_LOCK_POLY(p);
    int rc = poly_work(p);
    if (rc) return rc;
    cudaStream_t st = use_stream(lane().stream);
    char* d = (char*)staging_reserve(512);
    if (!d) return KZGB_ERR_CUDA;
    KZG_CUDA(cudaMemcpyAsync(d, x, 32, cudaMemcpyHostToDevice, st));
    KZG_CUDA(cudaMemcpyAsync(d + 32, y, 32, cudaMemcpyHostToDevice, st));
    rc = kzgb_witness_coeff_dev(srs, p->d_data, n, d, p->d_work, d + 256, st);
    if (rc < 0) return rc;
    KZG_CUDA(cudaMemcpy(out, d + 256, sizeof(kzgb_p1_affine), cudaMemcpyDeviceToHost));
    return rc;
}

// KZGProver::create_witness_batched (coeff_form.rs:83-111) on the first n coefficients; r_out / r_len as kzgb_witness_batched
KZGB_API int kzgb_poly_witness_batched(kzgb_srs* srs, kzgb_poly* p, size_t n, const kzgb_fr* xs, const kzgb_fr* ys, size_t m,
                                       kzgb_fr* r_out, size_t* r_len, kzgb_p1_affine* w_out) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    if (!srs || !p || !xs || !ys || !r_out || !r_len || !w_out || n == 0 || n > p->n || m == 0) {
        set_error("kzgb_poly_witness_batched: bad argument");
        return KZGB_ERR_INVALID_ARG;
    }
    KZG_LOCK_POLY(p);
    cudaStream_t st = use_stream(lane().stream);
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t mb = al(m * 32), rb = al((m + 2) * 32);
    char* d = (char*)staging_reserve(2 * mb + rb + 256);
    if (!d) return KZGB_ERR_CUDA;
    char *d_xs = d, *d_ys = d + mb, *d_r = d + 2 * mb, *d_out = d + 2 * mb + rb;
    KZG_CUDA(cudaMemcpyAsync(d_xs, xs, m * 32, cudaMemcpyHostToDevice, st));
    KZG_CUDA(cudaMemcpyAsync(d_ys, ys, m * 32, cudaMemcpyHostToDevice, st));
    int rc = kzgb_witness_batched_dev(srs, p->d_data, n, d_xs, d_ys, m, d_r, d_out, st);
    if (rc < 0) return rc;
    const size_t len = m == 1 ? 2 : m;
    KZG_CUDA(cudaMemcpy(r_out, d_r, len * 32, cudaMemcpyDeviceToHost));
    KZG_CUDA(cudaMemcpy(w_out, d_out, sizeof(kzgb_p1_affine), cudaMemcpyDeviceToHost));
    *r_len = len;
    return rc;
}

// KZGProverEvalForm::create_witness (eval_form.rs:124-140): the handle holds the 2^log_d evaluations
KZGB_API int kzgb_poly_witness_eval(kzgb_srs* lagrange_srs, kzgb_poly* evals, uint32_t log_d, size_t i, kzgb_p1_affine* out) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    if (log_d >= 32) return KZGB_DEGREE_TOO_LARGE;
    if (!lagrange_srs || !evals || !out || evals->n != ((size_t)1 << log_d) || i >= evals->n) {
        set_error("kzgb_poly_witness_eval: bad argument");
        return KZGB_ERR_INVALID_ARG;
    }
    KZG_LOCK_POLY(evals);
    int rc = poly_work(evals);
    if (rc) return rc;
    cudaStream_t st = use_stream(lane().stream);
    char* d_out = (char*)staging_reserve(256);
    if (!d_out) return KZGB_ERR_CUDA;
    rc = kzgb_witness_eval_dev(lagrange_srs, evals->d_data, log_d, i, evals->d_work, d_out, st);
    if (rc) return rc;
    KZG_CUDA(cudaMemcpyAsync(out, d_out, sizeof(kzgb_p1_affine), cudaMemcpyDeviceToHost, st));
    KZG_CUDA(cudaStreamSynchronize(st));
    return KZGB_OK;
}

// EvaluationDomain::fft / ifft (ft.rs:111-140) in place on the handle (len must be 2^log_n)
KZGB_API int kzgb_poly_ntt(kzgb_poly* p, uint32_t log_n, int inverse) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    if (log_n >= 32) return KZGB_DEGREE_TOO_LARGE;
    if (!p || p->n != ((size_t)1 << log_n)) { set_error("kzgb_poly_ntt: the handle must hold exactly 2^log_n elements"); return KZGB_ERR_INVALID_ARG; }
    KZG_LOCK_POLY(p);
    cudaStream_t st = use_stream(lane().stream);
    int rc = kzgb_ntt_fr_dev(p->d_data, log_n, inverse, st);
    if (rc) return rc;
    KZG_CUDA(cudaStreamSynchronize(st));
    return KZGB_OK;
}

}  // extern "C"
