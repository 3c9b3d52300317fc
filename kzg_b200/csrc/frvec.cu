// frvec.cu -- element-wise Fr vector helpers of EvaluationDomain (not on any prover path):
//   distribute_powers (ft.rs:142-166), coset_fft / icoset_fft (ft.rs:168-178), divide_by_z_on_coset (ft.rs:192-217),
//   mul_assign (ft.rs:220-244), sub_assign (ft.rs:247-271), and the transform pipeline of Polynomial::fft_mul
//   (polynomial.rs:167-183).  All memory-bound: one read-modify-write pass over 32-byte elements.
#include "common.cuh"
#include "field.cuh"
#include "fr_util.cuh"
#include "ntt.cuh"

namespace kzg {

KZG_DI Fr v_ld(const Fr* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
KZG_DI void v_st(Fr* p, const Fr& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

// op 0: a *= b   op 1: a -= b   op 3: a *= k
__global__ void fr_vec_binop_kernel(Fr* __restrict__ a, const Fr* __restrict__ b, size_t n, int op, Fr k) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr x = v_ld(&a[i]);
    if (op == 0) x = x * v_ld(&b[i]);
    else if (op == 1) x = x - v_ld(&b[i]);
    else x = x * k;
    v_st(&a[i], x);
}

// a[i] *= g^i
constexpr int DP_CHUNK = 16;
__global__ void fr_distribute_powers_kernel(Fr* __restrict__ a, size_t n, Fr g) {
    size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t lo = t * DP_CHUNK;
    if (lo >= n) return;
    size_t hi = min(lo + DP_CHUNK, n);
    Fr u = g.pow_u64(lo);
    for (size_t i = lo; i < hi; i++) {
        v_st(&a[i], v_ld(&a[i]) * u);
        u = u * g;
    }
}

int fr_vec_op_run(int op, Fr* d_a, const Fr* d_b, size_t n, const Fr& k, cudaStream_t stream) {
    if (n == 0) return KZGB_OK;
    if (op == 2) {
        fr_distribute_powers_kernel<<<div_up((n + DP_CHUNK - 1) / DP_CHUNK, 128), 128, 0, stream>>>(d_a, n, k);
    } else if (op == 0 || op == 1 || op == 3) {
        if (op != 3 && !d_b) { set_error("fr_vec_op: second operand missing"); return KZGB_ERR_INVALID_ARG; }
        fr_vec_binop_kernel<<<div_up(n, 256), 256, 0, stream>>>(d_a, d_b, n, op, k);
    } else {
        set_error("fr_vec_op: unknown op %d", op);
        return KZGB_ERR_INVALID_ARG;
    }
    KZG_LAUNCHED();
    return KZGB_OK;
}

static Fr fr_from_host(const FrHost& h) {
    Fr r;
    memcpy(r.l, h.l, 32);
    return r;
}

// coset transform on device data (tmp: scratch of the same size when log_n > 10)
int coset_ntt_run(Fr* d_a, Fr* d_tmp, uint32_t log_n, bool inverse, cudaStream_t stream) {
    if (log_n >= 32) return KZGB_DEGREE_TOO_LARGE;
    const size_t n = (size_t)1 << log_n;
    FrHost gen = fr_host_from_u64(7);  // Scalar::multiplicative_generator()
    int st;
    if (!inverse) {  // ft.rs:168-171
        if ((st = fr_vec_op_run(2, d_a, nullptr, n, fr_from_host(gen), stream))) return st;
        return ntt_run(d_a, d_tmp, log_n, false, stream);
    }
    // ft.rs:173-178
    if ((st = ntt_run(d_a, d_tmp, log_n, true, stream))) return st;
    return fr_vec_op_run(2, d_a, nullptr, n, fr_from_host(fr_host_inv(gen)), stream);
}

}  // namespace kzg

using namespace kzg;

extern "C" {

// op 0: a[i] *= b[i] (mul_assign)   1: a[i] -= b[i] (sub_assign)   2: a[i] *= k^i (distribute_powers)   3: a[i] *= k
KZGB_API int kzgb_fr_vec_op_dev(int op, void* d_a, const void* d_b, size_t n, const kzgb_fr* k, void* stream) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    if ((!d_a && n) || ((op == 2 || op == 3) && !k)) { set_error("kzgb_fr_vec_op_dev: bad argument"); return KZGB_ERR_INVALID_ARG; }
    cudaStream_t st = use_stream((cudaStream_t)stream);
    Fr kk = Fr::zero_host();
    if (k) memcpy(kk.l, k->l, 32);
    return fr_vec_op_run(op, (Fr*)d_a, (const Fr*)d_b, n, kk, st);
}

KZGB_API int kzgb_fr_vec_op(int op, kzgb_fr* a, const kzgb_fr* b, size_t n, const kzgb_fr* k) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    Context& c = ctx();
    (void)c;
    use_stream(lane().stream);
    if ((!a && n) || ((op == 0 || op == 1) && !b && n) || ((op == 2 || op == 3) && !k)) {
        set_error("kzgb_fr_vec_op: bad argument");
        return KZGB_ERR_INVALID_ARG;
    }
    if (n == 0) return KZGB_OK;
    size_t bytes = n * 32, pb = (bytes + 255) & ~(size_t)255;
    char* d = (char*)staging_reserve(2 * pb);
    if (!d) return KZGB_ERR_CUDA;
    KZG_CUDA(upload_async(d, a, bytes, lane().stream));
    if (op == 0 || op == 1) KZG_CUDA(upload_async(d + pb, b, bytes, lane().stream));
    Fr kk = Fr::zero_host();
    if (k) memcpy(kk.l, k->l, 32);
    int st = fr_vec_op_run(op, (Fr*)d, (const Fr*)(d + pb), n, kk, lane().stream);
    if (st) return st;
    KZG_CUDA(download_async(a, d, bytes, lane().stream));
    KZG_CUDA(cudaStreamSynchronize(lane().stream));
    return KZGB_OK;
}

// EvaluationDomain::coset_fft (ft.rs:168-171) / icoset_fft (ft.rs:173-178) on a host vector
KZGB_API int kzgb_coset_ntt_fr(kzgb_fr* inout, uint32_t log_n, int inverse) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    Context& c = ctx();
    (void)c;
    use_stream(lane().stream);
    if (log_n >= 32) return KZGB_DEGREE_TOO_LARGE;
    if (!inout) { set_error("kzgb_coset_ntt_fr: null pointer"); return KZGB_ERR_INVALID_ARG; }
    size_t bytes = (size_t)32 << log_n;
    char* d = (char*)staging_reserve(2 * bytes);
    if (!d) return KZGB_ERR_CUDA;
    KZG_CUDA(upload_async(d, inout, bytes, lane().stream));
    int st = coset_ntt_run((Fr*)d, (Fr*)(d + bytes), log_n, inverse != 0, lane().stream);
    if (st) return st;
    KZG_CUDA(download_async(inout, d, bytes, lane().stream));
    KZG_CUDA(cudaStreamSynchronize(lane().stream));
    return KZGB_OK;
}

// Product of two coefficient vectors through the transform pipeline of Polynomial::fft_mul (polynomial.rs:167-183):
// both padded to the power of two m >= na + nb, fft, pointwise product, ifft.  out must hold m elements; *m_out = m.
KZGB_API int kzgb_poly_mul(const kzgb_fr* a, size_t na, const kzgb_fr* b, size_t nb, kzgb_fr* out, size_t out_capacity,
                           size_t* m_out) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    Context& c = ctx();
    (void)c;
    use_stream(lane().stream);
    if (!a || !b || !out || !m_out || na == 0 || nb == 0) { set_error("kzgb_poly_mul: bad argument"); return KZGB_ERR_INVALID_ARG; }
    uint32_t log_m = 0;
    while (((size_t)1 << log_m) < na + nb) {
        log_m++;
        if (log_m >= 32) return KZGB_DEGREE_TOO_LARGE;  // from_coeffs(...).unwrap() in the reference
    }
    size_t m = (size_t)1 << log_m;
    if (out_capacity < m) { set_error("kzgb_poly_mul: output needs %zu elements", m); return KZGB_ERR_INVALID_ARG; }
    size_t bytes = m * 32;
    char* d = (char*)staging_reserve(3 * bytes);
    if (!d) return KZGB_ERR_CUDA;
    Fr* da = (Fr*)d;
    Fr* db = (Fr*)(d + bytes);
    Fr* tmp = (Fr*)(d + 2 * bytes);
    KZG_CUDA(cudaMemsetAsync(d, 0, 2 * bytes, lane().stream));
    KZG_CUDA(upload_async(da, a, na * 32, lane().stream));
    KZG_CUDA(upload_async(db, b, nb * 32, lane().stream));
    int st;
    if ((st = ntt_run(da, tmp, log_m, false, lane().stream))) return st;
    if ((st = ntt_run(db, tmp, log_m, false, lane().stream))) return st;
    if ((st = fr_vec_op_run(0, da, db, m, Fr::zero_host(), lane().stream))) return st;
    if ((st = ntt_run(da, tmp, log_m, true, lane().stream))) return st;
    KZG_CUDA(download_async(out, da, bytes, lane().stream));
    KZG_CUDA(cudaStreamSynchronize(lane().stream));
    *m_out = m;
    return KZGB_OK;
}

}  // extern "C"
