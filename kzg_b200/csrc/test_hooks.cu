// test_hooks.cu -- bulk element-wise entry points used by the parity tests to
// pin the device field arithmetic against the oracle.
#define KZG_FQ_DEDICATED_SQR 1
#include "common.cuh"
#include "field.cuh"

namespace kzg {

// op: 0 mul, 1 add, 2 sub, 3 inv(a), 4 to_mont(a), 5 from_mont(a), 6 sqr(a), 7 inv_fermat(a), 8 mul via the FP64 pipe,
//     9 Karatsuba mul, 10/11 lazy-reduction a*b - c*d, 12 inv by Pornin's binary GCD
template <class F>
__global__ void fe_binop_kernel(const F* __restrict__ a, const F* __restrict__ b, F* __restrict__ o, size_t n, int op) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    F x = a[i], y = b ? b[i] : F::zero(), r;
    switch (op) {
        case 0: r = x * y; break;
        case 1: r = x + y; break;
        case 2: r = x - y; break;
        case 3: r = x.inv(); break;
        case 4: r = x.to_mont(); break;
        case 5: r = x.from_mont(); break;
        case 7: r = x.inv_fermat(); break;
        case 8: r = F::mul_dfma(x, y); break;
        case 9: r = F::mul_kara(x, y); break;                      // Karatsuba product + ripple-free reduction
        case 10: r = F::mul_sub_mul(x, y, x + y, y); break;        // lazy reduction: x*y - (x+y)*y = -y^2
        case 11: r = F::mul_sub_mul(x, y, y, x.sqr()); break;      // x*y - y*x^2
        case 12: r = x.inv_bingcd(); break;                        // experiment: binary GCD on word-size approximations
        default: r = x.sqr(); break;
    }
    o[i] = r;
}

template <class F>
static int fe_binop(const void* a, const void* b, void* out, size_t n, int op) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    Context& c = ctx();
    (void)c;
    use_stream(lane().stream);
    if (n == 0) return KZGB_OK;
    if (!a || !out || ((op <= 2 || (op >= 8 && op <= 11)) && !b)) {
        set_error("test binop: null pointer");
        return KZGB_ERR_INVALID_ARG;
    }
    size_t bytes = n * sizeof(F);
    char* d = (char*)scratch_reserve(3 * bytes);
    if (!d) return KZGB_ERR_CUDA;
    KZG_CUDA(cudaMemcpyAsync(d, a, bytes, cudaMemcpyHostToDevice, lane().stream));
    if (b) KZG_CUDA(cudaMemcpyAsync(d + bytes, b, bytes, cudaMemcpyHostToDevice, lane().stream));
    fe_binop_kernel<F><<<div_up(n, 128), 128, 0, lane().stream>>>((const F*)d, b ? (const F*)(d + bytes) : nullptr,
                                                             (F*)(d + 2 * bytes), n, op);
    KZG_LAUNCHED();
    KZG_CUDA(cudaMemcpyAsync(out, d + 2 * bytes, bytes, cudaMemcpyDeviceToHost, lane().stream));
    KZG_CUDA(cudaStreamSynchronize(lane().stream));
    return KZGB_OK;
}

}  // namespace kzg

extern "C" {
KZGB_API int kzgb_test_fr_op(const kzgb_fr* a, const kzgb_fr* b, kzgb_fr* out, size_t n, int op) {
    return kzg::fe_binop<kzg::Fr>(a, b, out, n, op);
}
KZGB_API int kzgb_test_fp_op(const kzgb_fp* a, const kzgb_fp* b, kzgb_fp* out, size_t n, int op) {
    return kzg::fe_binop<kzg::Fq>(a, b, out, n, op);
}
}
