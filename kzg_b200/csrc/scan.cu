// scan.cu -- quotient by a linear factor as a parallel prefix (scan).
//
// Replaces `Polynomial::long_division` for the degree-1 divisor X - x
// (polynomial.rs:193-227) as called by `KZGProver::create_witness`
// (coeff_form.rs:66-71).  For divisor X - x schoolbook division is exactly the
// Horner recurrence
//        q_{k-1} = a_k + x * q_k   (k = n-1 .. 1),   rem = a_0 + x * q_0,
// including the reference's zero-run skips (shrink_degree, :218) and its early
// exit on a zero remainder (:209), which only ever leave zeros behind.
// In reversed index j = n-1-k it is h_j = b_j + x*h_{j-1}: a scan over the affine
// maps h -> x*h + b.  Three phases: (1) every thread folds SCAN_L coefficients,
// blocks scan their thread aggregates with strides x^(L*2^k) and keep every thread's
// in-block carry-in; (2) one block scans the block aggregates; (3) every thread
// completes its carry-in with x^(L*t) * (block carry) -- one product, not a second
// block scan -- replays its chunk and writes the quotient.  2.6 products per
// coefficient (4.0 when phase 3 repeated phase 1).  The constant `y` of create_witness (dividend.coeffs[0] -= y,
// coeff_form.rs:67-68) is folded into the load of a_0.
#include "common.cuh"
#include "field.cuh"
#include "scan.cuh"

namespace kzg {

constexpr int SCAN_L = 16;         // coefficients per thread
#ifndef KZG_SCAN_LOG_T
#define KZG_SCAN_LOG_T 7
#endif
constexpr int SCAN_LOG_T = KZG_SCAN_LOG_T;
constexpr int SCAN_THREADS = 1 << SCAN_LOG_T;  // threads per block of the fold / apply kernels (64 / 128 / 256 measured equal at 2^20, 128 best at 2^24; profiles/r02y_scan_restructure.log)
constexpr int SCAN2_THREADS = 256;             // threads of the one block that scans the block aggregates
constexpr int SCAN_POWS = 16;      // pw[]: x^(L*2^i), nine entries used

KZG_DI Fr ld_fr(const Fr* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
KZG_DI void st_fr(Fr* p, const Fr& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

// pw[i] = x^(SCAN_L * 2^i), i < SCAN_POWS_USED: the in-block scan reads pw[0..SCAN_LOG_T), the block-carry pass
// pw[SCAN_LOG_T].
// (One thread, a chain of dependent squarings at ~0.6 us each: computing all SCAN_POWS = 48 entries of the table cost
// 30 us of a 120 us division; the nine that are read cost 8.)
// ptab[t] = x^(SCAN_L * t), t < SCAN_THREADS: what a block's carry-in is worth at thread t (phase 3).
constexpr int SCAN_POWS_USED = SCAN_LOG_T + 1;
__global__ void __launch_bounds__(SCAN_THREADS) scan_powers_kernel(const Fr* __restrict__ x, Fr* __restrict__ pw,
                                                                   Fr* __restrict__ ptab) {
    Fr v = *x;
#pragma unroll 1
    for (int i = 1; i < SCAN_L; i <<= 1) v = v.sqr();   // x^SCAN_L (SCAN_L is a power of two)
    ptab[threadIdx.x] = v.pow_u64(threadIdx.x);
    if (threadIdx.x) return;
#pragma unroll 1
    for (int i = 0; i < SCAN_POWS_USED; i++) {
        pw[i] = v;
        v = v.sqr();
    }
}

// b_j = a_{n-1-j}, with y subtracted from a_0 (j = n-1)
KZG_DI Fr load_b(const Fr* __restrict__ a, size_t n, size_t j, const Fr* __restrict__ y) {
    size_t k = n - 1 - j;
    Fr v = ld_fr(&a[k]);
    if (k == 0 && y) v = v - *y;
    return v;
}

// Inclusive block scan of S_t <- sum_{u<=t} X^(t-u) * S_u  with X^(2^k) = pw[pw_off + k]
template <int NT>
KZG_DI Fr block_scan(Fr v, Fr* sh, const Fr* __restrict__ pw, int pw_off) {
    const int t = threadIdx.x;
    sh[t] = v;
    __syncthreads();
#pragma unroll 1
    for (int k = 0, d = 1; d < NT; k++, d <<= 1) {
        Fr other;
        bool has = t >= d;
        if (has) other = sh[t - d];
        __syncthreads();
        if (has) {
            v = v + other * pw[pw_off + k];
            sh[t] = v;
        }
        __syncthreads();
    }
    return v;
}

// phase 1: block aggregates, and every thread's carry-in as if the block started from zero (excl).
__global__ void __launch_bounds__(SCAN_THREADS) scan_fold_kernel(const Fr* __restrict__ a, size_t n,
                                                                 const Fr* __restrict__ x, const Fr* __restrict__ y,
                                                                 const Fr* __restrict__ pw, Fr* __restrict__ block_agg,
                                                                 Fr* __restrict__ excl) {
    __shared__ Fr sh[SCAN_THREADS];
    const size_t g = blockIdx.x * (size_t)SCAN_THREADS + threadIdx.x;
    const size_t j0 = g * SCAN_L;
    const Fr xv = *x;
    Fr h = Fr::zero();   // local Horner; past the end a chunk is SCAN_L zero coefficients: the zero map's value
    if (j0 < n) {
        size_t jend = min(j0 + SCAN_L, n);
        for (size_t j = j0; j < jend; j++) h = load_b(a, n, j, y) + h * xv;
        // a partial last chunk must still act as a full-length map for the threads after it; there are
        // none inside [0, n), so only its own value matters.
    }
    Fr incl = block_scan<SCAN_THREADS>(h, sh, pw, 0);
    if (threadIdx.x == SCAN_THREADS - 1) block_agg[blockIdx.x] = incl;
    if (j0 < n) st_fr(&excl[g], threadIdx.x == 0 ? Fr::zero() : sh[threadIdx.x - 1]);
}

// phase 3: carry-in = excl + x^(L*t) * (carry into the block), then the chunk is replayed and written.
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const Fr* __restrict__ a, size_t n,
                                                                  const Fr* __restrict__ x, const Fr* __restrict__ y,
                                                                  const Fr* __restrict__ ptab,
                                                                  const Fr* __restrict__ block_carry,
                                                                  const Fr* __restrict__ excl, Fr* __restrict__ q,
                                                                  Fr* __restrict__ rem) {
    const size_t g = blockIdx.x * (size_t)SCAN_THREADS + threadIdx.x;
    const size_t j0 = g * SCAN_L;
    if (j0 >= n) return;
    const Fr xv = *x;
    Fr hh = ld_fr(&block_carry[blockIdx.x]);
    if (threadIdx.x) hh = ld_fr(&excl[g]) + hh * ld_fr(&ptab[threadIdx.x]);
    size_t jend = min(j0 + SCAN_L, n);
    for (size_t j = j0; j < jend; j++) {
        hh = load_b(a, n, j, y) + hh * xv;
        if (j == n - 1) *rem = hh;
        else st_fr(&q[n - 2 - j], hh);
    }
}

// phase 2: carries of the block aggregates.  One block; thread t owns `per` consecutive aggregates.
// X_blk = pw[SCAN_LOG_T] = x^(L * SCAN_THREADS).
__global__ void __launch_bounds__(SCAN2_THREADS) scan_blocks_kernel(const Fr* __restrict__ agg, size_t nblocks,
                                                                   size_t per, const Fr* __restrict__ pw,
                                                                   Fr* __restrict__ carry) {
    __shared__ Fr sh[SCAN2_THREADS];
    __shared__ Fr spw[8];
    const Fr xb = pw[SCAN_LOG_T];
    const size_t lo = threadIdx.x * per;
    size_t hi = min(lo + per, nblocks);
    Fr h = Fr::zero();
    for (size_t i = lo; i < hi; i++) h = ld_fr(&agg[i]) + h * xb;
    if (lo >= nblocks) h = Fr::zero();
    // scan across threads with multiplier xb^per
    if (threadIdx.x == 0) {
        Fr m = xb.pow_u64(per);
        for (int k = 0; k < 8; k++) { spw[k] = m; m = m.sqr(); }
    }
    // threads past the end hold zero maps; pad them as `per` zero coefficients: value 0 is already right
    // because their inclusive values are never read.
    __syncthreads();
    Fr incl = block_scan<SCAN2_THREADS>(h, sh, spw, 0);
    (void)incl;
    Fr c = threadIdx.x == 0 ? Fr::zero() : sh[threadIdx.x - 1];
    for (size_t i = lo; i < hi; i++) {
        st_fr(&carry[i], c);
        c = ld_fr(&agg[i]) + c * xb;
    }
}

size_t scan_workspace_bytes(size_t n) {
    size_t nblocks = (n + (size_t)SCAN_L * SCAN_THREADS - 1) / ((size_t)SCAN_L * SCAN_THREADS);
    return (SCAN_POWS + SCAN_THREADS + 2 * nblocks + nblocks * SCAN_THREADS + 4) * sizeof(Fr);
}

// d_a: n coefficients (low -> high).  d_q: n-1 quotient coefficients.  d_rem: 1 element (= p(x) - y).
// d_x, d_y: device scalars (d_y may be null).  d_ws: scan_workspace_bytes(n).
int scan_divide_linear(const Fr* d_a, size_t n, const Fr* d_x, const Fr* d_y, Fr* d_q, Fr* d_rem, void* d_ws,
                       cudaStream_t stream) {
    if (n == 0) { set_error("divide_linear: empty polynomial"); return KZGB_ERR_INVALID_ARG; }
    StageScope sc(ST_SCAN, stream);
    size_t nblocks = (n + (size_t)SCAN_L * SCAN_THREADS - 1) / ((size_t)SCAN_L * SCAN_THREADS);
    Fr* pw = (Fr*)d_ws;
    Fr* ptab = pw + SCAN_POWS;
    Fr* agg = ptab + SCAN_THREADS;
    Fr* carry = agg + nblocks;
    Fr* excl = carry + nblocks;   // one per thread of phase 1
    scan_powers_kernel<<<1, SCAN_THREADS, 0, stream>>>(d_x, pw, ptab);
    KZG_LAUNCHED();
    scan_fold_kernel<<<(unsigned)nblocks, SCAN_THREADS, 0, stream>>>(d_a, n, d_x, d_y, pw, agg, excl);
    KZG_LAUNCHED();
    size_t per = (nblocks + SCAN2_THREADS - 1) / SCAN2_THREADS;
    scan_blocks_kernel<<<1, SCAN2_THREADS, 0, stream>>>(agg, nblocks, per, pw, carry);
    KZG_LAUNCHED();
    scan_apply_kernel<<<(unsigned)nblocks, SCAN_THREADS, 0, stream>>>(d_a, n, d_x, d_y, ptab, carry, excl, d_q, d_rem);
    KZG_LAUNCHED();
    return KZGB_OK;
}

}  // namespace kzg
