// batched.cu -- the batched-opening quotient of `KZGProver::create_witness_batched`
// (coeff_form.rs:83-111):
//     r   = interpolation through (x_i, y_i)                 (polynomial.rs:237-264, 303-365)
//     psi = (p - r) div Z,  Z = prod (X - x_i)                (polynomial.rs:443-460, 193-227)
//     Err(PointNotOnPolynomial) iff the remainder is non-zero (coeff_form.rs:96)
//
// Both results are unique polynomials, so they are computed by algebraically
// different, parallel formulations and compared bit-for-bit with the reference
// restatement in the tests:
//  * r_k = sum_{l>k} Z_l * P_{l-k-1},  P_e = sum_i c_i x_i^e,  c_i = y_i / Z'(x_i)
//    (O(m^2) work in three fully parallel steps instead of the reference's
//    recursive subproduct tree);
//  * psi by exact division on a coset: psi~ = iNTT( NTT(A.s^j) / NTT(Z.s^j) ) . s^-j with
//    A = p - r, N = 2^k >= len(A).  psi~ Z = A mod (X^N - s^N) and deg psi~ < N, hence
//    Z | A  <=>  the top m coefficients of psi~ vanish -- a deterministic test that
//    replaces the reference's "remainder is the zero polynomial".  The shift s is the
//    first power of 7 whose N-th power differs from every x_i^N, so Z has no root
//    on the coset.  (The reference's schoolbook loop costs (n-m)(m+1) sequential
//    multiply-subtracts; this is three length-N NTTs.)
#include "common.cuh"
#include "field.cuh"
#include "fr_util.cuh"
#include "msm.cuh"
#include "ntt.cuh"
#include "batched.cuh"
#include "batchinv.cuh"

namespace kzg {

KZG_DI Fr b_ld(const Fr* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
KZG_DI void b_st(Fr* p, const Fr& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

// ---- K11: interpolation ------------------------------------------------------
// These kernels work on m <= 16384 (typically a few hundred) values: a handful of warps each, so their cost is the LENGTH OF THE DEPENDENT CHAIN
// per thread (a dependent Fr product costs ~0.3-0.6 us), not arithmetic.  Every output is therefore computed by a
// group of IL = 32 adjacent lanes (a warp; 8 until the interpolation became the critical path of the batched opening) that split the chain and combine with shuffles (field sums and products are exact,
// so the grouping does not change a single bit of the result).
constexpr uint32_t IL = 32;
KZG_DI Fr fr_shfl_xor(const Fr& v, int mask) {
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = __shfl_xor_sync(0xffffffffu, v.l[i], mask);
    return r;
}

// Z(X) = prod (X - x_i): one block; two roots per step (new[k] = old[k-2] - (a+b) old[k-1] + ab old[k]), coefficients
// ping-pong between two buffers in shared memory (global memory for m too large to fit).
__global__ void __launch_bounds__(1024) interp_product_kernel(const Fr* __restrict__ xs, uint32_t m, Fr* buf0, Fr* buf1,
                                                              Fr* __restrict__ z_out, int use_smem) {
    extern __shared__ uint4 interp_sm[];
    Fr* cur = use_smem ? reinterpret_cast<Fr*>(interp_sm) : buf0;
    Fr* nxt = use_smem ? reinterpret_cast<Fr*>(interp_sm) + (m + 1) : buf1;
    for (uint32_t k = threadIdx.x; k <= m; k += blockDim.x) cur[k] = (k == 0) ? Fr::one() : Fr::zero();
    __syncthreads();
    uint32_t i = 0;
    for (; i + 1 < m; i += 2) {
        const Fr a = xs[i], b = xs[i + 1];
        const Fr sum = a + b, prod = a * b;
        // degree i -> i+2 (old[j] = 0 outside 0..i)
        for (uint32_t k = threadIdx.x; k <= i + 2; k += blockDim.x) {
            Fr v = (k >= 2) ? cur[k - 2] : Fr::zero();
            if (k >= 1 && k - 1 <= i) v = v - sum * cur[k - 1];
            if (k <= i) v = v + prod * cur[k];
            nxt[k] = v;
        }
        __syncthreads();
        Fr* t = cur; cur = nxt; nxt = t;
    }
    if (i < m) {  // odd m: one more root.  degree i -> i+1 :  new[k] = old[k-1] - x_i * old[k]
        const Fr xi = xs[i];
        for (uint32_t k = threadIdx.x; k <= i + 1; k += blockDim.x) {
            Fr v = (k >= 1) ? cur[k - 1] : Fr::zero();
            if (k <= i) v = v - xi * cur[k];
            nxt[k] = v;
        }
        __syncthreads();
        Fr* t = cur; cur = nxt; nxt = t;
    }
    for (uint32_t k = threadIdx.x; k <= m; k += blockDim.x) z_out[k] = cur[k];
}

// Z(X) = prod (X - x_i) through its values: E[j] = Z(w^j) on the N-point domain, N = 2^log_n > m, followed by ONE inverse
// NTT (deg Z = m < N, so the coefficients come back exactly).  The sequential product above is a chain of m/2 dependent
// steps (~4 us each: 557 us at m = 256, ~10 ms at m = 4096); here every value is an independent product of m factors
// split over IL lanes.  tw = the domain's half twiddle table (w^e, e < N/2).
__global__ void __launch_bounds__(128) interp_zeval_kernel(const Fr* __restrict__ xs, uint32_t m, const Fr* __restrict__ tw,
                                                           uint32_t log_n, Fr* __restrict__ E) {
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t j = gid / IL, r = gid % IL;
    const uint32_t N = 1u << log_n, half = N >> 1;
    const bool active = j < N;
    Fr d = Fr::one();
    if (active) {
        Fr w = b_ld(&tw[j & (half - 1)]);
        if (j & half) w = w.neg();
        for (uint32_t i = r; i < m; i += IL) d = d * (w - xs[i]);
    }
    for (int o = IL / 2; o > 0; o >>= 1) d = d * fr_shfl_xor(d, o);
    if (active && r == 0) b_st(&E[j], d);
}

// c_i = y_i / prod_{j != i} (x_i - x_j);   flag |= 2 when two x coincide (the reference panics:
// `c.invert().unwrap()`, polynomial.rs:260).  IL lanes per i: lane r takes the factors j = r (mod IL).
__global__ void __launch_bounds__(128) interp_weights_kernel(const Fr* __restrict__ xs, const Fr* __restrict__ ys, uint32_t m,
                                                             Fr* __restrict__ c, int* __restrict__ flag) {
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t i = gid / IL, r = gid % IL;
    const bool active = i < m;     // inactive lanes still take part in the shuffles
    Fr d = Fr::one();
    if (active) {
        const Fr xi = xs[i];
        for (uint32_t j = r; j < m; j += IL)
            if (j != i) d = d * (xi - xs[j]);
    }
    for (int o = IL / 2; o > 0; o >>= 1) d = d * fr_shfl_xor(d, o);
    // one inversion per block (batchinv.cuh): group leaders bring their product, every other lane brings 1.  A zero
    // product (two equal x) is replaced by 1 and flagged, so that it cannot zero the block's other results.
    __shared__ Fr sh[2 * 128 + 1];
    const bool lead = active && r == 0;
    const bool zero = lead && d.is_zero();
    if (zero) atomicOr(flag, 2);
    Fr dinv = block_inverse<128>((lead && !zero) ? d : Fr::one(), sh);
    if (lead) c[i] = zero ? Fr::zero() : ys[i] * dinv;
}

// T[e][i] = c_i * x_i^e.  IL lanes per i: lane r walks e = r (mod IL) with ratio x_i^IL.
__global__ void __launch_bounds__(128) interp_powers_kernel(const Fr* __restrict__ xs, const Fr* __restrict__ c, uint32_t m,
                                                            Fr* __restrict__ T) {
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t i = gid / IL, r = gid % IL;
    if (i >= m) return;
    const Fr xi = xs[i];
    Fr step = xi;
#pragma unroll
    for (uint32_t t = 1; t < IL; t <<= 1) step = step.sqr();   // x^IL (IL is a power of two)
    Fr v = c[i];
    v = v * xi.pow_u64(r);                                     // c * x^r
    for (uint32_t e = r; e < m; e += IL) {
        b_st(&T[(size_t)e * m + i], v);
        v = v * step;
    }
}

// P_e = sum_i T[e][i]
__global__ void __launch_bounds__(128) interp_powersum_kernel(const Fr* __restrict__ T, uint32_t m, Fr* __restrict__ P) {
    __shared__ Fr sh[128];
    const uint32_t e = blockIdx.x;
    Fr acc = Fr::zero();
    for (uint32_t i = threadIdx.x; i < m; i += 128) acc = acc + b_ld(&T[(size_t)e * m + i]);
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 64; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sh[threadIdx.x] = sh[threadIdx.x] + sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) P[e] = sh[0];
}

// r_k = sum_{l=k+1}^{m} Z_l * P_{l-k-1}.  IL lanes per k: lane r takes l = k+1+r (mod IL).
__global__ void __launch_bounds__(128) interp_conv_kernel(const Fr* __restrict__ Z, const Fr* __restrict__ P, uint32_t m,
                                                          Fr* __restrict__ r_out) {
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t k = gid / IL, r = gid % IL;
    const bool active = k < m;
    Fr acc = Fr::zero();
    if (active)
        for (uint32_t l = k + 1 + r; l <= m; l += IL) acc = acc + Z[l] * P[l - k - 1];
    for (int o = IL / 2; o > 0; o >>= 1) acc = acc + fr_shfl_xor(acc, o);
    if (active && r == 0) r_out[k] = acc;
}

// one-point case: r = X + (y - x)  (polynomial.rs:244-247 -- reproduced, not "fixed")
__global__ void interp_single_kernel(const Fr* __restrict__ xs, const Fr* __restrict__ ys, Fr* __restrict__ r,
                                     Fr* __restrict__ Z) {
    if (threadIdx.x || blockIdx.x) return;
    r[0] = ys[0] - xs[0];
    r[1] = Fr::one();
    Z[0] = xs[0].neg();
    Z[1] = Fr::one();
}

// ---- K10: exact division by Z on a coset ---------------------------------------
// s = 7^k for the smallest k >= 1 with (7^k)^N not in { x_i^N }.
// out[0] = s, out[1] = s^-1, out[2 + i] = s^(2^i), out[2 + CS_TAB + i] = s^-(2^i) (i < CS_TAB): the scaling kernels
// build s^lo from the table instead of a square-and-multiply ladder per thread.  7, 7^-1 and 7^N come from the host
// (fr_util.cuh): no inversion and no squaring ladder on this single-block kernel's critical path.
// The FIRST candidate is s = w_2N, the primitive 2N-th root of unity (s^N = -1, usable unless some x_i^N = -1): on that
// coset every power s^j, (s w_N^a)^k, s^-j is an entry of the 2N-domain's twiddle table, so the scaling kernels do one
// table lookup and one product per element instead of a power ladder and two products (out[CS_MODE].l[0] = 1).
constexpr int CS_TAB = 32;
constexpr int CS_MODE = 2 + 2 * CS_TAB;   // index of the mode word in the shift buffer
__global__ void __launch_bounds__(256) coset_shift_kernel(const Fr* __restrict__ xs, uint32_t m, uint32_t log_n,
                                                          Fr seven, Fr seven_inv, Fr gN, Fr w2n, Fr w2n_inv, uint32_t allow_root,
                                                          Fr* __restrict__ out) {
    __shared__ int bad, bad_root;
    __shared__ Fr s_sh[2];
    Fr cand = Fr::one();
    Fr s = Fr::one(), sinv = Fr::one();
    const Fr minus_one = Fr::one().neg();
    uint32_t mode = 0;
    for (uint32_t k = 1; k <= m + 1; k++) {
        cand = cand * gN;  // (7^k)^N
        s = s * seven;
        sinv = sinv * seven_inv;
        if (threadIdx.x == 0) { bad = 0; bad_root = 0; }
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) {
            Fr v = xs[i];
            for (uint32_t t = 0; t < log_n; t++) v = v.sqr();
            if (v == cand) bad = 1;
            if (v == minus_one) bad_root = 1;
        }
        __syncthreads();
        const int b = bad, br = bad_root;
        __syncthreads();
        if (k == 1 && allow_root && !br) {   // the root-of-unity coset is free of roots of Z
            s = w2n;
            sinv = w2n_inv;
            mode = 1;
            break;
        }
        if (!b) break;
    }
    if (threadIdx.x == 0) {
        out[0] = s;
        out[1] = sinv;
        Fr md = Fr::zero();
        md.l[0] = mode;
        out[CS_MODE] = md;
        s_sh[0] = s;
        s_sh[1] = sinv;
    }
    __syncthreads();
    if (threadIdx.x < 2 * CS_TAB) {   // thread i: i squarings of s, thread CS_TAB + i: of s^-1
        const uint32_t i = threadIdx.x % CS_TAB;
        Fr v = s_sh[threadIdx.x / CS_TAB];
        for (uint32_t t = 0; t < i; t++) v = v.sqr();
        out[2 + threadIdx.x] = v;
    }
}

// base^e from the table tab[i] = base^(2^i)
KZG_DI Fr pow_from_table(const Fr* __restrict__ tab, size_t e) {
    if (e == 0) return Fr::one();
    int i = 0;
    while (!(e & 1)) { e >>= 1; i++; }
    Fr r = tab[i];
    for (e >>= 1, i++; e; e >>= 1, i++)
        if (e & 1) r = r * tab[i];
    return r;
}

constexpr int CS_CHUNK = 16;
// dst[j] = (a[j] - b[j]) * s^j for j < len (a has na entries, b has nb entries; missing entries are zero),
// and 0 for len <= j < N.
// tw2 = half twiddle table of the 2N-point domain (w_2N^e, e < N); used when the shift is w_2N (mode word = 1)
__global__ void coset_scale_kernel(const Fr* __restrict__ a, size_t na, const Fr* __restrict__ b, size_t nb, size_t len,
                                   size_t N, const Fr* __restrict__ shift, const Fr* __restrict__ tw2, Fr* __restrict__ dst) {
    size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t lo = t * CS_CHUNK;
    if (lo >= N) return;
    size_t hi = min(lo + CS_CHUNK, N);
    if (shift[CS_MODE].l[0]) {   // s = w_2N: s^j is a table entry
        for (size_t j = lo; j < hi; j++) {
            Fr v = Fr::zero();
            if (j < len) {
                if (j < na) v = b_ld(&a[j]);
                if (b && j < nb) v = v - b_ld(&b[j]);
                v = v * b_ld(&tw2[j]);
            }
            b_st(&dst[j], v);
        }
        return;
    }
    const Fr s = shift[0];
    Fr pw = (lo < len) ? pow_from_table(shift + 2, lo) : Fr::zero();
    for (size_t j = lo; j < hi; j++) {
        Fr v = Fr::zero();
        if (j < len) {
            if (j < na) v = b_ld(&a[j]);
            if (b && j < nb) v = v - b_ld(&b[j]);
            v = v * pw;
            pw = pw * s;
        }
        b_st(&dst[j], v);
    }
}

// dst[j] -= r[j] * s^j for j < lr (the interpolant's share of the numerator, applied after the p-only scaling)
__global__ void coset_sub_scaled_kernel(Fr* __restrict__ dst, const Fr* __restrict__ r, size_t lr, const Fr* __restrict__ shift,
                                        const Fr* __restrict__ tw2) {
    size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t lo = t * CS_CHUNK;
    if (lo >= lr) return;
    size_t hi = min(lo + CS_CHUNK, lr);
    if (shift[CS_MODE].l[0]) {
        for (size_t j = lo; j < hi; j++) b_st(&dst[j], b_ld(&dst[j]) - b_ld(&r[j]) * b_ld(&tw2[j]));
        return;
    }
    const Fr s = shift[0];
    Fr pw = pow_from_table(shift + 2, lo);
    for (size_t j = lo; j < hi; j++) {
        b_st(&dst[j], b_ld(&dst[j]) - b_ld(&r[j]) * pw);
        pw = pw * s;
    }
}

// Z on the coset s*H_N through N/M transforms of size M = 2^log_m > m instead of one of size N: with j = a + (N/M) b,
// Z(s w_N^j) = sum_k [z_k (s w_N^a)^k] w_M^(b k), i.e. row a of dst is the M-point transform of Z's coefficients scaled
// by powers of s w_N^a (m + 1 of them; the rest of the row is zero: the caller clears dst).  log2 M instead of log2 N
// butterfly stages per element.  tw = half twiddle table of the N-point domain.
__global__ void __launch_bounds__(128) coset_fill_z_kernel(const Fr* __restrict__ Z, uint32_t m, const Fr* __restrict__ tw,
                                                           uint32_t log_cn, uint32_t log_m, const Fr* __restrict__ shift,
                                                           const Fr* __restrict__ tw2, Fr* __restrict__ dst) {
    const uint32_t chunks = (m + 1 + CS_CHUNK - 1) / CS_CHUNK;
    const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t a = idx / chunks;
    if (a >> log_cn) return;
    const uint32_t k0 = (uint32_t)(idx % chunks) * CS_CHUNK, k1 = min(k0 + CS_CHUNK, m + 1);
    if (shift[CS_MODE].l[0]) {   // s w_N^a = w_2N^(2a+1): its k-th power is the table entry (2a+1) k mod 2N (minus it above N)
        const uint32_t log_n = log_cn + log_m;
        const uint64_t N = (uint64_t)1 << log_n;
        for (uint32_t k = k0; k < k1; k++) {
            const uint64_t e = ((2 * (uint64_t)a + 1) * k) & (2 * N - 1);
            Fr w = b_ld(&tw2[e & (N - 1)]);
            Fr v = b_ld(&Z[k]) * w;
            b_st(&dst[(a << log_m) + k], e >= N ? v.neg() : v);
        }
        return;
    }
    const Fr base = shift[0] * b_ld(&tw[a]);   // a < N/M <= N/2
    Fr pw = base.pow_u64(k0);
    for (uint32_t k = k0; k < k1; k++) {
        b_st(&dst[(a << log_m) + k], b_ld(&Z[k]) * pw);
        pw = pw * base;
    }
}

#ifndef KZG_PD_BATCH
#define KZG_PD_BATCH 16
#endif
#ifndef KZG_PD_THREADS
#define KZG_PD_THREADS 128
#endif
#ifndef KZG_PD_MINB
#define KZG_PD_MINB 4   // 512 blocks at 2^20: all resident, so the one-thread inversions of all blocks overlap (profiles/r02ab_block_inverse.log)
#endif
constexpr int PD_BATCH = KZG_PD_BATCH;
constexpr int PD_THREADS = KZG_PD_THREADS;
// a[j] <- a[j] / z[j]: Montgomery's trick per thread (PD_BATCH elements) and across the block (batchinv.cuh): one
// inversion per PD_BATCH * PD_THREADS elements.  (One Kaliski inversion per 32 elements was 2/3 of this kernel's
// instructions: 227 us at 2^20.)
// z is either in natural order (log_cn = 0) or in the sub-coset layout of coset_fill_z_kernel: the value for index j sits
// at (j mod 2^log_cn) * 2^log_m + (j >> log_cn).
__global__ void __launch_bounds__(PD_THREADS, KZG_PD_MINB) pointwise_div_kernel(Fr* __restrict__ a, const Fr* __restrict__ z, size_t N,
                                                                                uint32_t log_cn, uint32_t log_m) {
    auto zi_of = [&](size_t j) { return ((j & (((size_t)1 << log_cn) - 1)) << log_m) + (j >> log_cn); };
    __shared__ Fr sh[2 * PD_THREADS + 1];
    size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t lo = t * PD_BATCH;
    int cnt = lo < N ? (int)min((size_t)PD_BATCH, N - lo) : 0;
    Fr prefix[PD_BATCH];
    Fr acc = Fr::one();
#pragma unroll
    for (int k = 0; k < PD_BATCH; k++) {
        if (k < cnt) {
            prefix[k] = acc;
            acc = acc * b_ld(&z[zi_of(lo + k)]);
        }
    }
    Fr inv = block_inverse<PD_THREADS>(acc, sh);
#pragma unroll
    for (int k = PD_BATCH - 1; k >= 0; k--) {
        if (k >= cnt) continue;
        Fr zi = inv * prefix[k];
        inv = inv * b_ld(&z[zi_of(lo + k)]);
        b_st(&a[lo + k], b_ld(&a[lo + k]) * zi);
    }
}

// psi[j] = q[j] * s^-j for j < keep; flag |= 1 if any q[j] != 0 for keep <= j < N
__global__ void coset_unscale_kernel(const Fr* __restrict__ q, size_t N, size_t keep, const Fr* __restrict__ shift,
                                     const Fr* __restrict__ tw2, Fr* __restrict__ psi, int* __restrict__ flag) {
    size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t lo = t * CS_CHUNK;
    if (lo >= N) return;
    size_t hi = min(lo + CS_CHUNK, N);
    if (shift[CS_MODE].l[0]) {   // s = w_2N: s^-j = w_2N^(2N - j) = -w_2N^(N - j) for 0 < j < N
        int nz = 0;
        for (size_t j = lo; j < hi; j++) {
            Fr v = b_ld(&q[j]);
            if (j < keep) {
                if (j) v = (v * b_ld(&tw2[N - j])).neg();
                b_st(&psi[j], v);
            } else if (!v.is_zero()) {
                nz = 1;
            }
        }
        if (nz) atomicOr(flag, 1);
        return;
    }
    const Fr sinv = shift[1];
    Fr pw = (lo < keep) ? pow_from_table(shift + 2 + CS_TAB, lo) : Fr::zero();
    int nz = 0;
    for (size_t j = lo; j < hi; j++) {
        Fr v = b_ld(&q[j]);
        if (j < keep) {
            b_st(&psi[j], v * pw);
            pw = pw * sinv;
        } else if (!v.is_zero()) {
            nz = 1;
        }
    }
    if (nz) atomicOr(flag, 1);
}

// flag |= 1 if (a - b) has a non-zero coefficient among the first len
__global__ void diff_nonzero_kernel(const Fr* __restrict__ a, size_t na, const Fr* __restrict__ b, size_t nb, size_t len,
                                    int* __restrict__ flag) {
    size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (j >= len) return;
    Fr v = Fr::zero();
    if (j < na) v = b_ld(&a[j]);
    if (j < nb) v = v - b_ld(&b[j]);
    if (!v.is_zero()) atomicOr(flag, 1);
}

// ---- host drivers --------------------------------------------------------------
static size_t al(size_t v) { return (v + 255) & ~(size_t)255; }
// Number of opening points of one batched witness.  The power-sum table of the interpolation is m x m field elements
// (32 m^2 bytes of scratch: 512 MiB at m = 4096, 8 GiB at the limit); the reference's SubProductTree has no such
// limit (coeff_form.rs:83-111), so larger batches must be split by the caller.
constexpr size_t KZG_MAX_BATCH_POINTS = 16384;

// m x m power table + five vectors of m + 1 + two of N <= 2m (Z's values and the NTT's scratch)
size_t interp_workspace_bytes(size_t m) { return al((m * m + 10 * (m + 2)) * sizeof(Fr)) + 256; }

// d_r receives max(m,2) coefficients (m == 1: the two coefficients of X + (y - x)); d_Z receives m+1.
int interpolate_run(const Fr* d_xs, const Fr* d_ys, size_t m, Fr* d_r, Fr* d_Z, int* d_flag, void* d_ws,
                    cudaStream_t stream) {
    if (m == 0 || m > KZG_MAX_BATCH_POINTS) {
        set_error("interpolate: number of points %zu out of range (1 .. %zu)", m, KZG_MAX_BATCH_POINTS);
        return KZGB_ERR_INVALID_ARG;
    }
    StageScope sc(ST_INTERP, stream);
    if (m == 1) {
        interp_single_kernel<<<1, 32, 0, stream>>>(d_xs, d_ys, d_r, d_Z);
        KZG_LAUNCHED();
        return KZGB_OK;
    }
    Fr* ws = (Fr*)d_ws;
    Fr* buf0 = ws;
    Fr* buf1 = buf0 + (m + 1);
    Fr* c = buf1 + (m + 1);
    Fr* P = c + (m + 1);
    Fr* T = P + (m + 1);
    uint32_t mm = (uint32_t)m;
    const size_t prod_smem = 2 * (m + 1) * sizeof(Fr);
    const int use_smem = prod_smem <= 96 * 1024;
    if (use_smem) {   // per-device kernel attribute: the flag lives in the context and is reset by kzgb_shutdown
        Context& cx = ctx();
        std::lock_guard<std::mutex> g(cx.cache_mu);
        if (!cx.interp_attr_set) {
            KZG_CUDA(cudaFuncSetAttribute(interp_product_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            cx.interp_attr_set = true;
        }
    }
    static const bool seq_product = getenv("KZGB_INTERP_SEQ_PRODUCT") != nullptr;   // the previous form, for A/B runs
    if (seq_product) {
        interp_product_kernel<<<1, 1024, use_smem ? prod_smem : 0, stream>>>(d_xs, mm, buf0, buf1, d_Z, use_smem);
        KZG_LAUNCHED();
    } else {
        uint32_t log_N = 1;
        while (((size_t)1 << log_N) < m + 1) log_N++;
        const size_t N = (size_t)1 << log_N;
        Fr* E = T + m * m;
        Fr* ntmp = E + N;
        const Fr* tw = nullptr;
        int st = ntt_get_twiddles(log_N, stream, &tw);
        if (st) return st;
        interp_zeval_kernel<<<div_up(N * IL, 128), 128, 0, stream>>>(d_xs, mm, tw, log_N, E);
        KZG_LAUNCHED();
        st = ntt_run(E, ntmp, log_N, true, stream);
        if (st) return st;
        KZG_CUDA(cudaMemcpyAsync(d_Z, E, (m + 1) * sizeof(Fr), cudaMemcpyDeviceToDevice, stream));
    }
    interp_weights_kernel<<<div_up(m * IL, 128), 128, 0, stream>>>(d_xs, d_ys, mm, c, d_flag);
    KZG_LAUNCHED();
    interp_powers_kernel<<<div_up(m * IL, 128), 128, 0, stream>>>(d_xs, c, mm, T);
    KZG_LAUNCHED();
    interp_powersum_kernel<<<mm, 128, 0, stream>>>(T, mm, P);
    KZG_LAUNCHED();
    interp_conv_kernel<<<div_up(m * IL, 128), 128, 0, stream>>>(d_Z, P, mm, d_r);
    KZG_LAUNCHED();
    return KZGB_OK;
}

size_t batch_div_workspace_bytes(size_t n, size_t m) {
    size_t L = n > m ? n : m;
    if (m == 1 && L < 2) L = 2;
    uint32_t log_n = 0;
    while (((size_t)1 << log_n) < L) log_n++;
    return 3 * al(((size_t)32) << log_n) + al((3 + 2 * CS_TAB) * sizeof(Fr)) + 256;
}

// psi = (p - r) div Z.  d_p: n coefficients; d_r: lr coefficients; d_Z: m+1 coefficients (monic).
// d_psi receives *psi_len coefficients.  d_flag |= 1 when Z does not divide p - r.
// r_ready != nullptr: d_r and d_Z are being produced on another stream and are complete once r_ready has fired; the
// work that needs neither (the coset shift and the scaling of p, half of this function's memory traffic) is issued first.
int batch_divide_run(const Fr* d_p, size_t n, const Fr* d_r, size_t lr, const Fr* d_Z, const Fr* d_xs, size_t m,
                     Fr* d_psi, size_t* psi_len, int* d_flag, void* d_ws, cudaStream_t stream, cudaEvent_t r_ready) {
    StageScope sc(ST_BATCH_DIV, stream);
    size_t L = n > lr ? n : lr;  // numerator length (Sub for &Polynomial, polynomial.rs:443-460)
    if (L <= m) {
        if (r_ready) KZG_CUDA(cudaStreamWaitEvent(stream, r_ready, 0));
        // deg(numerator) < deg Z: quotient is the zero polynomial; fine only if the numerator is zero
        // (polynomial.rs:194-199)
        diff_nonzero_kernel<<<div_up(L, 256), 256, 0, stream>>>(d_p, n, d_r, lr, L, d_flag);
        KZG_LAUNCHED();
        KZG_CUDA(cudaMemsetAsync(d_psi, 0, sizeof(Fr), stream));
        *psi_len = 1;
        return KZGB_OK;
    }
    uint32_t log_n = 0;
    while (((size_t)1 << log_n) < L) log_n++;
    if (log_n >= 32) return KZGB_DEGREE_TOO_LARGE;
    size_t N = (size_t)1 << log_n;
    char* w = (char*)d_ws;
    size_t nb = al(N * sizeof(Fr));
    Fr* A = (Fr*)w;
    Fr* Zb = (Fr*)(w + nb);
    Fr* tmp = (Fr*)(w + 2 * nb);
    Fr* shift = (Fr*)(w + 3 * nb);
    const Fr* tw2 = nullptr;   // half twiddle table of the 2N-point domain (root-of-unity coset)
    {
        static const FrHost h7 = fr_host_from_u64(7), h7inv = fr_host_inv(h7);
        const FrHost hgN = fr_host_pow(h7, (uint64_t)1 << log_n);   // 7^N (log_n < 32)
        Fr seven, seven_inv, gN;
        memcpy(seven.l, h7.l, 32);
        memcpy(seven_inv.l, h7inv.l, 32);
        memcpy(gN.l, hgN.l, 32);
        // first candidate: the 2N-th root of unity (needs the 2N-domain's twiddle table: log_n + 1 <= 31)
        const bool allow_root = log_n + 1 < 32 && !getenv("KZGB_BATCH_POWER_COSET");
        Fr w2n = seven, w2n_inv = seven;
        if (allow_root) {
            int st2 = ntt_get_twiddles(log_n + 1, stream, &tw2);
            if (st2) return st2;
            const FrHost hw = fr_host_domain_omega(log_n + 1), hwi = fr_host_inv(hw);
            memcpy(w2n.l, hw.l, 32);
            memcpy(w2n_inv.l, hwi.l, 32);
        }
        coset_shift_kernel<<<1, 256, 0, stream>>>(d_xs, (uint32_t)m, log_n, seven, seven_inv, gN, w2n, w2n_inv, allow_root ? 1u : 0u,
                                                  shift);
    }
    KZG_LAUNCHED();
    unsigned gsz = div_up((N + CS_CHUNK - 1) / CS_CHUNK, 128);
    if (r_ready) {
        coset_scale_kernel<<<gsz, 128, 0, stream>>>(d_p, n, nullptr, 0, L, N, shift, tw2, A);      // p * s^j alone
        KZG_LAUNCHED();
        KZG_CUDA(cudaStreamWaitEvent(stream, r_ready, 0));
        coset_sub_scaled_kernel<<<div_up((lr + CS_CHUNK - 1) / CS_CHUNK, 128), 128, 0, stream>>>(A, d_r, lr, shift, tw2);
        KZG_LAUNCHED();
    } else {
        coset_scale_kernel<<<gsz, 128, 0, stream>>>(d_p, n, d_r, lr, L, N, shift, tw2, A);
        KZG_LAUNCHED();
    }
    uint32_t log_m = 1;
    while (((size_t)1 << log_m) < m + 1) log_m++;
    const bool small_z = log_m + 1 <= log_n && (log_n - log_m) <= 15 && !getenv("KZGB_BATCH_FULL_Z_NTT");
    const uint32_t log_cn = small_z ? log_n - log_m : 0;
    int st;
    if (small_z) {   // N/M transforms of size M (see coset_fill_z_kernel)
        const Fr* twN = nullptr;
        st = ntt_get_twiddles(log_n, stream, &twN);
        if (st) return st;
        KZG_CUDA(cudaMemsetAsync(Zb, 0, N * sizeof(Fr), stream));
        const size_t work = ((size_t)1 << log_cn) * ((m + 1 + CS_CHUNK - 1) / CS_CHUNK);
        coset_fill_z_kernel<<<div_up(work, 128), 128, 0, stream>>>(d_Z, (uint32_t)m, twN, log_cn, log_m, shift, tw2, Zb);
        KZG_LAUNCHED();
    } else {
        coset_scale_kernel<<<gsz, 128, 0, stream>>>(d_Z, m + 1, nullptr, 0, m + 1, N, shift, tw2, Zb);
        KZG_LAUNCHED();
    }
    st = ntt_run(A, tmp, log_n, false, stream);
    if (st) return st;
    if (small_z) st = ntt_run(Zb, tmp, log_m, false, stream, 1u << log_cn);
    else st = ntt_run(Zb, tmp, log_n, false, stream, 1, nullptr, (uint32_t)(m + 1));   // Z has m + 1 coefficients
    if (st) return st;
    pointwise_div_kernel<<<div_up((N + PD_BATCH - 1) / PD_BATCH, PD_THREADS), PD_THREADS, 0, stream>>>(A, Zb, N, log_cn,
                                                                                                   small_z ? log_m : 0);
    KZG_LAUNCHED();
    st = ntt_run(A, tmp, log_n, true, stream);
    if (st) return st;
    size_t keep = L - m;
    coset_unscale_kernel<<<gsz, 128, 0, stream>>>(A, N, keep, shift, tw2, d_psi, d_flag);
    KZG_LAUNCHED();
    *psi_len = keep;
    return KZGB_OK;
}

}  // namespace kzg

using namespace kzg;

extern "C" {

// Lagrange interpolation through (xs[i], ys[i]) on host vectors.  r_out: max(m,2) entries; *r_len = m, or 2 when
// m == 1 (X + (y - x), polynomial.rs:244-247).  Distinct xs required (else KZGB_ERR_INVALID_ARG; the reference panics).
KZGB_API int kzgb_interpolate(const kzgb_fr* xs, const kzgb_fr* ys, size_t m, kzgb_fr* r_out, size_t* r_len) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    Context& c = ctx();
    (void)c;
    use_stream(lane().stream);
    if (!xs || !ys || !r_out || !r_len || m == 0 || m > KZG_MAX_BATCH_POINTS) {
        set_error("kzgb_interpolate: bad argument (1 <= m <= 16384; the workspace is 32 m^2 bytes)");
        return KZGB_ERR_INVALID_ARG;
    }
    size_t mb = al(m * 32), rb = al((m + 2) * 32);
    char* d = (char*)staging_reserve(2 * mb + 2 * rb + 256 + interp_workspace_bytes(m));
    if (!d) return KZGB_ERR_CUDA;
    Fr* d_xs = (Fr*)d;
    Fr* d_ys = (Fr*)(d + mb);
    Fr* d_r = (Fr*)(d + 2 * mb);
    Fr* d_Z = (Fr*)(d + 2 * mb + rb);
    int* d_flag = (int*)(d + 2 * mb + 2 * rb);
    char* ws = d + 2 * mb + 2 * rb + 256;
    KZG_CUDA(cudaMemcpyAsync(d_xs, xs, m * 32, cudaMemcpyHostToDevice, lane().stream));
    KZG_CUDA(cudaMemcpyAsync(d_ys, ys, m * 32, cudaMemcpyHostToDevice, lane().stream));
    KZG_CUDA(cudaMemsetAsync(d_flag, 0, 4, lane().stream));
    int st = interpolate_run(d_xs, d_ys, m, d_r, d_Z, d_flag, ws, lane().stream);
    if (st) return st;
    size_t len = m == 1 ? 2 : m;
    int flag = 0;
    KZG_CUDA(cudaMemcpyAsync(r_out, d_r, len * 32, cudaMemcpyDeviceToHost, lane().stream));
    KZG_CUDA(cudaMemcpyAsync(&flag, d_flag, 4, cudaMemcpyDeviceToHost, lane().stream));
    KZG_CUDA(cudaStreamSynchronize(lane().stream));
    if (flag & 2) { set_error("kzgb_interpolate: xs are not distinct"); return KZGB_ERR_INVALID_ARG; }
    *r_len = len;
    return KZGB_OK;
}

// create_witness_batched (coeff_form.rs:83-111) on device-resident inputs.
// d_p: n coefficients; d_xs, d_ys: m each; d_r_out: max(m,2) Fr; d_out: 96 B affine.  Synchronises the stream.
KZGB_API int kzgb_witness_batched_dev(kzgb_srs* srs, const void* d_p, size_t n, const void* d_xs, const void* d_ys,
                                      size_t m, void* d_r_out, void* d_out, void* stream) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    Context& c = ctx();
    (void)c;
    if (!srs || !d_p || !d_xs || !d_ys || !d_r_out || !d_out || n == 0 || m == 0 || m > KZG_MAX_BATCH_POINTS) {
        set_error("kzgb_witness_batched_dev: bad argument (1 <= m <= 16384; the workspace is 32 m^2 bytes)");
        return KZGB_ERR_INVALID_ARG;
    }
    KZG_LOCK_SRS(srs);
    cudaStream_t st = use_stream((cudaStream_t)stream);
    size_t lr = m == 1 ? 2 : m;
    size_t L = n > lr ? n : lr;
    size_t zb = al((m + 2) * 32), psib = al((L + 1) * 32);
    // The interpolation (a handful of small, latency-bound kernels) runs on the lane's side stream BESIDE the coset
    // shift and the scaling of p, which need neither r nor Z: two workspaces instead of one shared one.
    const size_t iws = al(interp_workspace_bytes(m)), dws = batch_div_workspace_bytes(n, m);
    char* d = (char*)scratch_reserve(zb + psib + 256 + iws + dws);
    if (!d) return KZGB_ERR_CUDA;
    Fr* d_Z = (Fr*)d;
    Fr* d_psi = (Fr*)(d + zb);
    int* d_flag = (int*)(d + zb + psib);
    char* ws_i = d + zb + psib + 256;
    char* ws_d = ws_i + iws;
    cudaStream_t side;
    cudaEvent_t fork_ev, join_ev;
    int rc = lane_side_stream(&side, &fork_ev, &join_ev);
    if (rc) return rc;
    KZG_CUDA(cudaMemsetAsync(d_flag, 0, 4, st));
    KZG_CUDA(cudaEventRecord(fork_ev, st));
    KZG_CUDA(cudaStreamWaitEvent(side, fork_ev, 0));
    rc = interpolate_run((const Fr*)d_xs, (const Fr*)d_ys, m, (Fr*)d_r_out, d_Z, d_flag, ws_i, side);
    if (rc) { cudaStreamSynchronize(side); return rc; }
    KZG_CUDA(cudaEventRecord(join_ev, side));
    size_t psi_len = 0;
    rc = batch_divide_run((const Fr*)d_p, n, (const Fr*)d_r_out, lr, d_Z, (const Fr*)d_xs, m, d_psi, &psi_len, d_flag, ws_d,
                          st, join_ev);
    if (rc) { cudaStreamSynchronize(side); return rc; }
    if (psi_len > srs->n) {
        set_error("kzgb_witness_batched: quotient has %zu coefficients but the SRS holds %zu points", psi_len, srs->n);
        return KZGB_ERR_INVALID_ARG;
    }
    const XYZZ* res = nullptr;
    rc = msm_run(srs, d_psi, psi_len, st, &res);
    if (rc) return rc;
    rc = msm_finish(res, d_out, 0, st);
    if (rc) return rc;
    int flag = 0;
    KZG_CUDA(cudaMemcpyAsync(&flag, d_flag, 4, cudaMemcpyDeviceToHost, st));
    KZG_CUDA(cudaStreamSynchronize(st));
    if (flag & 2) { set_error("kzgb_witness_batched: xs are not distinct"); return KZGB_ERR_INVALID_ARG; }
    return (flag & 1) ? KZGB_POINT_NOT_ON_POLYNOMIAL : KZGB_OK;
}

KZGB_API int kzgb_witness_batched(kzgb_srs* srs, const kzgb_fr* p, size_t n, const kzgb_fr* xs, const kzgb_fr* ys, size_t m,
                                  kzgb_fr* r_out, size_t* r_len, kzgb_p1_affine* w_out) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    Context& c = ctx();
    (void)c;
    use_stream(lane().stream);
    if (!srs || !p || !xs || !ys || !r_out || !r_len || !w_out || n == 0 || m == 0 || m > KZG_MAX_BATCH_POINTS) {
        set_error("kzgb_witness_batched: bad argument (1 <= m <= 16384; the workspace is 32 m^2 bytes)");
        return KZGB_ERR_INVALID_ARG;
    }
    KZG_LOCK_SRS(srs);
    size_t pb = al(n * 32), mb = al(m * 32), rb = al((m + 2) * 32);
    char* d = (char*)staging_reserve(pb + 2 * mb + rb + 256);
    if (!d) return KZGB_ERR_CUDA;
    Fr* d_p = (Fr*)d;
    Fr* d_xs = (Fr*)(d + pb);
    Fr* d_ys = (Fr*)(d + pb + mb);
    Fr* d_r = (Fr*)(d + pb + 2 * mb);
    char* d_out = d + pb + 2 * mb + rb;
    KZG_CUDA(upload_async(d_p, p, n * 32, lane().stream));
    KZG_CUDA(cudaMemcpyAsync(d_xs, xs, m * 32, cudaMemcpyHostToDevice, lane().stream));
    KZG_CUDA(cudaMemcpyAsync(d_ys, ys, m * 32, cudaMemcpyHostToDevice, lane().stream));
    int rc = kzgb_witness_batched_dev(srs, d_p, n, d_xs, d_ys, m, d_r, d_out, lane().stream);
    if (rc < 0) return rc;
    size_t len = m == 1 ? 2 : m;
    KZG_CUDA(cudaMemcpy(r_out, d_r, len * 32, cudaMemcpyDeviceToHost));
    KZG_CUDA(cudaMemcpy(w_out, d_out, sizeof(kzgb_p1_affine), cudaMemcpyDeviceToHost));
    *r_len = len;
    return rc;
}

}  // extern "C"
