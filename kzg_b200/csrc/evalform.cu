// evalform.cu -- evaluation-form quotient: replaces `div_by_omega_i`
// (eval_form.rs:58-84) as called by `KZGProverEvalForm::create_witness`
// (eval_form.rs:124-140).
//
// The reference computes, for the numerator f_j - f_m (eval_form.rs:125-129):
//     q_j = (f_j - f_m) / (w^j - w^m)                       j != m   (:78-79)
//     q_m = sum_{i != m} (f_i - f_m) * (a_m/a_i) / (w^m - w^i)        (:63-76)
// with a_i = d * w^-i, so a_m/a_i = w^(i-m) and q_m = - sum_{i != m} q_i * w^(i-m).
// It spends one modular inversion and one pow_vartime per element; here w^j comes
// from the domain's twiddle table, inversions are batched per thread (Montgomery's
// trick) and q_m is a block-reduced sum.
#include "common.cuh"
#include "field.cuh"
#include "batchinv.cuh"
#include "evalform.cuh"
#include "ntt.cuh"

namespace kzg {

#ifndef KZG_DVO_BATCH
#define KZG_DVO_BATCH 16
#endif
#ifndef KZG_DVO_MINB
#define KZG_DVO_MINB 4   // all 512 blocks of a 2^20 domain resident: the per-block inversions overlap (profiles/r02ab_block_inverse.log)
#endif
constexpr int DVO_BATCH = KZG_DVO_BATCH;
constexpr int DVO_THREADS = 128;

KZG_DI Fr dvo_ld(const Fr* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
KZG_DI void dvo_st(Fr* p, const Fr& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
// w^e from the half table (e < d)
KZG_DI Fr omega_pow(const Fr* __restrict__ tw, uint32_t e, uint32_t log_d) {
    const uint32_t half = 1u << (log_d - 1);
    Fr v = dvo_ld(&tw[e & (half - 1)]);
    return (e & half) ? v.neg() : v;
}

__global__ void __launch_bounds__(DVO_THREADS, KZG_DVO_MINB) div_by_omega_kernel(const Fr* __restrict__ f, uint32_t log_d, uint32_t m,
                                                                   const Fr* __restrict__ tw, Fr* __restrict__ q,
                                                                   Fr* __restrict__ block_sums) {
    __shared__ Fr sh[2 * DVO_THREADS + 1];
    const uint32_t d = 1u << log_d;
    const uint32_t t = blockIdx.x * DVO_THREADS + threadIdx.x;
    const uint32_t lo = t * DVO_BATCH;
    Fr part = Fr::zero();
    {
        const Fr fm = dvo_ld(&f[m]);
        const Fr wm = omega_pow(tw, m, log_d);
        Fr den[DVO_BATCH], prefix[DVO_BATCH];
        Fr acc = Fr::one();
#pragma unroll
        for (int k = 0; k < DVO_BATCH; k++) {
            uint32_t j = lo + k;
            if (j < d) {
                den[k] = omega_pow(tw, j, log_d) - wm;
                prefix[k] = acc;
                if (j != m) acc = acc * den[k];
            }
        }
        // one inversion per block (batchinv.cuh); threads past the end contribute 1
        Fr inv = block_inverse<DVO_THREADS>(acc, sh);
        __syncthreads();   // sh is reused for the block sum below
#pragma unroll
        for (int k = DVO_BATCH - 1; k >= 0; k--) {
            uint32_t j = lo + k;
            if (j < d && j != m) {
                Fr dinv = inv * prefix[k];
                inv = inv * den[k];
                Fr qj = (dvo_ld(&f[j]) - fm) * dinv;
                dvo_st(&q[j], qj);
                part = part + qj * omega_pow(tw, (j - m) & (d - 1), log_d);
            }
        }
    }
    sh[threadIdx.x] = part;
    __syncthreads();
    for (int s = DVO_THREADS / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sh[threadIdx.x] = sh[threadIdx.x] + sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) block_sums[blockIdx.x] = sh[0];
}

__global__ void __launch_bounds__(DVO_THREADS) div_by_omega_finish_kernel(const Fr* __restrict__ block_sums, uint32_t nblocks,
                                                                          uint32_t m, Fr* __restrict__ q) {
    __shared__ Fr sh[DVO_THREADS];
    Fr part = Fr::zero();
    for (uint32_t i = threadIdx.x; i < nblocks; i += DVO_THREADS) part = part + block_sums[i];
    sh[threadIdx.x] = part;
    __syncthreads();
    for (int s = DVO_THREADS / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sh[threadIdx.x] = sh[threadIdx.x] + sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) q[m] = sh[0].neg();
}

size_t div_by_omega_workspace_bytes(uint32_t log_d) {
    size_t d = (size_t)1 << log_d;
    size_t nblocks = (d + (size_t)DVO_BATCH * DVO_THREADS - 1) / ((size_t)DVO_BATCH * DVO_THREADS);
    return (nblocks + 1) * sizeof(Fr);
}

// q = div_by_omega_i(evals - evals[m], m); log_d >= 1
int div_by_omega_run(const Fr* d_f, uint32_t log_d, uint32_t m, Fr* d_q, void* d_ws, cudaStream_t stream) {
    const Fr* tw = nullptr;
    int st = ntt_get_twiddles(log_d, stream, &tw);
    if (st) return st;
    StageScope sc(ST_DIV_OMEGA, stream);
    size_t d = (size_t)1 << log_d;
    uint32_t nblocks = (uint32_t)((d + (size_t)DVO_BATCH * DVO_THREADS - 1) / ((size_t)DVO_BATCH * DVO_THREADS));
    div_by_omega_kernel<<<nblocks, DVO_THREADS, 0, stream>>>(d_f, log_d, m, tw, d_q, (Fr*)d_ws);
    KZG_LAUNCHED();
    div_by_omega_finish_kernel<<<1, DVO_THREADS, 0, stream>>>((const Fr*)d_ws, nblocks, m, d_q);
    KZG_LAUNCHED();
    return KZGB_OK;
}

}  // namespace kzg
