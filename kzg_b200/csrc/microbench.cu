// microbench.cu -- on-box peak measurements used as roofline denominators:
// INT32 multiply-add issue rates (IMAD / IMAD.HI / IMAD.WIDE), dependent Fq/Fr
// Montgomery-multiplication chains, and a device-to-device copy (HBM).
// "Peak figures must come from ... an on-box microbenchmark
// (dense independent IMAD.WIDE chains ...)".
#include "common.cuh"
#include "field.cuh"

namespace kzg {

constexpr int MB_CHAINS = 8;
constexpr int MB_UNROLL = 32;

template <int MODE>
__global__ void __launch_bounds__(256) mb_kernel(uint32_t* out, long long* cycles, int iters, uint32_t y, uint32_t z) {
    uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    long long t0 = clock64();
    if constexpr (MODE <= 3 || MODE == 6 || MODE == 7) {
        uint32_t x[MB_CHAINS];
        uint64_t w[MB_CHAINS];
#pragma unroll
        for (int k = 0; k < MB_CHAINS; k++) {
            x[k] = tid * 2654435761u + k;
            w[k] = ((uint64_t)x[k] << 32) | (tid + k);
        }
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int u = 0; u < MB_UNROLL; u++) {
#pragma unroll
                for (int k = 0; k < MB_CHAINS; k++) {
                    if constexpr (MODE == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(y), "r"(z));
                    if constexpr (MODE == 1) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(y), "r"(z));
                    if constexpr (MODE == 2) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[k]) : "r"((uint32_t)w[(k + 1) % MB_CHAINS]), "r"(y));
                    if constexpr (MODE == 3) {  // lo.cc + hi pair on one accumulator pair
                        uint32_t lo = (uint32_t)w[k], hi = (uint32_t)(w[k] >> 32);
                        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;"
                                     : "+r"(lo), "+r"(hi) : "r"(x[k]), "r"(y));
                        w[k] = ((uint64_t)hi << 32) | lo;
                    }
                    if constexpr (MODE == 6) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[k]) : "r"(x[(k + 1) % MB_CHAINS]));
                    if constexpr (MODE == 7) {
                        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[k]) : "r"((uint32_t)w[(k + 1) % MB_CHAINS]), "r"(y));
                        asm volatile("add.u32 %0, %0, %1;" : "+r"(x[k]) : "r"(x[(k + 1) % MB_CHAINS]));
                    }
                }
            }
        }
        uint32_t acc = 0;
#pragma unroll
        for (int k = 0; k < MB_CHAINS; k++) acc ^= x[k] ^ (uint32_t)w[k] ^ (uint32_t)(w[k] >> 32);
        if (acc == 0x12345678u) out[tid] = acc;
    } else if constexpr (MODE == 4) {
        Fq a, b;
#pragma unroll
        for (int i = 0; i < 12; i++) { a.l[i] = tid + i * y; b.l[i] = tid * 3 + i + z; }
        a.l[11] &= 0x0fffffff; b.l[11] &= 0x0fffffff;
        for (int it = 0; it < iters; it++) {
#pragma unroll 4
            for (int u = 0; u < MB_UNROLL; u++) { a = a * b; b = b * a; }
        }
        if (a.l[0] == 0x12345678u) out[tid] = b.l[1];
    } else if constexpr (MODE == 8 || MODE == 9) {
        // DFMA issue rate (8) and DFMA + IMAD.WIDE 1:1 (9: do the two pipes run concurrently?)
        double d[MB_CHAINS];
        uint64_t w[MB_CHAINS];
        const double m = __hiloint2double(0x3ff00000, y & 0xff), c = __hiloint2double(0x3e000000, z);
#pragma unroll
        for (int k = 0; k < MB_CHAINS; k++) { d[k] = (double)(tid + k); w[k] = ((uint64_t)tid << 32) | (tid + k); }
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int u = 0; u < MB_UNROLL; u++) {
#pragma unroll
                for (int k = 0; k < MB_CHAINS; k++) {
                    asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[k]) : "d"(m), "d"(c));
                    if constexpr (MODE == 9) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[k]) : "r"((uint32_t)w[(k + 1) % MB_CHAINS]), "r"(y));
                }
            }
        }
        double acc = 0;
        uint64_t wa = 0;
#pragma unroll
        for (int k = 0; k < MB_CHAINS; k++) { acc += d[k]; wa ^= w[k]; }
        if (acc == 0.12345 || wa == 0x12345678u) out[tid] = 1;
    } else if constexpr (MODE == 10) {
        Fq a, b;
#pragma unroll
        for (int i = 0; i < 12; i++) { a.l[i] = tid + i * y; b.l[i] = tid * 3 + i + z; }
        a.l[11] &= 0x0fffffff; b.l[11] &= 0x0fffffff;
        for (int it = 0; it < iters; it++) {
#pragma unroll 2
            for (int u = 0; u < MB_UNROLL; u++) { a = Fq::mul_dfma(a, b); b = Fq::mul_dfma(b, a); }
        }
        if (a.l[0] == 0x12345678u) out[tid] = b.l[1];
    } else if constexpr (MODE == 11) {
        Fr a, b;
#pragma unroll
        for (int i = 0; i < 8; i++) { a.l[i] = tid + i * y; b.l[i] = tid * 3 + i + z; }
        a.l[7] &= 0x3fffffff; b.l[7] &= 0x3fffffff;
        for (int it = 0; it < iters; it++) {
#pragma unroll 2
            for (int u = 0; u < MB_UNROLL; u++) { a = Fr::mul_dfma(a, b); b = Fr::mul_dfma(b, a); }
        }
        if (a.l[0] == 0x12345678u) out[tid] = b.l[1];
    } else if constexpr (MODE == 12) {
        // two independent chains per thread, one on each pipe
        Fq a, b, c, d;
#pragma unroll
        for (int i = 0; i < 12; i++) { a.l[i] = tid + i * y; b.l[i] = tid * 3 + i + z; c.l[i] = tid * 5 + i; d.l[i] = tid * 7 + i * z; }
        a.l[11] &= 0x0fffffff; b.l[11] &= 0x0fffffff; c.l[11] &= 0x0fffffff; d.l[11] &= 0x0fffffff;
        for (int it = 0; it < iters; it++) {
#pragma unroll 2
            for (int u = 0; u < MB_UNROLL / 2; u++) {
                a = Fq::mul_inline(a, b); c = Fq::mul_dfma(c, d);
                b = Fq::mul_inline(b, a); d = Fq::mul_dfma(d, c);
            }
        }
        if (a.l[0] == 0x12345678u) out[tid] = b.l[1] ^ c.l[0] ^ d.l[2];
    } else if constexpr (MODE == 13 || MODE == 15) {
        Fq a, b;
#pragma unroll
        for (int i = 0; i < 12; i++) { a.l[i] = tid + i * y; b.l[i] = tid * 3 + i + z; }
        a.l[11] &= 0x0fffffff; b.l[11] &= 0x0fffffff;
        for (int it = 0; it < iters; it++) {
#pragma unroll 4
            for (int u = 0; u < MB_UNROLL; u++) {
                if constexpr (MODE == 13) { a = Fq::mul_kara(a, b); b = Fq::mul_kara(b, a); }
                else { a = Fq::mul_sub_mul(a, b, b, b); b = Fq::mul_sub_mul(b, a, a, a); }   // counts as 2 products each
            }
        }
        if (a.l[0] == 0x12345678u) out[tid] = b.l[1];
    } else if constexpr (MODE == 14) {
        Fr a, b;
#pragma unroll
        for (int i = 0; i < 8; i++) { a.l[i] = tid + i * y; b.l[i] = tid * 3 + i + z; }
        a.l[7] &= 0x3fffffff; b.l[7] &= 0x3fffffff;
        for (int it = 0; it < iters; it++) {
#pragma unroll 4
            for (int u = 0; u < MB_UNROLL; u++) { a = Fr::mul_kara(a, b); b = Fr::mul_kara(b, a); }
        }
        if (a.l[0] == 0x12345678u) out[tid] = b.l[1];
    } else if constexpr (MODE == 16) {
        Fq a, b;
#pragma unroll
        for (int i = 0; i < 12; i++) { a.l[i] = tid + i * y; b.l[i] = tid * 3 + i + z; }
        a.l[11] &= 0x0fffffff; b.l[11] &= 0x0fffffff;
        for (int it = 0; it < iters; it++) {
#pragma unroll 4
            for (int u = 0; u < MB_UNROLL; u++) { a = Fq::sqr_inline(a) + b; b = Fq::sqr_inline(b) + a; }
        }
        if (a.l[0] == 0x12345678u) out[tid] = b.l[1];
    } else if constexpr (MODE == 5) {
        Fr a, b;
#pragma unroll
        for (int i = 0; i < 8; i++) { a.l[i] = tid + i * y; b.l[i] = tid * 3 + i + z; }
        a.l[7] &= 0x3fffffff; b.l[7] &= 0x3fffffff;
        for (int it = 0; it < iters; it++) {
#pragma unroll 4
            for (int u = 0; u < MB_UNROLL; u++) { a = a * b; b = b * a; }
        }
        if (a.l[0] == 0x12345678u) out[tid] = b.l[1];
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

__global__ void copy_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = src[i];
}

}  // namespace kzg

using namespace kzg;

extern "C" {

// mode: 0 IMAD.lo, 1 IMAD.HI, 2 IMAD.WIDE, 3 mad.lo.cc+madc.hi pair, 4 Fq mul chain, 5 Fr mul chain,
//       6 IADD, 7 IMAD.WIDE+IADD 1:1, 8 DFMA, 9 DFMA+IMAD.WIDE 1:1 (counts DFMA), 10/11 Fq/Fr mul_dfma chain,
//       12 Fq: one IMAD chain + one DFMA chain per thread, 13/14 Fq/Fr Karatsuba mul chain, 15 Fq lazy a*b - c*d chain
//       (counted as two products each), 16 Fq dedicated-square chain, 100 = HBM copy (bytes = iters MiB each way)
// out[0] = ops/s (mode 4/5: field muls/s; mode 100: bytes moved/s), out[1] = ops per SM-clock per SM, out[2] = ms
int kzgb_microbench(int mode, int blocks_per_sm, int threads, int iters, double* out) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    Context& c = ctx();
    (void)c;
    use_stream(lane().stream);
    cudaEvent_t e0, e1;
    KZG_CUDA(cudaEventCreate(&e0));
    KZG_CUDA(cudaEventCreate(&e1));
    float ms = 0;
    if (mode == 100) {
        size_t bytes = (size_t)iters << 20;
        void *a, *b;
        KZG_CUDA(cudaMalloc(&a, bytes));
        KZG_CUDA(cudaMalloc(&b, bytes));
        KZG_CUDA(cudaMemsetAsync(a, 1, bytes, lane().stream));
        float best = 1e30f;
        for (int rep = 0; rep < 6; rep++) {
            KZG_CUDA(cudaEventRecord(e0, lane().stream));
            copy_kernel<<<c.sm_count * 8, 512, 0, lane().stream>>>((const uint4*)a, (uint4*)b, bytes / 16);
            KZG_LAUNCHED();
            KZG_CUDA(cudaEventRecord(e1, lane().stream));
            KZG_CUDA(cudaEventSynchronize(e1));
            KZG_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < best) best = ms;
        }
        cudaFree(a);
        cudaFree(b);
        out[0] = 2.0 * bytes / (best * 1e-3);
        out[1] = 0;
        out[2] = best;
        return KZGB_OK;
    }
    int blocks = blocks_per_sm * c.sm_count;
    uint32_t* d_out;
    long long* d_cyc;
    KZG_CUDA(cudaMalloc(&d_out, (size_t)blocks * threads * 4));
    KZG_CUDA(cudaMalloc(&d_cyc, (size_t)blocks * 8));
    for (int rep = 0; rep < 3; rep++) {
        KZG_CUDA(cudaEventRecord(e0, lane().stream));
        switch (mode) {
#define MB_CASE(M) case M: mb_kernel<M><<<blocks, threads, 0, lane().stream>>>(d_out, d_cyc, iters, 0x9e3779b9u, 0x7f4a7c15u); break;
            MB_CASE(0) MB_CASE(1) MB_CASE(2) MB_CASE(3) MB_CASE(4) MB_CASE(5) MB_CASE(6) MB_CASE(7) MB_CASE(8) MB_CASE(9) MB_CASE(10) MB_CASE(11) MB_CASE(12)
            MB_CASE(13) MB_CASE(14) MB_CASE(15) MB_CASE(16)
#undef MB_CASE
            default: set_error("kzgb_microbench: bad mode %d", mode); return KZGB_ERR_INVALID_ARG;
        }
        KZG_LAUNCHED();
        KZG_CUDA(cudaEventRecord(e1, lane().stream));
        KZG_CUDA(cudaEventSynchronize(e1));
        KZG_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    }
    long long cyc0 = 0;
    KZG_CUDA(cudaMemcpy(&cyc0, d_cyc, 8, cudaMemcpyDeviceToHost));
    cudaFree(d_out);
    cudaFree(d_cyc);
    double per_thread;
    if (mode == 4 || mode == 5 || mode >= 10) per_thread = 2.0 * MB_UNROLL * iters;
    else per_thread = (double)MB_CHAINS * MB_UNROLL * iters;   // mode 7 counts IMAD.WIDE only
    double total = per_thread * threads * (double)blocks;
    out[0] = total / (ms * 1e-3);
    out[1] = per_thread * threads * blocks_per_sm / (double)cyc0;
    out[2] = ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return KZGB_OK;
}

}  // extern "C"
