// field.cuh -- BLS12-381 Fq (381-bit, 12 x u32) and Fr (255-bit, 8 x u32)
// Montgomery arithmetic for sm_100a.
//
// Replaces the arithmetic the reference takes from blstrs/blst (Cargo.toml:27;
// `Scalar` ops everywhere in src/, Fp ops inside `G1Projective::multi_exp`).
// Because 2^256 = (2^32)^8 and 2^384 = (2^32)^12, the 32-bit-limb Montgomery
// image of an element is bit-identical to blst's 64-bit-limb image, so blst_fr /
// blst_fp buffers are consumed with no conversion.
//
// Multiplication is an operand-scanning Montgomery product over two
// interleaved accumulators ("even"/"odd" limb alignment) so that every partial
// product a_j*b_i lands on a (lo,hi) register pair and each accumulator is one
// unbroken carry chain: mad.lo.cc / madc.hi.cc pairs that ptxas fuses into
// IMAD.WIDE.U32 (+carry) on the fma pipe.  4*N^2 32-bit mads per product
// (= 2*N^2 wide mads): 288 for Fq, 128 for Fr.
#pragma once
#include <cstdint>
#if defined(KZG_HOST_EMU) && !defined(__CUDACC__)
// plain g++ build of the host model (tests/cpp/field_host_model.cpp): CUDA keywords vanish, device intrinsics that
// the checked routines never reach are declared (not defined) so that the template bodies still parse
#define __host__
#define __device__
#define __constant__
#define __forceinline__ inline
#define __noinline__
double __hiloint2double(int, int);
int __double2loint(double);
int __double2hiint(double);
double __fma_rn(double, double, double);
unsigned __funnelshift_l(unsigned, unsigned, unsigned);
unsigned __funnelshift_r(unsigned, unsigned, unsigned);
#else
#include <cuda_runtime.h>
#endif

namespace kzg {

// --------------------------------------------------------------------------
// PTX carry-chain primitives
// --------------------------------------------------------------------------
#if defined(KZG_HOST_EMU) && !defined(__CUDA_ARCH__)
// Host emulation of the carry-chain primitives (tests/cpp/field_host_model.cu): the SAME limb-level routines are
// compiled for the host with the PTX condition-code register modelled by one thread-local bit, so that new
// multiplication schedules are checked bit-exact against a big-integer model without a GPU.  Never part of the library.
#define KZG_DI inline
namespace emu { static thread_local uint32_t cf = 0; }
inline uint32_t add_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b; emu::cf = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b + emu::cf; emu::cf = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a + b + emu::cf); }
inline uint32_t sub_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b; emu::cf = (uint32_t)(t >> 63); return (uint32_t)t; }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b - emu::cf; emu::cf = (uint32_t)(t >> 63); return (uint32_t)t; }
inline uint32_t subc(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a - b - emu::cf); }
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a * b); }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)mul_lo(a, b) + c; emu::cf = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)mul_lo(a, b) + c + emu::cf; emu::cf = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)mul_hi(a, b) + c + emu::cf; emu::cf = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return (uint32_t)((uint64_t)mul_hi(a, b) + c + emu::cf); }
#else
#define KZG_DI __device__ __forceinline__

KZG_DI uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
KZG_DI uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
KZG_DI uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
KZG_DI uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
KZG_DI uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
KZG_DI uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
KZG_DI uint32_t mul_lo(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
KZG_DI uint32_t mul_hi(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
KZG_DI uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
KZG_DI uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
KZG_DI uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
KZG_DI uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
#endif

KZG_DI uint32_t funnel_l(uint32_t lo, uint32_t hi, uint32_t sh) {   // (hi:lo << sh) >> 32, 0 < sh < 32
#if defined(KZG_HOST_EMU) && !defined(__CUDA_ARCH__)
    return (hi << sh) | (lo >> (32 - sh));
#else
    return __funnelshift_l(lo, hi, sh);
#endif
}

// --------------------------------------------------------------------------
// Field parameter packs (all values verified against the big-int oracle)
// --------------------------------------------------------------------------
struct FqParams {
    static constexpr int N = 12;
    static constexpr bool LOW_LIMBS_1_FFFFFFFF = false;
    // p = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
    __host__ __device__ static constexpr uint32_t mod(int i) {
        constexpr uint32_t t[12] = {0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u,
                                         0xf38512bfu, 0x64774b84u, 0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau};
        return t[i];
    }
    static constexpr uint32_t INV = 0xfffcfffdu;  // -p^-1 mod 2^32
    __host__ __device__ static constexpr uint32_t inv32() { return INV; }
    // R = 2^384 mod p (Montgomery one)
    __host__ __device__ static constexpr uint32_t one(int i) {
        constexpr uint32_t t[12] = {0x0002fffdu, 0x76090000u, 0xc40c0002u, 0xebf4000bu, 0x53c758bau, 0x5f489857u,
                                         0x70525745u, 0x77ce5853u, 0xa256ec6du, 0x5c071a97u, 0xfa80e493u, 0x15f65ec3u};
        return t[i];
    }
    // R^2 mod p
    __host__ __device__ static constexpr uint32_t r2(int i) {
        constexpr uint32_t t[12] = {0x1c341746u, 0xf4df1f34u, 0x09d104f1u, 0x0a76e6a6u, 0x4c95b6d5u, 0x8de5476cu,
                                        0x939d83c0u, 0x67eb88a9u, 0xb519952du, 0x9a793e85u, 0x92cae3aau, 0x11988fe5u};
        return t[i];
    }
};

// -r^-1 mod 2^32 happens to be 0xffffffff.  As an immediate, ptxas strength-reduces m = t0 * 0xffffffff to a
// negation and then no longer fuses the m*MOD products into IMAD.WIDE.U32.X (they become IMAD.HI + IMAD.X pairs:
// +40 % FMA-heavy pipe cycles per Fr product).  Reading the constant from constant memory keeps it opaque.
static __constant__ uint32_t FR_INV32_C[1] = {0xffffffffu};

struct FrParams {
    static constexpr int N = 8;
#ifdef KZG_FR_SPECIAL_REDUCE   // experiment, measured slower (profiles/experiments/README.md): off by default
    static constexpr bool LOW_LIMBS_1_FFFFFFFF = true;   // r = ...ffffffff00000001: see reduce_step
#else
    static constexpr bool LOW_LIMBS_1_FFFFFFFF = false;
#endif
    // r = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
    __host__ __device__ static constexpr uint32_t mod(int i) {
        constexpr uint32_t t[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u,
                                        0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
        return t[i];
    }
    static constexpr uint32_t INV = 0xffffffffu;  // -r^-1 mod 2^32
    KZG_DI static uint32_t inv32() {
#if defined(KZG_HOST_EMU) && !defined(__CUDA_ARCH__)
        return INV;
#else
        return FR_INV32_C[0];
#endif
    }
    __host__ __device__ static constexpr uint32_t one(int i) {
        constexpr uint32_t t[8] = {0xfffffffeu, 0x00000001u, 0x00034802u, 0x5884b7fau,
                                        0xecbc4ff5u, 0x998c4fefu, 0xacc5056fu, 0x1824b159u};
        return t[i];
    }
    __host__ __device__ static constexpr uint32_t r2(int i) {
        constexpr uint32_t t[8] = {0xf3f29c6du, 0xc999e990u, 0x87925c23u, 0x2b6cedcbu,
                                       0x7254398fu, 0x05d31496u, 0x9f59ff11u, 0x0748d9d9u};
        return t[i];
    }
};

// --------------------------------------------------------------------------
// Generic Montgomery field element
// --------------------------------------------------------------------------
template <class P>
struct Fe {
    static constexpr int N = P::N;
    uint32_t l[N];

    KZG_DI static Fe zero() {
        Fe r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = 0;
        return r;
    }
    __host__ static Fe zero_host() {
        Fe r;
        for (int i = 0; i < N; i++) r.l[i] = 0;
        return r;
    }
    KZG_DI static Fe one() {
        Fe r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = P::one(i);
        return r;
    }
    KZG_DI static Fe r2() {
        Fe r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = P::r2(i);
        return r;
    }
    KZG_DI bool is_zero() const {
        uint32_t v = 0;
#pragma unroll
        for (int i = 0; i < N; i++) v |= l[i];
        return v == 0;
    }
    KZG_DI bool operator==(const Fe& o) const {
        uint32_t v = 0;
#pragma unroll
        for (int i = 0; i < N; i++) v |= l[i] ^ o.l[i];
        return v == 0;
    }
    KZG_DI bool operator!=(const Fe& o) const { return !(*this == o); }

    // r = (a >= MOD) ? a - MOD : a      (a < 2*MOD)
    KZG_DI void final_sub() {
        uint32_t t[N];
        t[0] = sub_cc(l[0], P::mod(0));
#pragma unroll
        for (int i = 1; i < N; i++) t[i] = subc_cc(l[i], P::mod(i));
        uint32_t borrow = subc(0, 0);  // 0xffffffff if a < MOD
#pragma unroll
        for (int i = 0; i < N; i++) l[i] = borrow ? l[i] : t[i];
    }

    KZG_DI friend Fe operator+(const Fe& a, const Fe& b) {
        Fe r;
        r.l[0] = add_cc(a.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < N; i++) r.l[i] = addc_cc(a.l[i], b.l[i]);
        // top limb has spare bits (381 < 384, 255 < 256): no carry out for a,b < MOD
        r.final_sub();
        return r;
    }
    KZG_DI friend Fe operator-(const Fe& a, const Fe& b) {
        Fe r;
        r.l[0] = sub_cc(a.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < N; i++) r.l[i] = subc_cc(a.l[i], b.l[i]);
        uint32_t borrow = subc(0, 0);  // all ones if a < b
        if (borrow) {   // predicated adds of immediates: no AND per limb
            r.l[0] = add_cc(r.l[0], P::mod(0));
#pragma unroll
            for (int i = 1; i < N - 1; i++) r.l[i] = addc_cc(r.l[i], P::mod(i));
            r.l[N - 1] = addc(r.l[N - 1], P::mod(N - 1));
        }
        return r;
    }
    KZG_DI Fe neg() const { return zero() - *this; }
    KZG_DI Fe dbl() const { return *this + *this; }

    // Montgomery product a*b/R mod MOD.  With KZG_OUTLINE_FQ_MUL the 12-limb product is a real function call
    // (one copy of the ~700-instruction body per kernel) so that large kernels stay inside the instruction cache.
    KZG_DI friend Fe operator*(const Fe& a, const Fe& b) {
#ifdef KZG_OUTLINE_FQ_MUL
        if constexpr (N == 12) return mul_outlined(a, b);
#endif
#ifdef KZG_OUTLINE_FR_MUL
        if constexpr (N == 8) return mul_outlined(a, b);
#endif
#ifdef KZG_ROLLED_FQ_MUL
        if constexpr (N == 12) return mul_rolled(a, b);
#endif
#ifdef KZG_FR_KARATSUBA
        if constexpr (N == 8) return mul_kara(a, b);
#endif
        return mul_inline(a, b);
    }
    // One b-limb step of the interleaved product: X is the accumulator that currently plays "even", Y "odd".
    // After the step the roles are swapped (Y even, X odd, X shifted down by two limbs).
    KZG_DI static void mul_step(uint32_t* X, uint32_t* Y, const uint32_t* a, uint32_t bi) {
        Y[0] = add_cc(Y[0], X[1]);
#pragma unroll
        for (int j = 0; j < N - 2; j += 2) {
            X[j] = madc_lo_cc(a[j + 1], bi, X[j + 2]);
            X[j + 1] = madc_hi_cc(a[j + 1], bi, X[j + 3]);
        }
        X[N - 2] = madc_lo_cc(a[N - 1], bi, 0);
        X[N - 1] = madc_hi(a[N - 1], bi, 0);
        Y[0] = mad_lo_cc(a[0], bi, Y[0]);
        Y[1] = madc_hi_cc(a[0], bi, Y[1]);
#pragma unroll
        for (int j = 2; j < N; j += 2) {
            Y[j] = madc_lo_cc(a[j], bi, Y[j]);
            Y[j + 1] = madc_hi_cc(a[j], bi, Y[j + 1]);
        }
        X[N - 1] = addc(X[N - 1], 0);
        reduce_step(Y, X);
    }
    // Same product as mul_inline with the limb loop kept rolled (two b-limbs per trip, b rotated through
    // registers): ~10x less code, so kernels that inline many products stay inside the instruction cache.
    KZG_DI static Fe mul_rolled(const Fe& a, const Fe& b) {
        uint32_t X[N], Y[N], bb[N];
#pragma unroll
        for (int j = 0; j < N; j++) { X[j] = 0; Y[j] = 0; bb[j] = b.l[j]; }
#pragma unroll 1
        for (int i = 0; i < N; i += 2) {
            mul_step(X, Y, a.l, bb[0]);
            mul_step(Y, X, a.l, bb[1]);
#pragma unroll
            for (int j = 0; j < N - 2; j++) bb[j] = bb[j + 2];
        }
        Fe r;
        r.l[0] = add_cc(X[1], Y[0]);
#pragma unroll
        for (int j = 1; j < N - 1; j++) r.l[j] = addc_cc(X[j + 1], Y[j]);
        r.l[N - 1] = addc(0, Y[N - 1]);
        r.final_sub();
        return r;
    }
    __device__ __noinline__ static Fe mul_outlined(Fe a, Fe b) {
#ifdef KZG_FQ_KARATSUBA
        if constexpr (N == 12) return mul_kara(a, b);
#endif
        return mul_inline(a, b);
    }
    // FINAL = false: "lazy" form for operands and results in [0, 2*MOD) instead of [0, MOD): the conditional final
    // subtraction is skipped.  Valid where 4*MOD < 2^(32N) (Fq: p < 2^381): inputs < 2*MOD give (a*b + M*MOD)/R <
    // MOD*(4*MOD/R + 1) < 1.5*MOD, and the running value stays below 3*MOD.  See the *_lazy helpers below.
    template <bool FINAL = true>
    KZG_DI static Fe mul_inline(const Fe& a, const Fe& b) {
        uint32_t E[N], O[N];
        // ---- i = 0: plain products, no accumulation ----
        {
            const uint32_t bi = b.l[0];
#pragma unroll
            for (int j = 0; j < N; j += 2) {
                O[j] = mul_lo(a.l[j + 1], bi);
                O[j + 1] = mul_hi(a.l[j + 1], bi);
                E[j] = mul_lo(a.l[j], bi);
                E[j + 1] = mul_hi(a.l[j], bi);
            }
            reduce_step(E, O);
        }
#pragma unroll
        for (int i = 1; i < N; i++) {
            const uint32_t bi = b.l[i];
            // T/2^32: new even accumulator = old O (+ old E[1] folded into limb 0),
            // new odd accumulator = old E >> 64.
            uint32_t nE[N], nO[N];
#pragma unroll
            for (int j = 0; j < N; j++) nE[j] = O[j];
            nE[0] = add_cc(nE[0], E[1]);
#pragma unroll
            for (int j = 0; j < N - 2; j += 2) {
                nO[j] = madc_lo_cc(a.l[j + 1], bi, E[j + 2]);
                nO[j + 1] = madc_hi_cc(a.l[j + 1], bi, E[j + 3]);
            }
            nO[N - 2] = madc_lo_cc(a.l[N - 1], bi, 0);
            nO[N - 1] = madc_hi(a.l[N - 1], bi, 0);
            nE[0] = mad_lo_cc(a.l[0], bi, nE[0]);
            nE[1] = madc_hi_cc(a.l[0], bi, nE[1]);
#pragma unroll
            for (int j = 2; j < N; j += 2) {
                nE[j] = madc_lo_cc(a.l[j], bi, nE[j]);
                nE[j + 1] = madc_hi_cc(a.l[j], bi, nE[j + 1]);
            }
            nO[N - 1] = addc(nO[N - 1], 0);
            reduce_step(nE, nO);
#pragma unroll
            for (int j = 0; j < N; j++) { E[j] = nE[j]; O[j] = nO[j]; }
        }
        // result = E/2^32 + O   (E[0] == 0)
        Fe r;
        r.l[0] = add_cc(E[1], O[0]);
#pragma unroll
        for (int j = 1; j < N - 1; j++) r.l[j] = addc_cc(E[j + 1], O[j]);
        r.l[N - 1] = addc(0, O[N - 1]);
        if constexpr (FINAL) r.final_sub();
        return r;
    }

    // ------------------------------------------------------------------------------------------------------
    // FP64-pipe product (B200 has a full-rate DFMA pipe, 64 lanes/clk/SM, separate from the FMA-heavy pipe that
    // executes IMAD.WIDE).  a is taken as N 32-bit limbs, b as 2N 16-bit chunks; every partial product is
    // < 2^48 and a column of <= 12 of them plus its bias stays below 2^52, so a chain of DFMAs accumulates each
    // column EXACTLY in the mantissa of a double biased by 2^52.  Column j (weight 2^(32j)) has an "E" sum over
    // the low chunks and an "O" sum over the high chunks (weight 2^(32j+16)).  The raw bit patterns
    // 0x433<<52 | sum are then added as 64-bit integers at limb j; the exponent bits that come along are
    // cancelled by constants folded into the biases (limbs of -C mod 2^(64N), C = the sum of all exponent
    // patterns), so no masking is needed.  2*N*N DFMAs + 4N-2 integer adds + 2N funnel shifts replace the
    // N*N wide multiply-adds of the schoolbook product; the Montgomery reduction stays on the IMAD pipe.
    // Model and bounds: scripts/dfma_model.py.
    // ------------------------------------------------------------------------------------------------------
    KZG_DI static double dfma_bias_e(int j) {
        // 2^52 + limb j of -C
        return __hiloint2double(0x43300000, j == 0 ? 0u : (j == 1 ? 0xBCD00000u : 0xBCCFBCCFu));
    }
    KZG_DI static void mul_wide_dfma(uint32_t* T, const Fe& a, const Fe& b) {
        constexpr int M = 2 * N;
        const double two52 = __hiloint2double(0x43300000, 0);
        double ad[N], bl[N], bh[N];
#pragma unroll
        for (int i = 0; i < N; i++) {
            ad[i] = __hiloint2double(0x43300000, a.l[i]) - two52;
            bl[i] = __hiloint2double(0x43300000, b.l[i] & 0xffffu) - two52;
            bh[i] = __hiloint2double(0x43300000, b.l[i] >> 16) - two52;
        }
        uint32_t Elo[M - 1], Ehi[M - 1], Olo[M - 1], Ohi[M - 1];
#pragma unroll
        for (int j = 0; j < M - 1; j++) {
            double e = dfma_bias_e(j);
            // last O column also carries limb 2N-1 of -C (shifted into its own 2^16 frame)
            double o = (j == M - 2) ? __hiloint2double(0x43300000 | 0xBCCFu, 0xBCCF0000u) : two52;
#pragma unroll
            for (int i = 0; i < N; i++) {
                const int k = j - i;
                if (k >= 0 && k < N) {
                    e = __fma_rn(ad[i], bl[k], e);
                    o = __fma_rn(ad[i], bh[k], o);
                }
            }
            Elo[j] = (uint32_t)__double2loint(e); Ehi[j] = (uint32_t)__double2hiint(e);
            Olo[j] = (uint32_t)__double2loint(o); Ohi[j] = (uint32_t)__double2hiint(o);
        }
        uint32_t S1[M], S2[M];
        S1[0] = Elo[0];
        S1[1] = add_cc(Elo[1], Ehi[0]);
#pragma unroll
        for (int k = 2; k < M - 1; k++) S1[k] = addc_cc(Elo[k], Ehi[k - 1]);
        S1[M - 1] = addc(Ehi[M - 2], 0);
        S2[0] = Olo[0];
        S2[1] = add_cc(Olo[1], Ohi[0]);
#pragma unroll
        for (int k = 2; k < M - 1; k++) S2[k] = addc_cc(Olo[k], Ohi[k - 1]);
        S2[M - 1] = addc(Ohi[M - 2], 0);
        T[0] = add_cc(S1[0], S2[0] << 16);
#pragma unroll
        for (int k = 1; k < M - 1; k++) T[k] = addc_cc(S1[k], __funnelshift_l(S2[k - 1], S2[k], 16));
        T[M - 1] = addc(S1[M - 1], __funnelshift_l(S2[M - 2], S2[M - 1], 16));
    }
    // T (2N limbs, < MOD * 2^(32N)) -> T / 2^(32N) mod MOD: word-serial reduction on the interleaved (E, O) frame.
    KZG_DI static Fe mont_reduce_wide(const uint32_t* T) {
        uint32_t E[N], O[N];
#pragma unroll
        for (int j = 0; j < N; j++) { E[j] = T[j]; O[j] = 0; }
#pragma unroll
        for (int i = 0; i < N; i++) {
            reduce_step(E, O);
            // frame >>= 32: new E = O + (old E[1] at limb 0), new O = old E >> 64; then add T[N+i] at limb N-1
            uint32_t nE[N], nO[N];
            nE[0] = add_cc(O[0], E[1]);
#pragma unroll
            for (int j = 1; j < N - 1; j++) nE[j] = addc_cc(O[j], 0);
            nE[N - 1] = addc_cc(O[N - 1], T[N + i]);
            uint32_t top = addc(0, 0);
#pragma unroll
            for (int j = 0; j < N - 2; j++) nO[j] = E[j + 2];
            nO[N - 2] = 0;
            nO[N - 1] = top;   // carry out of limb N-1 has weight 2^(32N) = O[N-1]
#pragma unroll
            for (int j = 0; j < N; j++) { E[j] = nE[j]; O[j] = nO[j]; }
        }
        Fe r;
        r.l[0] = E[0];
        r.l[1] = add_cc(E[1], O[0]);
#pragma unroll
        for (int j = 2; j < N; j++) r.l[j] = addc_cc(E[j], O[j - 1]);
        r.final_sub();
        return r;
    }
    KZG_DI static Fe mul_dfma(const Fe& a, const Fe& b) {
        uint32_t T[2 * N];
        mul_wide_dfma(T, a, b);
        return mont_reduce_wide(T);
    }
    __device__ __noinline__ static Fe mul_dfma_outlined(Fe a, Fe b) { return mul_dfma(a, b); }
    KZG_DI Fe sqr() const {
#ifdef KZG_FQ_DEDICATED_SQR
        if constexpr (N == 12) return sqr_outlined(*this);
#endif
        return (*this) * (*this);
    }
    __device__ __noinline__ static Fe sqr_outlined(Fe a) { return sqr_inline(a); }

    // a^2/R mod MOD with N(N+1)/2 partial products for the square (off-diagonal terms once, doubled, plus the
    // diagonal) followed by a word-serial Montgomery reduction: 78 + 144 wide multiply-adds for Fq instead of 288.
    template <bool FINAL = true>
    KZG_DI static Fe sqr_inline(const Fe& a) {
        constexpr int M = 2 * N;
        // off-diagonal products a_i*a_j (i < j) in two accumulators by parity of i+j:
        //   A[k] has weight 2^(32k) (pairs at even k), B[k] has weight 2^(32(k+1)) (pairs at even k).
        uint32_t A[M], B[M];
#pragma unroll
        for (int k = 0; k < M; k++) { A[k] = 0; B[k] = 0; }
#pragma unroll
        for (int i = 0; i < N - 1; i++) {
            // j = i+1, i+3, ...: position i+j = 2i+1, 2i+3, ... (odd)  -> B at index 2i, 2i+2, ...
            {
                bool first = true;
#pragma unroll
                for (int j = i + 1; j < N; j += 2) {
                    const int k = i + j - 1;
                    if (first) { B[k] = mad_lo_cc(a.l[i], a.l[j], B[k]); first = false; }
                    else B[k] = madc_lo_cc(a.l[i], a.l[j], B[k]);
                    B[k + 1] = madc_hi_cc(a.l[i], a.l[j], B[k + 1]);
                }
                // absorb the carry in the next limb (fresh for this row: earlier rows end one pair lower)
                const int jl = i + 1 + 2 * ((N - 1 - (i + 1)) / 2);
                const int ke = i + jl - 1 + 2;
                if (ke < M) B[ke] = addc(B[ke], 0);
            }
            // j = i+2, i+4, ...: position i+j even -> A at index i+j
            if (i + 2 < N) {
                bool first = true;
#pragma unroll
                for (int j = i + 2; j < N; j += 2) {
                    const int k = i + j;
                    if (first) { A[k] = mad_lo_cc(a.l[i], a.l[j], A[k]); first = false; }
                    else A[k] = madc_lo_cc(a.l[i], a.l[j], A[k]);
                    A[k + 1] = madc_hi_cc(a.l[i], a.l[j], A[k + 1]);
                }
                const int jl = i + 2 + 2 * ((N - 1 - (i + 2)) / 2);
                const int ke = i + jl + 2;
                if (ke < M) A[ke] = addc(A[ke], 0);
            }
        }
        // T = 2*(A + B<<32) + sum a_i^2 2^(64i)
        uint32_t T[M];
        T[0] = A[0];
        T[1] = add_cc(A[1], B[0]);
#pragma unroll
        for (int k = 2; k < M - 1; k++) T[k] = addc_cc(A[k], B[k - 1]);
        T[M - 1] = addc(A[M - 1], B[M - 2]);
#pragma unroll
        for (int k = M - 1; k > 0; k--) T[k] = funnel_l(T[k - 1], T[k], 1);
        T[0] <<= 1;
        T[0] = mad_lo_cc(a.l[0], a.l[0], T[0]);
        T[1] = madc_hi_cc(a.l[0], a.l[0], T[1]);
#pragma unroll
        for (int i = 1; i < N; i++) {
            T[2 * i] = madc_lo_cc(a.l[i], a.l[i], T[2 * i]);
            if (i < N - 1) T[2 * i + 1] = madc_hi_cc(a.l[i], a.l[i], T[2 * i + 1]);
            else T[2 * i + 1] = madc_hi(a.l[i], a.l[i], T[2 * i + 1]);
        }
        return redc_wide<FINAL>(T);
    }

    // ------------------------------------------------------------------------------------------------------
    // Product / reduction split ("wide" schedule).  IMAD.WIDE issues at 32 lanes/clk/SM while IADD3/LOP3 issue at
    // 128: three of every four issue slots of a scheduler that is saturating the FMA-heavy pipe are free.  One level
    // of (subtractive) Karatsuba trades a quarter of the N*N partial products of a*b for additions on the ALU pipe:
    //   a*b = z0 + (z0 + z2 + (a0-a1)(b1-b0)) 2^(16N) + z2 2^(32N),  z0 = a0*b0, z2 = a1*b1      (3 (N/2)^2 products)
    // followed by a word-serial Montgomery reduction of the 2N-limb product: 108 + 144 wide multiply-adds per Fq
    // product instead of 288 (Fr: 48 + 64 instead of 128).  Keeping product and reduction apart also allows LAZY
    // reduction: a*b - c*d costs two wide products and ONE reduction (mul_sub_mul).
    // Bit-exactness of every routine below is checked on the host against a big-integer model
    // (tests/cpp/field_host_model.cu, run by tests/test_field_host_model.py) and on the device (test op 9..11).
    // ------------------------------------------------------------------------------------------------------
    // H x H limb schoolbook product -> 2H limbs.  Partial products a_j*b_i with i+j even accumulate in A (register
    // pair (k, k+1), k = i+j), those with i+j odd in B (pair index k = i+j-1, weight 2^(32(k+1))), so every product is
    // one mad.lo.cc/madc.hi.cc pair on an aligned pair and each row is an unbroken carry chain; T = A + (B << 32).
    template <int H>
    KZG_DI static void mul_half(uint32_t* T, const uint32_t* a, const uint32_t* b) {
        constexpr int M = 2 * H;
        uint32_t A[M], B[M];
#pragma unroll
        for (int k = 0; k < M; k++) { A[k] = 0; B[k] = 0; }
#pragma unroll
        for (int i = 0; i < H; i++) {
            {
                int klast = -1;
#pragma unroll
                for (int j = (i & 1); j < H; j += 2) {
                    const int k = i + j;
                    if (klast < 0) A[k] = mad_lo_cc(a[j], b[i], A[k]);
                    else A[k] = madc_lo_cc(a[j], b[i], A[k]);
                    A[k + 1] = madc_hi_cc(a[j], b[i], A[k + 1]);
                    klast = k;
                }
                // the carry out of the row lands on a limb that holds at most one earlier carry bit
                if (klast >= 0 && klast + 2 < M) A[klast + 2] = addc(A[klast + 2], 0);
            }
            {
                int klast = -1;
#pragma unroll
                for (int j = 1 - (i & 1); j < H; j += 2) {
                    const int k = i + j - 1;
                    if (klast < 0) B[k] = mad_lo_cc(a[j], b[i], B[k]);
                    else B[k] = madc_lo_cc(a[j], b[i], B[k]);
                    B[k + 1] = madc_hi_cc(a[j], b[i], B[k + 1]);
                    klast = k;
                }
                if (klast >= 0 && klast + 2 < M) B[klast + 2] = addc(B[klast + 2], 0);
            }
        }
        T[0] = A[0];
        T[1] = add_cc(A[1], B[0]);
#pragma unroll
        for (int k = 2; k < M - 1; k++) T[k] = addc_cc(A[k], B[k - 1]);
        T[M - 1] = addc(A[M - 1], B[M - 2]);
    }
    // x = |x0 - x1| over H limbs (x0 = lo[0..H), x1 = hi[0..H)); returns all-ones when x0 < x1
    template <int H>
    KZG_DI static uint32_t abs_diff(uint32_t* d, const uint32_t* lo, const uint32_t* hi) {
        d[0] = sub_cc(lo[0], hi[0]);
#pragma unroll
        for (int i = 1; i < H; i++) d[i] = subc_cc(lo[i], hi[i]);
        const uint32_t s = subc(0, 0);
        d[0] = sub_cc(d[0] ^ s, s);   // (d ^ s) - s: two's-complement negation when s is all ones
#pragma unroll
        for (int i = 1; i < H - 1; i++) d[i] = subc_cc(d[i] ^ s, s);
        d[H - 1] = subc(d[H - 1] ^ s, s);
        return s;
    }
    // T (2N limbs) = a * b, one Karatsuba level over N/2-limb halves
    KZG_DI static void mul_wide(uint32_t* T, const uint32_t* a, const uint32_t* b) {
        constexpr int H = N / 2;
        uint32_t m[N];
        uint32_t s;
        {
            uint32_t da[H], db[H];
            const uint32_t sa = abs_diff<H>(da, a, a + H);       // a0 - a1
            const uint32_t sb = abs_diff<H>(db, b + H, b);       // b1 - b0
            mul_half<H>(m, da, db);
            s = sa ^ sb;                                         // all ones: the cross term is -m
        }
        mul_half<H>(T, a, b);
        mul_half<H>(T + N, a + H, b + H);
        // t = z0 + z2 +- m  (N+1 limbs; the true value a0*b1 + a1*b0 is non-negative)
        uint32_t t[N + 1];
        t[0] = add_cc(T[0], T[N]);
#pragma unroll
        for (int i = 1; i < N; i++) t[i] = addc_cc(T[i], T[N + i]);
        t[N] = addc(0, 0);
        (void)add_cc(s, s);                                      // carry flag := s & 1 (the +1 of the negation)
        t[0] = addc_cc(t[0], m[0] ^ s);
#pragma unroll
        for (int i = 1; i < N; i++) t[i] = addc_cc(t[i], m[i] ^ s);
        t[N] = addc(t[N], s);                                    // sign extension
        T[H] = add_cc(T[H], t[0]);
#pragma unroll
        for (int i = 1; i <= N; i++) T[H + i] = addc_cc(T[H + i], t[i]);
#pragma unroll
        for (int k = H + N + 1; k < 2 * N - 1; k++) T[k] = addc_cc(T[k], 0);
        T[2 * N - 1] = addc(T[2 * N - 1], 0);
    }
    // T (2N limbs, < MOD * 2^(32N)) -> T / 2^(32N) mod MOD.  Only the low N limbs take part in the N word-serial rounds
    // ((T + M*MOD)/R = (T_lo + M*MOD)/R + T_hi exactly); they run on the interleaved (E, O) frame of mul_inline, where
    // the carry of the frame shift enters the next round's odd chain, so no round pays a carry ripple.
    template <bool FINAL = true>
    KZG_DI static Fe redc_wide(const uint32_t* T) {
        uint32_t E[N], O[N];
#pragma unroll
        for (int j = 0; j < N; j++) { E[j] = T[j]; O[j] = 0; }
        reduce_step(E, O);
#pragma unroll
        for (int i = 1; i < N; i++) {
            uint32_t nE[N], nO[N];
#pragma unroll
            for (int j = 0; j < N; j++) nE[j] = O[j];
            nE[0] = add_cc(nE[0], E[1]);                          // carry -> limb 0 of the new odd frame
            const uint32_t m = nE[0] * P::inv32();
#pragma unroll
            for (int j = 0; j < N - 2; j += 2) {
                nO[j] = madc_lo_cc(P::mod(j + 1), m, E[j + 2]);
                nO[j + 1] = madc_hi_cc(P::mod(j + 1), m, E[j + 3]);
            }
            nO[N - 2] = madc_lo_cc(P::mod(N - 1), m, 0);
            nO[N - 1] = madc_hi(P::mod(N - 1), m, 0);
            nE[0] = mad_lo_cc(P::mod(0), m, nE[0]);
            nE[1] = madc_hi_cc(P::mod(0), m, nE[1]);
#pragma unroll
            for (int j = 2; j < N; j += 2) {
                nE[j] = madc_lo_cc(P::mod(j), m, nE[j]);
                nE[j + 1] = madc_hi_cc(P::mod(j), m, nE[j + 1]);
            }
            nO[N - 1] = addc(nO[N - 1], 0);
#pragma unroll
            for (int j = 0; j < N; j++) { E[j] = nE[j]; O[j] = nO[j]; }
        }
        // U = E/2^32 + O (<= MOD), result = U + T_hi (< 2 MOD)
        Fe r;
        r.l[0] = add_cc(E[1], O[0]);
#pragma unroll
        for (int j = 1; j < N - 1; j++) r.l[j] = addc_cc(E[j + 1], O[j]);
        r.l[N - 1] = addc(0, O[N - 1]);
        r.l[0] = add_cc(r.l[0], T[N]);
#pragma unroll
        for (int j = 1; j < N - 1; j++) r.l[j] = addc_cc(r.l[j], T[N + j]);
        r.l[N - 1] = addc(r.l[N - 1], T[2 * N - 1]);
        if constexpr (FINAL) r.final_sub();
        return r;
    }
    KZG_DI static Fe mul_kara(const Fe& a, const Fe& b) {
        uint32_t T[2 * N];
        mul_wide(T, a.l, b.l);
        return redc_wide(T);
    }
    // a*b - c*d with one reduction: D = a*b - c*d (+ MOD*2^(32N) when negative) lies in [0, MOD*2^(32N))
    KZG_DI static Fe mul_sub_mul(const Fe& a, const Fe& b, const Fe& c, const Fe& d) {
        uint32_t T[2 * N], U[2 * N];
        mul_wide(T, a.l, b.l);
        mul_wide(U, c.l, d.l);
        T[0] = sub_cc(T[0], U[0]);
#pragma unroll
        for (int k = 1; k < 2 * N; k++) T[k] = subc_cc(T[k], U[k]);
        const uint32_t borrow = subc(0, 0);
        T[N] = add_cc(T[N], P::mod(0) & borrow);
#pragma unroll
        for (int k = 1; k < N - 1; k++) T[N + k] = addc_cc(T[N + k], P::mod(k) & borrow);
        T[2 * N - 1] = addc(T[2 * N - 1], P::mod(N - 1) & borrow);
        return redc_wide(T);
    }
    // (a*b + c*d) / R mod MOD in ONE interleaved Montgomery pass: every round adds the two product rows a*b_i and c*d_i
    // and performs one reduction row -- 2*N^2 + N^2 + N wide multiply-adds instead of 2*(2*N^2 + N) for two products,
    // with the carry-chain structure (and therefore the instruction mix) of mul_inline: no extra additions, one final
    // subtraction instead of two.  Inputs a, b, d < MOD and c <= MOD (so that MOD - y can be passed for a subtraction):
    // the running value stays below 3*MOD, which must fit the N-limb frame: 3*MOD < 2^(32N).  True for Fq (p < 2^381),
    // FALSE for Fr (3r > 2^256) -- Fq only (static_assert below; the host model shows the Fr overflow).  The result
    // (a*b + c*d + M*MOD)/R < MOD*(2*MOD/R + 1) < 1.25*MOD needs one final subtraction.
    KZG_DI static void mul_row_acc(uint32_t* nE, uint32_t* nO, const uint32_t* a, uint32_t bi) {
        nO[0] = mad_lo_cc(a[1], bi, nO[0]);
        nO[1] = madc_hi_cc(a[1], bi, nO[1]);
#pragma unroll
        for (int j = 2; j < N - 2; j += 2) {
            nO[j] = madc_lo_cc(a[j + 1], bi, nO[j]);
            nO[j + 1] = madc_hi_cc(a[j + 1], bi, nO[j + 1]);
        }
        nO[N - 2] = madc_lo_cc(a[N - 1], bi, nO[N - 2]);
        nO[N - 1] = madc_hi(a[N - 1], bi, nO[N - 1]);
        nE[0] = mad_lo_cc(a[0], bi, nE[0]);
        nE[1] = madc_hi_cc(a[0], bi, nE[1]);
#pragma unroll
        for (int j = 2; j < N; j += 2) {
            nE[j] = madc_lo_cc(a[j], bi, nE[j]);
            nE[j + 1] = madc_hi_cc(a[j], bi, nE[j + 1]);
        }
        nO[N - 1] = addc(nO[N - 1], 0);
    }
    template <bool FINAL = true>
    KZG_DI static Fe mul_add_mul(const Fe& a, const Fe& b, const Fe& c, const Fe& d) {
        static_assert(bits() + 2 <= 32 * N, "mul_add_mul: 3*MOD must fit in N limbs (Fq only)");
        uint32_t E[N], O[N];
        {   // round 0: the first row needs no accumulation
            const uint32_t bi = b.l[0];
#pragma unroll
            for (int j = 0; j < N; j += 2) {
                O[j] = mul_lo(a.l[j + 1], bi);
                O[j + 1] = mul_hi(a.l[j + 1], bi);
                E[j] = mul_lo(a.l[j], bi);
                E[j + 1] = mul_hi(a.l[j], bi);
            }
            mul_row_acc(E, O, c.l, d.l[0]);
            reduce_step(E, O);
        }
#pragma unroll
        for (int i = 1; i < N; i++) {
            const uint32_t bi = b.l[i];
            uint32_t nE[N], nO[N];
#pragma unroll
            for (int j = 0; j < N; j++) nE[j] = O[j];
            nE[0] = add_cc(nE[0], E[1]);
#pragma unroll
            for (int j = 0; j < N - 2; j += 2) {
                nO[j] = madc_lo_cc(a.l[j + 1], bi, E[j + 2]);
                nO[j + 1] = madc_hi_cc(a.l[j + 1], bi, E[j + 3]);
            }
            nO[N - 2] = madc_lo_cc(a.l[N - 1], bi, 0);
            nO[N - 1] = madc_hi(a.l[N - 1], bi, 0);
            nE[0] = mad_lo_cc(a.l[0], bi, nE[0]);
            nE[1] = madc_hi_cc(a.l[0], bi, nE[1]);
#pragma unroll
            for (int j = 2; j < N; j += 2) {
                nE[j] = madc_lo_cc(a.l[j], bi, nE[j]);
                nE[j + 1] = madc_hi_cc(a.l[j], bi, nE[j + 1]);
            }
            nO[N - 1] = addc(nO[N - 1], 0);
            mul_row_acc(nE, nO, c.l, d.l[i]);
            reduce_step(nE, nO);
#pragma unroll
            for (int j = 0; j < N; j++) { E[j] = nE[j]; O[j] = nO[j]; }
        }
        Fe r;
        r.l[0] = add_cc(E[1], O[0]);
#pragma unroll
        for (int j = 1; j < N - 1; j++) r.l[j] = addc_cc(E[j + 1], O[j]);
        r.l[N - 1] = addc(0, O[N - 1]);
        if constexpr (FINAL) r.final_sub();
        return r;
    }
    // MOD - a without the "a == 0" special case: a value in (0, MOD], acceptable as the `c` operand of mul_add_mul
    KZG_DI Fe neg_raw() const {
        Fe r;
        r.l[0] = sub_cc(P::mod(0), l[0]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.l[i] = subc_cc(P::mod(i), l[i]);
        r.l[N - 1] = subc(P::mod(N - 1), l[N - 1]);
        return r;
    }
    // ---- lazy range [0, 2*MOD) (Fq only: needs 8*MOD < 2^(32N)) -------------------------------------------------
    // The accumulate kernel keeps its running XYZZ sum in this range: the nine products of a mixed addition skip their
    // final subtraction, differences add 2*MOD instead of MOD, and only zero tests and stores look at canonical values.
    __host__ __device__ static constexpr uint32_t mod2(int i) {   // limb i of 2*MOD
        return (P::mod(i) << 1) | (i > 0 ? (P::mod(i - 1) >> 31) : 0u);
    }
    KZG_DI static Fe sub_lazy(const Fe& a, const Fe& b) {         // a - b (+ 2*MOD when negative), inputs and result < 2*MOD
        Fe r;
        r.l[0] = sub_cc(a.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < N; i++) r.l[i] = subc_cc(a.l[i], b.l[i]);
        const uint32_t borrow = subc(0, 0);
        r.l[0] = add_cc(r.l[0], mod2(0) & borrow);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.l[i] = addc_cc(r.l[i], mod2(i) & borrow);
        r.l[N - 1] = addc(r.l[N - 1], mod2(N - 1) & borrow);
        return r;
    }
    KZG_DI Fe dbl_lazy() const {                                   // 2a reduced below 2*MOD (a < 2*MOD)
        Fe r;
        r.l[0] = add_cc(l[0], l[0]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.l[i] = addc_cc(l[i], l[i]);
        r.l[N - 1] = addc(l[N - 1], l[N - 1]);
        uint32_t t[N];
        t[0] = sub_cc(r.l[0], mod2(0));
#pragma unroll
        for (int i = 1; i < N; i++) t[i] = subc_cc(r.l[i], mod2(i));
        const uint32_t borrow = subc(0, 0);
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = borrow ? r.l[i] : t[i];
        return r;
    }
    KZG_DI Fe neg_lazy() const {                                   // 2*MOD - a, in (0, 2*MOD] for a < 2*MOD
        Fe r;
        r.l[0] = sub_cc(mod2(0), l[0]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.l[i] = subc_cc(mod2(i), l[i]);
        r.l[N - 1] = subc(mod2(N - 1), l[N - 1]);
        return r;
    }
    KZG_DI bool is_zero_modp() const {                             // a in {0, MOD} (a < 2*MOD)
        uint32_t z = 0, m = 0;
#pragma unroll
        for (int i = 0; i < N; i++) { z |= l[i]; m |= l[i] ^ P::mod(i); }
        return z == 0 || m == 0;
    }
    KZG_DI Fe canon() const { Fe r = *this; r.final_sub(); return r; }
    __device__ __noinline__ static Fe mul_lazy_outlined(Fe a, Fe b) { return mul_inline<false>(a, b); }
    __device__ __noinline__ static Fe sqr_lazy_outlined(Fe a) { return sqr_inline<false>(a); }
    // a*b + c*d with c already negated by the caller (neg_lazy): result < 2*MOD, no final subtraction
    __device__ __noinline__ static Fe mul_add_mul_lazy_outlined(Fe a, Fe b, Fe c, Fe d) { return mul_add_mul<false>(a, b, c, d); }

    // a*b - c*d
    KZG_DI static Fe mul_sub_mul_fused(const Fe& a, const Fe& b, const Fe& c, const Fe& d) { return mul_add_mul(a, b, c.neg_raw(), d); }
    __device__ __noinline__ static Fe mul_sub_mul_fused_outlined(Fe a, Fe b, Fe c, Fe d) { return mul_add_mul(a, b, c.neg_raw(), d); }
    __device__ __noinline__ static Fe mul_kara_outlined(Fe a, Fe b) { return mul_kara(a, b); }
    __device__ __noinline__ static Fe mul_sub_mul_outlined(Fe a, Fe b, Fe c, Fe d) { return mul_sub_mul(a, b, c, d); }

    // E + O*2^32 += m*MOD with m chosen so that limb 0 of E becomes zero.
    KZG_DI static void reduce_step(uint32_t* E, uint32_t* O) {
        if constexpr (P::LOW_LIMBS_1_FFFFFFFF) {
            // MOD = 1 + 0xffffffff*2^32 + ... (Fr): m = -E0 and m*(1 + (2^32-1)*2^32) + E0 =
            // (E0+1)*2^32 + (~E0)*2^64 for E0 != 0 (and 0 for E0 == 0), so the two lowest partial products of
            // every round are three integer additions on the ALU pipe instead of two wide multiply-adds on the
            // FMA-heavy pipe: 6 instead of 8 wide mads per round, 120 instead of 136 per Fr product.  Bit-exact
            // (94 GPU tests pass with it), Fr multiply chain +3 %, but the NTT gets 5 % SLOWER (0.984 vs 0.936 ms at
            // 2^22): the compare/select/carry instructions lengthen the dependent chain.  Kept as an opt-in
            // experiment (-DKZG_FR_SPECIAL_REDUCE).  Limb-level model: scripts/fr_reduce_model.py.
            const uint32_t e0 = E[0];
            const uint32_t m = e0 * P::inv32();
            const uint32_t nz = e0 != 0 ? 1u : 0u;
            const uint32_t c = e0 == 0xffffffffu ? 1u : 0u;
            const uint32_t add1 = e0 + nz;        // limb 1 (wraps to 0 with c = 1 when e0 = 2^32-1)
            const uint32_t add2 = m - nz + c;     // limb 2
            O[1] = add_cc(O[1], add2);
#pragma unroll
            for (int j = 2; j < N - 2; j += 2) {
                O[j] = madc_lo_cc(P::mod(j + 1), m, O[j]);
                O[j + 1] = madc_hi_cc(P::mod(j + 1), m, O[j + 1]);
            }
            O[N - 2] = madc_lo_cc(P::mod(N - 1), m, O[N - 2]);
            O[N - 1] = madc_hi(P::mod(N - 1), m, O[N - 1]);
            E[0] = 0;
            E[1] = add_cc(E[1], add1);
#pragma unroll
            for (int j = 2; j < N; j += 2) {
                E[j] = madc_lo_cc(P::mod(j), m, E[j]);
                E[j + 1] = madc_hi_cc(P::mod(j), m, E[j + 1]);
            }
            O[N - 1] = addc(O[N - 1], 0);
            return;
        }
        const uint32_t m = E[0] * P::inv32();
        O[0] = mad_lo_cc(P::mod(1), m, O[0]);
        O[1] = madc_hi_cc(P::mod(1), m, O[1]);
#pragma unroll
        for (int j = 2; j < N - 2; j += 2) {
            O[j] = madc_lo_cc(P::mod(j + 1), m, O[j]);
            O[j + 1] = madc_hi_cc(P::mod(j + 1), m, O[j + 1]);
        }
        O[N - 2] = madc_lo_cc(P::mod(N - 1), m, O[N - 2]);
        O[N - 1] = madc_hi(P::mod(N - 1), m, O[N - 1]);
        E[0] = mad_lo_cc(P::mod(0), m, E[0]);
        E[1] = madc_hi_cc(P::mod(0), m, E[1]);
#pragma unroll
        for (int j = 2; j < N; j += 2) {
            E[j] = madc_lo_cc(P::mod(j), m, E[j]);
            E[j + 1] = madc_hi_cc(P::mod(j), m, E[j + 1]);
        }
        O[N - 1] = addc(O[N - 1], 0);
    }

    // Montgomery form <-> canonical integer limbs
    KZG_DI Fe to_mont() const { return (*this) * r2(); }
    KZG_DI Fe from_mont() const {
        Fe o = zero();
        o.l[0] = 1;
        return (*this) * o;
    }

    // a^e for a little-endian limb exponent known at run time (vartime)
    template <int EN>
    KZG_DI Fe pow_limbs(const uint32_t (&e)[EN]) const {
        Fe acc = one();
        Fe base = *this;
        for (int i = 0; i < EN * 32; i++) {
            if ((e[i >> 5] >> (i & 31)) & 1) acc = acc * base;
            base = base.sqr();
        }
        return acc;
    }
    KZG_DI Fe pow_u64(uint64_t e) const {
        Fe acc = one();
        Fe base = *this;
        while (e) {
            if (e & 1) acc = acc * base;
            base = base.sqr();
            e >>= 1;
        }
        return acc;
    }
    // Fermat inversion a^(MOD-2); inv(0) = 0.  (kept as the cross-check of inv())
    __device__ __noinline__ Fe inv_fermat() const {
        Fe acc = one();
        Fe base = *this;
#pragma unroll 1
        for (int i = 0; i < N * 32; i++) {
            if ((modm2_limb(i >> 5) >> (i & 31)) & 1) acc = acc * base;
            base = base.sqr();
        }
        return acc;
    }

    // Montgomery inverse by Kaliski's binary "almost inverse" (shifts and subtractions only, ~1.4*bits
    // iterations) followed by two Montgomery products for the 2^k correction: about 15x fewer cycles than the
    // Fermat chain.  Input and output in Montgomery form; inv(0) = 0.
    __device__ __noinline__ Fe inv() const {
        if (is_zero()) return zero();
        uint32_t u[N], v[N], r[N], s[N];
#pragma unroll
        for (int i = 0; i < N; i++) { u[i] = P::mod(i); v[i] = l[i]; r[i] = 0; s[i] = 0; }
        s[0] = 1;
        int k = 0;
        // Each pass removes ALL trailing zero bits at once (up to 31): a subtraction of two odd numbers is always followed
        // by at least one halving, so the subtract step and the halvings after it are one iteration.  Same sequence of
        // operations as the bit-at-a-time loop (same k, same r, s), about half as many trips.
#pragma unroll 1
        for (;;) {
            uint32_t vz = 0;
#pragma unroll
            for (int i = 0; i < N; i++) vz |= v[i];
            if (vz == 0) break;
            if ((u[0] & 1) == 0) {
                const int c = tz31(u[0]);
                shrk(u, c); shlk(s, c);
                k += c;
            } else if ((v[0] & 1) == 0) {
                const int c = tz31(v[0]);
                shrk(v, c); shlk(r, c);
                k += c;
            } else {
                uint32_t t[N];
                t[0] = sub_cc(v[0], u[0]);
#pragma unroll
                for (int i = 1; i < N; i++) t[i] = subc_cc(v[i], u[i]);
                uint32_t borrow = subc(0, 0);
                if (borrow) {  // u > v:  u = (u - v) / 2^c, r = r + s, s = 2^c s
                    u[0] = sub_cc(0, t[0]);
#pragma unroll
                    for (int i = 1; i < N; i++) u[i] = subc_cc(0, t[i]);
                    const int c = tz31(u[0]);
                    shrk(u, c);
                    addn(r, s); shlk(s, c);
                    k += c;
                } else {       // v >= u: v = (v - u) / 2^c, s = s + r, r = 2^c r
                    uint32_t tzv = 0;
#pragma unroll
                    for (int i = 0; i < N; i++) { v[i] = t[i]; tzv |= t[i]; }
                    const int c = tzv ? tz31(v[0]) : 1;   // v == u: exactly one halving, as in the bit-at-a-time loop
                    shrk(v, c);
                    addn(s, r); shlk(r, c);
                    k += c;
                }
            }
        }
        // r < 2*MOD holds a^-1 * 2^k up to sign: r = MOD - (r mod MOD)
        Fe x;
#pragma unroll
        for (int i = 0; i < N; i++) x.l[i] = r[i];
        x.final_sub();
        x = zero() - x;
        // x = (a R)^-1 2^k.  Multiply by 2^(2w-k) R with w = 32N:  mont(x, to_mont(2^e)), e = 2w - k.
        const int w = 32 * N;
        int e = 2 * w - k;
        constexpr int TOPBIT = bits() - 1;  // 2^TOPBIT < MOD
        int e0 = e < TOPBIT ? e : TOPBIT;
        Fe c = zero();
#pragma unroll
        for (int i = 0; i < N; i++) c.l[i] = ((e0 >> 5) == i) ? (1u << (e0 & 31)) : 0u;
        for (int i = e0; i < e; i++) c = c.dbl();
        c = c.to_mont();
        return x * c;
    }
    // ---- inversion by Pornin's optimised binary GCD ("Optimized Binary GCD for Modular Inversion", 2020) -------------
    // EXPERIMENT, not used by the library (test op 12, host model): same result as inv() (the inverse is unique).  The
    // 30 inner steps of a round run on two 62-bit approximations (low 30 bits + top 32 bits of a and b) and only the
    // round's 2x2 update matrix touches the N-limb numbers.  Built to shorten the dependent chain where ONE thread inverts
    // while a block or a stream waits (msm_finish_kernel, block_inverse); measured SLOWER on B200 than Kaliski's loop
    // (msm_finish_kernel 80 vs 68 us, pointwise_div_kernel 165 vs 146 us): with 32-bit registers the 64-bit inner steps
    // and the four signed N-limb multiply-accumulate chains per round cost as many dependent instructions as the 1.4*bits
    // shift/subtract iterations they replace (profiles/experiments/README.md).
    //   a = y, b = MOD, u = 1, v = 0;  invariants a = u*y, b = v*y (mod MOD);  after ROUNDS rounds a = 0, b = gcd = 1
    //   and v = y^-1.  Parity decisions use exact low bits, so every division by 2^30 is exact; a wrong size comparison
    //   (the approximations are not exact) only produces a negative value, which is negated together with its row.
    // If the final b is not 1 (cannot happen for 0 < y < MOD; kept as a guard) the result comes from inv().
    static constexpr int BG_K = 30;
    KZG_DI static uint32_t bg_limb(const uint32_t* x, int idx) {   // x[idx] without dynamic register indexing; 0 outside
        uint32_t r = 0;
#pragma unroll
        for (int i = 0; i < N; i++) r = (i == idx) ? x[i] : r;
        return r;
    }
    KZG_DI static int bg_bitlen(const uint32_t* x) {
        int top = -1;
        uint32_t w = 0;
#pragma unroll
        for (int i = 0; i < N; i++)
            if (x[i]) { top = i; w = x[i]; }
        if (top < 0) return 0;
        int lz = 0;
        while (!(w & 0x80000000u)) { w <<= 1; lz++; }
        return 32 * top + 32 - lz;
    }
    KZG_DI static uint64_t bg_approx(const uint32_t* x, int n) {   // n >= 62: low 30 bits | bits [n-32, n) << 30
        const int pos = n - 32, li = pos >> 5, off = pos & 31;
        const uint32_t lo = bg_limb(x, li), hi = bg_limb(x, li + 1);
        const uint32_t top = off ? ((lo >> off) | (hi << (32 - off))) : lo;
        return (uint64_t)(x[0] & 0x3fffffffu) | ((uint64_t)top << 30);
    }
    // T = x*f + y*g as N+1 two's-complement limbs, returns the (sign-extended) top part; |f| + |g| <= 2^30
    KZG_DI static int64_t bg_lincomb(const uint32_t* x, const uint32_t* y, int32_t f, int32_t g, uint32_t* T) {
        const uint32_t af = f < 0 ? (uint32_t)(-f) : (uint32_t)f, ag = g < 0 ? (uint32_t)(-g) : (uint32_t)g;
        int64_t acc = 0;
#pragma unroll
        for (int i = 0; i < N; i++) {
            const int64_t pf = (int64_t)((uint64_t)x[i] * af), pg = (int64_t)((uint64_t)y[i] * ag);   // < 2^62 each
            acc += (f < 0 ? -pf : pf) + (g < 0 ? -pg : pg);
            T[i] = (uint32_t)acc;
            acc >>= 32;
        }
        T[N] = (uint32_t)acc;
        return acc;
    }
    KZG_DI static void bg_shift30(const uint32_t* T, uint32_t* out) {   // out = low N limbs of (T >> 30)
#pragma unroll
        for (int i = 0; i < N; i++) out[i] = (T[i] >> 30) | (T[i + 1] << 2);
    }
    KZG_DI static void bg_negate(uint32_t* x) {
        uint32_t c = 1;
#pragma unroll
        for (int i = 0; i < N; i++) {
            const uint32_t t = ~x[i] + c;
            c = (c && t == 0) ? 1u : 0u;
            x[i] = t;
        }
    }
#if defined(KZG_HOST_EMU) && !defined(__CUDA_ARCH__)
    static long& bingcd_fallbacks() { static long n = 0; return n; }
#endif
    __device__ __noinline__ Fe inv_bingcd() const {
        if (is_zero()) return zero();
        uint32_t a[N], b[N], u[N], v[N];
#pragma unroll
        for (int i = 0; i < N; i++) { a[i] = l[i]; b[i] = P::mod(i); u[i] = 0; v[i] = 0; }
        u[0] = 1;
        const uint32_t minv = P::inv32();   // -MOD^-1 mod 2^32
        constexpr int BG_ROUNDS = (2 * bits() - 1 + BG_K - 1) / BG_K + 1;   // Pornin: 2*len - 1 inner steps suffice; one spare round
#pragma unroll 1
        for (int round = 0; round < BG_ROUNDS; round++) {
            int n = bg_bitlen(a), nb = bg_bitlen(b);
            n = n > nb ? n : nb;
            n = n > 62 ? n : 62;
            uint64_t xa = bg_approx(a, n), xb = bg_approx(b, n);
            int32_t f0 = 1, g0 = 0, f1 = 0, g1 = 1;
#pragma unroll 1
            for (int j = 0; j < BG_K; j++) {
                if (xa & 1) {
                    if (xa < xb) {
                        const uint64_t tx = xa; xa = xb; xb = tx;
                        int32_t t = f0; f0 = f1; f1 = t;
                        t = g0; g0 = g1; g1 = t;
                    }
                    xa -= xb; f0 -= f1; g0 -= g1;
                }
                xa >>= 1; f1 <<= 1; g1 <<= 1;
            }
            // (a, b) <- ((a f0 + b g0) / 2^30, (a f1 + b g1) / 2^30), made non-negative together with their rows
            uint32_t T[N + 1], na[N], nbv[N];
            if (bg_lincomb(a, b, f0, g0, T) < 0) { bg_shift30(T, na); bg_negate(na); f0 = -f0; g0 = -g0; }
            else bg_shift30(T, na);
            if (bg_lincomb(a, b, f1, g1, T) < 0) { bg_shift30(T, nbv); bg_negate(nbv); f1 = -f1; g1 = -g1; }
            else bg_shift30(T, nbv);
            // (u, v) <- the same combinations divided by 2^30 modulo MOD: add t*MOD with t = -T/MOD mod 2^30 first
            uint32_t nu[N], nv[N];
#pragma unroll
            for (int row = 0; row < 2; row++) {
                int64_t hi = bg_lincomb(u, v, row ? f1 : f0, row ? g1 : g0, T);
                const uint32_t t = (T[0] * minv) & 0x3fffffffu;
                uint64_t c = 0;
#pragma unroll
                for (int i = 0; i < N; i++) {
                    c += (uint64_t)t * P::mod(i) + T[i];
                    T[i] = (uint32_t)c;
                    c >>= 32;
                }
                hi += (int64_t)c;
                T[N] = (uint32_t)hi;
                uint32_t* out = row ? nv : nu;
                bg_shift30(T, out);                  // value in (-MOD, 2 MOD), two's complement if negative
                Fe w;
#pragma unroll
                for (int i = 0; i < N; i++) w.l[i] = out[i];
                if (hi < 0) {                        // += MOD
                    w.l[0] = add_cc(w.l[0], P::mod(0));
#pragma unroll
                    for (int i = 1; i < N - 1; i++) w.l[i] = addc_cc(w.l[i], P::mod(i));
                    w.l[N - 1] = addc(w.l[N - 1], P::mod(N - 1));
                } else {
                    w.final_sub();
                }
#pragma unroll
                for (int i = 0; i < N; i++) out[i] = w.l[i];
            }
#pragma unroll
            for (int i = 0; i < N; i++) { a[i] = na[i]; b[i] = nbv[i]; u[i] = nu[i]; v[i] = nv[i]; }
        }
        uint32_t rest = b[0] ^ 1u;
#pragma unroll
        for (int i = 1; i < N; i++) rest |= b[i];
#pragma unroll
        for (int i = 0; i < N; i++) rest |= a[i];
        if (rest) {                                  // guard: not reached for 0 < y < MOD (the host model counts it)
#if defined(KZG_HOST_EMU) && !defined(__CUDA_ARCH__)
            bingcd_fallbacks()++;
#endif
            return inv();
        }
        // v = (x R)^-1 as a plain integer; (x R)^-1 * R^3 / R = x^-1 R
        Fe r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = v[i];
        return r * (r2() * r2());
    }
    __host__ __device__ static constexpr int bits() {
        int b = 32 * N;
        uint32_t top = P::mod(N - 1);
        while (!(top & 0x80000000u)) { top <<= 1; b--; }
        return b;
    }
    KZG_DI static void shr1(uint32_t* a) {
#pragma unroll
        for (int i = 0; i < N - 1; i++) a[i] = __funnelshift_r(a[i], a[i + 1], 1);
        a[N - 1] >>= 1;
    }
    KZG_DI static void shl1(uint32_t* a) {
#pragma unroll
        for (int i = N - 1; i > 0; i--) a[i] = __funnelshift_l(a[i - 1], a[i], 1);
        a[0] <<= 1;
    }
    // trailing zero bits of a non-zero-or-zero low limb, capped at 31 (a zero limb: 31 -- the value stays even)
    KZG_DI static int tz31(uint32_t w) {
#if defined(KZG_HOST_EMU) && !defined(__CUDA_ARCH__)
        const int c = w ? __builtin_ctz(w) : 31;
#else
        const int c = w ? (__ffs((int)w) - 1) : 31;
#endif
        return c > 31 ? 31 : c;
    }
    KZG_DI static void shrk(uint32_t* a, int c) {   // a >>= c, 0 < c < 32
#pragma unroll
        for (int i = 0; i < N - 1; i++) a[i] = __funnelshift_r(a[i], a[i + 1], c);
        a[N - 1] >>= c;
    }
    KZG_DI static void shlk(uint32_t* a, int c) {   // a <<= c, 0 < c < 32
#pragma unroll
        for (int i = N - 1; i > 0; i--) a[i] = __funnelshift_l(a[i - 1], a[i], c);
        a[0] <<= c;
    }
    KZG_DI static void addn(uint32_t* a, const uint32_t* b) {
        a[0] = add_cc(a[0], b[0]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) a[i] = addc_cc(a[i], b[i]);
        a[N - 1] = addc(a[N - 1], b[N - 1]);
    }
    // limb i of MOD-2 (borrow propagated at compile time)
    __host__ __device__ static constexpr uint32_t modm2_const(int i) {
        uint32_t borrow = 2;
        uint32_t v = 0;
        for (int k = 0; k <= i; k++) {
            uint32_t m = P::mod(k);
            v = m - borrow;
            borrow = (m < borrow) ? 1u : 0u;
        }
        return v;
    }
    KZG_DI static uint32_t mod_limb(int i) {  // run-time indexed limb of the modulus
        uint32_t v = 0;
#pragma unroll
        for (int k = 0; k < N; k++) v = (i == k) ? P::mod(k) : v;
        return v;
    }
    KZG_DI static uint32_t modm2_limb(int i) {
        uint32_t v = 0;
#pragma unroll
        for (int k = 0; k < N; k++) v = (i == k) ? modm2_const(k) : v;
        return v;
    }
};

using Fq = Fe<FqParams>;
using Fr = Fe<FrParams>;

}  // namespace kzg
