// msm.cu -- G1 multi-scalar multiplication (Pippenger) for sm_100a.
//
// Replaces `G1Projective::multi_exp(points, scalars)` of the blstrs fork
// (Cargo.toml:27) at its five prover call sites: coeff_form.rs:61,78,102 and
// eval_form.rs:118,136.  Like multi_exp it sums over min(len) pairs.
//
// Design (B200-first, see DESIGN.md "MSM"):
//  * The SRS is immutable for a prover's lifetime (the prover holds
//    &'params KZGParams, coeff_form.rs:39), so the handle stores, once, the
//    affine table T[w][i] = 2^(c*w) * P_i for every window w (HBM is 180 GB).
//    All windows then share ONE bucket set: no per-window reduction and no
//    Horner doublings remain on the per-call path.
//  * per call:  K2 scalar prep (Montgomery -> canonical, signed c-bit digits)
//               K3 CUB radix sort of (bucket, table-index|sign) pairs
//               K4 bucket accumulation: every thread walks a fixed-length chunk
//                  of the sorted list with one XYZZ accumulator (mixed adds),
//                  so the load is balanced for ANY scalar distribution
//                  (full-width, or the reference benches' single-u64 scalars
//                  where most digits collide into a few buckets)
//               K4b segmented fix-up of chunk-boundary partial sums
//               K5 bucket reduction sum_b (b+1)*B[b]: chunked running sums,
//                  small scalar-mul per chunk, tree sum
//               K6 XYZZ -> affine / Jacobian
#include <cub/device/device_radix_sort.cuh>

#ifndef KZG_MSM_NO_DEDICATED_SQR
#define KZG_FQ_DEDICATED_SQR 1
#endif
#if defined(KZG_MSM_ROLLED_MUL)
#define KZG_ROLLED_FQ_MUL 1
#elif !defined(KZG_MSM_INLINE_MUL)
#define KZG_OUTLINE_FQ_MUL 1
#endif

// The next table point of the accumulation is staged through shared memory with cp.async (default); the round-1
// register-staged software pipeline is kept behind -DKZG_MSM_REG_PREFETCH (measured: 4.89 vs 4.84 ms, r02j logs).
#ifndef KZG_MSM_REG_PREFETCH
#define KZG_MSM_SMEM_PREFETCH 1
#endif
// The accumulator of the accumulate kernel lives in the lazy range [0, 2p) (field.cuh): default; -DKZG_MSM_CANONICAL keeps
// every intermediate below p (measured: 4.533 vs 4.498 ms, profiles/r02p_msm_lazy_range.log).
#ifndef KZG_MSM_CANONICAL
#define KZG_MSM_LAZY2P 1
#endif

#include <cstdlib>

#include "common.cuh"
#include "ec.cuh"
#include "msm.cuh"

namespace kzg {

// --------------------------------------------------------------------------
// K2: scalars (Montgomery Fr) -> signed digits -> (key, value) pairs
//   key = |d|-1 (bucket), or `nbuckets` for d == 0 (sorts to the end)
//   val = (w*n + i) | (d < 0) << 31
// --------------------------------------------------------------------------
// C = 0: window width and count at run time (any c); C > 0: compile-time instance for the two production widths (limb
// positions and shifts become constants, the loop unrolls: 68 -> see profiles/r02ad_digits.log).
template <int C>
__global__ void __launch_bounds__(256) msm_digits_kernel(const Fr* __restrict__ scalars, uint32_t m, uint32_t n,
                                                         int c_rt, int nwin_rt, uint32_t* __restrict__ keys,
                                                         uint32_t* __restrict__ vals,
                                                         unsigned long long* __restrict__ zero_count) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t zeros = 0;
    if (i < m) {
        Fr s;
        {   // two 16-byte loads (device scalar vectors are 32-byte aligned)
            const uint4* q = reinterpret_cast<const uint4*>(&scalars[i]);
            const uint4 a = q[0], b = q[1];
            s.l[0] = a.x; s.l[1] = a.y; s.l[2] = a.z; s.l[3] = a.w;
            s.l[4] = b.x; s.l[5] = b.y; s.l[6] = b.z; s.l[7] = b.w;
        }
        s = s.from_mont();
        const int c = C ? C : c_rt;
        const int nwin = C ? (255 + C) / C : nwin_rt;   // ceil(256 / c): 255 scalar bits + the carry
        const uint32_t half = 1u << (c - 1);
        uint32_t carry = 0;
#pragma unroll
        for (int w = 0; w < (C ? (255 + C) / C : 1); w++) {
            if (C == 0) break;
            const int bit = w * C;
            const int limb = bit >> 5, off = bit & 31;
            uint64_t v = s.l[limb < 8 ? limb : 7];
            if (limb >= 8) v = 0;
            if (limb + 1 < 8) v |= (uint64_t)s.l[limb + 1 < 8 ? limb + 1 : 7] << 32;
            uint32_t d = (uint32_t)((v >> off) & ((1u << C) - 1)) + carry;
            uint32_t neg = 0;
            if (d > half) { d = (1u << C) - d; neg = 1; carry = 1; } else { carry = 0; }
            uint32_t key = d ? d - 1 : half;
            zeros += (d == 0);
            keys[(size_t)w * m + i] = key;
            vals[(size_t)w * m + i] = ((uint32_t)w * n + i) | (neg << 31);
        }
        for (int w = 0; C == 0 && w < nwin; w++) {
            int bit = w * c;
            int limb = bit >> 5, off = bit & 31;
            uint64_t v = 0;
            // dynamic limb select without local-memory indexing
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (k == limb) v |= (uint64_t)s.l[k];
                if (k == limb + 1) v |= (uint64_t)s.l[k] << 32;
            }
            uint32_t d = (uint32_t)((v >> off) & ((1u << c) - 1)) + carry;
            uint32_t neg = 0;
            if (d > half) {
                d = (1u << c) - d;
                neg = 1;
                carry = 1;
            } else {
                carry = 0;
            }
            uint32_t key = d ? d - 1 : half;
            zeros += (d == 0);
            keys[(size_t)w * m + i] = key;
            vals[(size_t)w * m + i] = ((uint32_t)w * n + i) | (neg << 31);
        }
    }
    // one atomic per warp
    for (int o = 16; o > 0; o >>= 1) zeros += __shfl_down_sync(0xffffffffu, zeros, o);
    if ((threadIdx.x & 31) == 0 && zeros) atomicAdd(zero_count, (unsigned long long)zeros);
}

// --------------------------------------------------------------------------
// K4: chunked bucket accumulation
// --------------------------------------------------------------------------
constexpr uint32_t SLOT_EMPTY = 0xffffffffu;

// The P == 0 cases (the incoming point equals the accumulator or its negative).  Inline on purpose: as an out-of-line
// function taking `acc` by reference it forced the accumulator into a 336-byte stack frame, and the hot loop then moved
// it through local memory on every addition (ncu, round 1/2: 42 M LDL + 66 M STL warp instructions per launch, 5.5 % of
// the kernel's instructions).  With the field products out of line the inlined doubling is a dozen CALLs, not 60 KB.
KZG_DI void xyzz_madd_rare(XYZZ& acc, const Fq& x2, const Fq& y2, bool same) {
    if (same) acc = xyzz_dbl_affine(x2, y2);
    else acc = XYZZ::inf();
}

// hot-path mixed add; the P == 0 cases (equal / opposite points) are rare
#ifdef KZG_MSM_LAZY2P
// Lazy range: the accumulator's coordinates live in [0, 2p) between additions (field.cuh "lazy range"): the nine field
// products of the addition skip their conditional final subtraction, differences add 2p instead of p.  The table point
// (x2, y2) is canonical; zero tests compare against {0, p}; st_acc canonicalises before a bucket or partial is stored.
KZG_DI void madd_hot(XYZZ& acc, const Fq& x2, const Fq& y2) {
    if (acc.is_inf()) {
        acc.x = x2; acc.y = y2; acc.zz = Fq::one(); acc.zzz = Fq::one();
        return;
    }
#if defined(KZG_MSM_INLINE_PRODUCTS) && KZG_MSM_INLINE_PRODUCTS >= 4   // experiment: some products inline (no argument marshalling, more code)
    Fq P = Fq::sub_lazy(Fq::mul_inline<false>(x2, acc.zz), acc.x);
    Fq R = Fq::sub_lazy(Fq::mul_inline<false>(y2, acc.zzz), acc.y);
#else
    Fq P = Fq::sub_lazy(Fq::mul_lazy_outlined(x2, acc.zz), acc.x);
    Fq R = Fq::sub_lazy(Fq::mul_lazy_outlined(y2, acc.zzz), acc.y);
#endif
    if (P.is_zero_modp()) {
        xyzz_madd_rare(acc, x2, y2, R.is_zero_modp());
        return;
    }
    Fq PP = Fq::sqr_lazy_outlined(P);
    Fq PPP = Fq::mul_lazy_outlined(P, PP);
    Fq Q = Fq::mul_lazy_outlined(acc.x, PP);
#if defined(KZG_MSM_INLINE_PRODUCTS) && KZG_MSM_INLINE_PRODUCTS >= 2
    acc.zz = Fq::mul_inline<false>(acc.zz, PP);
    acc.zzz = Fq::mul_inline<false>(acc.zzz, PPP);
#else
    acc.zz = Fq::mul_lazy_outlined(acc.zz, PP);
    acc.zzz = Fq::mul_lazy_outlined(acc.zzz, PPP);
#endif
    Fq X3 = Fq::sub_lazy(Fq::sub_lazy(Fq::sqr_lazy_outlined(R), PPP), Q.dbl_lazy());
    acc.y = Fq::mul_add_mul_lazy_outlined(R, Fq::sub_lazy(Q, X3), acc.y.neg_lazy(), PPP);
    acc.x = X3;
}
KZG_DI void st_acc(XYZZ* dst, const XYZZ& acc) {
    XYZZ c{acc.x.canon(), acc.y.canon(), acc.zz.canon(), acc.zzz.canon()};
    st_xyzz(dst, c);
}
#else
KZG_DI void madd_hot(XYZZ& acc, const Fq& x2, const Fq& y2) {
    if (acc.is_inf()) {
        acc.x = x2; acc.y = y2; acc.zz = Fq::one(); acc.zzz = Fq::one();
        return;
    }
    Fq P = x2 * acc.zz - acc.x;
    Fq R = y2 * acc.zzz - acc.y;
    if (P.is_zero()) {
        xyzz_madd_rare(acc, x2, y2, R.is_zero());
        return;
    }
    Fq PP = P.sqr();
    Fq PPP = P * PP;
    Fq Q = acc.x * PP;
    acc.zz = acc.zz * PP;
    acc.zzz = acc.zzz * PPP;
    Fq X3 = R.sqr() - PPP - Q.dbl();
#if !defined(KZG_MSM_SEPARATE_Y) && !defined(KZG_MSM_LAZY_Y)
    // Y3 = R*(Q - X3) - Y1*PPP: both products in ONE interleaved Montgomery pass (Fe::mul_add_mul: two product rows and one
    // reduction row per round -- 444 instead of 600 wide multiply-adds, same carry-chain structure, one final
    // subtraction).  Accumulate kernel 4.84 -> 4.53 ms (profiles/r02m_msm_fused_y.log).
    acc.y = Fq::mul_sub_mul_fused_outlined(R, Q - X3, acc.y, PPP);
#elif defined(KZG_MSM_LAZY_Y)
    acc.y = Fq::mul_sub_mul_outlined(R, Q - X3, acc.y, PPP);   // two wide products, ONE Montgomery reduction
#else
    acc.y = R * (Q - X3) - acc.y * PPP;
#endif
    acc.x = X3;
}
KZG_DI void st_acc(XYZZ* dst, const XYZZ& acc) { st_xyzz(dst, acc); }
#endif

struct PartialMeta {
    uint32_t bucket;   // SLOT_EMPTY if unused
    uint32_t is_start; // the bucket's segment begins inside the producing thread's chunk
};

#ifdef KZG_MSM_SMEM_PREFETCH
// The next table point travels global -> shared memory with cp.async (LDGSTS: no register staging) while the current
// one is added, instead of sitting in 24 registers across the ten out-of-line field products of the addition.
KZG_DI void cp_async16(void* smem_dst, const void* gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
KZG_DI void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
KZG_DI void cp_async_wait_all_but_one() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
#endif

#ifndef KZG_MSM_ACC_BLOCKS
#define KZG_MSM_ACC_BLOCKS 3
#endif
#ifndef KZG_MSM_ACC_THREADS
#define KZG_MSM_ACC_THREADS 128
#endif
#ifdef KZG_MSM_ACC_MAXREG   // experiments: an explicit register cap instead of the launch-bounds heuristic
__global__ void __maxnreg__(KZG_MSM_ACC_MAXREG)
#else
__global__ void __launch_bounds__(KZG_MSM_ACC_THREADS, KZG_MSM_ACC_BLOCKS)
#endif
msm_accumulate_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                      const Affine* __restrict__ table, uint64_t total, const unsigned long long* __restrict__ zero_count,
                      uint32_t chunk, XYZZ* __restrict__ buckets, XYZZ* __restrict__ partials,
                      PartialMeta* __restrict__ meta, uint32_t nthreads) {
#ifdef KZG_MSM_SMEM_PREFETCH
    __shared__ uint4 pstage[2][6][KZG_MSM_ACC_THREADS];   // [buffer][16-byte word of the point][thread]: 24 KiB, conflict-free columns
#endif
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nthreads) return;
    const uint64_t nvalid = total - *zero_count;
    uint64_t start = (uint64_t)t * chunk;
    uint64_t end = start + chunk;
    if (end > nvalid) end = nvalid;
    PartialMeta m0{SLOT_EMPTY, 0}, m1{SLOT_EMPTY, 0};
    if (start < end) {
        uint32_t cur = keys[start];
        bool cur_started_here = (start == 0) || (keys[start - 1] != cur);
        bool first = true;
        XYZZ acc = XYZZ::inf();
        // software pipeline: the next point is in flight while the current one is added
        uint32_t v = vals[start];
#ifdef KZG_MSM_SMEM_PREFETCH
        auto prefetch = [&](int buf, uint32_t val) {
            const uint4* src = reinterpret_cast<const uint4*>(&table[val & 0x7fffffffu]);
#pragma unroll
            for (int w = 0; w < 6; w++) cp_async16(&pstage[buf][w][threadIdx.x], src + w);
        };
        prefetch(0, v);
        cp_async_commit();
        int buf = 0;
#else
        Affine p = ldcs_affine(&table[v & 0x7fffffffu]);
#endif
        for (uint64_t pos = start; pos < end; pos++) {
            uint32_t k = cur, vn = 0;
            bool more = pos + 1 < end;
#ifdef KZG_MSM_SMEM_PREFETCH
            if (more) {
                k = keys[pos + 1];
                vn = vals[pos + 1];
                prefetch(buf ^ 1, vn);
            }
            cp_async_commit();             // one group per iteration (empty after the last point)
            cp_async_wait_all_but_one();   // => the group of THIS iteration's point has landed
            Affine p;
#pragma unroll
            for (int w = 0; w < 3; w++) {
                const uint4 a = pstage[buf][w][threadIdx.x], b = pstage[buf][3 + w][threadIdx.x];
                p.x.l[4 * w] = a.x; p.x.l[4 * w + 1] = a.y; p.x.l[4 * w + 2] = a.z; p.x.l[4 * w + 3] = a.w;
                p.y.l[4 * w] = b.x; p.y.l[4 * w + 1] = b.y; p.y.l[4 * w + 2] = b.z; p.y.l[4 * w + 3] = b.w;
            }
            buf ^= 1;
#else
            Affine pn = p;
            if (more) {
                k = keys[pos + 1];
                vn = vals[pos + 1];
                pn = ldcs_affine(&table[vn & 0x7fffffffu]);
            }
#endif
            if (!p.is_inf()) {
                Fq y = (v >> 31) ? p.y.neg() : p.y;
                madd_hot(acc, p.x, y);
            }
            if (more && k != cur) {
                // bucket `cur` ends inside this chunk
                if (cur_started_here) {
                    st_acc(&buckets[cur], acc);
                } else {  // can only be the first bucket of the chunk
                    st_acc(&partials[2 * (size_t)t], acc);
                    m0 = PartialMeta{cur, 0};
                }
                acc = XYZZ::inf();
                cur = k;
                cur_started_here = true;
                first = false;
            }
            v = vn;
#ifndef KZG_MSM_SMEM_PREFETCH
            p = pn;
#endif
        }
        bool ended_here = (end == nvalid) || (keys[end] != cur);
        if (cur_started_here && ended_here) {
            st_acc(&buckets[cur], acc);
        } else if (first) {
            st_acc(&partials[2 * (size_t)t], acc);
            m0 = PartialMeta{cur, cur_started_here ? 1u : 0u};
        } else {
            st_acc(&partials[2 * (size_t)t + 1], acc);
            m1 = PartialMeta{cur, 1u};  // not the first bucket => it started here
        }
    }
    meta[2 * (size_t)t] = m0;
    meta[2 * (size_t)t + 1] = m1;
}


// --------------------------------------------------------------------------
// 4-lane cooperative point arithmetic for the latency-bound tail kernels.
// The 14 field multiplications of an XYZZ addition have dependency depth 4 (and the 9 of a doubling depth 3), but a
// single thread issues them back to back (~0.6 us each).  Four adjacent lanes hold the SAME operands, each computes
// one product per step and the results are exchanged with shuffles: ~3x lower latency per point operation at 4x the
// lanes -- the right trade where a kernel has far fewer additions than the chip has lanes.
// --------------------------------------------------------------------------
KZG_DI Fq grp_bcast(const Fq& v, int src, unsigned gmask) {
    Fq r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = __shfl_sync(gmask, v.l[i], src, 4);
    return r;
}
KZG_DI Fq sel4(int r, const Fq& a0, const Fq& a1, const Fq& a2, const Fq& a3) {
    Fq o;
#pragma unroll
    for (int i = 0; i < 12; i++) o.l[i] = r == 0 ? a0.l[i] : (r == 1 ? a1.l[i] : (r == 2 ? a2.l[i] : a3.l[i]));
    return o;
}
KZG_DI unsigned grp_mask() { return 0xfu << (threadIdx.x & 28); }

// acc += q; all four lanes of the group pass identical acc and q and all receive the result
KZG_DI void xyzz_add_coop(XYZZ& acc, const XYZZ& q) {
    if (q.is_inf()) return;
    if (acc.is_inf()) { acc = q; return; }
    const unsigned gm = grp_mask();
    const int r = threadIdx.x & 3;
    Fq m = sel4(r, acc.x, q.x, acc.y, q.y) * sel4(r, q.zz, acc.zz, q.zzz, acc.zzz);
    Fq U1 = grp_bcast(m, 0, gm), U2 = grp_bcast(m, 1, gm), S1 = grp_bcast(m, 2, gm), S2 = grp_bcast(m, 3, gm);
    Fq P = U2 - U1, R = S2 - S1;
    if (P.is_zero()) {  // equal or opposite points: rare, take the single-thread path (uniform across the group)
        xyzz_add(acc, q);
        return;
    }
    m = sel4(r, P, R, acc.zz, acc.zzz) * sel4(r, P, R, q.zz, q.zzz);
    Fq PP = grp_bcast(m, 0, gm), RR = grp_bcast(m, 1, gm), ZZ12 = grp_bcast(m, 2, gm), ZZZ12 = grp_bcast(m, 3, gm);
    m = sel4(r, P, U1, ZZ12, ZZ12) * PP;
    Fq PPP = grp_bcast(m, 0, gm), Q = grp_bcast(m, 1, gm);
    acc.zz = grp_bcast(m, 2, gm);
    Fq X3 = RR - PPP - Q.dbl();
    m = sel4(r, R, S1, ZZZ12, ZZZ12) * sel4(r, Q - X3, PPP, PPP, PPP);
    Fq t1 = grp_bcast(m, 0, gm), t2 = grp_bcast(m, 1, gm);
    acc.zzz = grp_bcast(m, 2, gm);
    acc.x = X3;
    acc.y = t1 - t2;
}

// p = 2p, cooperative
KZG_DI void xyzz_dbl_coop(XYZZ& p) {
    if (p.is_inf() || p.y.is_zero()) { p = XYZZ::inf(); return; }
    const unsigned gm = grp_mask();
    const int r = threadIdx.x & 3;
    Fq U = p.y.dbl();
    Fq a = sel4(r, U, p.x, U, U);
    Fq m = a * a;
    Fq V = grp_bcast(m, 0, gm), XX = grp_bcast(m, 1, gm);
    Fq M = XX.dbl() + XX;
    m = sel4(r, U, p.x, V, M) * sel4(r, V, V, p.zz, M);
    Fq W = grp_bcast(m, 0, gm), S = grp_bcast(m, 1, gm), ZZ3 = grp_bcast(m, 2, gm), MM = grp_bcast(m, 3, gm);
    Fq X3 = MM - S.dbl();
    m = sel4(r, M, W, W, W) * sel4(r, S - X3, p.y, p.zzz, p.zzz);
    Fq t1 = grp_bcast(m, 0, gm), t2 = grp_bcast(m, 1, gm);
    p.zzz = grp_bcast(m, 2, gm);
    p.zz = ZZ3;
    p.x = X3;
    p.y = t1 - t2;
}

// k * p for k < 2^nbits (k uniform across the group).  The loop is the same for every group of the warp (fixed trip
// count, the addend is the identity where the bit is clear) so that groups do not serialise on divergent calls.
KZG_DI XYZZ xyzz_mul_small_coop(const XYZZ& p, uint32_t k, int nbits) {
    XYZZ acc = XYZZ::inf();
#pragma unroll 1
    for (int i = nbits - 1; i >= 0; i--) {
        xyzz_dbl_coop(acc);
        XYZZ q = ((k >> i) & 1) ? p : XYZZ::inf();
        xyzz_add_coop(acc, q);
    }
    return acc;
}

// tree sum of one value per 4-lane group over a block of NT threads (NT/4 groups); result in every lane of group 0
template <int NT>
KZG_DI XYZZ block_tree_sum_coop(XYZZ v, XYZZ* sh /* NT/4 entries */) {
    const int g = threadIdx.x >> 2, r = threadIdx.x & 3;
    if (r == 0) sh[g] = v;
    __syncthreads();
    for (int s = NT / 8; s > 0; s >>= 1) {
        if (g < s) {
            XYZZ a = sh[g];
            XYZZ b = sh[g + s];
            xyzz_add_coop(a, b);
            v = a;
        }
        __syncthreads();
        if (g < s && r == 0) sh[g] = v;
        __syncthreads();
    }
    return sh[0];
}

// K4b: hierarchical segmented reduction of the chunk-boundary partial sums.
// The slot list (2 per producing thread, in sorted-bucket order) is cut into runs: a run starts at a slot with
// is_start and continues over following slots of the same bucket (EMPTY slots are skipped).  Each thread owns
// FIX_F consecutive slots; runs that start and end inside its range are final and go to buckets[]; the (at most
// two) runs that cross its range boundary become the next level's slots.  Every level shrinks the list
// FIX_F/2-fold, so a bucket holding half of all points (reference-style u64 scalars) costs O(log) levels instead
// of one thread walking thousands of partials.
// Producers emit two slots each (2a: continuation of the previous producer's last bucket, 2a+1: own last bucket), so
// the natural run pairs are (2a+1, 2a+2): ranges start at odd slots ([0, F+1), [F+1, 2F+1), ...) to keep such pairs
// inside one thread.  A bounded look-ahead past the range end decides whether the last run is finished; if it cannot
// tell, the run is forwarded to the next level.
constexpr uint32_t FIX_F = 16;        // slots per thread from level 1 on
constexpr uint32_t FIX_F_MIN = 4;     // level 0: slots per thread are chosen per launch (see msm_run), in [4, 30]
constexpr uint32_t FIX_F_MAX = 30;
constexpr uint32_t FIX_PEEK = 4;

template <bool COOP>
#ifndef KZG_FIX_MINB
#define KZG_FIX_MINB 1
#endif
__global__ void __launch_bounds__(128, KZG_FIX_MINB) msm_fixup_level_kernel(const XYZZ* __restrict__ pin,
                                                              const PartialMeta* __restrict__ min_, uint32_t nslots_in,
                                                              XYZZ* __restrict__ buckets, XYZZ* __restrict__ pout,
                                                              PartialMeta* __restrict__ mout, uint32_t nthreads,
                                                              uint32_t F) {
    // COOP: one 4-lane group per range (the four lanes run the same control flow on the same data, lane 0 stores) --
    // used from level 1 on, where partials are sparse and latency is all that matters; level 0 is dense and uses
    // one thread per range.
    const uint32_t gt = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t t = COOP ? gt >> 2 : gt;
    const bool lead = COOP ? (threadIdx.x & 3) == 0 : true;
    if (t >= nthreads) return;
    uint32_t lo = t ? t * F + 1 : 0;
    uint32_t hi = min((t + 1) * F + 1, nslots_in);
    PartialMeta m0{SLOT_EMPTY, 0}, m1{SLOT_EMPTY, 0};
    XYZZ acc = XYZZ::inf();
    uint32_t cur = SLOT_EMPTY, cur_start = 0;
    bool flushed_any = false;
    // valid slots of the range as a bit mask (F + 1 <= 32 slots); lanes then walk their k-th valid slot
    // together, so sparse levels pay one point addition per round instead of one per slot position
    uint32_t mask = 0;
    for (uint32_t j = lo; j < hi; j++)
        if (min_[j].bucket != SLOT_EMPTY) mask |= 1u << (j - lo);
    while (mask) {
        uint32_t j = lo + (uint32_t)__ffs((int)mask) - 1;
        mask &= mask - 1;
        PartialMeta mj = min_[j];
        if (cur != SLOT_EMPTY && (mj.bucket != cur || mj.is_start)) {
            // run `cur` ends here
            if (cur_start) {
                if (lead) st_xyzz(&buckets[cur], acc);
            } else {  // continues a run from an earlier thread: only possible for the first run
                if (lead) st_xyzz(&pout[2 * (size_t)t], acc);
                m0 = PartialMeta{cur, 0};
            }
            flushed_any = true;
            acc = XYZZ::inf();
            cur = SLOT_EMPTY;
        }
        if (cur == SLOT_EMPTY) { cur = mj.bucket; cur_start = mj.is_start; }
        XYZZ q = ld_xyzz(&pin[j]);
        if (COOP) xyzz_add_coop(acc, q);
        else xyzz_add(acc, q);
    }
    if (cur != SLOT_EMPTY) {
        // does the run end at the range boundary?
        bool ended = false, known = false;
        for (uint32_t j = hi; j < hi + FIX_PEEK; j++) {
            if (j >= nslots_in) { ended = true; known = true; break; }
            PartialMeta mj = min_[j];
            if (mj.bucket == SLOT_EMPTY) continue;
            ended = (mj.bucket != cur) || mj.is_start;
            known = true;
            break;
        }
        if (known && ended && cur_start) {
            if (lead) st_xyzz(&buckets[cur], acc);
        } else if (!flushed_any) {
            if (lead) st_xyzz(&pout[2 * (size_t)t], acc);
            m0 = PartialMeta{cur, cur_start};
        } else {
            if (lead) st_xyzz(&pout[2 * (size_t)t + 1], acc);
            m1 = PartialMeta{cur, cur_start};
        }
    }
    if (lead) {
        mout[2 * (size_t)t] = m0;
        mout[2 * (size_t)t + 1] = m1;
    }
}

// --------------------------------------------------------------------------
// K5: bucket reduction  S = sum_{b=0}^{nb-1} (b+1) * B[b], two-level.
// Buckets form a rows x cols matrix (b = hi*cols + lo), so
//     S = cols * sum_hi hi * R[hi]  +  sum_lo (lo+1) * C[lo],
// with R = row sums and C = column sums.  Every step is a short chain (<= ~16 point operations per thread):
// (1) 2*nb additions spread over (rows+cols) blocks, (2) one small scalar multiplication per row/column and a
// block tree sum, (3) a handful of additions and log2(cols) doublings.  No thread ever walks a long run of
// buckets -- single-thread point additions cost ~8 us each, so the serial chain length IS the latency.
// --------------------------------------------------------------------------
constexpr int RC_THREADS = 64;
constexpr uint32_t MSM_COOP_THRESHOLD = 4096;  // additions per level below which lanes are plentiful (x4)

KZG_DI XYZZ block_tree_sum(XYZZ acc, XYZZ* sh, int nthreads) {
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = nthreads / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) {
            XYZZ a = sh[threadIdx.x];
            XYZZ b = sh[threadIdx.x + s];
            xyzz_add(a, b);
            sh[threadIdx.x] = a;
        }
        __syncthreads();
    }
    return sh[0];
}

// Row sums and column sums by level-by-level pairwise additions: every lane of every warp performs one useful
// addition per level (a serial chunk + shared-memory tree leaves half the lanes idle on most steps, and these
// kernels are bound by warp-time on the FMA pipe).
//   level 0:  Rb[hi][lo] = B[hi][lo] + B[hi][lo + cols/2]   (lo < cols/2)
//             Cb[hi][lo] = B[hi][lo] + B[hi + rows/2][lo]   (hi < rows/2)
//   level k:  Rb[hi][lo] += Rb[hi][lo + wr],  Cb[hi][lo] += Cb[hi + hc][lo]
// The row stride of Rb stays cols/2 and that of Cb stays cols; R[hi] = Rb[hi][0], C[lo] = Cb[0][lo].
__global__ void __launch_bounds__(128, 3) msm_pairsum_first_kernel(const XYZZ* __restrict__ B, uint32_t rows, uint32_t cols,
                                                                XYZZ* __restrict__ Rb, XYZZ* __restrict__ Cb) {
    const uint32_t half = (rows * cols) >> 1;
    uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < half) {  // R part
        uint32_t hw = cols >> 1;
        uint32_t hi = idx / hw, lo = idx % hw;
        XYZZ a = ld_xyzz(&B[(size_t)hi * cols + lo]);
        XYZZ b = ld_xyzz(&B[(size_t)hi * cols + lo + hw]);
        xyzz_add(a, b);
        st_xyzz(&Rb[idx], a);
    } else if (idx < 2 * half) {  // C part
        uint32_t j = idx - half;
        uint32_t hi = j / cols, lo = j % cols;
        XYZZ a = ld_xyzz(&B[(size_t)hi * cols + lo]);
        XYZZ b = ld_xyzz(&B[(size_t)(hi + (rows >> 1)) * cols + lo]);
        xyzz_add(a, b);
        st_xyzz(&Cb[j], a);
    }
}
// wr: number of live columns of Rb AFTER this level (0 = R is finished); hc likewise for the rows of Cb
__global__ void __launch_bounds__(128, 3) msm_pairsum_level_kernel(XYZZ* __restrict__ Rb, uint32_t rows, uint32_t rstride,
                                                                uint32_t wr, XYZZ* __restrict__ Cb, uint32_t cols,
                                                                uint32_t hc) {
    uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t nr = rows * wr;
    if (idx < nr) {
        uint32_t hi = idx / wr, lo = idx % wr;
        XYZZ a = ld_xyzz(&Rb[(size_t)hi * rstride + lo]);
        XYZZ b = ld_xyzz(&Rb[(size_t)hi * rstride + lo + wr]);
        xyzz_add(a, b);
        st_xyzz(&Rb[(size_t)hi * rstride + lo], a);
    } else if (idx < nr + hc * cols) {
        uint32_t j = idx - nr;
        XYZZ a = ld_xyzz(&Cb[j]);
        XYZZ b = ld_xyzz(&Cb[(size_t)hc * cols + j]);
        xyzz_add(a, b);
        st_xyzz(&Cb[j], a);
    }
}

// Same level as msm_pairsum_level_kernel with one 4-lane group per addition (used once a level has too few additions
// to fill the chip: latency, not throughput, is what such a level costs).
__global__ void __launch_bounds__(128) msm_pairsum_level_coop_kernel(XYZZ* __restrict__ Rb, uint32_t rows, uint32_t rstride,
                                                                     uint32_t wr, XYZZ* __restrict__ Cb, uint32_t cols,
                                                                     uint32_t hc) {
    uint32_t idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const bool lead = (threadIdx.x & 3) == 0;
    const uint32_t nr = rows * wr;
    if (idx < nr) {
        uint32_t hi = idx / wr, lo = idx % wr;
        XYZZ a = ld_xyzz(&Rb[(size_t)hi * rstride + lo]);
        XYZZ b = ld_xyzz(&Rb[(size_t)hi * rstride + lo + wr]);
        xyzz_add_coop(a, b);
        if (lead) st_xyzz(&Rb[(size_t)hi * rstride + lo], a);
    } else if (idx < nr + hc * cols) {
        uint32_t j = idx - nr;
        XYZZ a = ld_xyzz(&Cb[j]);
        XYZZ b = ld_xyzz(&Cb[(size_t)hc * cols + j]);
        xyzz_add_coop(a, b);
        if (lead) st_xyzz(&Cb[j], a);
    }
}

// One 4-lane group per row / column: hi * R[hi] or (lo+1) * C[lo] by cooperative double-and-add, then a cooperative
// tree sum over the block's 32 groups.  blocks [0, nbr): rows, blocks [nbr, nbr+nbc): columns.
constexpr int WS_THREADS = 128;
constexpr int WS_GROUPS = WS_THREADS / 4;
__global__ void __launch_bounds__(WS_THREADS) msm_weighted_sum_kernel(const XYZZ* __restrict__ R, uint32_t rows,
                                                                      uint32_t rstride, const XYZZ* __restrict__ C,
                                                                      uint32_t cols, uint32_t nbr,
                                                                      XYZZ* __restrict__ partial) {
    __shared__ XYZZ sh[WS_GROUPS];
    const uint32_t g = threadIdx.x >> 2;
    XYZZ base = XYZZ::inf();
    uint32_t k = 0;
    int nbits;
    if (blockIdx.x < nbr) {
        uint32_t hi = blockIdx.x * WS_GROUPS + g;
        if (hi < rows && hi > 0) { base = ld_xyzz(&R[(size_t)hi * rstride]); k = hi; }
        nbits = 32 - __clz(rows | 1);
    } else {
        uint32_t lo = (blockIdx.x - nbr) * WS_GROUPS + g;
        if (lo < cols) { base = ld_xyzz(&C[lo]); k = lo + 1; }
        nbits = 32 - __clz(cols | 1);
    }
    XYZZ acc = xyzz_mul_small_coop(base, k, nbits);
    XYZZ tot = block_tree_sum_coop<WS_THREADS>(acc, sh);
    if (threadIdx.x == 0) st_xyzz(&partial[blockIdx.x], tot);
}

// result = 2^log_cols * (sum of the first nbr partials) + (sum of the remaining nbc partials); one block, cooperative
__global__ void __launch_bounds__(WS_THREADS) msm_reduce_final_kernel(const XYZZ* __restrict__ partial, uint32_t nbr,
                                                                      uint32_t nbc, uint32_t log_cols,
                                                                      XYZZ* __restrict__ out) {
    __shared__ XYZZ sh[WS_GROUPS];
    __shared__ XYZZ vr_s;
    const uint32_t g = threadIdx.x >> 2;
    XYZZ acc = XYZZ::inf();
    for (uint32_t i = g; i < nbr; i += WS_GROUPS) { XYZZ q = ld_xyzz(&partial[i]); xyzz_add_coop(acc, q); }
    XYZZ vr = block_tree_sum_coop<WS_THREADS>(acc, sh);
    __syncthreads();
    if (g == 0) {
        for (uint32_t i = 0; i < log_cols; i++) xyzz_dbl_coop(vr);
        if (threadIdx.x == 0) vr_s = vr;
    }
    __syncthreads();
    acc = XYZZ::inf();
    for (uint32_t i = g; i < nbc; i += WS_GROUPS) { XYZZ q = ld_xyzz(&partial[nbr + i]); xyzz_add_coop(acc, q); }
    XYZZ vc = block_tree_sum_coop<WS_THREADS>(acc, sh);
    if (g == 0) {
        XYZZ v = vr_s;
        xyzz_add_coop(v, vc);
        if (threadIdx.x == 0) st_xyzz(out, v);
    }
}

// The same two steps with every weight cut into a low and a high half (k = k_hi * 2^h + k_lo): the double-and-add chains
// are half as long (they are the longest serial piece of the tail, ~7 us per point operation), the two halves run in
// different blocks, and the factor 2^h is applied ONCE, to the sum of the high halves.  Blocks of the weighted kernel:
// [rows low | rows high | columns low | columns high], nbr, nbr, nbc, nbc of them; partial[] in the same order.
__global__ void __launch_bounds__(WS_THREADS) msm_weighted_sum_split_kernel(const XYZZ* __restrict__ R, uint32_t rows,
                                                                            uint32_t rstride, const XYZZ* __restrict__ C,
                                                                            uint32_t cols, uint32_t nbr, uint32_t nbc,
                                                                            uint32_t h_r, uint32_t bits_r, uint32_t h_c,
                                                                            uint32_t bits_c, XYZZ* __restrict__ partial) {
    __shared__ XYZZ sh[WS_GROUPS];
    const uint32_t g = threadIdx.x >> 2;
    XYZZ base = XYZZ::inf();
    uint32_t k = 0, h, bits, part;
    uint32_t b = blockIdx.x;
    if (b < 2 * nbr) {
        part = b >= nbr;
        uint32_t hi = (b - part * nbr) * WS_GROUPS + g;
        if (hi < rows && hi > 0) { base = ld_xyzz(&R[(size_t)hi * rstride]); k = hi; }
        h = h_r; bits = bits_r;
    } else {
        b -= 2 * nbr;
        part = b >= nbc;
        uint32_t lo = (b - part * nbc) * WS_GROUPS + g;
        if (lo < cols) { base = ld_xyzz(&C[lo]); k = lo + 1; }
        h = h_c; bits = bits_c;
    }
    const uint32_t kp = part ? k >> h : k & ((1u << h) - 1u);
    XYZZ acc = xyzz_mul_small_coop(base, kp, part ? (int)(bits - h) : (int)h);
    XYZZ tot = block_tree_sum_coop<WS_THREADS>(acc, sh);
    if (threadIdx.x == 0) st_xyzz(&partial[blockIdx.x], tot);
}

// result = 2^log_cols * (2^h_r * sum(rows high) + sum(rows low)) + 2^h_c * sum(columns high) + sum(columns low).
// One block of four 64-thread segments (16 groups each: 256 threads keep the full register budget), one per partial list; the four tree sums run side by side, then the row chain
// (segment 1) and the column chain (segment 3) run in different warps.
constexpr int RF_GROUPS = 16;               // 4-lane groups per segment
constexpr int RF_SEG = 4 * RF_GROUPS;       // threads per segment
constexpr int RF_THREADS = 4 * RF_SEG;
__global__ void __launch_bounds__(RF_THREADS) msm_reduce_final_split_kernel(const XYZZ* __restrict__ partial, uint32_t nbr,
                                                                            uint32_t nbc, uint32_t h_r, uint32_t h_c,
                                                                            uint32_t log_cols, XYZZ* __restrict__ out) {
    __shared__ XYZZ sh[4][RF_GROUPS];
    const uint32_t seg = threadIdx.x / RF_SEG, g = (threadIdx.x % RF_SEG) >> 2, r = threadIdx.x & 3;
    const uint32_t count = seg < 2 ? nbr : nbc;
    const XYZZ* src = partial + (seg == 0 ? 0 : seg == 1 ? nbr : seg == 2 ? 2 * nbr : 2 * nbr + nbc);
    XYZZ v = XYZZ::inf();
    for (uint32_t i = g; i < count; i += RF_GROUPS) { XYZZ q = ld_xyzz(&src[i]); xyzz_add_coop(v, q); }
    if (r == 0) sh[seg][g] = v;
    __syncthreads();
    for (uint32_t s = RF_GROUPS / 2; s > 0; s >>= 1) {
        if (g < s) {
            XYZZ a = sh[seg][g];
            XYZZ b = sh[seg][g + s];
            xyzz_add_coop(a, b);
            v = a;
        }
        __syncthreads();
        if (g < s && r == 0) sh[seg][g] = v;
        __syncthreads();
    }
    if (g == 0 && (seg == 1 || seg == 3)) {   // high halves: shift, then add the low halves
        XYZZ t = sh[seg][0];
        const uint32_t h = seg == 1 ? h_r : h_c;
        for (uint32_t i = 0; i < h; i++) xyzz_dbl_coop(t);
        XYZZ lo = sh[seg - 1][0];
        xyzz_add_coop(t, lo);
        if (seg == 1)
            for (uint32_t i = 0; i < log_cols; i++) xyzz_dbl_coop(t);
        v = t;
    }
    __syncthreads();
    if (g == 0 && seg == 3 && r == 0) sh[3][0] = v;
    __syncthreads();
    if (g == 0 && seg == 1) {
        XYZZ c = sh[3][0];
        xyzz_add_coop(v, c);
        if (r == 0) st_xyzz(out, v);
    }
}

// K6: result conversion.  mode 0: affine (blst_p1_affine), 1: Jacobian (blst_p1), 2: raw XYZZ
__global__ void msm_finish_kernel(const XYZZ* __restrict__ in, void* __restrict__ out, int mode) {
    if (threadIdx.x) return;
    // one block per result (gridDim.x = 1 everywhere except the batched multi-GPU combine)
    XYZZ p = ld_xyzz(in + blockIdx.x);
    if (mode == 0) {
        Affine a = xyzz_to_affine(p);
        stg_affine((Affine*)out + blockIdx.x, a);
    } else if (mode == 1) {
        Jacobian j = xyzz_to_jacobian(p);
        Jacobian* o = (Jacobian*)out + blockIdx.x;
        stg_fq(&o->x, j.x); stg_fq(&o->y, j.y); stg_fq(&o->z, j.z);
    } else {
        st_xyzz((XYZZ*)out + blockIdx.x, p);
    }
}

// sum of `count` XYZZ partial results (multi-GPU combine): block b sums in[g * rank_stride + b], g < count (one block
// per item of a batch; a single item has stride 1).  One 4-lane group per partial, cooperative tree sum: log2(count)
// cooperative additions (~7 us each) instead of count - 1 single-thread ones (~10 us each) on the critical path of
// every sharded commit.
constexpr int COMBINE_THREADS = 64;   // 16 groups = NTT_MAX_PEERS ranks in one round
__global__ void __launch_bounds__(COMBINE_THREADS) msm_combine_kernel(const XYZZ* __restrict__ in, uint32_t count,
                                                                      XYZZ* __restrict__ out, uint32_t rank_stride) {
    __shared__ XYZZ sh[COMBINE_THREADS / 4];
    const uint32_t g = threadIdx.x >> 2;
    XYZZ acc = XYZZ::inf();
    for (uint32_t r = g; r < count; r += COMBINE_THREADS / 4) {
        XYZZ q = ld_xyzz(&in[(size_t)r * rank_stride + blockIdx.x]);
        xyzz_add_coop(acc, q);
    }
    XYZZ tot = block_tree_sum_coop<COMBINE_THREADS>(acc, sh);
    if (threadIdx.x == 0) st_xyzz(out + blockIdx.x, tot);
}

// --------------------------------------------------------------------------
// SRS table precomputation: T[w][i] = 2^(c*w) * P_i, affine
// --------------------------------------------------------------------------
constexpr int MSM_MAX_WINDOWS = 32;

__global__ void __launch_bounds__(128) msm_precompute_kernel(Affine* __restrict__ table, uint32_t n, int c, int nwin) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Affine base = ldg_affine(&table[i]);
    if (base.is_inf()) {
        for (int w = 1; w < nwin; w++) stg_affine(&table[(size_t)w * n + i], Affine::inf());
        return;
    }
    Fq zarr[MSM_MAX_WINDOWS], prefix[MSM_MAX_WINDOWS];
    Jacobian q{base.x, base.y, Fq::one()};
    Fq accz = Fq::one();
    prefix[0] = accz;
    for (int w = 1; w < nwin; w++) {
        for (int d = 0; d < c; d++) q = jac_dbl(q);
        stg_affine(&table[(size_t)w * n + i], Affine{q.x, q.y});
        zarr[w] = q.z;
        accz = accz * q.z;
        prefix[w] = accz;
    }
    Fq inv = accz.inv();
    for (int w = nwin - 1; w >= 1; w--) {
        Fq zinv = inv * prefix[w - 1];
        inv = inv * zarr[w];
        Fq zinv2 = zinv.sqr();
        Affine* slot = &table[(size_t)w * n + i];
        Fq X = ldg_fq(&slot->x), Y = ldg_fq(&slot->y);
        stg_affine(slot, Affine{X * zinv2, Y * (zinv2 * zinv)});
    }
}

// Jacobian (blst_p1) -> affine, batched inversion inside each thread
constexpr int TOAFF_BATCH = 8;
__global__ void __launch_bounds__(128) jac_to_affine_kernel(const Jacobian* __restrict__ in, Affine* __restrict__ out,
                                                            size_t n) {
    size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t lo = t * TOAFF_BATCH;
    if (lo >= n) return;
    size_t cnt = min((size_t)TOAFF_BATCH, n - lo);
    Fq z[TOAFF_BATCH], prefix[TOAFF_BATCH];
    Fq acc = Fq::one();
#pragma unroll
    for (int j = 0; j < TOAFF_BATCH; j++) {
        if ((size_t)j < cnt) {
            z[j] = ldg_fq(&in[lo + j].z);
            prefix[j] = acc;
            if (!z[j].is_zero()) acc = acc * z[j];
        }
    }
    Fq inv = acc.inv();
#pragma unroll
    for (int j = TOAFF_BATCH - 1; j >= 0; j--) {
        if ((size_t)j < cnt) {
            if (z[j].is_zero()) {
                stg_affine(&out[lo + j], Affine::inf());
            } else {
                Fq zinv = inv * prefix[j];
                inv = inv * z[j];
                Fq zinv2 = zinv.sqr();
                Fq X = ldg_fq(&in[lo + j].x), Y = ldg_fq(&in[lo + j].y);
                stg_affine(&out[lo + j], Affine{X * zinv2, Y * (zinv2 * zinv)});
            }
        }
    }
}

// --------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------
int msm_choose_window(size_t n) {
    // cost model in Fq multiplications, calibrated on B200 at n = 2^20: n*W mixed adds (10 each) + 2^(c-1) buckets
    // * 30 (two full adds plus the latency-bound share of fix-up/reduction); table memory W*n*96 B is capped.
    int best = 8;
    double best_cost = 1e300;
    for (int c = 6; c <= 22; c++) {
        int w = (256 + c - 1) / c;
        if (w > MSM_MAX_WINDOWS) continue;
        double cost = (double)n * w * 10.0 + (double)(1u << (c - 1)) * 30.0;
        if ((double)n * w * 96.0 > 40e9) continue;
        if (cost < best_cost) { best_cost = cost; best = c; }
    }
    return best;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int srs_build_table(kzgb_srs* s, cudaStream_t stream) {
    // s->table[0..n) already holds the affine base points
    if (s->nwin > 1) {
        msm_precompute_kernel<<<div_up(s->n, 128), 128, 0, stream>>>(s->table, (uint32_t)s->n, s->c, s->nwin);
        KZG_LAUNCHED();
    }
    return KZGB_OK;
}

int srs_alloc(kzgb_srs** out, size_t n, int c) {
    if (n == 0 || n > ((size_t)1 << 26)) {
        set_error("srs: n = %zu out of range (1 .. 2^26)", n);
        return KZGB_ERR_INVALID_ARG;
    }
    kzgb_srs* s = new kzgb_srs();
    s->n = n;
    s->c = c > 0 ? c : msm_choose_window(n);
    if (s->c < 2) s->c = 2;
    if (s->c > 24) s->c = 24;
    s->nwin = (256 + s->c - 1) / s->c;
    if (s->nwin > MSM_MAX_WINDOWS) {
        set_error("srs: window %d gives %d windows (max %d)", s->c, s->nwin, MSM_MAX_WINDOWS);
        delete s;
        return KZGB_ERR_INVALID_ARG;
    }
    // a sorted value is (w*n + i) | sign << 31 (msm_digits_kernel): the table index must stay below 2^31
    if ((uint64_t)s->nwin * n > ((uint64_t)1 << 31)) {
        set_error("srs: %d windows x %zu points = %llu table entries exceed the 2^31 the sort payload can index; "
                  "pass a wider window_bits", s->nwin, n, (unsigned long long)s->nwin * n);
        delete s;
        return KZGB_ERR_INVALID_ARG;
    }
    cudaError_t e = cudaMalloc((void**)&s->table, (size_t)s->nwin * n * sizeof(Affine));
    if (e != cudaSuccess) {
        delete s;
        return cuda_fail(e, "cudaMalloc(srs table)", __FILE__, __LINE__);
    }
    srs_register(s);
    *out = s;
    return KZGB_OK;
}

static void pipe_destroy(MsmPipe* p) {
    if (!p) return;
    cudaStream_t all[4] = {p->heavy, p->copy, p->tail_s[0], p->tail_s[1]};
    for (cudaStream_t st : all)
        if (st) cudaStreamSynchronize(st);
    for (int k = 0; k < 2; k++) {
        if (p->handoff[k]) cudaEventDestroy(p->handoff[k]);
        if (p->done[k]) cudaEventDestroy(p->done[k]);
        if (p->staged[k]) cudaEventDestroy(p->staged[k]);
        if (p->d_stage[k]) cudaFree(p->d_stage[k]);
    }
    for (cudaStream_t st : all)
        if (st) cudaStreamDestroy(st);
    if (p->d_out) cudaFree(p->d_out);
    if (p->h_out) cudaFreeHost(p->h_out);
    delete p;
}

// Frees everything the handle holds on the device (the caller has made sure nothing is in flight on it).
void srs_release_device(kzgb_srs* s) {
    if (!s || s->dead) return;
    pipe_destroy(s->pipe);
    s->pipe = nullptr;
    if (s->table) cudaFree(s->table);
    if (s->work) cudaFree(s->work);
    for (int k = 0; k < 2; k++) {
        if (s->pipe_work[k]) cudaFree(s->pipe_work[k]);
        s->pipe_work[k] = nullptr;
        s->pipe_work_bytes[k] = 0;
    }
    s->table = nullptr;
    s->work = nullptr;
    s->work_bytes = 0;
    s->dead = true;
}

void srs_free(kzgb_srs* s) {
    if (!s) return;
    if (!s->dead) {
        srs_unregister(s);
        srs_release_device(s);
    }
    delete s;
}

int msm_pipe_get(kzgb_srs* s, MsmPipe** out) {
    if (!s->pipe) {
        MsmPipe* p = new MsmPipe();
        int lo = 0, hi = 0;  // lo = least priority (numerically greatest), hi = greatest priority
        cudaError_t e = cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&p->heavy, cudaStreamNonBlocking, lo);
        if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&p->copy, cudaStreamNonBlocking, lo);
        for (int k = 0; k < 2 && e == cudaSuccess; k++) {
            e = cudaStreamCreateWithPriority(&p->tail_s[k], cudaStreamNonBlocking, hi);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->handoff[k], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->done[k], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->staged[k], cudaEventDisableTiming);
        }
        if (e != cudaSuccess) {
            pipe_destroy(p);
            return cuda_fail(e, "batch pipeline streams/events", __FILE__, __LINE__);
        }
        s->pipe = p;
    }
    *out = s->pipe;
    return KZGB_OK;
}

// Layout of the per-SRS work area for an MSM over up to n scalars
struct MsmPlan {
    size_t total;       // n_scalars * nwin
    uint32_t nb;        // buckets
    uint32_t chunk, nthreads, nslots, fix_slots;
    size_t off_fixp0, off_fixp1, off_fixm0, off_fixm1;
    uint32_t rows, cols, log_cols, nbr, nbc;
    size_t off_keys0, off_keys1, off_vals0, off_vals1, off_zero, off_buckets, off_partials, off_meta, off_red0,
        off_red1, off_result, off_cub, cub_bytes, bytes;
};

static int msm_plan(const kzgb_srs* s, size_t m, MsmPlan& p) {
    Context& c = ctx();
    p.total = m * (size_t)s->nwin;
    p.nb = 1u << (s->c - 1);
    // threads: a whole number of waves of (SMs * 3 blocks * 128 threads)
    size_t wave = (size_t)c.sm_count * KZG_MSM_ACC_BLOCKS * KZG_MSM_ACC_THREADS;
    // entries per accumulate thread: 48 (five waves at 2^20 x 13 windows) balances the accumulate kernel best; from ~2^21
    // points on, the boundary partials (two per thread) make the fix-up the larger term and longer chunks win
    // (2^21 points: 11.14 ms at 96 vs 11.35 at 48; 2^24 points: 74.5 ms at 128 vs 75.9 at 80;
    // profiles/r02e_msm_tail_knobs.log, r02u_msm_chunk_sweep.log)
    size_t target_chunk = p.total > ((size_t)1 << 27) ? 128 : (p.total > ((size_t)1 << 24) ? 96 : 48);
    if (const char* e = getenv("KZGB_MSM_CHUNK")) target_chunk = (size_t)atoi(e) > 0 ? (size_t)atoi(e) : target_chunk;
    size_t waves = (p.total + wave * target_chunk - 1) / (wave * target_chunk);
    if (waves < 1) waves = 1;
    size_t nthreads = waves * wave;
    size_t chunk = (p.total + nthreads - 1) / nthreads;
    if (chunk < 8) {  // tiny problems: fewer threads
        chunk = 8;
        nthreads = (p.total + chunk - 1) / chunk;
    }
    p.chunk = (uint32_t)chunk;
    p.nthreads = (uint32_t)nthreads;
    p.log_cols = (uint32_t)(s->c - 1) / 2;
    p.cols = 1u << p.log_cols;
    p.rows = p.nb >> p.log_cols;
    p.nbr = (p.rows + WS_GROUPS - 1) / WS_GROUPS;
    p.nbc = (p.cols + WS_GROUPS - 1) / WS_GROUPS;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
    p.off_keys0 = take(p.total * 4);
    p.off_keys1 = take(p.total * 4);
    p.off_vals0 = take(p.total * 4);
    p.off_vals1 = take(p.total * 4);
    p.off_zero = take(8);
    p.off_buckets = take((size_t)p.nb * sizeof(XYZZ));
    p.nslots = 2 * p.nthreads;
    p.off_partials = take((size_t)p.nslots * sizeof(XYZZ));
    p.off_meta = take((size_t)p.nslots * sizeof(PartialMeta));
    p.fix_slots = 2 * (p.nslots / FIX_F_MIN + 1) + 2;
    p.off_fixp0 = take((size_t)p.fix_slots * sizeof(XYZZ));
    p.off_fixp1 = take((size_t)p.fix_slots * sizeof(XYZZ));
    p.off_fixm0 = take((size_t)p.fix_slots * sizeof(PartialMeta));
    p.off_fixm1 = take((size_t)p.fix_slots * sizeof(PartialMeta));
    p.off_red0 = take(((size_t)p.nb + 2) * sizeof(XYZZ));  // Rb (nb/2) then Cb (nb/2)
    p.off_red1 = take((size_t)2 * (p.nbr + p.nbc) * sizeof(XYZZ));   // low and high weight halves
    p.off_result = take(sizeof(XYZZ) * 2);
    p.cub_bytes = 0;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, p.cub_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                                    (const uint32_t*)nullptr, (uint32_t*)nullptr, p.total, 0, s->c);
    if (e != cudaSuccess) return cuda_fail(e, "cub SortPairs size query", __FILE__, __LINE__);
    p.off_cub = take(p.cub_bytes);
    p.bytes = o;
    return KZGB_OK;
}

// Runs the MSM of the first m table points against d_scalars (device, Montgomery Fr).
// The XYZZ result is left at *d_result_xyzz (inside the SRS work area).
int msm_run(kzgb_srs* s, const void* d_scalars, size_t m, cudaStream_t stream, const XYZZ** d_result_xyzz, int slot,
            cudaStream_t tail_stream, cudaEvent_t handoff) {
    if (m > s->n) {
        set_error("msm: %zu scalars but the SRS holds %zu points", m, s->n);
        return KZGB_ERR_INVALID_ARG;
    }
    if (s->dead) { set_error("msm: the SRS handle outlived kzgb_shutdown"); return KZGB_ERR_INVALID_ARG; }
    MsmPlan p;
    // The work area is laid out for the handle's FULL size, not for this call's m: a prover alternates between
    // commit (m = n) and create_witness (m = n - 1) on one handle, and a layout per m would free and reallocate the
    // area in the middle of a pipeline.  It is allocated once per slot and never grows afterwards.
    int st = msm_plan(s, s->n, p);
    if (st) return st;
    const size_t full_bytes = p.bytes;
    st = msm_plan(s, m ? m : 1, p);
    if (st) return st;
    void*& work = slot < 0 ? s->work : s->pipe_work[slot & 1];
    size_t& work_bytes = slot < 0 ? s->work_bytes : s->pipe_work_bytes[slot & 1];
    if (slot < 0) {   // main work area: the previous call may still be running on another stream
        if (s->have_last && s->last_stream != stream) KZG_CUDA(cudaStreamSynchronize(s->last_stream));
        s->last_stream = stream;
        s->have_last = true;
    }
    const size_t want = p.bytes > full_bytes ? p.bytes : full_bytes;
    if (want > work_bytes) {
        if (work) {   // only reachable if the layout of a smaller m exceeds the full one (alignment); drain first
            KZG_CUDA(cudaStreamSynchronize(stream));
            if (tail_stream) KZG_CUDA(cudaStreamSynchronize(tail_stream));
            cudaFree(work);
            work = nullptr;
            work_bytes = 0;
        }
        KZG_CUDA(cudaMalloc(&work, want));
        work_bytes = want;
    }
    char* w = (char*)work;
    XYZZ* result = (XYZZ*)(w + p.off_result);
    *d_result_xyzz = result;
    if (m == 0) {
        KZG_CUDA(cudaMemsetAsync(result, 0, sizeof(XYZZ), stream));
        if (tail_stream && tail_stream != stream) {
            KZG_CUDA(cudaEventRecord(handoff, stream));
            KZG_CUDA(cudaStreamWaitEvent(tail_stream, handoff, 0));
        }
        return KZGB_OK;
    }
    uint32_t* keys0 = (uint32_t*)(w + p.off_keys0);
    uint32_t* keys1 = (uint32_t*)(w + p.off_keys1);
    uint32_t* vals0 = (uint32_t*)(w + p.off_vals0);
    uint32_t* vals1 = (uint32_t*)(w + p.off_vals1);
    unsigned long long* zero = (unsigned long long*)(w + p.off_zero);
    XYZZ* buckets = (XYZZ*)(w + p.off_buckets);
    XYZZ* partials = (XYZZ*)(w + p.off_partials);
    PartialMeta* meta = (PartialMeta*)(w + p.off_meta);
    XYZZ* red0 = (XYZZ*)(w + p.off_red0);
    XYZZ* red1 = (XYZZ*)(w + p.off_red1);

    KZG_CUDA(cudaMemsetAsync(zero, 0, 8, stream));
    KZG_CUDA(cudaMemsetAsync(buckets, 0, (size_t)p.nb * sizeof(XYZZ), stream));
    stage_begin(ST_MSM_DIGITS, stream);
    {
        const unsigned g = div_up(m, 256);
        const Fr* sc = (const Fr*)d_scalars;
        if (s->c == 20) msm_digits_kernel<20><<<g, 256, 0, stream>>>(sc, (uint32_t)m, (uint32_t)s->n, s->c, s->nwin, keys0, vals0, zero);
        else if (s->c == 22) msm_digits_kernel<22><<<g, 256, 0, stream>>>(sc, (uint32_t)m, (uint32_t)s->n, s->c, s->nwin, keys0, vals0, zero);
        else msm_digits_kernel<0><<<g, 256, 0, stream>>>(sc, (uint32_t)m, (uint32_t)s->n, s->c, s->nwin, keys0, vals0, zero);
    }
    KZG_LAUNCHED();
    stage_end(ST_MSM_DIGITS, stream);
    stage_begin(ST_MSM_SORT, stream);
    size_t cub_bytes = p.cub_bytes;
    KZG_CUDA(cub::DeviceRadixSort::SortPairs(w + p.off_cub, cub_bytes, keys0, keys1, vals0, vals1, p.total, 0, s->c,
                                             stream));
    ctx().launches += 4;  // CUB onesweep: histogram + scan + passes (library kernels; approximate count)
    stage_end(ST_MSM_SORT, stream);
    stage_begin(ST_MSM_ACCUMULATE, stream);
    msm_accumulate_kernel<<<div_up(p.nthreads, KZG_MSM_ACC_THREADS), KZG_MSM_ACC_THREADS, 0, stream>>>(keys1, vals1, s->table, p.total, zero, p.chunk,
                                                                      buckets, partials, meta, p.nthreads);
    KZG_LAUNCHED();
    stage_end(ST_MSM_ACCUMULATE, stream);
    if (tail_stream && tail_stream != stream) {  // pipelined batch: the tail goes to its own (high-priority) stream
        KZG_CUDA(cudaEventRecord(handoff, stream));
        KZG_CUDA(cudaStreamWaitEvent(tail_stream, handoff, 0));
        stream = tail_stream;
    }
    stage_begin(ST_MSM_FIXUP, stream);
    {
        const XYZZ* pin = partials;
        const PartialMeta* min_ = meta;
        uint32_t nslots_in = p.nslots;
        XYZZ* pbuf[2] = {(XYZZ*)(w + p.off_fixp0), (XYZZ*)(w + p.off_fixp1)};
        PartialMeta* mbuf[2] = {(PartialMeta*)(w + p.off_fixm0), (PartialMeta*)(w + p.off_fixm1)};
        int which = 0;
        // Level 0 is dense (two partials per accumulate thread): its cost is the chain of full additions per thread.
        // Measured on B200 (profiles/r02e_msm_tail_knobs.log): 8 slots per thread (4 additions) beats 16 at every size
        // (2^20: 0.33 vs 0.39 ms, 2^21: 0.73 vs 0.77 ms); sparser levels keep 16.  Must be even: the natural pairs
        // (2a+1, 2a+2) may not straddle a range boundary at t*F + 1.
        uint32_t F0 = 8;
        if (const char* e = getenv("KZGB_MSM_FIX_F")) {
            uint32_t f = (uint32_t)atoi(e);
            f += f & 1;
            if (f >= FIX_F_MIN && f <= FIX_F_MAX) F0 = f;
        }
        for (int level = 0;; level++) {
            const uint32_t F = level == 0 ? F0 : FIX_F;
            uint32_t nth = nslots_in > 1 ? (nslots_in - 1 + F - 1) / F : 1;
            if (level == 0)
                msm_fixup_level_kernel<false><<<div_up(nth, 128), 128, 0, stream>>>(pin, min_, nslots_in, buckets,
                                                                                    pbuf[which], mbuf[which], nth, F);
            else
                msm_fixup_level_kernel<true><<<div_up((size_t)nth * 4, 128), 128, 0, stream>>>(pin, min_, nslots_in, buckets,
                                                                                         pbuf[which], mbuf[which], nth, F);
            KZG_LAUNCHED();
            if (nth == 1) break;
            pin = pbuf[which];
            min_ = mbuf[which];
            nslots_in = 2 * nth;
            which ^= 1;
        }
    }
    stage_end(ST_MSM_FIXUP, stream);
    stage_begin(ST_MSM_BUCKET_REDUCE, stream);
    {
        XYZZ* Rb = red0;
        XYZZ* Cb = red0 + (p.nb >> 1) + 1;
        const XYZZ* Rsrc = buckets;   // degenerate shapes: a single column / a single row need no summing
        const XYZZ* Csrc = buckets;
        uint32_t rstride = p.cols;
        if (p.nb >= 2 && p.cols >= 2 && p.rows >= 2) {
            msm_pairsum_first_kernel<<<div_up(p.nb, 128), 128, 0, stream>>>(buckets, p.rows, p.cols, Rb, Cb);
            KZG_LAUNCHED();
            rstride = p.cols >> 1;
            uint32_t wr = p.cols >> 1, hc = p.rows >> 1;  // live extents after level 0
            while (wr > 1 || hc > 1) {
                uint32_t nwr = wr > 1 ? wr >> 1 : 0, nhc = hc > 1 ? hc >> 1 : 0;
                uint32_t work = p.rows * nwr + nhc * p.cols;
                if (work >= 4 * MSM_COOP_THRESHOLD)
                    msm_pairsum_level_kernel<<<div_up(work, 128), 128, 0, stream>>>(Rb, p.rows, rstride, nwr, Cb, p.cols,
                                                                                    nhc);
                else
                    msm_pairsum_level_coop_kernel<<<div_up((size_t)work * 4, 128), 128, 0, stream>>>(Rb, p.rows, rstride,
                                                                                                nwr, Cb, p.cols, nhc);
                KZG_LAUNCHED();
                if (nwr) wr = nwr;
                if (nhc) hc = nhc;
            }
            Rsrc = Rb;
            Csrc = Cb;
        } else if (p.cols == 1) {
            // rows x 1: R[hi] = B[hi], C[0] = sum of all -> let the weighted kernel treat B as R only
            rstride = 1;
        }
        if (p.cols == 1 || p.rows == 1) {
            // tiny windows (c <= 2): fall back to weights on the flat bucket array: S = sum (b+1) B[b]
            msm_weighted_sum_kernel<<<(p.nb + WS_GROUPS - 1) / WS_GROUPS, WS_THREADS, 0, stream>>>(
                buckets, 0, 1, buckets, p.nb, 0, red1);
            KZG_LAUNCHED();
            msm_reduce_final_kernel<<<1, WS_THREADS, 0, stream>>>(red1, 0, (p.nb + WS_GROUPS - 1) / WS_GROUPS, 0, result);
            KZG_LAUNCHED();
        } else {
            // weights: row hi < rows (log2 rows bits), column lo + 1 <= cols (log2 cols + 1 bits); each cut in two halves
            const uint32_t bits_r = 31 - __builtin_clz(p.rows), bits_c = p.log_cols + 1;
            const uint32_t h_r = (bits_r + 1) / 2, h_c = (bits_c + 1) / 2;
            if (getenv("KZGB_MSM_UNSPLIT_WEIGHTS")) {   // the previous form, kept for A/B runs
                msm_weighted_sum_kernel<<<p.nbr + p.nbc, WS_THREADS, 0, stream>>>(Rsrc, p.rows, rstride, Csrc, p.cols, p.nbr,
                                                                                red1);
                KZG_LAUNCHED();
                msm_reduce_final_kernel<<<1, WS_THREADS, 0, stream>>>(red1, p.nbr, p.nbc, p.log_cols, result);
                KZG_LAUNCHED();
            } else {
                msm_weighted_sum_split_kernel<<<2 * (p.nbr + p.nbc), WS_THREADS, 0, stream>>>(
                    Rsrc, p.rows, rstride, Csrc, p.cols, p.nbr, p.nbc, h_r, bits_r, h_c, bits_c, red1);
                KZG_LAUNCHED();
                msm_reduce_final_split_kernel<<<1, RF_THREADS, 0, stream>>>(red1, p.nbr, p.nbc, h_r, h_c, p.log_cols, result);
                KZG_LAUNCHED();
            }
        }
    }
    stage_end(ST_MSM_BUCKET_REDUCE, stream);
    stage_begin(ST_MSM_FINISH, stream);
    return KZGB_OK;
}

int msm_finish(const XYZZ* d_xyzz, void* d_out, int mode, cudaStream_t stream, uint32_t items) {
    msm_finish_kernel<<<items, 32, 0, stream>>>(d_xyzz, d_out, mode);
    KZG_LAUNCHED();
    stage_end(ST_MSM_FINISH, stream);
    return KZGB_OK;
}

int msm_combine(const void* d_partials_xyzz, uint32_t count, void* d_out_xyzz, cudaStream_t stream, uint32_t items) {
    msm_combine_kernel<<<items, COMBINE_THREADS, 0, stream>>>((const XYZZ*)d_partials_xyzz, count, (XYZZ*)d_out_xyzz,
                                                              items > 1 ? items : 1);
    KZG_LAUNCHED();
    return KZGB_OK;
}

int jac_to_affine_dev(const void* d_jac, void* d_aff, size_t n, cudaStream_t stream) {
    size_t threads = (n + TOAFF_BATCH - 1) / TOAFF_BATCH;
    jac_to_affine_kernel<<<div_up(threads, 128), 128, 0, stream>>>((const Jacobian*)d_jac, (Affine*)d_aff, n);
    KZG_LAUNCHED();
    return KZGB_OK;
}

}  // namespace kzg
