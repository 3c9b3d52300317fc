// ntt.cu -- radix-2 number-theoretic transform over Fr for sm_100a.
//
// Replaces `best_fft` / `serial_fft` / `parallel_fft` (ft.rs:274-387) as reached
// from `EvaluationDomain::fft` (ft.rs:111-113) and `ifft` (ft.rs:115-140):
//     out[k] = sum_j a[j] * omega^(j*k),  natural order in, natural order out,
// and for the inverse omega^-1 followed by a multiplication by n^-1.
//
// Decomposition: N = R_1 * ... * R_P (P <= 3, R_j <= 2^11).  Pass j runs, for
// every value of the other digits, an R_j-point transform over digit j of the
// index (stride S_j = R_{j+1}*...*R_P) inside shared memory, then multiplies by
// the inter-pass twiddle omega_N^((N/N_j) * k_j * low).  A thread block stages a
// tile of R_j x T elements (T consecutive low indices => T*32 B contiguous
// segments in HBM).  The last pass stages T rows of R_P contiguous elements and
// writes with the digits reversed (k_1 fastest) so the result is in natural
// order; its tile is T consecutive k_1, so those writes are T*32 B contiguous
// too.  One table omega_N^i (i < N/2) per domain serves both directions
// (omega^-e = -omega^(N/2-e)), and every stage of every pass.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "field.cuh"
#include "fr_util.cuh"
#include "ntt.cuh"

namespace kzg {

constexpr int NTT_THREADS = 256;
constexpr int NTT_MAX_TILE_LOG = 11;  // 2048 elements = 64 KiB of shared memory per block

struct NttPassParams {
    uint32_t log_n;      // transform size
    uint32_t log_r;      // this pass's radix
    uint32_t log_t;      // tile width (columns / rows per block)
    uint32_t log_s;      // stride of this pass's digit (non-last passes)
    uint32_t log_nj;     // log2(R_j * S_j)
    uint32_t log_r1;     // last pass: radix of pass 1 (0 when P == 1)
    uint32_t log_r2;     // last pass: radix of pass 2 when P == 3, else 0
    uint32_t npass;      // P
    uint32_t inverse;
    uint32_t tw_shift;   // inter-pass twiddle exponent = (k_j * (low_offset + low)) << tw_shift
    uint32_t low_offset; // global index of this slab's first column (four-step, multi-GPU); 0 otherwise
    uint32_t apply_scale; // last pass: multiply by `minv`
    uint32_t swz;        // non-last passes: XOR-swizzle the shared-memory slots (see ntt_pass_kernel)
    uint32_t fold;       // inter-pass twiddles folded into the NEXT pass's butterflies (see ntt_pass_kernel)
    uint64_t batch_stride; // elements between consecutive vectors of a batch (gridDim.y vectors)
    // twiddle lookup constants (host-computed so that the kernel reads them as constant-bank operands instead of
    // rebuilding them with uniform-datapath instructions at every lookup): N - 1, N/2 - 1, log2 N - 1
    uint32_t tw_mask, tw_half_m1, tw_top_shift;
};
static void ntt_fill_tw_consts(NttPassParams& p) {
    p.tw_mask = (uint32_t)(((uint64_t)1 << p.log_n) - 1);
    p.tw_half_m1 = (uint32_t)(((uint64_t)1 << (p.log_n - 1)) - 1);
    p.tw_top_shift = p.log_n - 1;
}

KZG_DI Fr lds_fr(const Fr* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
KZG_DI void sts_fr(Fr* p, const Fr& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
KZG_DI Fr ldg_fr(const Fr* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}

// Shared-memory tile layout: the low and high 16-byte halves of element e live in separate planes
// (plane h at byte offset h * 16 * E), so consecutive lanes touch consecutive 16-byte slots: with 32-byte AoS
// elements every LDS.128/STS.128 of a warp had a 32-byte lane stride and was a 2-way bank conflict.
KZG_DI Fr sm_ld(const uint4* sm, uint32_t e, uint32_t E) {
    uint4 a = sm[e], b = sm[E + e];
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
KZG_DI void sm_st(uint4* sm, uint32_t e, uint32_t E, const Fr& v) {
    sm[e] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    sm[E + e] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

// omega_N^e (or omega_N^-e) from the half table; `neg` = the value is the negative of the returned element
KZG_DI Fr twiddle(const Fr* __restrict__ tw, uint32_t e, uint32_t log_n, uint32_t inverse, bool& neg) {
    const uint32_t mask = (log_n >= 32) ? 0xffffffffu : ((1u << log_n) - 1);
    e &= mask;
    if (inverse) e = (0u - e) & mask;
    const uint32_t half = 1u << (log_n - 1);
    neg = (e & half) != 0;
    return ldg_fr(&tw[e & (half - 1)]);
}

// The same lookup for the butterfly stages, whose exponents are already below N (e1, e2, e3 < N by construction): the
// masks are computed once per kernel, not per twiddle (the general form above cost ~25 instructions per lookup, three
// lookups per radix-4 group: 7 % of the loop body).
KZG_DI Fr twiddle_fast(const Fr* __restrict__ tw, uint32_t e, const NttPassParams& p, bool& neg) {
    if (p.inverse) e = (0u - e) & p.tw_mask;
    neg = ((e >> p.tw_top_shift) & 1u) != 0;
    return ldg_fr(&tw[e & p.tw_half_m1]);
}

// Fused four-step exchange (multi-GPU): instead of storing the column pass's result back into the local slab and
// handing it to an all-to-all, the epilogue stores element (k1, col) straight into the row matrix of the rank
// that owns row k1 -- peer memory mapped over NVLink (CUDA IPC) -- at rows[k1 mod rpr][rank*C + col], i.e. already
// transposed into the layout the row pass wants.  Each (k1, tile) is a contiguous T*32-byte peer store.
template <bool LAST, bool SCATTER>
__global__ void __launch_bounds__(NTT_THREADS, 3) ntt_pass_kernel(const Fr* in, Fr* out, const Fr* __restrict__ tw,
                                                               NttPassParams p, Fr minv, NttScatter sc) {
    extern __shared__ uint4 sm[];
    const uint32_t tid = threadIdx.x, nthr = blockDim.x;
    const uint32_t R = 1u << p.log_r, T = 1u << p.log_t;
    const uint32_t E = R * T;
    // Last pass only: consecutive lanes stage consecutive d of ONE row t, i.e. bit-reversed rows dr at the same
    // column -- every lane of the warp on the same 16-byte bank group (32-way conflict).  XOR the column with the top
    // three bits of the row: the 8 lanes of a quarter-warp then hit 8 different groups, and accesses along a row
    // (the butterfly stages, the final read) stay conflict-free because the XOR permutes each aligned group of 8.
    const uint32_t sw_shift = (LAST && p.log_r >= 3 && p.log_t >= 3) ? p.log_r - 3 : 0;
    const uint32_t sw_mask = (LAST && p.log_r >= 3 && p.log_t >= 3) ? 7u : 0u;
    // Non-last passes: a tile row holds only T = 2^log_t consecutive elements (T = 4 for the usual 2^8 x 4 tile, T = 1
    // for a 2^10-point pass), so a quarter-warp touches SEVERAL rows per 128-byte bank phase and the row pattern decides
    // the conflicts: bit-reversed rows while staging, rows 4*gg in the first radix-4 step -- all multiples of 32 or 4
    // slots, i.e. the same bank group (ncu, round 1: 3.5 M conflicts per pass).  Folding the upper slot bits onto the low
    // three (slot ^= slot>>3 ^ slot>>6 ^ slot>>9, mod 8) makes every power-of-two stride conflict-free.
    // (an experiment that changed nothing measurable; it is compiled out by default -- as a run-time flag it cost three
    // instructions in every slot computation of the hot loop: build with -DKZG_NTT_GENERIC_SWZ to get KZGB_NTT_SWZ=1 back)
#ifdef KZG_NTT_GENERIC_SWZ
    const bool gswz = !LAST && p.swz;
#else
    constexpr bool gswz = false;
#endif
    auto SW = [&](uint32_t row, uint32_t col) {
        uint32_t slot = (row << p.log_t) + (col ^ ((row >> sw_shift) & sw_mask));
        if (gswz) slot ^= ((slot >> 3) ^ (slot >> 6) ^ (slot >> 9)) & 7u;
        return slot;
    };
    const uint64_t tile = blockIdx.x;
    // batched transforms: blockIdx.y selects one of gridDim.y independent length-2^log_n vectors
    in += (size_t)blockIdx.y * p.batch_stride;
    out += (size_t)blockIdx.y * p.batch_stride;
    uint64_t base = 0;      // non-last: address of (d = 0, tl = 0)
    uint64_t low0 = 0;
    uint64_t k10 = 0, k2 = 0;
    // Folded twiddles (p.fold): instead of multiplying every output of pass j by omega_N^(k_j * low) and starting the
    // next pass with unit twiddles, pass j+1 transforms its digit as a COSET transform with the shifted root
    // g = omega_N^(S * kappa), kappa = k_1 + R_1 k_2 + ... = the output digits already fixed: stage s (m = 2^s) then uses
    //      e = j * N/(2m) + kappa * N_j/(2m),   N_j = R_j * S_j
    // as its twiddle exponent -- one table lookup as before -- and the separate multiplication disappears: N fewer
    // products per pass boundary, N/2 more because the first stage of the later passes is no longer trivial
    // (2^22 = 2^8 2^8 2^6: 10.5 N instead of 11.5 N products).  Index algebra checked against the naive DFT for every
    // 1-3 pass plan up to 2^8 in scripts/ntt_fold_model.py.
    uint32_t kap_tile = 0;  // non-last passes: kappa of the whole tile (= hi: 0 in pass 1, k_1 in pass 2 of 3)
    if (!LAST) {
        const uint64_t tiles_per_hi = (1ull << p.log_s) >> p.log_t;
        const uint64_t hi = tile / tiles_per_hi;
        if (p.fold) kap_tile = (uint32_t)hi;
        low0 = (tile % tiles_per_hi) << p.log_t;
        base = (hi << p.log_nj) + low0;
        for (uint32_t e = tid; e < E; e += nthr) {
            uint32_t d = e >> p.log_t, tl = e & (T - 1);
            uint32_t dr = __brev(d) >> (32 - p.log_r);
            if (p.log_r == 0) dr = 0;
            sm_st(sm, SW(dr, tl), E, lds_fr(&in[base + ((uint64_t)d << p.log_s) + tl]));
        }
    } else {
        if (p.npass > 1) {
            const uint64_t ntile_k1 = (1ull << p.log_r1) >> p.log_t;
            k2 = tile / ntile_k1;
            k10 = (tile % ntile_k1) << p.log_t;
        }
        const uint32_t log_s1 = p.log_n - p.log_r1;  // stride of digit 1
        for (uint32_t e = tid; e < E; e += nthr) {
            uint32_t t = e >> p.log_r, d = e & (R - 1);
            uint32_t dr = p.log_r ? (__brev(d) >> (32 - p.log_r)) : 0;
            uint64_t addr = (p.npass > 1) ? (((k10 + t) << log_s1) + (k2 << p.log_r) + d) : d;
            sm_st(sm, SW(dr, t), E, lds_fr(&in[addr]));
        }
    }
    __syncthreads();
    // DIT stages over the R dimension: one radix-2 stage when log_r is odd, then radix-4 steps (two stages per
    // trip through shared memory, the four operands of a step held in registers).
    uint32_t s = 0;
    if (p.log_r & 1) {
        for (uint32_t b = tid; b < E / 2; b += nthr) {
            uint32_t tl = b & (T - 1), bb = b >> p.log_t;
            uint32_t i0 = SW(bb << 1, tl);   // m = 1: pairs (2bb, 2bb+1), twiddle 1 (folded: g^(R/2))
            uint32_t i1 = SW((bb << 1) + 1, tl);
            Fr u = sm_ld(sm, i0, E);
            Fr t = sm_ld(sm, i1, E);
            const uint32_t kap = LAST ? (p.fold ? (uint32_t)(k10 + tl) + ((uint32_t)k2 << p.log_r1) : 0u) : kap_tile;
            const uint32_t e0 = (kap << p.log_nj) >> 1;
            bool neg = false;
            if (LAST || kap_tile != 0) {   // block-uniform test (a lane-wise `e0 != 0` diverges and saves nothing)
                Fr w = twiddle_fast(tw, e0, p, neg);
                t = t * w;
            }
            // a twiddle in the upper half of the circle is minus a table entry: the two results trade PLACES (one select
            // on the slot index instead of eight on the limbs of each value)
            sm_st(sm, neg ? i1 : i0, E, u + t);
            sm_st(sm, neg ? i0 : i1, E, u - t);
        }
        __syncthreads();
        s = 1;
    }
    for (; s < p.log_r; s += 2) {
        const uint32_t m = 1u << s;
        for (uint32_t g = tid; g < E / 4; g += nthr) {
            uint32_t tl = g & (T - 1), gg = g >> p.log_t;
            uint32_t j = gg & (m - 1);
            uint32_t k = (gg >> s) << (s + 2);
            const uint32_t i0 = SW(k + j, tl), i1 = SW(k + j + m, tl), i2 = SW(k + j + 2 * m, tl), i3 = SW(k + j + 3 * m, tl);
            Fr x0 = sm_ld(sm, i0, E), x1 = sm_ld(sm, i1, E), x2 = sm_ld(sm, i2, E), x3 = sm_ld(sm, i3, E);
            const uint32_t kap = LAST ? (p.fold ? (uint32_t)(k10 + tl) + ((uint32_t)k2 << p.log_r1) : 0u) : kap_tile;
            const uint32_t kbase = kap << p.log_nj;     // kappa * N_j  (< N)
            // A twiddle in the upper half of the circle is MINUS a table entry (twiddle() reports it as `neg`): sum and
            // difference of that butterfly trade places.  The swap is carried on INDICES -- which exponent the next
            // stage's pair uses, which slot a result is stored to -- not on the eight limbs of every value (16 selects
            // per butterfly, 9 % of the non-multiply instructions of a pass).
            // stage s: (x0,x1) and (x2,x3), twiddle w^(j * N/2m + kappa * N_j/2m)
            // Unit twiddles are skipped only where the test is uniform over the block: the first stage (m = 1, j = 0) of a
            // tile with kappa = 0 (every tile of the first pass).  Elsewhere a lane whose exponent happens to be 0
            // multiplies by table[0] = 1: exact, and cheaper than a divergent branch around the product.
            const bool unit_first = !LAST && kap_tile == 0 && s == 0;
            bool negA = false;
            if (!unit_first) {
                const uint32_t e1 = (j << (p.log_n - s - 1)) + (kbase >> (s + 1));
                Fr w = twiddle_fast(tw, e1, p, negA);
                x1 = x1 * w;
                x3 = x3 * w;
            }
            const Fr y0 = x0 + x1, y1 = x0 - x1, y2 = x2 + x3, y3 = x2 - x3;   // logical (Y0,Y1,Y2,Y3) = negA ? (y1,y0,y3,y2) : (y0,y1,y2,y3)
            // stage s+1: logical pair (Y0,Y2) with w^(j * N/4m + kappa * N_j/4m) -> slots (i0,i2); logical pair (Y1,Y3) with
            // w^((j+m) * N/4m + kappa * N_j/4m) -> slots (i1,i3).  Registers (y0,y2) are the first pair unless negA.
            const uint32_t e2 = (j << (p.log_n - s - 2)) + (kbase >> (s + 2));
            const uint32_t e3 = ((j + m) << (p.log_n - s - 2)) + (kbase >> (s + 2));
            {
                const uint32_t ep = negA ? e3 : e2;
                const uint32_t lo = negA ? i1 : i0, hi = negA ? i3 : i2;
                bool neg = false;
                Fr t = y2;
                if (!unit_first) t = t * twiddle_fast(tw, ep, p, neg);   // unit_first: negA = false, ep = e2 = 0
                sm_st(sm, neg ? hi : lo, E, y0 + t);
                sm_st(sm, neg ? lo : hi, E, y0 - t);
            }
            {
                const uint32_t ep = negA ? e2 : e3;
                const uint32_t lo = negA ? i0 : i1, hi = negA ? i2 : i3;
                bool neg = false;
                Fr t = y3 * twiddle_fast(tw, ep, p, neg);
                sm_st(sm, neg ? hi : lo, E, y1 + t);
                sm_st(sm, neg ? lo : hi, E, y1 - t);
            }
        }
        __syncthreads();
    }
    if (!LAST) {
        for (uint32_t e = tid; e < E; e += nthr) {
            uint32_t kj = e >> p.log_t, tl = e & (T - 1);
            Fr v = sm_ld(sm, SW(kj, tl), E);
            uint64_t ex = p.fold ? 0 : ((uint64_t)kj * (p.low_offset + low0 + tl)) << p.tw_shift;
            if (ex != 0) {
                bool neg;
                Fr w = twiddle(tw, (uint32_t)ex, p.log_n, p.inverse, neg);
                v = v * w;
                if (neg) v = v.neg();
            }
            Fr* dst;
            if (SCATTER) {
                const uint32_t h = kj >> sc.log_rpr, r = kj & ((1u << sc.log_rpr) - 1);
                dst = sc.peer[h] + ((((uint64_t)r << sc.log_world) + sc.rank) << sc.log_c) + low0 + tl;
            } else {
                dst = &out[base + ((uint64_t)kj << p.log_s) + tl];
            }
            sts_fr(dst, v);  // 2 x 16-byte global stores
        }
    } else {
        for (uint32_t e = tid; e < E; e += nthr) {
            uint32_t kp = e >> p.log_t, t = e & (T - 1);
            Fr v = sm_ld(sm, SW(kp, t), E);
            if (p.apply_scale) v = v * minv;
            uint64_t idx = (p.npass > 1)
                               ? ((k10 + t) + (k2 << p.log_r1) + ((uint64_t)kp << (p.log_n - p.log_r)))
                               : kp;
            sts_fr(&out[idx], v);
        }
    }
}

// First pass of a transform whose input is zero beyond its first nz entries (nz <= the pass's stride S): every
// R-point column transform then has at most ONE non-zero input (digit 0), so its R outputs are that input times the
// inter-pass twiddle, out[k*S + low] = a[low] * omega_N^(k*low) for low < nz, and zero elsewhere (the caller clears the
// destination).  Used for the vanishing polynomial Z (m + 1 coefficients in a 2^20-point transform, batched.cu).
__global__ void __launch_bounds__(256) ntt_sparse_first_pass_kernel(const Fr* __restrict__ in, Fr* __restrict__ out,
                                                                    const Fr* __restrict__ tw, uint32_t log_n, uint32_t log_r,
                                                                    uint32_t nz, uint32_t inverse, uint32_t fold) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t k = idx / nz, low = idx % nz;
    if (k >= (1u << log_r)) return;
    Fr v = ldg_fr(&in[low]);
    const uint32_t ex = fold ? 0u : k * low;  // < 2^log_r * S = N: no wrap before the mask in twiddle(); folded plans
                                              // leave the twiddle to the next pass: a plain broadcast
    if (ex != 0) {
        bool neg;
        Fr w = twiddle(tw, ex, log_n, inverse, neg);
        v = v * w;
        if (neg) v = v.neg();
    }
    sts_fr(&out[((size_t)k << (log_n - log_r)) + low], v);
}

// tw[i] = omega^i, i < count
__global__ void ntt_twiddle_kernel(Fr* __restrict__ out, Fr omega, size_t count) {
    constexpr int CH = 16;
    size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t lo = t * CH;
    if (lo >= count) return;
    Fr cur = omega.pow_u64(lo);
    size_t hi = min(lo + CH, count);
    for (size_t i = lo; i < hi; i++) {
        out[i] = cur;
        cur = cur * omega;
    }
}

// ---- host side ---------------------------------------------------------------
// Twiddle tables are a per-device cache shared by every lane and stream (Context::ntt_tw): a table is built once, on
// the requesting stream, and that stream is drained before the pointer is published, so later users on any stream
// read finished data.
int ntt_get_twiddles(uint32_t log_n, cudaStream_t stream, const Fr** out) {
    if (log_n == 0 || log_n >= 32) { *out = nullptr; return KZGB_OK; }
    Context& c = ctx();
    std::lock_guard<std::mutex> g(c.cache_mu);
    if (!c.ntt_tw[log_n]) {
        size_t count = (size_t)1 << (log_n - 1);
        Fr* t = nullptr;
        KZG_CUDA(cudaMalloc((void**)&t, count * sizeof(Fr)));
        FrHost w = fr_host_domain_omega(log_n);
        Fr dw;
        memcpy(dw.l, w.l, 32);
        ntt_twiddle_kernel<<<div_up((count + 15) / 16, 128), 128, 0, stream>>>(t, dw, count);
        c.launches++;
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) { cudaFree(t); return cuda_fail(e, "twiddle kernel", __FILE__, __LINE__); }
        c.ntt_tw[log_n] = t;
    }
    *out = (const Fr*)c.ntt_tw[log_n];
    return KZGB_OK;
}

// Experiment knob KZGB_NTT_SWZ=1: XOR-swizzled slots in the non-last passes.  Bit-exact, removes the bank conflicts ncu
// reports for those passes, and changes nothing measurable (2^22: 0.921 vs 0.914 ms, 2^20: 0.244 vs 0.248 ms;
// profiles/r02i_ntt_sweep.log): shared-memory bandwidth is not what limits the passes.  Off by default.
static uint32_t ntt_swizzle_enabled() {
    const char* e = getenv("KZGB_NTT_SWZ");
    return (e && *e == '1') ? 1u : 0u;
}

static uint32_t ntt_fold_enabled() {   // KZGB_NTT_FOLD=0: round-1 scheme (explicit inter-pass twiddle multiplication)
    const char* e = getenv("KZGB_NTT_FOLD");
    return (e && *e == '0') ? 0u : 1u;
}

struct NttPlan {
    int npass;
    uint32_t log_r[3];
};

static NttPlan ntt_plan(uint32_t log_n) {
    NttPlan pl;
    if (const char* e = getenv("KZGB_NTT_PLAN")) {  // experiment knob: "a,b[,c]" radices (must sum to log_n)
        unsigned a = 0, b = 0, c = 0;
        int k = sscanf(e, "%u,%u,%u", &a, &b, &c);
        if (k >= 1 && a + b + c == log_n && a <= 11 && b <= 11 && c <= 11) {
            pl.npass = k;
            pl.log_r[0] = a; pl.log_r[1] = b; pl.log_r[2] = c;
            return pl;
        }
    }
    // radix-4 steps cover two stages at a time, so even radices avoid the extra radix-2 stage: make every radix
    // even where the total allows it (2^22 -> 8,8,6: 0.935 ms vs 0.969 ms for 8,7,7).
    // Up to 2^20 the grid of a pass is only a wave or two of tiles, so tile count matters as much as arithmetic:
    // 512-element tiles (ntt_run) and, from 2^18 on, three passes of radix <= 2^8 instead of two of radix 2^9..2^10
    // (measured, profiles/r02i_ntt_sweep_b.log: 2^16 0.033 vs 0.055 ms, 2^18 0.074 vs 0.078, 2^19 0.130 vs 0.147,
    // 2^20 0.224 vs 0.248; from 2^21 on the round-1 plan and 1024-element tiles stay best).
    pl.log_r[0] = pl.log_r[1] = pl.log_r[2] = 0;
    if (log_n >= 18 && log_n <= 20) {
        pl.npass = 3;
        pl.log_r[0] = pl.log_r[1] = log_n == 18 ? 6 : 8;
        pl.log_r[2] = log_n - 2 * pl.log_r[0];
        return pl;
    }
    pl.npass = log_n <= 10 ? 1 : (log_n <= 20 ? 2 : 3);
    uint32_t remaining = log_n;
    for (int j = 0; j < pl.npass; j++) {
        uint32_t left = (uint32_t)(pl.npass - j);
        uint32_t r = (remaining + left - 1) / left;
        if (j < pl.npass - 1 && (r & 1) && r + 1 <= 10 && remaining - (r + 1) >= left - 1) r += 1;
        pl.log_r[j] = r;
        remaining -= r;
    }
    return pl;
}

static int ntt_set_attrs() {
    Context& c = ctx();
    std::lock_guard<std::mutex> g(c.cache_mu);
    if (!c.ntt_attr_set) {
        KZG_CUDA(cudaFuncSetAttribute(ntt_pass_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        KZG_CUDA(cudaFuncSetAttribute(ntt_pass_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        KZG_CUDA(cudaFuncSetAttribute(ntt_pass_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        c.ntt_attr_set = true;
    }
    return KZGB_OK;
}

// Four-step, step 1 (multi-GPU NTT): this rank's slab is [N1 rows][C columns] (row-major) holding
// a[i1*N2 + col_offset + c].  In place: N1-point transforms down every column, then the twiddle
// omega_N^(k1 * (col_offset + c)).  One shared-memory pass (N1 <= 2^11).
// With `scatter` != nullptr the slab is left untouched and the result goes to the peers' row matrices instead.
int ntt_columns_run(Fr* d_slab, uint32_t log_n, uint32_t log_n1, uint32_t log_c, uint32_t col_offset, bool inverse,
                    cudaStream_t stream, const NttScatter* scatter) {
    if (log_n >= 32) return KZGB_DEGREE_TOO_LARGE;
    if (log_n1 > NTT_MAX_TILE_LOG || log_n1 > log_n) {
        set_error("ntt_columns: log_n1 = %u out of range", log_n1);
        return KZGB_ERR_INVALID_ARG;
    }
    if (log_n == 0) return KZGB_OK;
    const Fr* tw = nullptr;
    int st = ntt_get_twiddles(log_n, stream, &tw);
    if (st) return st;
    st = ntt_set_attrs();
    if (st) return st;
    StageScope sc(ST_NTT, stream);
    NttPassParams p;
    p.log_n = log_n;
    p.log_r = log_n1;
    p.log_s = log_c;
    p.log_nj = log_n1 + log_c;
    p.log_r1 = p.log_r2 = 0;
    p.npass = 2;
    p.inverse = inverse ? 1 : 0;
    p.tw_shift = 0;
    p.low_offset = col_offset;
    p.apply_scale = 0;
    p.batch_stride = 0;
    ntt_fill_tw_consts(p);
    p.swz = ntt_swizzle_enabled();
    p.fold = 0;   // four-step column pass: the row pass that follows is a batch of PLAIN transforms, so the twiddle is applied here
    p.log_t = std::min<uint32_t>(NTT_MAX_TILE_LOG - p.log_r, log_c);
    size_t tiles = ((size_t)1 << log_c) >> p.log_t;
    size_t smem = (size_t)32 << (p.log_r + p.log_t);
    Fr one;
    for (int i = 0; i < 8; i++) one.l[i] = FrParams::one(i);
    // stage twiddles: omega_N1^(j*N1/2m) = omega_N^(j*N/2m), i.e. the same ladder the kernel uses for any pass
    if (scatter)
        ntt_pass_kernel<false, true><<<dim3((unsigned)tiles, 1), NTT_THREADS, smem, stream>>>(d_slab, d_slab, tw, p, one,
                                                                                              *scatter);
    else
        ntt_pass_kernel<false, false><<<dim3((unsigned)tiles, 1), NTT_THREADS, smem, stream>>>(d_slab, d_slab, tw, p, one,
                                                                                               NttScatter{});
    KZG_LAUNCHED();
    return KZGB_OK;
}

// In-place (for the caller) transform of d_data[0 .. 2^log_n); d_tmp is a same-size scratch buffer
// (unused when the plan has one pass).
int ntt_run(Fr* d_data, Fr* d_tmp, uint32_t log_n, bool inverse, cudaStream_t stream, uint32_t batch,
            const Fr* scale, uint32_t nonzero_prefix) {
    if (log_n >= 32) return KZGB_DEGREE_TOO_LARGE;  // ft.rs:67-69
    if (log_n == 0 && !scale) return KZGB_OK;        // size-1 transform is the identity (minv = 1)
    if (batch == 0) return KZGB_OK;
    if (batch > 65535) { set_error("ntt: batch %u too large", batch); return KZGB_ERR_INVALID_ARG; }
    const Fr* tw = nullptr;
    int st = ntt_get_twiddles(log_n, stream, &tw);
    if (st) return st;
    st = ntt_set_attrs();
    if (st) return st;
    NttPlan pl = ntt_plan(log_n);
    Fr minv;
    for (int i = 0; i < 8; i++) minv.l[i] = FrParams::one(i);
    bool apply_scale = inverse;
    if (scale) {
        minv = *scale;
        apply_scale = true;
    } else if (inverse) {
        FrHost mi = fr_host_inv(fr_host_from_u64(1ull << log_n));
        memcpy(minv.l, mi.l, 32);
    }
    StageScope sc(ST_NTT, stream);
    // 1024-element tiles (32 KiB) with 128 threads: six resident blocks per SM instead of three 64 KiB ones -- same
    // occupancy, but finer-grained blocks (shorter barrier phases, smaller last-wave loss): 2^22 0.94 -> 0.91 ms,
    // 2^20 0.27 -> 0.25 ms, 2^24 3.95 -> 3.77 ms.  Knob for experiments: KZGB_NTT_TILE="tile_log,threads".
    uint32_t tile_log = log_n <= 20 ? 9 : 10, threads = 128;
    if (const char* e = getenv("KZGB_NTT_TILE")) {
        unsigned a = 0, b = 0;
        if (sscanf(e, "%u,%u", &a, &b) == 2 && a >= 8 && a <= 11 && (b == 64 || b == 128 || b == 256)) { tile_log = a; threads = b; }
    }
    uint32_t log_s = log_n;
    for (int j = 0; j < pl.npass; j++) {
        NttPassParams p;
        p.log_n = log_n;
        p.log_r = pl.log_r[j];
        log_s -= p.log_r;
        p.log_s = log_s;
        p.log_nj = p.log_r + log_s;
        p.log_r1 = pl.npass > 1 ? pl.log_r[0] : 0;
        p.log_r2 = pl.npass > 2 ? pl.log_r[1] : 0;
        p.npass = pl.npass;
        p.inverse = inverse ? 1 : 0;
        p.tw_shift = log_n - p.log_nj;
        p.low_offset = 0;
        p.swz = ntt_swizzle_enabled();
        p.fold = ntt_fold_enabled();
        p.batch_stride = (uint64_t)1 << log_n;
        ntt_fill_tw_consts(p);
        bool last = (j == pl.npass - 1);
        p.apply_scale = (last && apply_scale) ? 1 : 0;
        uint32_t log_t;
        const uint32_t tl = std::max<uint32_t>(tile_log, p.log_r);
        if (!last) log_t = std::min<uint32_t>(tl - p.log_r, log_s);
        else log_t = pl.npass > 1 ? std::min<uint32_t>(tl - p.log_r, p.log_r1) : 0;
        p.log_t = log_t;
        size_t tiles = ((size_t)1 << log_n) >> (p.log_r + log_t);
        size_t smem = ((size_t)32 << (p.log_r + log_t));
        const Fr* src;
        Fr* dst;
        if (pl.npass == 1) { src = d_data; dst = d_data; }
        else if (j == 0) { src = d_data; dst = d_tmp; }
        else if (last) { src = d_tmp; dst = d_data; }
        else { src = d_tmp; dst = d_tmp; }
        if (j == 0 && !last && batch == 1 && nonzero_prefix > 0 && nonzero_prefix <= (1u << log_s)) {
            // sparse input (only the first nonzero_prefix entries are non-zero): the first pass degenerates
            KZG_CUDA(cudaMemsetAsync(dst, 0, (size_t)32 << log_n, stream));
            const uint32_t work = nonzero_prefix << p.log_r;
            ntt_sparse_first_pass_kernel<<<div_up(work, 256), 256, 0, stream>>>(src, dst, tw, log_n, p.log_r, nonzero_prefix,
                                                                                p.inverse, p.fold);
            KZG_LAUNCHED();
            continue;
        }
        dim3 grid((unsigned)tiles, batch);
        if (last) ntt_pass_kernel<true, false><<<grid, threads, smem, stream>>>(src, dst, tw, p, minv, NttScatter{});
        else ntt_pass_kernel<false, false><<<grid, threads, smem, stream>>>(src, dst, tw, p, minv, NttScatter{});
        KZG_LAUNCHED();
    }
    return KZGB_OK;
}

// Four-step, after the all-to-all: recv[g][r][c] (G blocks of rows x C) -> rows[r][g*C + c]
__global__ void ntt4_unpack_kernel(const uint4* __restrict__ recv, uint4* __restrict__ rows, uint32_t G, uint32_t nrows,
                                   uint32_t C) {
    size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;  // over uint4 halves of Fr elements
    size_t total = (size_t)G * nrows * C * 2;
    if (idx >= total) return;
    size_t e = idx >> 1, half = idx & 1;
    uint32_t c = (uint32_t)(e % C);
    size_t t = e / C;
    uint32_t r = (uint32_t)(t % nrows);
    uint32_t g = (uint32_t)(t / nrows);
    rows[(((size_t)r * G + g) * C + c) * 2 + half] = recv[idx];
}

}  // namespace kzg

using namespace kzg;

extern "C" {

KZGB_API int kzgb_ntt_fr_dev(void* d_inout, uint32_t log_n, int inverse, void* stream) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    Context& c = ctx();
    (void)c;
    if (log_n >= 32) return KZGB_DEGREE_TOO_LARGE;
    if (!d_inout) { set_error("kzgb_ntt_fr_dev: null pointer"); return KZGB_ERR_INVALID_ARG; }
    cudaStream_t st = use_stream((cudaStream_t)stream);
    size_t bytes = (size_t)32 << log_n;
    Fr* tmp = nullptr;
    if (log_n > 10) {
        tmp = (Fr*)scratch_reserve(bytes);
        if (!tmp) return KZGB_ERR_CUDA;
    }
    return ntt_run((Fr*)d_inout, tmp, log_n, inverse != 0, st);
}

// ---- four-step building blocks for the multi-GPU NTT (N >= 2^24) ----
KZGB_API int kzgb_ntt4_columns_dev(void* d_slab, uint32_t log_n, uint32_t log_n1, uint32_t log_c, uint32_t col_offset,
                                   int inverse, void* stream) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    if (!d_slab) { set_error("kzgb_ntt4_columns_dev: null pointer"); return KZGB_ERR_INVALID_ARG; }
    cudaStream_t st = use_stream((cudaStream_t)stream);
    return ntt_columns_run((Fr*)d_slab, log_n, log_n1, log_c, col_offset, inverse != 0, st);
}

// Fused column pass + exchange: like kzgb_ntt4_columns_dev, but element (k1, c) of the result is stored straight
// into peer_rows[k1 / (N1/world)] (device pointers valid on THIS device: local memory or CUDA-IPC mappings of the
// peers' row matrices, host array of `world` entries) at rows[k1 mod (N1/world)][rank*C + c].  The slab is only
// read.  The caller orders the ranks (a barrier) between this call and the row pass.
KZGB_API int kzgb_ntt4_columns_scatter_dev(const void* d_slab, void* const* peer_rows, uint32_t world, uint32_t rank,
                                           uint32_t log_n, uint32_t log_n1, uint32_t log_c, uint32_t col_offset,
                                           int inverse, void* stream) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    if (!d_slab || !peer_rows) { set_error("kzgb_ntt4_columns_scatter_dev: null pointer"); return KZGB_ERR_INVALID_ARG; }
    uint32_t lw = 0;
    while ((1u << lw) < world) lw++;
    if (world == 0 || (1u << lw) != world || world > (uint32_t)NTT_MAX_PEERS || rank >= world || lw > log_n1) {
        set_error("kzgb_ntt4_columns_scatter_dev: world = %u (power of two <= %d, <= N1), rank = %u", world, NTT_MAX_PEERS,
                  rank);
        return KZGB_ERR_INVALID_ARG;
    }
    NttScatter sc{};
    for (uint32_t i = 0; i < world; i++) {
        if (!peer_rows[i]) { set_error("kzgb_ntt4_columns_scatter_dev: peer %u is null", i); return KZGB_ERR_INVALID_ARG; }
        sc.peer[i] = (Fr*)peer_rows[i];
    }
    sc.log_rpr = log_n1 - lw;
    sc.rank = rank;
    sc.log_world = lw;
    sc.log_c = log_c;
    cudaStream_t st = use_stream((cudaStream_t)stream);
    return ntt_columns_run((Fr*)d_slab, log_n, log_n1, log_c, col_offset, inverse != 0, st, &sc);
}

// ---- peer memory for the fused exchange: plain cudaMalloc + CUDA IPC handles (one process per GPU) ----
KZGB_API int kzgb_peer_alloc(size_t bytes, void** d_ptr, unsigned char handle[64]) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    if (!d_ptr || !handle || bytes == 0) { set_error("kzgb_peer_alloc: bad argument"); return KZGB_ERR_INVALID_ARG; }
    void* p = nullptr;
    KZG_CUDA(cudaMalloc(&p, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return cuda_fail(e, "cudaIpcGetMemHandle", __FILE__, __LINE__); }
    memcpy(handle, &h, 64);
    *d_ptr = p;
    return KZGB_OK;
}
KZGB_API int kzgb_peer_open(const unsigned char handle[64], void** d_ptr) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    if (!d_ptr || !handle) { set_error("kzgb_peer_open: bad argument"); return KZGB_ERR_INVALID_ARG; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    KZG_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return KZGB_OK;
}
KZGB_API int kzgb_peer_close(void* d_ptr) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    KZG_CUDA(cudaDeviceSynchronize());
    KZG_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return KZGB_OK;
}
KZGB_API int kzgb_peer_free(void* d_ptr) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    KZG_CUDA(cudaDeviceSynchronize());
    KZG_CUDA(cudaFree(d_ptr));
    return KZGB_OK;
}

KZGB_API int kzgb_ntt4_unpack_dev(const void* d_recv, void* d_rows, uint32_t world, uint32_t nrows, uint32_t cols,
                                  void* stream) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    if (!d_recv || !d_rows) { set_error("kzgb_ntt4_unpack_dev: null pointer"); return KZGB_ERR_INVALID_ARG; }
    cudaStream_t st = use_stream((cudaStream_t)stream);
    size_t total = (size_t)world * nrows * cols * 2;
    ntt4_unpack_kernel<<<div_up(total, 256), 256, 0, st>>>((const uint4*)d_recv, (uint4*)d_rows, world, nrows, cols);
    KZG_LAUNCHED();
    return KZGB_OK;
}

// `batch` independent contiguous transforms of 2^log_n elements each (rows of the four-step matrix); when
// scale != NULL every output is multiplied by *scale (Montgomery Fr, host pointer) instead of the default
// (1 for forward, 2^-log_n for inverse).  d_tmp: batch * 2^log_n elements of scratch (unused for log_n <= 10).
KZGB_API int kzgb_ntt_batch_dev(void* d_inout, void* d_tmp, uint32_t log_n, uint32_t batch, int inverse,
                                const kzgb_fr* scale, void* stream) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    if (log_n >= 32) return KZGB_DEGREE_TOO_LARGE;
    if (!d_inout || (log_n > 10 && !d_tmp)) { set_error("kzgb_ntt_batch_dev: null pointer"); return KZGB_ERR_INVALID_ARG; }
    cudaStream_t st = use_stream((cudaStream_t)stream);
    Fr sc;
    if (scale) memcpy(sc.l, scale->l, 32);
    return ntt_run((Fr*)d_inout, (Fr*)d_tmp, log_n, inverse != 0, st, batch, scale ? &sc : nullptr);
}

KZGB_API int kzgb_ntt_fr(kzgb_fr* inout, uint32_t log_n, int inverse) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    Context& c = ctx();
    (void)c;
    use_stream(lane().stream);
    if (log_n >= 32) return KZGB_DEGREE_TOO_LARGE;
    if (!inout) { set_error("kzgb_ntt_fr: null pointer"); return KZGB_ERR_INVALID_ARG; }
    size_t bytes = (size_t)32 << log_n;
    char* d = (char*)scratch_reserve(2 * bytes);
    if (!d) return KZGB_ERR_CUDA;
    KZG_CUDA(upload_async(d, inout, bytes, lane().stream));
    int st = ntt_run((Fr*)d, (Fr*)(d + bytes), log_n, inverse != 0, lane().stream);
    if (st) return st;
    KZG_CUDA(download_async(inout, d, bytes, lane().stream));
    KZG_CUDA(cudaStreamSynchronize(lane().stream));
    return KZGB_OK;
}

}  // extern "C"
