// srs_gen.cu -- synthetic inputs generated on the device.
//
//  * monomial SRS  gs[i] = s^i * G          -- the same points `setup` builds by
//    n sequential scalar-muls (lib.rs:38-46), computed as an Fr power table +
//    fixed-base windowed scalar multiplication of the generator;
//  * Lagrange SRS  L_i(s) * G, natural omega^i order -- the same points
//    `compute_lagrange_basis` returns (eval_form.rs:254-280, O(d^3) there), via
//    the known-secret closed form L_i(s) = (s^d - 1) * w^i / (d * (s - w^i));
//  * the deterministic splitmix64 Fr stream shared with the oracle.
// A range [first, first+count) of a size-`total` SRS can be generated so that
// each rank of a multi-GPU job builds only its own shard.
#include "common.cuh"
#include "ec.cuh"
#include "msm.cuh"
#include "fr_util.cuh"
#include "ntt.cuh"

namespace kzg {

// ---- synthetic Fr stream ----------------------------------------------------
__device__ __forceinline__ uint64_t splitmix64_at(uint64_t idx) {
    uint64_t z = (idx + 1) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

__global__ void fr_stream_kernel(Fr* __restrict__ out, uint64_t seed, uint64_t start, size_t n, int mode) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t c = (seed << 40) + 4 * (start + i);
    Fr v;
    uint64_t w0 = splitmix64_at(c);
    v.l[0] = (uint32_t)w0; v.l[1] = (uint32_t)(w0 >> 32);
    if (mode == 1) {
#pragma unroll
        for (int k = 2; k < 8; k++) v.l[k] = 0;
    } else {
#pragma unroll
        for (int k = 1; k < 4; k++) {
            uint64_t w = splitmix64_at(c + k);
            v.l[2 * k] = (uint32_t)w; v.l[2 * k + 1] = (uint32_t)(w >> 32);
        }
        // 2^256 / r < 2.3: at most two subtractions
        v.final_sub();
        v.final_sub();
    }
    out[i] = v.to_mont();
}

// ---- Fr power table: out[i] = base^(first + i) ---------------------------------
constexpr int POW_CHUNK = 16;
__global__ void fr_powers_kernel(Fr* __restrict__ out, Fr base, uint64_t first, size_t n) {
    size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t lo = t * POW_CHUNK;
    if (lo >= n) return;
    Fr cur = base.pow_u64(first + lo);
    size_t hi = min(lo + POW_CHUNK, n);
    for (size_t i = lo; i < hi; i++) {
        out[i] = cur;
        cur = cur * base;
    }
}

// ---- Lagrange scalars: out[i] = (s^d - 1) * w^j / (d * (s - w^j)), j = first + i ---
constexpr int LAG_CHUNK = 8;
__global__ void lagrange_scalars_kernel(Fr* __restrict__ out, Fr s, Fr omega, Fr zs_over_d, uint64_t first, size_t n) {
    size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t lo = t * LAG_CHUNK;
    if (lo >= n) return;
    int cnt = (int)min((size_t)LAG_CHUNK, n - lo);
    Fr w = omega.pow_u64(first + lo);
    Fr wj[LAG_CHUNK], den[LAG_CHUNK], prefix[LAG_CHUNK];
    Fr acc = Fr::one();
#pragma unroll
    for (int j = 0; j < LAG_CHUNK; j++) {
        if (j < cnt) {
            wj[j] = w;
            den[j] = s - w;   // zero iff s is a d-th root of unity: then (s^d - 1) == 0 too; caller rejects
            prefix[j] = acc;
            acc = acc * den[j];
            w = w * omega;
        }
    }
    Fr inv = acc.inv();
#pragma unroll
    for (int j = LAG_CHUNK - 1; j >= 0; j--) {
        if (j < cnt) {
            Fr dinv = inv * prefix[j];
            inv = inv * den[j];
            out[lo + j] = zs_over_d * wj[j] * dinv;
        }
    }
}

// ---- fixed-base table of the generator: T[k][d] = d * 2^(8k) * G, d = 1..255 ----
__constant__ uint32_t G1_GEN_X[12] = {0xfd530c16u, 0x5cb38790u, 0x9976fff5u, 0x7817fc67u, 0x143ba1c1u, 0x154f95c7u,
                                      0xf3d0e747u, 0xf0ae6acdu, 0x21dbf440u, 0xedce6eccu, 0x9e0bfb75u, 0x12017741u};
__constant__ uint32_t G1_GEN_Y[12] = {0x0ce72271u, 0xbaac93d5u, 0x7918fd8eu, 0x8c22631au, 0x570725ceu, 0xdd595f13u,
                                      0x50405194u, 0x51ac5829u, 0xad0059c0u, 0x0e1c8c3fu, 0x5008a26au, 0x0bbc3efcu};

__global__ void gen_table_kernel(Affine* __restrict__ table /* [32][256] */) {
    int k = blockIdx.x;       // byte position
    int d = threadIdx.x;      // digit
    Affine g;
#pragma unroll
    for (int i = 0; i < 12; i++) { g.x.l[i] = G1_GEN_X[i]; g.y.l[i] = G1_GEN_Y[i]; }
    XYZZ b = XYZZ::from_affine(g);
    for (int i = 0; i < 8 * k; i++) b = xyzz_dbl(b);
    XYZZ r = xyzz_mul_small(b, (uint32_t)d);
    table[k * 256 + d] = xyzz_to_affine(r);  // d == 0 -> identity (all zero)
}

__global__ void __launch_bounds__(128) fixed_base_kernel(const Fr* __restrict__ scalars, size_t n,
                                                         const Affine* __restrict__ table, XYZZ* __restrict__ out) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr s = scalars[i].from_mont();
    XYZZ acc = XYZZ::inf();
#pragma unroll 1
    for (int k = 0; k < 32; k++) {
        uint32_t limb = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) limb = (j == (k >> 2)) ? s.l[j] : limb;
        uint32_t d = (limb >> (8 * (k & 3))) & 0xff;
        if (d) {
            Affine p = ldg_affine(&table[k * 256 + d]);
            xyzz_madd(acc, p);
        }
    }
    st_xyzz(&out[i], acc);
}

constexpr int NORM_BATCH = 8;
__global__ void __launch_bounds__(128) xyzz_to_affine_batch_kernel(const XYZZ* __restrict__ in, Affine* __restrict__ out,
                                                                   size_t n) {
    size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t lo = t * NORM_BATCH;
    if (lo >= n) return;
    int cnt = (int)min((size_t)NORM_BATCH, n - lo);
    Fq zz[NORM_BATCH], zzz[NORM_BATCH], prefix[NORM_BATCH];
    Fq acc = Fq::one();
#pragma unroll
    for (int j = 0; j < NORM_BATCH; j++) {
        if (j < cnt) {
            zz[j] = ldg_fq(&in[lo + j].zz);
            zzz[j] = ldg_fq(&in[lo + j].zzz);
            prefix[j] = acc;
            if (!zz[j].is_zero()) acc = acc * (zz[j] * zzz[j]);
        }
    }
    Fq inv = acc.inv();
#pragma unroll
    for (int j = NORM_BATCH - 1; j >= 0; j--) {
        if (j < cnt) {
            if (zz[j].is_zero()) {
                stg_affine(&out[lo + j], Affine::inf());
            } else {
                Fq tinv = inv * prefix[j];
                inv = inv * (zz[j] * zzz[j]);
                Fq X = ldg_fq(&in[lo + j].x), Y = ldg_fq(&in[lo + j].y);
                stg_affine(&out[lo + j], Affine{X * (tinv * zzz[j]), Y * (tinv * zz[j])});
            }
        }
    }
}


// --------------------------------------------------------------------------
// Secret-free Lagrange-basis SRS: the points `compute_lagrange_basis(params)` derives from the monomial SRS alone
// (eval_form.rs:254-280; O(d^3) there).  L_i(X) = (1/d) sum_j w^(-ij) X^j, hence
//     L_i(s) * G = (1/d) * sum_j w^(-ij) * gs[j]  =  inverse NTT of the vector gs OVER THE GROUP:
// radix-2 DIT butterflies (a, b) -> (a + w*b, a - w*b) on XYZZ points, w*b a variable-base scalar multiplication by a
// power of omega^-1, then a final multiplication by d^-1 and a batched conversion to affine.
// --------------------------------------------------------------------------
KZG_DI XYZZ xyzz_scalar_mul(const XYZZ& p, const Fr& k_canonical) {
    XYZZ acc = XYZZ::inf();
    bool started = false;
#pragma unroll 1
    for (int i = 254; i >= 0; i--) {
        uint32_t limb = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) limb = (j == (i >> 5)) ? k_canonical.l[j] : limb;
        bool bit = (limb >> (i & 31)) & 1;
        if (started) acc = xyzz_dbl(acc);
        if (bit) { xyzz_add(acc, p); started = true; }
    }
    return acc;
}

__global__ void group_bitrev_load_kernel(const Affine* __restrict__ in, XYZZ* __restrict__ out, uint32_t log_d) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (1u << log_d)) return;
    uint32_t r = log_d ? (__brev(i) >> (32 - log_d)) : 0;
    st_xyzz(&out[r], XYZZ::from_affine(ldg_affine(&in[i])));
}

// one DIT stage with half-size m = 2^s; tw[e] = omega^e (e < d/2), inverse twiddle omega^-e = -omega^(d/2 - e)
__global__ void __launch_bounds__(128) group_ntt_stage_kernel(XYZZ* __restrict__ a, uint32_t log_d, uint32_t s,
                                                              const Fr* __restrict__ tw) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t half_d = 1u << (log_d - 1);
    if (b >= half_d) return;
    const uint32_t m = 1u << s;
    uint32_t j = b & (m - 1);
    uint32_t k = (b >> s) << (s + 1);
    XYZZ u = ld_xyzz(&a[k + j]);
    XYZZ t = ld_xyzz(&a[k + j + m]);
    if (j != 0) {
        uint32_t e = j << (log_d - s - 1);            // exponent of omega; inverse: omega^-e = -omega^(d/2 - e)
        Fr w = tw[half_d - e].from_mont();
        t = xyzz_scalar_mul(t, w);
        t.y = t.y.neg();
    }
    XYZZ hi = u, lo = u;
    xyzz_add(hi, t);
    t.y = t.y.neg();
    xyzz_add(lo, t);
    st_xyzz(&a[k + j], hi);
    st_xyzz(&a[k + j + m], lo);
}

__global__ void __launch_bounds__(128) group_scale_kernel(XYZZ* __restrict__ a, uint32_t d, Fr k_canonical) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d) return;
    st_xyzz(&a[i], xyzz_scalar_mul(ld_xyzz(&a[i]), k_canonical));
}

// scalars (device, Montgomery Fr) -> scalars[i] * G as affine points
static int fixed_base_mul(const Fr* d_scalars, size_t n, Affine* d_out, cudaStream_t stream) {
    Affine* table = nullptr;
    XYZZ* tmp = nullptr;
    const size_t slab = (size_t)1 << 20;
    KZG_CUDA(cudaMalloc((void**)&table, 32 * 256 * sizeof(Affine)));
    cudaError_t e = cudaMalloc((void**)&tmp, std::min(n, slab) * sizeof(XYZZ));
    if (e != cudaSuccess) { cudaFree(table); return cuda_fail(e, "cudaMalloc(fixed-base tmp)", __FILE__, __LINE__); }
    gen_table_kernel<<<32, 256, 0, stream>>>(table);
    ctx().launches++;
    for (size_t off = 0; off < n; off += slab) {
        size_t cnt = std::min(slab, n - off);
        fixed_base_kernel<<<div_up(cnt, 128), 128, 0, stream>>>(d_scalars + off, cnt, table, tmp);
        ctx().launches++;
        xyzz_to_affine_batch_kernel<<<div_up((cnt + NORM_BATCH - 1) / NORM_BATCH, 128), 128, 0, stream>>>(tmp, d_out + off,
                                                                                                        cnt);
        ctx().launches++;
    }
    e = cudaStreamSynchronize(stream);
    cudaFree(table);
    cudaFree(tmp);
    if (e != cudaSuccess) return cuda_fail(e, "fixed_base_mul", __FILE__, __LINE__);
    return KZGB_OK;
}

}  // namespace kzg

using namespace kzg;

extern "C" {

KZGB_API int kzgb_fr_stream_dev(void* d_out, uint64_t seed, uint64_t start, size_t n, int mode, void* stream) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    if (!d_out && n) { set_error("kzgb_fr_stream_dev: null output"); return KZGB_ERR_INVALID_ARG; }
    if (n == 0) return KZGB_OK;
    cudaStream_t st = use_stream((cudaStream_t)stream);
    fr_stream_kernel<<<div_up(n, 256), 256, 0, st>>>((Fr*)d_out, seed, start, n, mode);
    KZG_LAUNCHED();
    return KZGB_OK;
}

// kind 0: monomial SRS s^i*G (lib.rs:38-46); kind 1: Lagrange SRS L_i(s)*G over the size-`total` domain
// (eval_form.rs:254-280; total must be a power of two).  Generates points [first, first+count).
KZGB_API int kzgb_srs_generate(int kind, const kzgb_fr* secret, size_t total, size_t first, size_t count,
                               int window_bits, kzgb_srs** out) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    Context& c = ctx();
    (void)c;
    use_stream(lane().stream);
    if (!secret || !out || count == 0 || first + count > total || (kind != 0 && kind != 1)) {
        set_error("kzgb_srs_generate: bad argument");
        return KZGB_ERR_INVALID_ARG;
    }
    Fr s;
    memcpy(s.l, secret->l, 32);
    kzgb_srs* h = nullptr;
    int st = srs_alloc(&h, count, window_bits);
    if (st) return st;
    Fr* d_sc = nullptr;
    cudaError_t e = cudaMalloc((void**)&d_sc, count * sizeof(Fr));
    if (e != cudaSuccess) { srs_free(h); return cuda_fail(e, "cudaMalloc(srs scalars)", __FILE__, __LINE__); }
    if (kind == 0) {
        fr_powers_kernel<<<div_up((count + POW_CHUNK - 1) / POW_CHUNK, 128), 128, 0, lane().stream>>>(d_sc, s, first, count);
        c.launches++;
    } else {
        if (total & (total - 1)) {
            cudaFree(d_sc); srs_free(h);
            set_error("kzgb_srs_generate: Lagrange domain size %zu is not a power of two", total);
            return KZGB_ERR_INVALID_ARG;
        }
        uint32_t log_d = 0;
        while (((size_t)1 << log_d) < total) log_d++;
        if (log_d >= 32) { cudaFree(d_sc); srs_free(h); return KZGB_DEGREE_TOO_LARGE; }
        FrHost hs = fr_host_from_device_layout(secret->l);
        FrHost omega = fr_host_domain_omega(log_d);
        FrHost zs = fr_host_sub(fr_host_pow(hs, total), fr_host_one());
        FrHost dinv = fr_host_inv(fr_host_from_u64(total));
        FrHost zs_over_d = fr_host_mul(zs, dinv);
        if (fr_host_is_zero(zs)) {
            cudaFree(d_sc); srs_free(h);
            set_error("kzgb_srs_generate: secret is a root of unity of the domain");
            return KZGB_ERR_INVALID_ARG;
        }
        Fr d_omega, d_zd;
        memcpy(d_omega.l, omega.l, 32);
        memcpy(d_zd.l, zs_over_d.l, 32);
        lagrange_scalars_kernel<<<div_up((count + LAG_CHUNK - 1) / LAG_CHUNK, 128), 128, 0, lane().stream>>>(
            d_sc, s, d_omega, d_zd, first, count);
        c.launches++;
    }
    st = fixed_base_mul(d_sc, count, h->table, lane().stream);
    cudaFree(d_sc);
    if (!st) st = srs_build_table(h, lane().stream);
    if (!st && (e = cudaStreamSynchronize(lane().stream)) != cudaSuccess) st = cuda_fail(e, "sync", __FILE__, __LINE__);
    if (st) { srs_free(h); return st; }
    *out = h;
    return KZGB_OK;
}

// Lagrange-basis SRS from a monomial SRS, without the secret (eval_form.rs:254-280 `compute_lagrange_basis`):
// out[i] = L_i(s)*G for the size-2^log_d domain, natural omega^i order, from mono.gs[0 .. 2^log_d).
KZGB_API int kzgb_srs_lagrange_from_monomial(const kzgb_srs* mono, uint32_t log_d, int window_bits, kzgb_srs** out) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    Context& c = ctx();
    (void)c;
    use_stream(lane().stream);
    if (!mono || !out) { set_error("kzgb_srs_lagrange_from_monomial: null argument"); return KZGB_ERR_INVALID_ARG; }
    if (log_d >= 32) return KZGB_DEGREE_TOO_LARGE;
    const size_t d = (size_t)1 << log_d;
    if (d > mono->n) {
        set_error("kzgb_srs_lagrange_from_monomial: domain 2^%u larger than the SRS (%zu points)", log_d, mono->n);
        return KZGB_ERR_INVALID_ARG;
    }
    KZG_LOCK_SRS(mono);
    kzgb_srs* h = nullptr;
    int st = srs_alloc(&h, d, window_bits);
    if (st) return st;
    XYZZ* work = nullptr;
    cudaError_t e = cudaMalloc((void**)&work, d * sizeof(XYZZ));
    if (e != cudaSuccess) { srs_free(h); return cuda_fail(e, "cudaMalloc(group ntt)", __FILE__, __LINE__); }
    do {
        group_bitrev_load_kernel<<<div_up(d, 256), 256, 0, lane().stream>>>(mono->table, work, log_d);
        c.launches++;
        if (log_d > 0) {
            const Fr* tw = nullptr;
            if ((st = ntt_get_twiddles(log_d, lane().stream, &tw))) break;
            for (uint32_t s = 0; s < log_d; s++) {
                group_ntt_stage_kernel<<<div_up(d / 2, 128), 128, 0, lane().stream>>>(work, log_d, s, tw);
                c.launches++;
            }
            FrHost dinv = fr_host_inv(fr_host_from_u64((uint64_t)d));
            FrHost one_canon;
            memset(one_canon.l, 0, 32);
            one_canon.l[0] = 1;
            FrHost dinv_canon = fr_host_mul(dinv, one_canon);   // out of Montgomery form
            Fr k;
            memcpy(k.l, dinv_canon.l, 32);
            group_scale_kernel<<<div_up(d, 128), 128, 0, lane().stream>>>(work, (uint32_t)d, k);
            c.launches++;
        }
        xyzz_to_affine_batch_kernel<<<div_up((d + NORM_BATCH - 1) / NORM_BATCH, 128), 128, 0, lane().stream>>>(work, h->table, d);
        c.launches++;
        if ((e = cudaGetLastError()) != cudaSuccess) { st = cuda_fail(e, "group ntt launch", __FILE__, __LINE__); break; }
        if ((st = srs_build_table(h, lane().stream))) break;
        if ((e = cudaStreamSynchronize(lane().stream)) != cudaSuccess) st = cuda_fail(e, "group ntt", __FILE__, __LINE__);
    } while (0);
    cudaFree(work);
    if (st) { srs_free(h); return st; }
    *out = h;
    return KZGB_OK;
}

}  // extern "C"
