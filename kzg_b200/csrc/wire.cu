// wire.cu -- wire/disk encodings: ZCash/IETF compressed G1 (48 bytes: big-endian x, bit 7 = compressed,
// bit 6 = infinity, bit 5 = y is the lexicographically larger root -- what blstrs `G1Affine::to_compressed` /
// `from_compressed_unchecked` use) and canonical little-endian Fr (blstrs `Scalar::to_bytes_le`).  The reference's own
// serde support is dead code (lib.rs:13, ft.rs:389-404); these are the encodings the parity checks compare and what a
// ceremony SRS file contains, so an SRS can be loaded straight into device memory.
#include "common.cuh"
#include "ec.cuh"
#include "msm.cuh"

namespace kzg {

__global__ void g1_compress_kernel(const Affine* __restrict__ pts, size_t n, uint8_t* __restrict__ out) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    Affine p = ldg_affine(&pts[i]);
    uint8_t* o = out + 48 * i;
    if (p.is_inf()) {
        o[0] = 0xc0;
        for (int k = 1; k < 48; k++) o[k] = 0;
        return;
    }
    Fq x = p.x.from_mont(), y = p.y.from_mont();
    Fq ny = p.y.neg().from_mont();
    bool y_big = false;  // y > -y  <=>  y > (p-1)/2
    for (int k = 11; k >= 0; k--) {
        if (y.l[k] != ny.l[k]) { y_big = y.l[k] > ny.l[k]; break; }
    }
    for (int k = 0; k < 12; k++) {
        uint32_t w = x.l[11 - k];
        o[4 * k] = (uint8_t)(w >> 24); o[4 * k + 1] = (uint8_t)(w >> 16); o[4 * k + 2] = (uint8_t)(w >> 8); o[4 * k + 3] = (uint8_t)w;
    }
    o[0] |= 0x80 | (y_big ? 0x20 : 0);
}

// (p + 1) / 4, little-endian 32-bit limbs: y = (x^3 + 4)^((p+1)/4) is a square root when one exists (p = 3 mod 4)
__constant__ uint32_t SQRT_EXP[12] = {0xffffeaabu, 0xee7fbfffu, 0xac54ffffu, 0x07aaffffu, 0x3dac3d89u, 0xd9cc34a8u,
                                      0x3ce144afu, 0xd91dd2e1u, 0x90d2eb35u, 0x92c6e9edu, 0x8e5ff9a6u, 0x0680447au};

// status[i]: 0 ok, 1 malformed flags / x >= p, 2 x not on the curve
__global__ void __launch_bounds__(128) g1_decompress_kernel(const uint8_t* __restrict__ in, size_t n, Affine* __restrict__ out,
                                                            int* __restrict__ bad) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* b = in + 48 * i;
    const uint8_t flags = b[0];
    if (!(flags & 0x80)) { atomicMax(bad, 1); stg_affine(&out[i], Affine::inf()); return; }  // uncompressed form not accepted
    if (flags & 0x40) {
        bool zero = (flags & 0x3f) == 0;
        for (int k = 1; k < 48; k++) zero &= b[k] == 0;
        if (!zero) atomicMax(bad, 1);
        stg_affine(&out[i], Affine::inf());
        return;
    }
    Fq x;
    for (int k = 0; k < 12; k++) {
        uint32_t w = ((uint32_t)b[4 * k] << 24) | ((uint32_t)b[4 * k + 1] << 16) | ((uint32_t)b[4 * k + 2] << 8) | b[4 * k + 3];
        if (k == 0) w &= 0x1fffffffu;
        x.l[11 - k] = w;
    }
    // x < p ?
    bool lt = false;
    for (int k = 11; k >= 0; k--) {
        uint32_t m = Fq::mod_limb(k);
        if (x.l[k] != m) { lt = x.l[k] < m; break; }
    }
    if (!lt) { atomicMax(bad, 1); stg_affine(&out[i], Affine::inf()); return; }
    x = x.to_mont();
    Fq four = Fq::zero();
    four.l[0] = 4;
    four = four.to_mont();
    Fq rhs = x.sqr() * x + four;
    Fq y = Fq::one(), base = rhs;
#pragma unroll 1
    for (int k = 0; k < 384; k++) {
        if ((SQRT_EXP[k >> 5] >> (k & 31)) & 1) y = y * base;
        base = base.sqr();
    }
    if (y.sqr() != rhs) { atomicMax(bad, 2); stg_affine(&out[i], Affine::inf()); return; }
    Fq yc = y.from_mont(), nyc = y.neg().from_mont();
    bool y_big = false;
    for (int k = 11; k >= 0; k--) {
        if (yc.l[k] != nyc.l[k]) { y_big = yc.l[k] > nyc.l[k]; break; }
    }
    if (y_big != ((flags & 0x20) != 0)) y = y.neg();
    stg_affine(&out[i], Affine{x, y});
}

// Prime-order subgroup membership: [r]P == O (r = the scalar-field modulus).  E(Fq) has order h*r with cofactor
// h = 0x396c8c005555e1568c00aaab0000aaab, so an on-curve point from an untrusted file need not lie in G1; blstrs'
// `from_compressed` performs this check, `from_compressed_unchecked` does not.  Plain double-and-add over the 255 bits
// of r: one-off import cost (~0.13 s per 2^20 points), never on the prover path.
__global__ void __launch_bounds__(128) g1_subgroup_check_kernel(const Affine* __restrict__ pts, size_t n, int* __restrict__ bad) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    Affine p = ldg_affine(&pts[i]);
    if (p.is_inf()) return;
    XYZZ acc = XYZZ::inf();
#pragma unroll 1
    for (int k = 254; k >= 0; k--) {
        acc = xyzz_dbl(acc);
        if ((Fr::mod_limb(k >> 5) >> (k & 31)) & 1) xyzz_madd(acc, p.x, p.y);
    }
    if (!acc.is_inf()) atomicMax(bad, 3);
}

// canonical little-endian bytes <-> Montgomery limbs
__global__ void fr_bytes_kernel(const Fr* __restrict__ in, Fr* __restrict__ out, size_t n, int to_bytes, int* __restrict__ bad) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr v = in[i];
    if (to_bytes) {
        out[i] = v.from_mont();
    } else {
        bool lt = false;  // must be < r (blstrs `from_bytes_le` rejects non-canonical input)
        for (int k = 7; k >= 0; k--) {
            uint32_t m = Fr::mod_limb(k);
            if (v.l[k] != m) { lt = v.l[k] < m; break; }
        }
        if (!lt) { atomicMax(bad, 1); out[i] = Fr::zero(); return; }
        out[i] = v.to_mont();
    }
}

}  // namespace kzg

using namespace kzg;

static size_t al_w(size_t v) { return (v + 255) & ~(size_t)255; }

extern "C" {

// affine points (host) -> n x 48 compressed bytes (host)
KZGB_API int kzgb_g1_compress(const kzgb_p1_affine* points, size_t n, uint8_t* out48) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    Context& c = ctx();
    (void)c;
    use_stream(lane().stream);
    if ((!points || !out48) && n) { set_error("kzgb_g1_compress: null argument"); return KZGB_ERR_INVALID_ARG; }
    if (n == 0) return KZGB_OK;
    size_t pb = al_w(n * 96);
    char* d = (char*)staging_reserve(pb + al_w(n * 48));
    if (!d) return KZGB_ERR_CUDA;
    KZG_CUDA(cudaMemcpyAsync(d, points, n * 96, cudaMemcpyHostToDevice, lane().stream));
    g1_compress_kernel<<<div_up(n, 128), 128, 0, lane().stream>>>((const Affine*)d, n, (uint8_t*)(d + pb));
    KZG_LAUNCHED();
    KZG_CUDA(cudaMemcpyAsync(out48, d + pb, n * 48, cudaMemcpyDeviceToHost, lane().stream));
    KZG_CUDA(cudaStreamSynchronize(lane().stream));
    return KZGB_OK;
}

// SRS points [start, start+count) as compressed bytes
KZGB_API int kzgb_srs_export_compressed(const kzgb_srs* s, size_t start, size_t count, uint8_t* out48) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    Context& c = ctx();
    (void)c;
    use_stream(lane().stream);
    if (!s || !out48 || start + count > s->n) { set_error("kzgb_srs_export_compressed: bad range"); return KZGB_ERR_INVALID_ARG; }
    if (count == 0) return KZGB_OK;
    KZG_LOCK_SRS(s);
    uint8_t* d = (uint8_t*)staging_reserve(al_w(count * 48));
    if (!d) return KZGB_ERR_CUDA;
    g1_compress_kernel<<<div_up(count, 128), 128, 0, lane().stream>>>(s->table + start, count, d);
    KZG_LAUNCHED();
    KZG_CUDA(cudaMemcpyAsync(out48, d, count * 48, cudaMemcpyDeviceToHost, lane().stream));
    KZG_CUDA(cudaStreamSynchronize(lane().stream));
    return KZGB_OK;
}

static int decompress_to(const uint8_t* in48, size_t n, Affine* d_out, cudaStream_t stream, bool subgroup_check = false) {
    Context& c = ctx();
    char* d = (char*)staging_reserve(al_w(n * 48) + 256);
    if (!d) return KZGB_ERR_CUDA;
    int* d_bad = (int*)(d + al_w(n * 48));
    KZG_CUDA(cudaMemcpyAsync(d, in48, n * 48, cudaMemcpyHostToDevice, stream));
    KZG_CUDA(cudaMemsetAsync(d_bad, 0, 4, stream));
    g1_decompress_kernel<<<div_up(n, 128), 128, 0, stream>>>((const uint8_t*)d, n, d_out, d_bad);
    KZG_LAUNCHED();
    if (subgroup_check) {
        g1_subgroup_check_kernel<<<div_up(n, 128), 128, 0, stream>>>(d_out, n, d_bad);
        KZG_LAUNCHED();
    }
    int bad = 0;
    KZG_CUDA(cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, stream));
    KZG_CUDA(cudaStreamSynchronize(stream));
    (void)c;
    if (bad) {
        set_error(bad == 1 ? "g1 decompress: malformed encoding (flags, or x >= p)"
                  : (bad == 2 ? "g1 decompress: x is not on the curve"
                              : "g1 decompress: a point is on the curve but not in the prime-order subgroup G1"));
        return KZGB_ERR_INVALID_ARG;
    }
    return KZGB_OK;
}

// n x 48 compressed bytes (host) -> affine points (host).  On-curve is checked; subgroup membership is not
// (like blstrs `from_compressed_unchecked`).
KZGB_API int kzgb_g1_decompress(const uint8_t* in48, size_t n, kzgb_p1_affine* out) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    Context& c = ctx();
    (void)c;
    use_stream(lane().stream);
    if ((!in48 || !out) && n) { set_error("kzgb_g1_decompress: null argument"); return KZGB_ERR_INVALID_ARG; }
    if (n == 0) return KZGB_OK;
    Affine* d_out = (Affine*)scratch_reserve(n * sizeof(Affine));
    if (!d_out) return KZGB_ERR_CUDA;
    int st = decompress_to(in48, n, d_out, lane().stream);
    if (st) return st;
    KZG_CUDA(cudaMemcpy(out, d_out, n * sizeof(Affine), cudaMemcpyDeviceToHost));
    return KZGB_OK;
}

// n affine points (host) -> *all_in_g1 = 1 iff every point is the identity or lies in the prime-order subgroup
KZGB_API int kzgb_g1_subgroup_check(const kzgb_p1_affine* points, size_t n, int* all_in_g1) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    use_stream(lane().stream);
    if ((!points && n) || !all_in_g1) { set_error("kzgb_g1_subgroup_check: null argument"); return KZGB_ERR_INVALID_ARG; }
    *all_in_g1 = 1;
    if (n == 0) return KZGB_OK;
    char* d = (char*)staging_reserve(al_w(n * sizeof(Affine)) + 256);
    if (!d) return KZGB_ERR_CUDA;
    int* d_bad = (int*)(d + al_w(n * sizeof(Affine)));
    KZG_CUDA(cudaMemcpyAsync(d, points, n * sizeof(Affine), cudaMemcpyHostToDevice, lane().stream));
    KZG_CUDA(cudaMemsetAsync(d_bad, 0, 4, lane().stream));
    g1_subgroup_check_kernel<<<div_up(n, 128), 128, 0, lane().stream>>>((const Affine*)d, n, d_bad);
    KZG_LAUNCHED();
    int bad = 0;
    KZG_CUDA(cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, lane().stream));
    KZG_CUDA(cudaStreamSynchronize(lane().stream));
    *all_in_g1 = bad ? 0 : 1;
    return KZGB_OK;
}

// SRS handle straight from a compressed-point file image.  Checked like blstrs `from_compressed`: well-formed, on the
// curve AND in the prime-order subgroup; the _unchecked variant skips the subgroup test (`from_compressed_unchecked`)
// for inputs that were validated elsewhere.
static int srs_from_compressed(const uint8_t* in48, size_t n, int window_bits, kzgb_srs** out, bool subgroup_check) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    Context& c = ctx();
    (void)c;
    use_stream(lane().stream);
    if (!in48 || !out) { set_error("kzgb_srs_create_g1_compressed: null argument"); return KZGB_ERR_INVALID_ARG; }
    kzgb_srs* h = nullptr;
    int st = srs_alloc(&h, n, window_bits);
    if (st) return st;
    st = decompress_to(in48, n, h->table, lane().stream, subgroup_check);
    if (!st) st = srs_build_table(h, lane().stream);
    if (!st) {
        cudaError_t e = cudaStreamSynchronize(lane().stream);
        if (e != cudaSuccess) st = cuda_fail(e, "sync", __FILE__, __LINE__);
    }
    if (st) { srs_free(h); return st; }
    *out = h;
    return KZGB_OK;
}
KZGB_API int kzgb_srs_create_g1_compressed(const uint8_t* in48, size_t n, int window_bits, kzgb_srs** out) {
    return srs_from_compressed(in48, n, window_bits, out, true);
}
KZGB_API int kzgb_srs_create_g1_compressed_unchecked(const uint8_t* in48, size_t n, int window_bits, kzgb_srs** out) {
    return srs_from_compressed(in48, n, window_bits, out, false);
}

// Fr vectors: Montgomery limbs -> canonical little-endian 32-byte strings (to_bytes = 1) or back (to_bytes = 0;
// non-canonical input, i.e. a value >= r, is rejected)
KZGB_API int kzgb_fr_bytes_le(const void* in, void* out, size_t n, int to_bytes) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    Context& c = ctx();
    (void)c;
    use_stream(lane().stream);
    if ((!in || !out) && n) { set_error("kzgb_fr_bytes_le: null argument"); return KZGB_ERR_INVALID_ARG; }
    if (n == 0) return KZGB_OK;
    size_t pb = al_w(n * 32);
    char* d = (char*)staging_reserve(2 * pb + 256);
    if (!d) return KZGB_ERR_CUDA;
    int* d_bad = (int*)(d + 2 * pb);
    KZG_CUDA(upload_async(d, in, n * 32, lane().stream));
    KZG_CUDA(cudaMemsetAsync(d_bad, 0, 4, lane().stream));
    fr_bytes_kernel<<<div_up(n, 256), 256, 0, lane().stream>>>((const Fr*)d, (Fr*)(d + pb), n, to_bytes, d_bad);
    KZG_LAUNCHED();
    int bad = 0;
    KZG_CUDA(download_async(out, d + pb, n * 32, lane().stream));
    KZG_CUDA(cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, lane().stream));
    KZG_CUDA(cudaStreamSynchronize(lane().stream));
    if (bad) { set_error("kzgb_fr_bytes_le: non-canonical scalar (>= r)"); return KZGB_ERR_INVALID_ARG; }
    return KZGB_OK;
}

}  // extern "C"
