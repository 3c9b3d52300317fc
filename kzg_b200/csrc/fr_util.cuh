// fr_util.cuh -- host-side Fr scalars for domain parameters (omega, omega^-1,
// n^-1, s^d ...): the handful of field elements the reference derives when it
// builds an EvaluationDomain (ft.rs:55-109).  Montgomery form, same limb image
// as the device Fr / blst_fr.  Not a compute path: O(log n) multiplications.
#pragma once
#include <cstdint>
#include <cstring>

namespace kzg {

struct FrHost {
    uint32_t l[8];  // device layout (8 x u32 LE == 4 x u64 LE on little-endian hosts)
};

namespace frh {
typedef unsigned __int128 u128;
static const uint64_t MOD[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL,
                                0x73eda753299d7d48ULL};
static const uint64_t INV = 0xfffffffeffffffffULL;
static const uint64_t ONE[4] = {0x00000001fffffffeULL, 0x5884b7fa00034802ULL, 0x998c4fefecbc4ff5ULL,
                                0x1824b159acc5056fULL};
static const uint64_t R2[4] = {0xc999e990f3f29c6dULL, 0x2b6cedcb87925c23ULL, 0x05d314967254398fULL,
                               0x0748d9d99f59ff11ULL};
// 2^32-th primitive root of unity 7^((r-1)/2^32) (Scalar::root_of_unity(), ft.rs:73), Montgomery form
static const uint64_t ROOT[4] = {0xb9b58d8c5f0e466aULL, 0x5b1b4c801819d7ecULL, 0x0af53ae352a31e64ULL,
                                 0x5bf3adda19e9b27bULL};

static inline void load(uint64_t* o, const FrHost& a) { memcpy(o, a.l, 32); }
static inline FrHost store(const uint64_t* a) { FrHost r; memcpy(r.l, a, 32); return r; }
static inline bool geq(const uint64_t* a, const uint64_t* b) {
    for (int i = 3; i >= 0; i--) {
        if (a[i] != b[i]) return a[i] > b[i];
    }
    return true;
}
static inline void sub_mod_raw(uint64_t* a) {  // a -= MOD
    u128 bw = 0;
    for (int i = 0; i < 4; i++) {
        u128 t = (u128)a[i] - MOD[i] - (uint64_t)bw;
        a[i] = (uint64_t)t;
        bw = (t >> 64) & 1;
    }
}
static inline void mul(uint64_t* r, const uint64_t* a, const uint64_t* b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)a[j] * b[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        uint64_t q = t[0] * INV;
        c = ((u128)q * MOD[0] + t[0]) >> 64;
        for (int j = 1; j < 4; j++) {
            c += (u128)q * MOD[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    if (t[4] || geq(t, MOD)) sub_mod_raw(t);
    memcpy(r, t, 32);
}
}  // namespace frh

static inline FrHost fr_host_from_device_layout(const uint64_t* limbs) { return frh::store(limbs); }
static inline FrHost fr_host_one() { return frh::store(frh::ONE); }
static inline bool fr_host_is_zero(const FrHost& a) {
    uint32_t v = 0;
    for (int i = 0; i < 8; i++) v |= a.l[i];
    return v == 0;
}
static inline FrHost fr_host_mul(const FrHost& a, const FrHost& b) {
    uint64_t x[4], y[4], r[4];
    frh::load(x, a); frh::load(y, b);
    frh::mul(r, x, y);
    return frh::store(r);
}
static inline FrHost fr_host_sub(const FrHost& a, const FrHost& b) {
    uint64_t x[4], y[4], r[4];
    frh::load(x, a); frh::load(y, b);
    frh::u128 bw = 0;
    for (int i = 0; i < 4; i++) {
        frh::u128 t = (frh::u128)x[i] - y[i] - (uint64_t)bw;
        r[i] = (uint64_t)t;
        bw = (t >> 64) & 1;
    }
    if (bw) {
        frh::u128 c = 0;
        for (int i = 0; i < 4; i++) {
            c += (frh::u128)r[i] + frh::MOD[i];
            r[i] = (uint64_t)c;
            c >>= 64;
        }
    }
    return frh::store(r);
}
static inline FrHost fr_host_neg(const FrHost& a) {
    FrHost z;
    memset(z.l, 0, 32);
    return fr_host_sub(z, a);
}
static inline FrHost fr_host_from_u64(uint64_t v) {
    uint64_t x[4] = {v, 0, 0, 0}, r[4];
    frh::mul(r, x, frh::R2);
    return frh::store(r);
}
static inline FrHost fr_host_pow(FrHost base, uint64_t e) {
    FrHost acc = fr_host_one();
    while (e) {
        if (e & 1) acc = fr_host_mul(acc, base);
        base = fr_host_mul(base, base);
        e >>= 1;
    }
    return acc;
}
static inline FrHost fr_host_inv(const FrHost& a) {  // a^(r-2)
    uint64_t e[4];
    memcpy(e, frh::MOD, 32);
    e[0] -= 2;
    FrHost acc = fr_host_one(), base = a;
    for (int i = 0; i < 256; i++) {
        if ((e[i >> 6] >> (i & 63)) & 1) acc = fr_host_mul(acc, base);
        base = fr_host_mul(base, base);
    }
    return acc;
}
// omega of the 2^exp domain: root_of_unity^(2^(32-exp))  (ft.rs:73)
static inline FrHost fr_host_domain_omega(uint32_t exp) {
    FrHost w = frh::store(frh::ROOT);
    for (uint32_t i = exp; i < 32; i++) w = fr_host_mul(w, w);
    return w;
}

}  // namespace kzg
