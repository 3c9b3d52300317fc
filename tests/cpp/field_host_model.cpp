// field_host_model.cpp -- host-side bit-exactness check of the limb-level multiplication schedules in
// kzg_b200/csrc/field.cuh (Karatsuba wide product, ripple-free word-serial reduction, lazy a*b - c*d, dedicated
// squaring).  The routines are the SAME template code the GPU runs, compiled for the host with the PTX carry flag
// emulated (KZG_HOST_EMU); the model they are compared with is an independent 64-bit-limb CIOS Montgomery product.
// Test infrastructure only (run by tests/test_field_host_model.py); never linked into libkzg_b200.so.
#define KZG_HOST_EMU 1
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../kzg_b200/csrc/field.cuh"

// the two funnel shifts Fe::inv() (Kaliski) uses, for the host build
unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh) { sh &= 31; return sh ? (hi << sh) | (lo >> (32 - sh)) : hi; }
unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) { sh &= 31; return sh ? (lo >> sh) | (hi << (32 - sh)) : lo; }

using namespace kzg;
typedef unsigned __int128 u128;

static uint64_t rng_state = 0x9e3779b97f4a7c15ULL;
static uint64_t rnd() {  // splitmix64
    uint64_t z = (rng_state += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

template <class P>
struct Model {
    static constexpr int N = P::N, W = P::N / 2;
    static void mod64(uint64_t* m) { for (int i = 0; i < W; i++) m[i] = (uint64_t)P::mod(2 * i) | ((uint64_t)P::mod(2 * i + 1) << 32); }
    static uint64_t inv64() {  // -mod^-1 mod 2^64 by Newton iteration
        uint64_t m[W]; mod64(m);
        uint64_t x = 1;
        for (int i = 0; i < 7; i++) x *= 2 - m[0] * x;
        return 0 - x;
    }
    static bool geq(const uint64_t* a, const uint64_t* b) {
        for (int i = W - 1; i >= 0; i--) if (a[i] != b[i]) return a[i] > b[i];
        return true;
    }
    static void sub(uint64_t* a, const uint64_t* b) {
        u128 bw = 0;
        for (int i = 0; i < W; i++) { u128 t = (u128)a[i] - b[i] - (uint64_t)bw; a[i] = (uint64_t)t; bw = (t >> 64) & 1; }
    }
    // r = a*b/2^(64W) mod m (CIOS)
    static void mul(uint64_t* r, const uint64_t* a, const uint64_t* b) {
        uint64_t m[W]; mod64(m);
        const uint64_t inv = inv64();
        uint64_t t[W + 2];
        memset(t, 0, sizeof(t));
        for (int i = 0; i < W; i++) {
            u128 c = 0;
            for (int j = 0; j < W; j++) { c += (u128)a[j] * b[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
            c += t[W]; t[W] = (uint64_t)c; t[W + 1] = (uint64_t)(c >> 64);
            uint64_t q = t[0] * inv;
            c = ((u128)q * m[0] + t[0]) >> 64;
            for (int j = 1; j < W; j++) { c += (u128)q * m[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
            c += t[W]; t[W - 1] = (uint64_t)c; t[W] = t[W + 1] + (uint64_t)(c >> 64);
        }
        if (t[W] || geq(t, m)) sub(t, m);
        memcpy(r, t, 8 * W);
    }
    static void submod(uint64_t* r, const uint64_t* a, const uint64_t* b) {  // a - b mod m
        uint64_t m[W]; mod64(m);
        memcpy(r, a, 8 * W);
        if (!geq(r, b)) {  // r += m first (may wrap past 2^(64W) only transiently: do it as r - b + m in wide arithmetic)
            u128 c = 0;
            for (int i = 0; i < W; i++) { c += (u128)r[i] + m[i]; r[i] = (uint64_t)c; c >>= 64; }
        }
        sub(r, b);
    }
};

template <class P>
static void random_fe(Fe<P>& x, int kind) {
    constexpr int N = P::N;
    uint64_t m[N / 2]; Model<P>::mod64(m);
    for (;;) {
        uint64_t v[N / 2];
        for (int i = 0; i < N / 2; i++) v[i] = rnd();
        if (kind == 1) for (int i = 0; i < N / 2; i++) v[i] = 0;                         // zero
        if (kind == 2) { memcpy(v, m, sizeof(v)); v[0] -= 1; }                           // p - 1
        if (kind == 3) for (int i = N / 4; i < N / 2; i++) v[i] = v[i - N / 4];            // a0 == a1 (zero cross difference)
        if (kind == 4) for (int i = 0; i < N / 2; i++) v[i] = (rnd() & 1) ? ~0ULL : 0ULL;  // long carry runs
        if (kind == 5) { for (int i = 0; i < N / 2; i++) v[i] = 0; v[0] = 1; }            // one
        v[N / 2 - 1] &= (~0ULL) >> (64 - (Fe<P>::bits() - 64 * (N / 2 - 1)));
        if (Model<P>::geq(v, m)) { if (kind == 2 || kind == 1 || kind == 5) { } else continue; }
        if (Model<P>::geq(v, m)) continue;
        memcpy(x.l, v, sizeof(v));
        return;
    }
}

template <class P>
static int run(const char* name, int iters) {
    typedef Fe<P> F;
    constexpr int W = P::N / 2;
    int bad = 0;
    for (int it = 0; it < iters && bad < 5; it++) {
        F a, b, c, d;
        int k = it < 64 ? it % 8 : 0;
        random_fe<P>(a, k % 6); random_fe<P>(b, (k / 2) % 6); random_fe<P>(c, (it < 64) ? (it / 8) % 6 : 0); random_fe<P>(d, 0);
        if (it % 7 == 3) c = a;           // a*b - a*d and a*b - a*b (exact zero / negative differences)
        if (it % 11 == 5) { c = a; d = b; }
        uint64_t ma[W], mb[W], mc[W], md[W], ab[W], cd[W], df[W], aa[W];
        memcpy(ma, a.l, 8 * W); memcpy(mb, b.l, 8 * W); memcpy(mc, c.l, 8 * W); memcpy(md, d.l, 8 * W);
        Model<P>::mul(ab, ma, mb);
        Model<P>::mul(cd, mc, md);
        Model<P>::mul(aa, ma, ma);
        Model<P>::submod(df, ab, cd);
        F r0 = F::mul_inline(a, b);     // validates the emulation itself against the model
        F r1 = F::mul_kara(a, b);
        F r2 = F::mul_sub_mul(a, b, c, d);
        F r3 = F::sqr_inline(a);
        F r4;                                        // one interleaved pass for a*b + (MOD - c)*d: Fq only (3*MOD must fit)
        if constexpr (P::N == 12) r4 = F::mul_sub_mul_fused(a, b, c, d);
        else memcpy(r4.l, df, 8 * W);
        if (memcmp(r0.l, ab, 8 * W)) { printf("%s: mul_inline mismatch at %d\n", name, it); bad++; }
        if (memcmp(r1.l, ab, 8 * W)) { printf("%s: mul_kara mismatch at %d\n", name, it); bad++; }
        if (memcmp(r2.l, df, 8 * W)) { printf("%s: mul_sub_mul mismatch at %d\n", name, it); bad++; }
        if (memcmp(r3.l, aa, 8 * W)) { printf("%s: sqr_inline mismatch at %d\n", name, it); bad++; }
        if (memcmp(r4.l, df, 8 * W)) { printf("%s: mul_sub_mul_fused mismatch at %d\n", name, it); bad++; }
    }
    // ---- lazy range [0, 2*MOD) (Fq only): operands are canonical values with MOD added at random; every lazy routine
    // must return a value below 2*MOD that is congruent to the model's canonical result
    if constexpr (P::N == 12) {
        uint64_t m[W];
        Model<P>::mod64(m);
        auto lift = [&](F x) {   // x or x + MOD
            if (rnd() & 1) {
                u128 c = 0;
                uint64_t v[W];
                memcpy(v, x.l, 8 * W);
                for (int i = 0; i < W; i++) { c += (u128)v[i] + m[i]; v[i] = (uint64_t)c; c >>= 64; }
                memcpy(x.l, v, 8 * W);
            }
            return x;
        };
        auto below_2p = [&](const F& x) {
            uint64_t v[W], two[W];
            memcpy(v, x.l, 8 * W);
            u128 c = 0;
            for (int i = 0; i < W; i++) { c += (u128)m[i] + m[i]; two[i] = (uint64_t)c; c >>= 64; }
            return !Model<P>::geq(v, two);
        };
        for (int it = 0; it < iters && bad < 5; it++) {
            F a, b, c, d;
            random_fe<P>(a, it < 64 ? it % 6 : 0); random_fe<P>(b, it < 64 ? (it / 6) % 6 : 0); random_fe<P>(c, 0); random_fe<P>(d, 0);
            uint64_t ma[W], mb[W], mc[W], md[W], ab[W], cd[W], df[W], aa[W], amb[W], a2[W], zero[W] = {0};
            memcpy(ma, a.l, 8 * W); memcpy(mb, b.l, 8 * W); memcpy(mc, c.l, 8 * W); memcpy(md, d.l, 8 * W);
            Model<P>::mul(ab, ma, mb); Model<P>::mul(cd, mc, md); Model<P>::mul(aa, ma, ma);
            Model<P>::submod(df, ab, cd); Model<P>::submod(amb, ma, mb);
            uint64_t na[W];
            Model<P>::submod(na, zero, ma); Model<P>::submod(a2, ma, na);          // a - (-a) = 2a
            F la = lift(a), lb = lift(b), lc = lift(c), ld = lift(d);
            F r1 = F::template mul_inline<false>(la, lb);
            F r2 = F::template sqr_inline<false>(la);
            F r3 = F::template mul_add_mul<false>(la, lb, lc.neg_lazy(), ld);
            F r4 = F::sub_lazy(la, lb);
            F r5 = la.dbl_lazy();
            struct { const char* n; F v; uint64_t* want; } chk[5] = {{"mul_lazy", r1, ab}, {"sqr_lazy", r2, aa}, {"fused_lazy", r3, df},
                                                                   {"sub_lazy", r4, amb}, {"dbl_lazy", r5, a2}};
            for (auto& k : chk) {
                F cv = k.v.canon();
                if (!below_2p(k.v) || memcmp(cv.l, k.want, 8 * W)) { printf("%s: %s mismatch at %d\n", name, k.n, it); bad++; }
            }
            F z = F::sub_lazy(la, la), pm;
            memcpy(pm.l, m, 8 * W);
            if (!z.is_zero_modp() || !pm.is_zero_modp() || (!a.is_zero() && la.is_zero_modp())) { printf("%s: is_zero_modp at %d\n", name, it); bad++; }
        }
    }
    // ---- inversion: Pornin's binary GCD (inv_bingcd, an experiment) against Kaliski's almost-inverse (inv) and against x * x^-1 = 1
    {
        F one = F::one();
        uint64_t m[W];
        Model<P>::mod64(m);
        const int n_inv = iters / 20 + 200;
        for (int it = 0; it < n_inv && bad < 5; it++) {
            F a;
            random_fe<P>(a, it < 60 ? it % 6 : 0);
            if (it >= 60 && it < 60 + 32 * P::N) {          // single bits, and MOD minus single bits
                memset(a.l, 0, sizeof(a.l));
                a.l[(it - 60) / 32] = 1u << ((it - 60) % 32);
                uint64_t v[W];
                memcpy(v, a.l, 8 * W);
                if (Model<P>::geq(v, m)) continue;
                if (it & 1) { uint64_t z[W] = {0}, r[W]; Model<P>::submod(r, z, v); memcpy(a.l, r, 8 * W); }
            }
            F g = a.inv_bingcd(), k = a.inv();
            uint64_t prod[W], ma[W], mg[W];
            memcpy(ma, a.l, 8 * W); memcpy(mg, g.l, 8 * W);
            Model<P>::mul(prod, ma, mg);
            const bool ok = a.is_zero() ? g.is_zero() : (memcmp(prod, one.l, 8 * W) == 0);
            if (!ok || memcmp(g.l, k.l, 8 * W)) { printf("%s: inv (binary GCD) mismatch at %d\n", name, it); bad++; }
        }
        printf("%s: %d inversions checked, %ld fell back to Kaliski\n", name, n_inv, F::bingcd_fallbacks());
        if (F::bingcd_fallbacks()) bad++;
    }
    printf("%s: %d iterations, %d mismatches\n", name, iters, bad);
    return bad;
}

int main(int argc, char** argv) {
    int iters = argc > 1 ? atoi(argv[1]) : 200000;
    int bad = run<FqParams>("Fq", iters) + run<FrParams>("Fr", iters);
    return bad ? 1 : 0;
}
