// ec.cuh -- BLS12-381 G1 (y^2 = x^3 + 4) point arithmetic for the MSM kernels.
//
// Replaces blst's G1 add/double reached through `G1Projective::multi_exp`
// (reference call sites coeff_form.rs:61,78,102; eval_form.rs:118,136).
// Bucket accumulators use XYZZ coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2):
// mixed add 8M+2S, full add 12M+2S, double 6M+3S, no inversions.  The identity
// is ZZ == 0 (an all-zero XYZZ is the identity, so cudaMemset initialises
// bucket arrays).  Affine identity is all-zero, as in blst_p1_affine.
#pragma once
#include "field.cuh"

namespace kzg {

struct Affine {
    Fq x, y;
    KZG_DI bool is_inf() const { return x.is_zero() && y.is_zero(); }
    KZG_DI static Affine inf() { return Affine{Fq::zero(), Fq::zero()}; }
};

struct Jacobian {  // == blst_p1
    Fq x, y, z;
    KZG_DI bool is_inf() const { return z.is_zero(); }
};

struct XYZZ {
    Fq x, y, zz, zzz;
    KZG_DI bool is_inf() const { return zz.is_zero(); }
    KZG_DI static XYZZ inf() { return XYZZ{Fq::zero(), Fq::zero(), Fq::zero(), Fq::zero()}; }
    KZG_DI static XYZZ from_affine(const Affine& p) {
        if (p.is_inf()) return inf();
        return XYZZ{p.x, p.y, Fq::one(), Fq::one()};
    }
};

// ---- vectorised global-memory access (16-byte transactions) ----------------
KZG_DI Fq ldg_fq(const Fq* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    Fq r;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        uint4 v = __ldg(q + i);
        r.l[4 * i] = v.x; r.l[4 * i + 1] = v.y; r.l[4 * i + 2] = v.z; r.l[4 * i + 3] = v.w;
    }
    return r;
}
KZG_DI void stg_fq(Fq* p, const Fq& a) {
    uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
    for (int i = 0; i < 3; i++) q[i] = make_uint4(a.l[4 * i], a.l[4 * i + 1], a.l[4 * i + 2], a.l[4 * i + 3]);
}
KZG_DI Affine ldg_affine(const Affine* p) { return Affine{ldg_fq(&p->x), ldg_fq(&p->y)}; }
// streaming gather (evict-first): table points are touched once per MSM, keep them from displacing the sequentially
// read key/value lines in L1
KZG_DI Fq ldcs_fq(const Fq* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    Fq r;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        uint4 v = __ldcs(q + i);
        r.l[4 * i] = v.x; r.l[4 * i + 1] = v.y; r.l[4 * i + 2] = v.z; r.l[4 * i + 3] = v.w;
    }
    return r;
}
KZG_DI Affine ldcs_affine(const Affine* p) { return Affine{ldcs_fq(&p->x), ldcs_fq(&p->y)}; }
KZG_DI void stg_affine(Affine* p, const Affine& a) { stg_fq(&p->x, a.x); stg_fq(&p->y, a.y); }
KZG_DI XYZZ ld_xyzz(const XYZZ* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    XYZZ r;
    uint32_t* w = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int i = 0; i < 12; i++) {
        uint4 v = q[i];
        w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
    }
    return r;
}
KZG_DI void st_xyzz(XYZZ* p, const XYZZ& a) {
    uint4* q = reinterpret_cast<uint4*>(p);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(&a);
#pragma unroll
    for (int i = 0; i < 12; i++) q[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
}

// ---- XYZZ arithmetic ---------------------------------------------------------
// 2*(x,y) for an affine point (mdbl-2008-s-1)
KZG_DI XYZZ xyzz_dbl_affine(const Fq& x, const Fq& y) {
    if (y.is_zero()) return XYZZ::inf();
    Fq U = y.dbl();
    Fq V = U.sqr();
    Fq W = U * V;
    Fq S = x * V;
    Fq M = x.sqr();
    M = M.dbl() + M;
    XYZZ r;
    r.x = M.sqr() - S.dbl();
    r.y = M * (S - r.x) - W * y;
    r.zz = V;
    r.zzz = W;
    return r;
}

// acc = 2*acc (dbl-2008-s-1)
KZG_DI XYZZ xyzz_dbl(const XYZZ& p) {
    if (p.is_inf() || p.y.is_zero()) return XYZZ::inf();
    Fq U = p.y.dbl();
    Fq V = U.sqr();
    Fq W = U * V;
    Fq S = p.x * V;
    Fq M = p.x.sqr();
    M = M.dbl() + M;
    XYZZ r;
    r.x = M.sqr() - S.dbl();
    r.y = M * (S - r.x) - W * p.y;
    r.zz = V * p.zz;
    r.zzz = W * p.zzz;
    return r;
}

// acc += (x2, y2)   affine, not the identity (madd-2008-s)
KZG_DI void xyzz_madd(XYZZ& acc, const Fq& x2, const Fq& y2) {
    if (acc.is_inf()) {
        acc.x = x2; acc.y = y2; acc.zz = Fq::one(); acc.zzz = Fq::one();
        return;
    }
    Fq P = x2 * acc.zz - acc.x;
    Fq R = y2 * acc.zzz - acc.y;
    if (P.is_zero()) {
        if (R.is_zero()) acc = xyzz_dbl_affine(x2, y2);
        else acc = XYZZ::inf();
        return;
    }
    Fq PP = P.sqr();
    Fq PPP = P * PP;
    Fq Q = acc.x * PP;
    acc.zz = acc.zz * PP;
    acc.zzz = acc.zzz * PPP;
    Fq X3 = R.sqr() - PPP - Q.dbl();
    acc.y = R * (Q - X3) - acc.y * PPP;
    acc.x = X3;
}
KZG_DI void xyzz_madd(XYZZ& acc, const Affine& p) {
    if (p.is_inf()) return;
    xyzz_madd(acc, p.x, p.y);
}

// acc += q   (add-2008-s)
KZG_DI void xyzz_add(XYZZ& acc, const XYZZ& q) {
    if (q.is_inf()) return;
    if (acc.is_inf()) { acc = q; return; }
    Fq U1 = acc.x * q.zz;
    Fq S1 = acc.y * q.zzz;
    Fq P = q.x * acc.zz - U1;
    Fq R = q.y * acc.zzz - S1;
    if (P.is_zero()) {
        if (R.is_zero()) acc = xyzz_dbl(acc);
        else acc = XYZZ::inf();
        return;
    }
    Fq PP = P.sqr();
    Fq PPP = P * PP;
    Fq Q = U1 * PP;
    acc.zz = acc.zz * q.zz * PP;
    acc.zzz = acc.zzz * q.zzz * PPP;
    Fq X3 = R.sqr() - PPP - Q.dbl();
#ifdef KZG_OUTLINE_FQ_MUL
    acc.y = Fq::mul_sub_mul_fused_outlined(R, Q - X3, S1, PPP);   // one interleaved pass for both products (field.cuh)
#else
    acc.y = R * (Q - X3) - S1 * PPP;
#endif
    acc.x = X3;
}

// k * p by double-and-add, k < 2^bits (vartime)
KZG_DI XYZZ xyzz_mul_small(const XYZZ& p, uint32_t k) {
    XYZZ acc = XYZZ::inf();
    for (int i = 31 - __clz(k | 1); i >= 0; i--) {
        acc = xyzz_dbl(acc);
        if ((k >> i) & 1) xyzz_add(acc, p);
    }
    return acc;
}

// XYZZ -> affine (one inversion)
KZG_DI Affine xyzz_to_affine(const XYZZ& p) {
    if (p.is_inf()) return Affine::inf();
    Fq t = (p.zz * p.zzz).inv();
    Affine r;
    r.x = p.x * (t * p.zzz);
    r.y = p.y * (t * p.zz);
    return r;
}
// XYZZ -> Jacobian without inversion: Z = ZZ*ZZZ
KZG_DI Jacobian xyzz_to_jacobian(const XYZZ& p) {
    Jacobian r;
    if (p.is_inf()) { r.x = Fq::zero(); r.y = Fq::zero(); r.z = Fq::zero(); return r; }
    Fq z = p.zz * p.zzz;            // Z
    Fq zzz2 = p.zzz.sqr();          // ZZZ^2
    r.x = p.x * p.zz * zzz2;        // x * Z^2 = X*ZZ*ZZZ^2
    r.y = p.y * p.zz.sqr() * p.zz * zzz2;  // y * Z^3 = Y*ZZ^3*ZZZ^2
    r.z = z;
    return r;
}

// ---- Jacobian (blst_p1 interop, table precomputation) ------------------------
KZG_DI Jacobian jac_dbl(const Jacobian& p) {  // dbl-2009-l, a = 0
    if (p.is_inf()) return p;
    Fq A = p.x.sqr();
    Fq B = p.y.sqr();
    Fq C = B.sqr();
    Fq t = p.x + B;
    Fq D = (t.sqr() - A - C).dbl();
    Fq E = A.dbl() + A;
    Fq F = E.sqr();
    Jacobian r;
    r.x = F - D.dbl();
    r.z = (p.y * p.z).dbl();
    Fq C8 = C.dbl().dbl().dbl();
    r.y = E * (D - r.x) - C8;
    return r;
}
KZG_DI XYZZ jac_to_xyzz(const Jacobian& p) {
    if (p.is_inf()) return XYZZ::inf();
    XYZZ r;
    r.x = p.x; r.y = p.y;
    r.zz = p.z.sqr();
    r.zzz = r.zz * p.z;
    return r;
}

}  // namespace kzg
