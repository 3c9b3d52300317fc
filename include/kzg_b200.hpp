// kzg_b200.hpp -- header-only C++17 host mirror of the reference's prover API over the C ABI (kzg_b200.h).
//
// The reference is a Rust crate; where no Rust toolchain exists this is the compiled-language host side of the
// drop-in boundary.  Names, argument meaning and error behaviour follow the reference:
//   KZGParams / setup            lib.rs:12-55        Polynomial            polynomial.rs:24-165
//   KZGProver                    coeff_form.rs:38-112 EvaluationDomain      ft.rs:17-140
//   KZGProverEvalForm            eval_form.rs:40-147
// Scalars are blst_fr images (Montgomery), points blst_p1 / blst_p1_affine.  Failures that the reference turns into
// panics (slice out of range, failed asserts, CUDA errors here) throw std::runtime_error; KZGError variants are
// returned through kzg::Result-like optionals exactly where the reference returns Result.
#pragma once
#include <cstring>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "kzg_b200.h"

namespace kzg_b200 {

using Scalar = kzgb_fr;
using G1Affine = kzgb_p1_affine;
using G1Projective = kzgb_p1;

enum class KZGError { NoPolynomial, PointNotOnPolynomial, BatchOpeningZeroRemainder, PolynomialDegreeTooLarge };  // lib.rs:26-36

inline void check(int status, const char* what) {
    if (status < 0) throw std::runtime_error(std::string(what) + ": " + kzgb_last_error_message());
}
inline bool is_zero(const Scalar& s) { return (s.l[0] | s.l[1] | s.l[2] | s.l[3]) == 0; }

// polynomial.rs:24-27 -- `degree` is authoritative: hot-path readers use coeffs[..degree+1]
struct Polynomial {
    size_t degree = 0;
    std::vector<Scalar> coeffs;

    static size_t compute_degree(const std::vector<Scalar>& c, size_t upper_bound) {  // polynomial.rs:94-105
        size_t i = upper_bound;
        for (;;) {
            if (i == 0) return 0;
            if (!is_zero(c[i])) return i;
            i--;
        }
    }
    static Polynomial new_from_coeffs(std::vector<Scalar> c, size_t degree) { return Polynomial{degree, std::move(c)}; }
    static Polynomial make(std::vector<Scalar> c) {  // Polynomial::new, polynomial.rs:83-87
        size_t d = compute_degree(c, c.size() - 1);
        return Polynomial{d, std::move(c)};
    }
    size_t num_coeffs() const { return degree + 1; }
    const Scalar* slice_coeffs() const { return coeffs.data(); }  // first num_coeffs() entries
    bool is_zero_poly() const { return degree == 0 && is_zero(coeffs[0]); }
    void shrink_degree() { degree = compute_degree(coeffs, degree); }
};

// lib.rs:12-19 (G1 half).  Owns the device-resident SRS; freed with the last copy.
class KZGParams {
public:
    explicit KZGParams(const std::vector<G1Projective>& gs, int window_bits = 0) {
        kzgb_srs* h = nullptr;
        check(kzgb_srs_create_g1(gs.data(), gs.size(), window_bits, &h), "kzgb_srs_create_g1");
        srs_.reset(h, [](kzgb_srs* p) { kzgb_srs_destroy(p); });
    }
    // lib.rs:38-55 `setup`: gs[i] = s^i * G (same points, generated on the device)
    static KZGParams setup(const Scalar& s, size_t num_coeffs, int window_bits = 0) {
        kzgb_srs* h = nullptr;
        check(kzgb_srs_generate(0, &s, num_coeffs, 0, num_coeffs, window_bits, &h), "kzgb_srs_generate");
        return KZGParams(h);
    }
    size_t len() const { return kzgb_srs_size(srs_.get()); }
    kzgb_srs* handle() const { return srs_.get(); }

private:
    explicit KZGParams(kzgb_srs* h) : srs_(h, [](kzgb_srs* p) { kzgb_srs_destroy(p); }) {}
    std::shared_ptr<kzgb_srs> srs_;
};

struct KZGBatchWitness {  // coeff_form.rs:14-35
    Polynomial r;
    G1Affine w;
};

template <class T>
struct Result {
    std::optional<T> value;
    KZGError error{};
    bool ok() const { return value.has_value(); }
};

class KZGProver {  // coeff_form.rs:38-112
public:
    explicit KZGProver(const KZGParams& parameters) : parameters_(parameters) {}
    const KZGParams& parameters() const { return parameters_; }

    G1Affine commit(const Polynomial& p) const {  // coeff_form.rs:59-64
        if (p.num_coeffs() > parameters_.len()) throw std::runtime_error("range end index out of range for slice");
        G1Affine out;
        check(kzgb_msm_g1(parameters_.handle(), p.slice_coeffs(), p.num_coeffs(), &out), "kzgb_msm_g1");
        return out;
    }
    // commit (coeff_form.rs:59-64) of several polynomials in one pipelined call (kzgb_msm_g1_batch): same results as a
    // loop of commit(); the tail of one MSM overlaps the copy / sort / accumulation of the next
    std::vector<G1Affine> commit_many(const std::vector<const Polynomial*>& ps) const {
        std::vector<const kzgb_fr*> ptrs;
        std::vector<size_t> lens;
        for (const Polynomial* p : ps) {
            if (p->num_coeffs() > parameters_.len()) throw std::runtime_error("range end index out of range for slice");
            ptrs.push_back(p->slice_coeffs());
            lens.push_back(p->num_coeffs());
        }
        std::vector<G1Affine> out(ps.size());
        check(kzgb_msm_g1_batch(parameters_.handle(), ptrs.data(), lens.data(), ps.size(), out.data()), "kzgb_msm_g1_batch");
        return out;
    }
    Result<G1Affine> create_witness(const Polynomial& p, const Scalar& x, const Scalar& y) const {  // coeff_form.rs:66-81
        G1Affine out;
        int st = kzgb_witness_coeff(parameters_.handle(), p.slice_coeffs(), p.num_coeffs(), &x, &y, &out);
        check(st, "kzgb_witness_coeff");
        if (st == KZGB_POINT_NOT_ON_POLYNOMIAL) return {std::nullopt, KZGError::PointNotOnPolynomial};
        return {out, {}};
    }
    // create_witness at every (xs[i], ys[i]) in one pipelined call (kzgb_witness_coeff_many); like collecting the
    // per-point Results: the first point off the polynomial makes the whole call PointNotOnPolynomial
    Result<std::vector<G1Affine>> create_witness_many(const Polynomial& p, const std::vector<Scalar>& xs,
                                                      const std::vector<Scalar>& ys) const {
        if (xs.size() != ys.size()) throw std::runtime_error("xs and ys differ in length");
        std::vector<G1Affine> out(xs.size());
        std::vector<int> status(xs.size() ? xs.size() : 1);
        check(kzgb_witness_coeff_many(parameters_.handle(), p.slice_coeffs(), p.num_coeffs(), xs.data(), ys.data(), xs.size(),
                                      out.data(), status.data()),
              "kzgb_witness_coeff_many");
        for (size_t i = 0; i < xs.size(); i++)
            if (status[i] == KZGB_POINT_NOT_ON_POLYNOMIAL) return {std::nullopt, KZGError::PointNotOnPolynomial};
        return {out, {}};
    }
    Result<KZGBatchWitness> create_witness_batched(const Polynomial& p, const std::vector<Scalar>& xs,
                                                   const std::vector<Scalar>& ys) const {  // coeff_form.rs:83-111
        if (xs.size() != ys.size()) throw std::runtime_error("assertion failed: xs.len() == ys.len()");
        std::vector<Scalar> r(xs.size() < 2 ? 2 : xs.size());
        size_t r_len = 0;
        G1Affine w;
        int st = kzgb_witness_batched(parameters_.handle(), p.slice_coeffs(), p.num_coeffs(), xs.data(), ys.data(),
                                      xs.size(), r.data(), &r_len, &w);
        check(st, "kzgb_witness_batched");
        if (st == KZGB_POINT_NOT_ON_POLYNOMIAL) return {std::nullopt, KZGError::PointNotOnPolynomial};
        r.resize(r_len);
        // the shape the reference's tree recursion gives r (polynomial.rs:303-365): from 258 points on, the top-level
        // best_mul takes the fft_mul path (polynomial.rs:167-191): vector padded to 2^k >= m + 1, degree re-measured
        if (xs.size() >= 258) {
            size_t size = 1;
            while (size < xs.size() + 1) size *= 2;
            r.resize(size, Scalar{});
            size_t deg = r_len - 1;
            auto is_zero = [](const Scalar& v) { return (v.l[0] | v.l[1] | v.l[2] | v.l[3]) == 0; };
            while (deg > 0 && is_zero(r[deg])) deg--;
            return {KZGBatchWitness{Polynomial::new_from_coeffs(std::move(r), deg), w}, {}};
        }
        return {KZGBatchWitness{Polynomial::new_from_coeffs(std::move(r), r_len - 1), w}, {}};
    }

private:
    KZGParams parameters_;
};

// ft.rs:17-25 (the fields the hot path reads) with fft / ifft on the device
struct EvaluationDomain {
    std::vector<Scalar> coeffs;
    size_t d = 1;
    uint32_t exp = 0;

    static Result<EvaluationDomain> from_coeffs(std::vector<Scalar> c) {  // ft.rs:94-109
        size_t m = 1;
        uint32_t exp = 0;
        while (m < c.size()) {
            m *= 2;
            exp += 1;
            if (exp >= 32) return {std::nullopt, KZGError::PolynomialDegreeTooLarge};  // ft.rs:67-69
        }
        Scalar zero{};
        c.resize(m, zero);
        return {EvaluationDomain{std::move(c), m, exp}, {}};
    }
    size_t len() const { return coeffs.size(); }
    void fft() { check(kzgb_ntt_fr(coeffs.data(), exp, 0), "kzgb_ntt_fr"); }   // ft.rs:111-113
    void ifft() { check(kzgb_ntt_fr(coeffs.data(), exp, 1), "kzgb_ntt_fr"); }  // ft.rs:115-140
};

class KZGProverEvalForm {  // eval_form.rs:40-46, 86-147
public:
    // lagrange_basis_g: device SRS of L_i(s)*G in natural omega^i order (eval_form.rs:254-280)
    KZGProverEvalForm(const KZGParams& parameters, const KZGParams& lagrange_basis_g)
        : parameters_(parameters), lagrange_(lagrange_basis_g) {
        size_t m = 1;
        while (m < parameters.len()) { m *= 2; exp_ += 1; }  // compute_omega(parameters.gs.len()), eval_form.rs:92
        d_ = m;
    }
    size_t degree() const { return d_; }
    G1Affine commit(const EvaluationDomain& evals) const {  // eval_form.rs:114-122
        if (d_ != evals.d) throw std::runtime_error("assertion failed: self.d == evals.d");
        G1Affine out;
        check(kzgb_msm_g1(lagrange_.handle(), evals.coeffs.data(), evals.len(), &out), "kzgb_msm_g1");
        return out;
    }
    G1Affine create_witness(const EvaluationDomain& evals, size_t i) const {  // eval_form.rs:124-140 (infallible)
        if (i >= evals.len()) throw std::runtime_error("index out of bounds");
        G1Affine out;
        check(kzgb_witness_eval(lagrange_.handle(), evals.coeffs.data(), evals.exp, i, &out), "kzgb_witness_eval");
        return out;
    }
    G1Affine create_witness_all() const { return G1Affine{}; }  // eval_form.rs:142-146: identity

private:
    KZGParams parameters_, lagrange_;
    size_t d_ = 1;
    uint32_t exp_ = 0;
};

}  // namespace kzg_b200
