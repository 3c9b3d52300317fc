//! Replacement bodies for the hot-path seams of proxima-one/kzg.  Each function names the reference lines it
//! replaces; signatures of the public API do not change.
use crate::ffi::{self, Srs};
use crate::polynomial::Polynomial;
use crate::{KZGError, KZGWitness};
use blstrs::{G1Affine, G1Projective, Scalar};

/// `G1Projective::multi_exp(&gs[..scalars.len()], scalars)` — coeff_form.rs:61,78,102; eval_form.rs:118,136.
/// Returns a `G1Projective`, so the caller's `.to_affine()` (coeff_form.rs:63) stays as it is.
pub fn multi_exp(srs: &Srs, scalars: &[Scalar]) -> G1Projective {
    let mut out = G1Projective::default();
    ffi::check(unsafe { ffi::kzgb_msm_g1_jacobian(srs.raw(), scalars.as_ptr(), scalars.len(), &mut out) }, "kzgb_msm_g1_jacobian");
    out
}

/// `commit` (coeff_form.rs:59-64) of several polynomials in one call: an addition to the reference API
/// (`KZGProver::commit_many`), same results as a loop of `commit`.  The library pipelines the MSMs over two work
/// areas so that the latency-bound tail of one overlaps the copy / sort / accumulation of the next.
pub fn multi_exp_many(srs: &Srs, polys: &[&[Scalar]]) -> Vec<G1Affine> {
    let ptrs: Vec<*const Scalar> = polys.iter().map(|p| p.as_ptr()).collect();
    let lens: Vec<usize> = polys.iter().map(|p| p.len()).collect();
    let mut out = vec![G1Affine::default(); polys.len()];
    ffi::check(unsafe { ffi::kzgb_msm_g1_batch(srs.raw(), ptrs.as_ptr(), lens.as_ptr(), polys.len(), out.as_mut_ptr()) }, "kzgb_msm_g1_batch");
    out
}

/// Body of `best_fft(a, omega, log_n)` for the domain's own omega / omegainv — ft.rs:274-288 as called from
/// `EvaluationDomain::fft` (ft.rs:112) and `ifft` (ft.rs:116; the `* minv` loop at :118-139 is fused, so `ifft`
/// calls this with `inverse = true` and drops its scaling loop).
pub fn fft_in_place(coeffs: &mut [Scalar], log_n: u32, inverse: bool) -> Result<(), KZGError> {
    assert_eq!(coeffs.len(), 1usize << log_n); // ft.rs:302
    match ffi::check(unsafe { ffi::kzgb_ntt_fr(coeffs.as_mut_ptr(), log_n, inverse as i32) }, "kzgb_ntt_fr") {
        ffi::KZGB_DEGREE_TOO_LARGE => Err(KZGError::PolynomialDegreeTooLarge), // ft.rs:67-69
        _ => Ok(()),
    }
}

/// `KZGProver::create_witness` — coeff_form.rs:66-81, whole body (clone, `coeffs[0] -= y`, long_division by
/// X - x, remainder check, MSM, to_affine) in one call.
pub fn create_witness(srs: &Srs, polynomial: &Polynomial, (x, y): (Scalar, Scalar)) -> Result<KZGWitness, KZGError> {
    let p = polynomial.slice_coeffs();
    let mut out = G1Affine::default();
    match ffi::check(unsafe { ffi::kzgb_witness_coeff(srs.raw(), p.as_ptr(), p.len(), &x, &y, &mut out) }, "kzgb_witness_coeff") {
        ffi::KZGB_POINT_NOT_ON_POLYNOMIAL => Err(KZGError::PointNotOnPolynomial), // coeff_form.rs:74
        _ => Ok(out),
    }
}

/// `create_witness` (coeff_form.rs:66-81) at every point of `points` in one pipelined call: an addition to the
/// reference API (`KZGProver::create_witness_many`).  Equivalent to collecting the per-point `Result`s.
pub fn create_witness_many(srs: &Srs, polynomial: &Polynomial, points: &[(Scalar, Scalar)]) -> Result<Vec<KZGWitness>, KZGError> {
    let p = polynomial.slice_coeffs();
    let xs: Vec<Scalar> = points.iter().map(|q| q.0).collect();
    let ys: Vec<Scalar> = points.iter().map(|q| q.1).collect();
    let mut out = vec![G1Affine::default(); points.len()];
    let mut status = vec![0i32; points.len()];
    ffi::check(unsafe {
        ffi::kzgb_witness_coeff_many(srs.raw(), p.as_ptr(), p.len(), xs.as_ptr(), ys.as_ptr(), points.len(), out.as_mut_ptr(), status.as_mut_ptr())
    }, "kzgb_witness_coeff_many");
    if status.iter().any(|&s| s == ffi::KZGB_POINT_NOT_ON_POLYNOMIAL) {
        return Err(KZGError::PointNotOnPolynomial); // coeff_form.rs:74
    }
    Ok(out)
}

/// `KZGProver::create_witness_batched` — coeff_form.rs:83-111.  Returns (r, w); the caller wraps them in
/// `KZGBatchWitness { r, w }` (coeff_form.rs:105-108).
pub fn create_witness_batched(srs: &Srs, polynomial: &Polynomial, xs: &[Scalar], ys: &[Scalar]) -> Result<(Polynomial, G1Affine), KZGError> {
    assert_eq!(xs.len(), ys.len()); // polynomial.rs:242
    let p = polynomial.slice_coeffs();
    let mut r = vec![Scalar::default(); xs.len().max(2)];
    let mut r_len = 0usize;
    let mut w = G1Affine::default();
    let st = unsafe {
        ffi::kzgb_witness_batched(srs.raw(), p.as_ptr(), p.len(), xs.as_ptr(), ys.as_ptr(), xs.len(), r.as_mut_ptr(), &mut r_len, &mut w)
    };
    match ffi::check(st, "kzgb_witness_batched") {
        ffi::KZGB_POINT_NOT_ON_POLYNOMIAL => Err(KZGError::PointNotOnPolynomial), // coeff_form.rs:96
        _ => {
            r.truncate(r_len);
            Ok((interpolant_shape(r, xs.len()), w))
        }
    }
}

/// `KZGProverEvalForm::create_witness` — eval_form.rs:124-140 (numerator, div_by_omega_i, MSM, to_affine).
pub fn create_witness_eval(lagrange: &Srs, evals: &[Scalar], log_d: u32, i: usize) -> KZGWitness {
    let mut out = G1Affine::default();
    ffi::check(unsafe { ffi::kzgb_witness_eval(lagrange.raw(), evals.as_ptr(), log_d, i, &mut out) }, "kzgb_witness_eval");
    out
}

/// The shape the reference's tree recursion gives the interpolant (polynomial.rs:303-365): below 258 points exactly m
/// coefficients with declared degree m - 1 (m == 1: X + (y - x), degree 1, polynomial.rs:244-247); from 258 points on the
/// top-level `best_mul` takes the `fft_mul` path (polynomial.rs:167-191), whose result keeps the whole padded transform
/// (2^k >= m + 1 coefficients) with a re-measured degree (`Polynomial::new`).
fn interpolant_shape(mut r: Vec<Scalar>, m: usize) -> Polynomial {
    if m >= 258 {
        r.resize((m + 1).next_power_of_two(), Scalar::default());
        Polynomial::new(r)
    } else {
        let d = r.len() - 1;
        Polynomial::new_from_coeffs(r, d)
    }
}

/// `commit` over a point-sharded SRS (one process per GPU, `ffi::kzgb_comm_init` done at start-up): every rank passes
/// its slice of the coefficients; the library all-gathers the 192-byte partial sums over NCCL and every rank returns
/// the full commitment.  Same bytes as `commit` on one GPU.
pub fn multi_exp_sharded(shard: &Srs, scalars_shard: &[Scalar]) -> G1Affine {
    let mut out = G1Affine::default();
    ffi::check(unsafe { ffi::kzgb_msm_g1_sharded(shard.raw(), scalars_shard.as_ptr(), scalars_shard.len(), &mut out) }, "kzgb_msm_g1_sharded");
    out
}

/// `commit` / `create_witness` on a resident copy of the polynomial (`ffi::DevicePoly`): no per-call upload.
pub fn commit_resident(srs: &Srs, poly: &ffi::DevicePoly, num_coeffs: usize) -> G1Affine {
    let mut out = G1Affine::default();
    ffi::check(unsafe { ffi::kzgb_poly_commit(srs.raw(), poly.raw(), num_coeffs, &mut out) }, "kzgb_poly_commit");
    out
}
pub fn create_witness_resident(srs: &Srs, poly: &ffi::DevicePoly, num_coeffs: usize, (x, y): (Scalar, Scalar)) -> Result<KZGWitness, KZGError> {
    let mut out = G1Affine::default();
    match ffi::check(unsafe { ffi::kzgb_poly_witness(srs.raw(), poly.raw(), num_coeffs, &x, &y, &mut out) }, "kzgb_poly_witness") {
        ffi::KZGB_POINT_NOT_ON_POLYNOMIAL => Err(KZGError::PointNotOnPolynomial),
        _ => Ok(out),
    }
}
