//! Raw bindings to libkzg_b200.so (include/kzg_b200.h) and a safe SRS handle.
//!
//! Layouts: `blstrs::Scalar` is `#[repr(transparent)]` over `blst_fr`, `G1Projective` over `blst_p1`,
//! `G1Affine` over `blst_p1_affine`, so slices of them are passed as-is (no conversion, no copy on the host).
#![allow(non_camel_case_types)]
use blstrs::{G1Affine, G1Projective, Scalar};
use std::os::raw::{c_char, c_int, c_void};
use std::sync::Once;

#[repr(C)]
pub struct kzgb_srs {
    _private: [u8; 0],
}
#[repr(C)]
pub struct kzgb_poly {
    _private: [u8; 0],
}
#[repr(C)]
pub struct kzgb_ntt4_plan {
    _private: [u8; 0],
}
pub const KZGB_COMM_ID_BYTES: usize = 128;

pub const KZGB_OK: c_int = 0;
pub const KZGB_POINT_NOT_ON_POLYNOMIAL: c_int = 1;
pub const KZGB_DEGREE_TOO_LARGE: c_int = 2;

extern "C" {
    pub fn kzgb_init(device: c_int) -> c_int;
    pub fn kzgb_last_error_message() -> *const c_char;
    pub fn kzgb_srs_create_g1(points: *const G1Projective, n: usize, window_bits: c_int, out: *mut *mut kzgb_srs) -> c_int;
    pub fn kzgb_srs_destroy(srs: *mut kzgb_srs) -> c_int;
    pub fn kzgb_msm_g1_jacobian(srs: *mut kzgb_srs, scalars: *const Scalar, m: usize, out: *mut G1Projective) -> c_int;
    pub fn kzgb_msm_g1(srs: *mut kzgb_srs, scalars: *const Scalar, m: usize, out: *mut G1Affine) -> c_int;
    pub fn kzgb_msm_g1_batch(srs: *mut kzgb_srs, scalars: *const *const Scalar, ms: *const usize, count: usize, out: *mut G1Affine) -> c_int;
    pub fn kzgb_ntt_fr(inout: *mut Scalar, log_n: u32, inverse: c_int) -> c_int;
    pub fn kzgb_divide_linear(p: *const Scalar, n: usize, x: *const Scalar, y: *const Scalar, q_out: *mut Scalar, rem_out: *mut Scalar) -> c_int;
    pub fn kzgb_witness_coeff(srs: *mut kzgb_srs, p: *const Scalar, n: usize, x: *const Scalar, y: *const Scalar, out: *mut G1Affine) -> c_int;
    pub fn kzgb_witness_coeff_many(srs: *mut kzgb_srs, p: *const Scalar, n: usize, xs: *const Scalar, ys: *const Scalar, m: usize,
                                   out: *mut G1Affine, status: *mut c_int) -> c_int;
    pub fn kzgb_witness_batched(srs: *mut kzgb_srs, p: *const Scalar, n: usize, xs: *const Scalar, ys: *const Scalar, m: usize,
                                r_out: *mut Scalar, r_len: *mut usize, w_out: *mut G1Affine) -> c_int;
    pub fn kzgb_div_by_omega_i(evals: *const Scalar, log_d: u32, i: usize, q_out: *mut Scalar) -> c_int;
    pub fn kzgb_witness_eval(lagrange_srs: *mut kzgb_srs, evals: *const Scalar, log_d: u32, i: usize, out: *mut G1Affine) -> c_int;
    // device-resident polynomial handles (one upload, many calls)
    pub fn kzgb_poly_upload(coeffs: *const Scalar, n: usize, out: *mut *mut kzgb_poly) -> c_int;
    pub fn kzgb_poly_free(poly: *mut kzgb_poly) -> c_int;
    pub fn kzgb_poly_commit(srs: *mut kzgb_srs, poly: *mut kzgb_poly, m: usize, out: *mut G1Affine) -> c_int;
    pub fn kzgb_poly_witness(srs: *mut kzgb_srs, poly: *mut kzgb_poly, n: usize, x: *const Scalar, y: *const Scalar, out: *mut G1Affine) -> c_int;
    pub fn kzgb_poly_witness_batched(srs: *mut kzgb_srs, poly: *mut kzgb_poly, n: usize, xs: *const Scalar, ys: *const Scalar, m: usize,
                                     r_out: *mut Scalar, r_len: *mut usize, w_out: *mut G1Affine) -> c_int;
    pub fn kzgb_poly_witness_eval(lagrange_srs: *mut kzgb_srs, evals: *mut kzgb_poly, log_d: u32, i: usize, out: *mut G1Affine) -> c_int;
    // multi-GPU (one process per GPU; NCCL inside the library)
    pub fn kzgb_comm_unique_id(id: *mut u8) -> c_int;
    pub fn kzgb_comm_init(rank: c_int, world: c_int, id: *const u8) -> c_int;
    pub fn kzgb_comm_destroy() -> c_int;
    pub fn kzgb_srs_generate(kind: c_int, secret: *const Scalar, total: usize, first: usize, count: usize, window_bits: c_int,
                             out: *mut *mut kzgb_srs) -> c_int;
    pub fn kzgb_msm_g1_sharded(shard: *mut kzgb_srs, scalars: *const Scalar, m: usize, out: *mut G1Affine) -> c_int;
    pub fn kzgb_msm_g1_sharded_batch(shard: *mut kzgb_srs, scalars: *const *const Scalar, ms: *const usize, count: usize,
                                     out: *mut G1Affine) -> c_int;
    pub fn kzgb_ntt4_plan_create(log_n: u32, log_n1: u32, out: *mut *mut kzgb_ntt4_plan) -> c_int;
    pub fn kzgb_ntt4_plan_destroy(plan: *mut kzgb_ntt4_plan) -> c_int;
    pub fn kzgb_ntt_fr_sharded(plan: *mut kzgb_ntt4_plan, input: *const Scalar, out: *mut Scalar, inverse: c_int, gather: c_int) -> c_int;
    #[allow(dead_code)]
    pub fn kzgb_shutdown() -> c_int;
    #[allow(dead_code)]
    pub fn kzgb_profile_enable(on: c_int) -> c_int;
    #[allow(dead_code)]
    pub fn kzgb_ntt_fr_dev(d_inout: *mut c_void, log_n: u32, inverse: c_int, stream: *mut c_void) -> c_int;
}

static INIT: Once = Once::new();

/// Negative statuses are programming / device errors: the reference panics in the same situations
/// (slice out of range at coeff_form.rs:60, failed asserts at eval_form.rs:115), so the shim panics too.
pub fn check(status: c_int, what: &str) -> c_int {
    if status < 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(kzgb_last_error_message()) };
        panic!("{}: {}", what, msg.to_string_lossy());
    }
    status
}

/// Device-resident copy of a G1 SRS (affine + window-multiple table).  Created once per prover and shared by
/// clones through an `Arc`; safe to cache because provers borrow `&'params KZGParams` (coeff_form.rs:39).
pub struct Srs(*mut kzgb_srs);
unsafe impl Send for Srs {}
unsafe impl Sync for Srs {} // calls on one handle serialise on the handle's own lock; different handles run concurrently

impl Srs {
    pub fn new(points: &[G1Projective]) -> Self {
        INIT.call_once(|| {
            let dev = std::env::var("KZG_B200_DEVICE").ok().and_then(|v| v.parse().ok()).unwrap_or(0);
            check(unsafe { kzgb_init(dev) }, "kzgb_init");
        });
        let mut h = std::ptr::null_mut();
        check(unsafe { kzgb_srs_create_g1(points.as_ptr(), points.len(), 0, &mut h) }, "kzgb_srs_create_g1");
        Srs(h)
    }
    pub fn raw(&self) -> *mut kzgb_srs {
        self.0
    }
}

impl Drop for Srs {
    fn drop(&mut self) {
        unsafe { kzgb_srs_destroy(self.0) };
    }
}

/// Device-resident copy of a polynomial's coefficients (`kzgb_poly`): `commit` -> `create_witness` ->
/// `create_witness_batched` on the same `&Polynomial` then upload it once instead of once per call.
pub struct DevicePoly(*mut kzgb_poly, usize);
unsafe impl Send for DevicePoly {}
unsafe impl Sync for DevicePoly {}
impl DevicePoly {
    pub fn new(coeffs: &[Scalar]) -> Self {
        let mut h = std::ptr::null_mut();
        check(unsafe { kzgb_poly_upload(coeffs.as_ptr(), coeffs.len(), &mut h) }, "kzgb_poly_upload");
        DevicePoly(h, coeffs.len())
    }
    pub fn raw(&self) -> *mut kzgb_poly {
        self.0
    }
    pub fn len(&self) -> usize {
        self.1
    }
}
impl Drop for DevicePoly {
    fn drop(&mut self) {
        unsafe { kzgb_poly_free(self.0) };
    }
}
