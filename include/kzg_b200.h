/*
 * kzg_b200.h -- C ABI of libkzg_b200.so: the B200 (sm_100a) implementation of
 * the KZG prover hot path of proxima-one/kzg.
 *
 * The reference has no FFI of its own (pure Rust over blstrs/blst); the seam
 * this library fills is INSIDE the bodies of the reference's public methods
 * (SURVEY.md 8b).  Each entry point names the reference code it replaces
 * (file:line relative to the reference root).  INTEGRATION.md shows the Rust
 * (`extern "C"` / bindgen) and Python (ctypes) bindings.
 *
 * Data layout = blstrs/blst in-memory layout, passed without conversion:
 *   kzgb_fr        == blst_fr        4 x u64 LE limbs, Montgomery R = 2^256 mod r
 *   kzgb_fp        == blst_fp        6 x u64 LE limbs, Montgomery R = 2^384 mod p
 *   kzgb_p1        == blst_p1        Jacobian {x,y,z}, 144 B, identity: z == 0
 *   kzgb_p1_affine == blst_p1_affine {x,y}, 96 B, identity: all-zero
 *
 * Conventions: every function returns an int status (never unwinds):
 *   0  KZGB_OK
 *   1  KZGB_POINT_NOT_ON_POLYNOMIAL   (KZGError::PointNotOnPolynomial, coeff_form.rs:74,96)
 *   2  KZGB_DEGREE_TOO_LARGE          (KZGError::PolynomialDegreeTooLarge, ft.rs:67-69)
 *  <0  invalid argument / CUDA failure; kzgb_last_error_message() has the text.
 *      The Rust shim maps these to panic!, as the reference panics on
 *      out-of-range slices and failed asserts (coeff_form.rs:60, eval_form.rs:115).
 * Host-pointer entry points copy in/out through the library's pinned staging
 * buffers.  `_dev` entry points take device pointers (same layouts) and run on
 * the caller-supplied CUDA stream (a cudaStream_t passed as void*; NULL = the
 * library's own stream).  One process drives one GPU (kzgb_init(device)).
 * Calls are serialised per process by an internal mutex (the reference's
 * provers are Send + Sync; concurrent callers are safe, not parallel).
 */
#ifndef KZG_B200_H
#define KZG_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define KZGB_API __attribute__((visibility("default")))
#else
#define KZGB_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint64_t l[4]; } kzgb_fr;
typedef struct { uint64_t l[6]; } kzgb_fp;
typedef struct { kzgb_fp x, y, z; } kzgb_p1;
typedef struct { kzgb_fp x, y; } kzgb_p1_affine;
typedef struct kzgb_srs kzgb_srs; /* opaque: device-resident SRS (affine, + precomputed tables) */

#define KZGB_OK 0
#define KZGB_POINT_NOT_ON_POLYNOMIAL 1
#define KZGB_DEGREE_TOO_LARGE 2
#define KZGB_ERR_INVALID_ARG (-1)
#define KZGB_ERR_CUDA (-2)
#define KZGB_ERR_NOT_INITIALISED (-3)

/* ---- lifecycle ------------------------------------------------------------ */
KZGB_API int kzgb_init(int device);
KZGB_API int kzgb_shutdown(void);
KZGB_API const char* kzgb_last_error_message(void);
/* number of kernels this library has launched since kzgb_init (bench.py's gpu_launches) */
KZGB_API uint64_t kzgb_launch_count(void);

/* ---- on-box peak measurements (roofline denominators) ------------ */
/* mode: 0 IMAD.lo, 1 IMAD.HI, 2 IMAD.WIDE, 3 mad.lo.cc+madc.hi pair, 4 Fq mul chain, 5 Fr mul chain,
 *       6 IADD, 7 IMAD.WIDE+IADD, 100 = D2D copy of `iters` MiB.
 * out[0] = ops/s, out[1] = ops per clock per SM, out[2] = ms of the last repetition. */
KZGB_API int kzgb_microbench(int mode, int blocks_per_sm, int threads, int iters, double* out);

/* ---- test hooks: element-wise device field arithmetic on host vectors ------- */
/* op: 0 a*b, 1 a+b, 2 a-b, 3 a^-1, 4 to_mont(a), 5 from_mont(a), 6 a^2 (b may be NULL for op >= 3).
 * Pins the device Fr/Fp arithmetic (which replaces blst's, Cargo.toml:27) against the oracle. */
KZGB_API int kzgb_test_fr_op(const kzgb_fr* a, const kzgb_fr* b, kzgb_fr* out, size_t n, int op);
KZGB_API int kzgb_test_fp_op(const kzgb_fp* a, const kzgb_fp* b, kzgb_fp* out, size_t n, int op);

#ifdef __cplusplus
}
#endif
#endif /* KZG_B200_H */
