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
 * Host-pointer entry points copy in/out on the library's own stream and return
 * when the result is in the caller's buffer.  Pinned memory gives full PCIe
 * speed; ordinary pageable memory (a Rust Vec<Fr>) is accepted everywhere, and
 * transfers of 4 MiB or more are relayed through a pinned ring by a few host
 * threads (KZGB_UPLOAD_THREADS, default 4; 0 leaves the copy to the driver).
 * `_dev` entry points take device pointers (same layouts) and enqueue
 * on the caller-supplied CUDA stream (a cudaStream_t passed as void*; NULL =
 * the CUDA default stream) without synchronising, except where a status must
 * be read back (the witness calls).  One process drives one GPU
 * (kzgb_init(device)); every entry point selects that device on the calling
 * thread for the duration of the call and restores the caller's.
 * Thread safety (the reference's provers are Send + Sync plain structs,
 * coeff_form.rs:37-41): any entry point may be called from any host thread.
 * A call runs on one of a small pool of lanes (stream + scratch arenas), so
 * calls from different threads overlap on the device; calls that use the SAME
 * kzgb_srs handle serialise on that handle's lock (its work area is per
 * handle).  Arenas and work areas are reused between calls: when consecutive
 * users run on different streams the earlier stream is drained first.
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

/* ---- SRS handles (KZGParams.gs, lib.rs:14-19; created once per prover) ----- */
/* Uploads the G1 half of KZGParams (Vec<G1Projective> == contiguous blst_p1), converts it to affine ONCE (the
 * reference's multi_exp redoes this conversion on every call) and precomputes the window-multiple table
 * T[w][i] = 2^(c*w) * gs[i].  window_bits = 0 lets the library choose c from n. */
KZGB_API int kzgb_srs_create_g1(const kzgb_p1* points, size_t n, int window_bits, kzgb_srs** out);
/* Same from affine points (host pointer, or device pointer when on_device != 0). */
KZGB_API int kzgb_srs_create_g1_affine(const kzgb_p1_affine* points, size_t n, int window_bits, int on_device,
                                       kzgb_srs** out);
KZGB_API int kzgb_srs_destroy(kzgb_srs* srs);
KZGB_API size_t kzgb_srs_size(const kzgb_srs* srs);
KZGB_API int kzgb_srs_window_bits(const kzgb_srs* srs);
KZGB_API size_t kzgb_srs_device_bytes(const kzgb_srs* srs);
KZGB_API const void* kzgb_srs_device_points(const kzgb_srs* srs); /* device ptr to n affine points */
KZGB_API int kzgb_srs_read_affine(const kzgb_srs* srs, size_t start, size_t count, kzgb_p1_affine* out);
/* Synthetic SRS built on the device from a KNOWN secret (benchmarks/tests):
 * kind 0: gs[i] = s^i*G, the points lib.rs:38-46 `setup` produces;
 * kind 1: L_i(s)*G in natural omega^i order, the points eval_form.rs:254-280 `compute_lagrange_basis` produces
 *         (total must be a power of two).  Only [first, first+count) of the size-`total` SRS is generated,
 *         so each rank of a multi-GPU job builds its own shard. */
KZGB_API int kzgb_srs_generate(int kind, const kzgb_fr* secret, size_t total, size_t first, size_t count,
                               int window_bits, kzgb_srs** out);

/* Lagrange-basis SRS from a monomial SRS WITHOUT the secret: the points eval_form.rs:254-280
 * `compute_lagrange_basis(params)` returns (O(d^3) there), computed as an inverse NTT over the group:
 * out[i] = (1/d) sum_j omega^(-ij) gs[j] = L_i(s)*G, natural omega^i order, d = 2^log_d <= size of `monomial`. */
KZGB_API int kzgb_srs_lagrange_from_monomial(const kzgb_srs* monomial, uint32_t log_d, int window_bits, kzgb_srs** out);

/* ---- G1 MSM: replaces G1Projective::multi_exp(gs[..m], scalars) ---------------
 * reference call sites: coeff_form.rs:61 (commit), :78 (create_witness), :102 (create_witness_batched),
 * eval_form.rs:118 (commit), :136 (create_witness).  Uses the first m points of the SRS (m <= n), as the
 * reference slices gs[..num_coeffs] (coeff_form.rs:60,77,101).  Scalars are Montgomery-form blst_fr. */
KZGB_API int kzgb_msm_g1(kzgb_srs* srs, const kzgb_fr* scalars, size_t m, kzgb_p1_affine* out);   /* + .to_affine() (coeff_form.rs:63) */
KZGB_API int kzgb_msm_g1_jacobian(kzgb_srs* srs, const kzgb_fr* scalars, size_t m, kzgb_p1* out); /* G1Projective result */
/* Device-pointer variant.  out_mode: 0 affine (96 B), 1 Jacobian (144 B), 2 raw XYZZ partial sum (192 B). */
KZGB_API int kzgb_msm_g1_dev(kzgb_srs* srs, const void* d_scalars, size_t m, void* d_out, int out_mode, void* stream);
/* Multi-GPU combine: sums `count` raw XYZZ partial sums (gathered from the ranks) into one point. */
/* `count` commits against one SRS in one call (a prover committing to several polynomials; each item is
 * coeff_form.rs:59-64).  Same results as `count` separate calls; consecutive MSMs are pipelined over two work
 * areas so that one MSM's latency-bound tail overlaps the next one's copy, sort and accumulation.
 * Host variant: scalars[i] -> ms[i] host scalars, out[i] affine.  Device variant: d_scalars is a HOST array of
 * device pointers, d_out a device array of `count` results (out_mode 0 affine 96 B / 1 Jacobian 144 B / 2 XYZZ 192 B),
 * ordered after and before the work on `stream`. */
KZGB_API int kzgb_msm_g1_batch(kzgb_srs* srs, const kzgb_fr* const* scalars, const size_t* ms, size_t count,
                               kzgb_p1_affine* out);
KZGB_API int kzgb_msm_g1_batch_dev(kzgb_srs* srs, const void* const* d_scalars, const size_t* ms, size_t count,
                                   void* d_out, int out_mode, void* stream);
/* create_witness (coeff_form.rs:66-81) of one polynomial at m points in one pipelined call; p uploaded once.
 * status[i] = 0 or 1 (PointNotOnPolynomial: out[i] unspecified).  Same bytes as m calls of kzgb_witness_coeff. */
KZGB_API int kzgb_witness_coeff_many(kzgb_srs* srs, const kzgb_fr* p, size_t n, const kzgb_fr* xs, const kzgb_fr* ys,
                                     size_t m, kzgb_p1_affine* out, int* status);
/* eval-form create_witness (eval_form.rs:124-140) of one evaluation vector at m indices, pipelined; evals uploaded once */
KZGB_API int kzgb_witness_eval_many(kzgb_srs* lagrange_srs, const kzgb_fr* evals, uint32_t log_d, const size_t* indices,
                                    size_t m, kzgb_p1_affine* out);
/* host scalars in, device results out (multi-GPU commit from host buffers); returns when the batch is complete */
KZGB_API int kzgb_msm_g1_batch_h2d(kzgb_srs* srs, const kzgb_fr* const* scalars, const size_t* ms, size_t count,
                                   void* d_out, int out_mode, void* stream);
KZGB_API int kzgb_g1_combine_dev(const void* d_partials, uint32_t count, void* d_out, int out_mode, void* stream);
/* K polynomials at once: d_partials = all-gathered [count ranks][items][192 B]; d_out = `items` results */
KZGB_API int kzgb_g1_combine_batch_dev(const void* d_partials, uint32_t count, uint32_t items, void* d_out, int out_mode,
                                       void* stream);

/* ---- Fr NTT: replaces the bodies of EvaluationDomain::fft (ft.rs:111-113) and ifft (ft.rs:115-140),
 * i.e. best_fft/serial_fft/parallel_fft (ft.rs:274-387).  In place over 2^log_n Montgomery-form elements,
 * natural order in and out; out[k] = sum_j a[j]*omega^(jk) with omega = root_of_unity^(2^(32-log_n))
 * (ft.rs:73); inverse uses omega^-1 and multiplies by n^-1.  log_n >= 32 -> KZGB_DEGREE_TOO_LARGE (ft.rs:67). */
KZGB_API int kzgb_ntt_fr(kzgb_fr* inout, uint32_t log_n, int inverse);
KZGB_API int kzgb_ntt_fr_dev(void* d_inout, uint32_t log_n, int inverse, void* stream);

/* ---- quotient by a linear factor: Polynomial::long_division with divisor X - x (polynomial.rs:193-227) as used
 * by create_witness (coeff_form.rs:70-71).  p: n >= 1 coefficients low -> high (= slice_coeffs()); computes
 * q = (p - y) div (X - x) (n-1 coefficients, not degree-shrunk, as polynomial.rs:202-205) and rem = p(x) - y.
 * y may be NULL (y = 0). */
KZGB_API int kzgb_divide_linear(const kzgb_fr* p, size_t n, const kzgb_fr* x, const kzgb_fr* y, kzgb_fr* q_out,
                                kzgb_fr* rem_out);
KZGB_API int kzgb_divide_linear_dev(const void* d_p, size_t n, const void* d_xy, int use_y, void* d_q, void* d_rem,
                                    void* stream);

/* ---- KZGProver::create_witness (coeff_form.rs:66-81): psi = (p - y)/(X - x), w = MSM(gs[..n-1], psi).to_affine().
 * Returns KZGB_POINT_NOT_ON_POLYNOMIAL when the remainder is non-zero (coeff_form.rs:74).  n = degree + 1. */
KZGB_API int kzgb_witness_coeff(kzgb_srs* srs, const kzgb_fr* p, size_t n, const kzgb_fr* x, const kzgb_fr* y,
                                kzgb_p1_affine* out);
/* device variant: d_xy = {x, y}; d_work = n Fr of scratch (receives psi); d_out = 96 B; synchronises the stream. */
KZGB_API int kzgb_witness_coeff_dev(kzgb_srs* srs, const void* d_p, size_t n, const void* d_xy, void* d_work,
                                    void* d_out, void* stream);

/* ---- KZGProver::create_witness_batched (coeff_form.rs:83-111):
 *   r = Polynomial::lagrange_interpolation_with_tree(xs, ys) (polynomial.rs:237-264; tree :303-365),
 *   psi = (p - r) div prod(X - x_i) (polynomial.rs:443-460, 193-227), w = MSM(gs, psi).to_affine().
 * r_out receives the coefficients of r: m of them (declared degree m-1), or for m == 1 the TWO coefficients of
 * X + (y - x) -- the reference's one-point quirk (polynomial.rs:244-247).  KZGB_POINT_NOT_ON_POLYNOMIAL when the
 * remainder is non-zero (coeff_form.rs:96).  xs must be distinct (the reference panics; here KZGB_ERR_INVALID_ARG).
 * 1 <= m <= 16384: the interpolation keeps an m x m table (32 m^2 bytes of device scratch: 512 MiB at 4096, 8 GiB at
 * the limit), unlike the reference's SubProductTree; larger batches return KZGB_ERR_INVALID_ARG and must be split.
 * r_out always holds exactly *r_len = m coefficients; for m >= 258 the reference's fft_mul path (polynomial.rs:167-191)
 * returns the same coefficients in a vector zero-padded to 2^k >= m + 1 -- the host mirrors reproduce that shape. */
KZGB_API int kzgb_witness_batched(kzgb_srs* srs, const kzgb_fr* p, size_t n, const kzgb_fr* xs, const kzgb_fr* ys,
                                  size_t m, kzgb_fr* r_out, size_t* r_len, kzgb_p1_affine* w_out);
KZGB_API int kzgb_witness_batched_dev(kzgb_srs* srs, const void* d_p, size_t n, const void* d_xs, const void* d_ys,
                                      size_t m, void* d_r_out, void* d_out, void* stream);
/* interpolation alone (for tests): r_out has max(m,2) entries */
KZGB_API int kzgb_interpolate(const kzgb_fr* xs, const kzgb_fr* ys, size_t m, kzgb_fr* r_out, size_t* r_len);

/* ---- evaluation form: div_by_omega_i(evals - evals[i], i) (eval_form.rs:58-84, 125-131) and
 * KZGProverEvalForm::create_witness (eval_form.rs:124-140) over the Lagrange-basis SRS.  d = 2^log_d. */
KZGB_API int kzgb_div_by_omega_i(const kzgb_fr* evals, uint32_t log_d, size_t i, kzgb_fr* q_out);
KZGB_API int kzgb_witness_eval(kzgb_srs* lagrange_srs, const kzgb_fr* evals, uint32_t log_d, size_t i,
                               kzgb_p1_affine* out);
KZGB_API int kzgb_witness_eval_dev(kzgb_srs* lagrange_srs, const void* d_evals, uint32_t log_d, size_t i, void* d_work,
                                   void* d_out, void* stream);

/* ---- device-resident polynomial / evaluation-vector handles: the reference's provers borrow
 * &Polynomial (polynomial.rs:24-27) / &EvaluationDomain (ft.rs:17-25) and typically run commit -> create_witness ->
 * create_witness_batched on the SAME polynomial; a kzgb_poly keeps it in HBM (one H2D at upload), so the per-call
 * traffic is the opening point in and 96 bytes out.  `n` / `m` = polynomial.num_coeffs() (<= uploaded length). */
typedef struct kzgb_poly kzgb_poly;
KZGB_API int kzgb_poly_upload(const kzgb_fr* coeffs, size_t n, kzgb_poly** out);
KZGB_API int kzgb_poly_free(kzgb_poly* poly);
KZGB_API size_t kzgb_poly_len(const kzgb_poly* poly);
KZGB_API const void* kzgb_poly_device_ptr(const kzgb_poly* poly);
KZGB_API int kzgb_poly_download(kzgb_poly* poly, kzgb_fr* out);
KZGB_API int kzgb_poly_commit(kzgb_srs* srs, kzgb_poly* poly, size_t m, kzgb_p1_affine* out);            /* coeff_form.rs:59-64, eval_form.rs:114-122 */
KZGB_API int kzgb_poly_witness(kzgb_srs* srs, kzgb_poly* poly, size_t n, const kzgb_fr* x, const kzgb_fr* y,
                               kzgb_p1_affine* out);                                                     /* coeff_form.rs:66-81 */
KZGB_API int kzgb_poly_witness_batched(kzgb_srs* srs, kzgb_poly* poly, size_t n, const kzgb_fr* xs, const kzgb_fr* ys,
                                       size_t m, kzgb_fr* r_out, size_t* r_len, kzgb_p1_affine* w_out);   /* coeff_form.rs:83-111 */
KZGB_API int kzgb_poly_witness_eval(kzgb_srs* lagrange_srs, kzgb_poly* evals, uint32_t log_d, size_t i,
                                    kzgb_p1_affine* out);                                                /* eval_form.rs:124-140 */
KZGB_API int kzgb_poly_ntt(kzgb_poly* poly, uint32_t log_n, int inverse);                                /* ft.rs:111-140, in place */

/* ---- stage timers (CUDA events on the launching stream; bench.py's roofline) --- */
KZGB_API int kzgb_profile_enable(int on);
KZGB_API int kzgb_profile_read(double* sum_ms, uint64_t* counts); /* arrays of kzgb_profile_stage_count() */
KZGB_API int kzgb_profile_stage_count(void);
KZGB_API const char* kzgb_profile_stage_name(int stage);
/* per-launch trace (diagnostics): warm in-stream time of every kernel launch, keyed by the source line of the launch */
KZGB_API int kzgb_trace_enable(int on);
KZGB_API int kzgb_trace_read(int* lines, float* ms, int cap);

/* ---- four-step building blocks of the multi-GPU NTT (N >= 2^24, one all-to-all; structural
 * precedent: parallel_fft, ft.rs:336-387).  View the vector as N1 x N2 (index i1*N2 + i2).  Rank g holds the slab
 * [N1][C] of columns i2 in [g*C, (g+1)*C), C = N2/G.
 *   kzgb_ntt4_columns_dev: in place, N1-point transforms down the columns + twiddle omega_N^(k1*i2);
 *   (all-to-all of row blocks between ranks: done by the caller over NCCL)
 *   kzgb_ntt4_unpack_dev : recv[g][r][c] -> rows[r][g*C + c];
 *   kzgb_ntt_batch_dev   : N2-point transforms of the N1/G local rows; row r of rank h then holds
 *                          X[(h*N1/G + r) + N1*k2], k2 = 0..N2-1. */
KZGB_API int kzgb_ntt4_columns_dev(void* d_slab, uint32_t log_n, uint32_t log_n1, uint32_t log_c, uint32_t col_offset,
                                   int inverse, void* stream);
KZGB_API int kzgb_ntt4_unpack_dev(const void* d_recv, void* d_rows, uint32_t world, uint32_t nrows, uint32_t cols,
                                  void* stream);
KZGB_API int kzgb_ntt_batch_dev(void* d_inout, void* d_tmp, uint32_t log_n, uint32_t batch, int inverse,
                                const kzgb_fr* scale, void* stream);
/* Fused column pass + exchange over NVLink peer memory (replaces columns + all-to-all + unpack): element (k1, c)
 * of the column pass is stored directly into the row matrix of the rank that owns row k1, already transposed:
 * peer_rows[k1 / (N1/world)][(k1 mod (N1/world)) * N2 + rank*C + c].  peer_rows: host array of `world` device
 * pointers valid on this device (peer_rows[rank] local, the others CUDA-IPC mappings: kzgb_peer_*).  The slab is
 * only read.  The caller puts a barrier between this call and kzgb_ntt_batch_dev on the rows. */
KZGB_API int kzgb_ntt4_columns_scatter_dev(const void* d_slab, void* const* peer_rows, uint32_t world, uint32_t rank,
                                           uint32_t log_n, uint32_t log_n1, uint32_t log_c, uint32_t col_offset,
                                           int inverse, void* stream);
/* peer memory: cudaMalloc + a 64-byte CUDA IPC handle to send to the other ranks of the node; open / close a
 * peer's handle (enables peer access lazily); free what kzgb_peer_alloc returned */
KZGB_API int kzgb_peer_alloc(size_t bytes, void** d_ptr, unsigned char handle[64]);
KZGB_API int kzgb_peer_open(const unsigned char handle[64], void** d_ptr);
KZGB_API int kzgb_peer_close(void* d_ptr);
KZGB_API int kzgb_peer_free(void* d_ptr);

/* ---- multi-GPU entry points (one process per GPU on one NVLink/NVSwitch node; NCCL bound at run time) ---------------
 * SURVEY 8(e) / north_star: the MSM shards by points with the per-GPU partial sums reduced over NCCL; NTTs >= 2^24 use
 * the four-step split.  Every rank of the job makes the same sequence of calls.  Bring-up: rank 0 calls
 * kzgb_comm_unique_id and ships the 128 bytes to the other ranks (any transport the host program has: MPI, a file, a
 * broadcast of its own launcher); every rank then calls kzgb_comm_init after kzgb_init(local device).  With world == 1
 * (or no communicator) the sharded entry points degenerate to their single-GPU twins. */
#define KZGB_COMM_ID_BYTES 128
KZGB_API int kzgb_comm_unique_id(unsigned char id[KZGB_COMM_ID_BYTES]);
KZGB_API int kzgb_comm_init(int rank, int world, const unsigned char id[KZGB_COMM_ID_BYTES]);
KZGB_API int kzgb_comm_destroy(void);
KZGB_API int kzgb_comm_rank(void);
KZGB_API int kzgb_comm_world(void);
KZGB_API int kzgb_comm_barrier(void* stream); /* stream-ordered (a 4-byte all-reduce): no host synchronisation */
/* KZGProver::commit (coeff_form.rs:59-64) over a point-sharded SRS: `shard` holds gs[lo_g, hi_g) (built by this rank
 * with kzgb_srs_create_* / kzgb_srs_generate(first = lo_g)), `scalars` the matching slice coeffs[lo_g, hi_g).  Every
 * rank receives the full result.  Exchange: one all-gather of world x 192 B. */
KZGB_API int kzgb_msm_g1_sharded(kzgb_srs* shard, const kzgb_fr* scalars, size_t m, kzgb_p1_affine* out);
KZGB_API int kzgb_msm_g1_sharded_dev(kzgb_srs* shard, const void* d_scalars, size_t m, void* d_out, int out_mode, void* stream);
/* `count` sharded commits in one call (local MSMs pipelined as in kzgb_msm_g1_batch, ONE all-gather of all partials) */
KZGB_API int kzgb_msm_g1_sharded_batch(kzgb_srs* shard, const kzgb_fr* const* scalars, const size_t* ms, size_t count,
                                       kzgb_p1_affine* out);
KZGB_API int kzgb_msm_g1_sharded_batch_dev(kzgb_srs* shard, const void* const* d_scalars, const size_t* ms, size_t count,
                                           void* d_out, int out_mode, void* stream);
/* EvaluationDomain::fft / ifft (ft.rs:111-140) of ONE 2^log_n vector split over the ranks: four-step with the
 * exchange fused into the column pass (peer stores over NVLink into IPC-mapped row matrices, one stream-ordered
 * barrier, row pass).  The plan owns the peer-mapped matrices; create/destroy are collective.  log_n1 = 0: automatic.
 * Layouts: input slab [N1][C] with slab[i1][c] = a[i1*N2 + rank*C + c], C = N2/world; result block [N1/world][N2] with
 * rows[r][k2] = X[(rank*N1/world + r) + N1*k2]. */
typedef struct kzgb_ntt4_plan kzgb_ntt4_plan;
KZGB_API int kzgb_ntt4_plan_create(uint32_t log_n, uint32_t log_n1, kzgb_ntt4_plan** out);
KZGB_API int kzgb_ntt4_plan_destroy(kzgb_ntt4_plan* plan);
KZGB_API int kzgb_ntt4_plan_shape(const kzgb_ntt4_plan* plan, uint32_t* log_n1, uint32_t* log_n2, uint32_t* log_c);
KZGB_API int kzgb_ntt_fr_sharded_dev(kzgb_ntt4_plan* plan, const void* d_slab, int inverse, void** d_rows_out, void* stream);
/* host vectors in natural order (every rank passes the full input; gather != 0: every rank receives the full output,
 * else only its own elements of `out` are written) */
KZGB_API int kzgb_ntt_fr_sharded(kzgb_ntt4_plan* plan, const kzgb_fr* in, kzgb_fr* out, int inverse, int gather);

/* ---- EvaluationDomain helpers off the prover path: element-wise passes and the coset transforms.
 * op 0: a[i] *= b[i]  (mul_assign, ft.rs:220-244)      op 1: a[i] -= b[i]  (sub_assign, ft.rs:247-271)
 * op 2: a[i] *= k^i   (distribute_powers, ft.rs:142-166)  op 3: a[i] *= k (e.g. divide_by_z_on_coset, ft.rs:192-217) */
KZGB_API int kzgb_fr_vec_op(int op, kzgb_fr* a, const kzgb_fr* b, size_t n, const kzgb_fr* k);
KZGB_API int kzgb_fr_vec_op_dev(int op, void* d_a, const void* d_b, size_t n, const kzgb_fr* k, void* stream);
/* coset_fft (ft.rs:168-171: distribute_powers(7) then fft) / icoset_fft (ft.rs:173-178: ifft then distribute_powers(7^-1)) */
KZGB_API int kzgb_coset_ntt_fr(kzgb_fr* inout, uint32_t log_n, int inverse);
/* transform pipeline of Polynomial::fft_mul (polynomial.rs:167-183): pad to m = 2^k >= na + nb, fft, multiply, ifft */
KZGB_API int kzgb_poly_mul(const kzgb_fr* a, size_t na, const kzgb_fr* b, size_t nb, kzgb_fr* out, size_t out_capacity,
                           size_t* m_out);

/* ---- wire / disk encodings: ZCash/IETF compressed G1 (48 B; blstrs to_compressed /
 * from_compressed_unchecked: on-curve checked, subgroup membership not) and canonical little-endian Fr
 * (blstrs to_bytes_le / from_bytes_le).  The reference's own serde support is dead code (lib.rs:13, ft.rs:389-404). */
KZGB_API int kzgb_g1_compress(const kzgb_p1_affine* points, size_t n, uint8_t* out48);
KZGB_API int kzgb_g1_decompress(const uint8_t* in48, size_t n, kzgb_p1_affine* out);
/* SRS import is checked like blstrs `from_compressed` (well-formed, on the curve, in the prime-order subgroup: [r]P = O,
 * ~0.13 s per 2^20 points, once); the _unchecked variant skips the subgroup test for pre-validated input. */
KZGB_API int kzgb_srs_create_g1_compressed(const uint8_t* in48, size_t n, int window_bits, kzgb_srs** out);
KZGB_API int kzgb_srs_create_g1_compressed_unchecked(const uint8_t* in48, size_t n, int window_bits, kzgb_srs** out);
KZGB_API int kzgb_g1_subgroup_check(const kzgb_p1_affine* points, size_t n, int* all_in_g1);
KZGB_API int kzgb_srs_export_compressed(const kzgb_srs* srs, size_t start, size_t count, uint8_t* out48);
KZGB_API int kzgb_fr_bytes_le(const void* in, void* out, size_t n, int to_bytes);

/* ---- synthetic inputs: splitmix64 Fr stream, Montgomery form, on the device ----
 * element i of stream `seed` = 4 consecutive outputs as a 256-bit LE integer mod r (mode 0), or a single
 * u64 (mode 1, the distribution of benches/commit_coeff_form.rs:16-20). */
KZGB_API int kzgb_fr_stream_dev(void* d_out, uint64_t seed, uint64_t start, size_t n, int mode, void* stream);

/* ---- test hooks: element-wise device field arithmetic on host vectors ------- */
/* op: 0 a*b, 1 a+b, 2 a-b, 3 a^-1, 4 to_mont(a), 5 from_mont(a), 6 a^2 (b may be NULL for op >= 3).
 * Pins the device Fr/Fp arithmetic (which replaces blst's, Cargo.toml:27) against the oracle. */
KZGB_API int kzgb_test_fr_op(const kzgb_fr* a, const kzgb_fr* b, kzgb_fr* out, size_t n, int op);
KZGB_API int kzgb_test_fp_op(const kzgb_fp* a, const kzgb_fp* b, kzgb_fp* out, size_t n, int op);

#ifdef __cplusplus
}
#endif
#endif /* KZG_B200_H */
