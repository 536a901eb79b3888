/*
 * dplonk.h - C ABI of the B200-native worker hot path of MengLing-L/distributed_plonk.
 *
 * The reference worker (Rust) has no FFI today: its RPC method bodies call arkworks directly
 * (SURVEY.md §8b).  Each entry point below is what the body of one `PlonkSlave` / `PlonkPeer`
 * method (src/hello_world.capnp:15-52, implemented in src/worker.rs:125-439) binds instead of the
 * arkworks call it makes today; INTEGRATION.md shows the Rust `extern "C"` block and the patched
 * method bodies.  Wire formats are exactly what the reference puts in its `Data` blobs
 * (src/utils.rs:27-43 = raw in-memory Rust structs):
 *
 *   Fr            32 B  ark-ff Fp256, Montgomery form (R = 2^256), 4 x u64 little-endian
 *   BigInteger256 32 B  canonical scalar (Fr::into_repr), 4 x u64 little-endian
 *   G1Affine     104 B  x (48 B Fq Montgomery) | y (48 B) | infinity flag (1 B) | 7 B padding
 *   G1Projective 144 B  Jacobian X | Y | Z, each 48 B Fq Montgomery; identity has Z = 0
 *
 * Conventions: every function returns DP_OK (0) or a negative DP_E_* code and never throws or
 * aborts across the boundary; dp_last_error() gives the text.  All buffers are caller-owned,
 * borrowed only for the duration of the call, may be unaligned, and are HOST memory unless a
 * parameter says "dev".  One exception, for page-locked (cudaHostAlloc / cudaHostRegister) row
 * buffers handed to dp_fft1 / dp_fft1_rows: their copy-in is truly asynchronous, so they must stay
 * unchanged until dp_fft2 of that task (or dp_sync) returns; ordinary pageable memory - what a
 * Cap'n Proto message gives the reference worker - is staged before the call returns.  A context is bound to one CUDA device and must be used from one thread
 * at a time (the reference worker is single-threaded: worker.rs:441,453).  There is no CPU
 * fallback: dp_create fails with DP_E_CUDA when no sm_100 device is usable.
 */
#ifndef DPLONK_H
#define DPLONK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DP_OK 0
#define DP_E_ARG (-1)   /* bad argument (size mismatch, non-canonical scalar, unknown id ...) */
#define DP_E_STATE (-2) /* call out of order (fft2 before fft2_prepare, rows missing ...)      */
#define DP_E_OOM (-3)   /* device or host allocation failed                                   */
#define DP_E_CUDA (-4)  /* CUDA runtime / kernel error                                        */
#define DP_E_COMM (-5)  /* exchange needed but no peer transport attached                     */

#define DP_FR_BYTES 32
#define DP_G1_AFFINE_BYTES 104
#define DP_G1_COMPRESSED_BYTES 48 /* ark-serialize 0.3.0 compressed GroupAffine */
#define DP_G1_PROJECTIVE_BYTES 144

typedef struct dp_ctx dp_ctx;

/* src/utils.rs:3-19 FftWorkload (src/hello_world.capnp:8-13) */
typedef struct dp_fft_workload {
    uint64_t row_start, row_end, col_start, col_end;
} dp_fft_workload;

/* ---- lifetime -------------------------------------------------------------------------------
 * Replaces the `State` construction in worker main (src/worker.rs:455-472).  `me` is the worker
 * index (argv[1], worker.rs:443-449), `n_workers` the number of workers sharing distributed NTTs. */
int dp_create(int cuda_device, uint64_t me, uint64_t n_workers, dp_ctx **out);
int dp_destroy(dp_ctx *ctx);
const char *dp_last_error(const dp_ctx *ctx); /* valid until the next call on ctx; ctx may be NULL */
const char *dp_version(void);

/* ---- PlonkSlave.init (src/worker.rs:126-157) -------------------------------------------------
 * Stores the SRS bases on the device and builds the (r, c) split domains for `domain_size` and
 * `quot_domain_size` (Radix2EvaluationDomain::new rounds both up to powers of two).
 * bases: n_bases raw G1Affine (104 B each), the concatenation of the `bases` Data chunks; host memory, or
 * device memory of the context's GPU (a resident SRS: the copy is then device-to-device). */
int dp_init(dp_ctx *ctx, const void *bases, size_t n_bases, uint64_t domain_size, uint64_t quot_domain_size);

/* "next" row: the same, from the canonical encoding SRS files hold - what the
 * reference's dispatcher deserialises on the CPU before shipping raw structs (ark-serialize 0.3.0
 * `GroupAffine::deserialize` / `deserialize_unchecked`; jellyfish `UnivariateUniversalParams`).
 * bases48: n_bases x 48 B = canonical x little-endian, bit 7 of the last byte = (y > -y), bit 6 =
 * infinity.  Decompression (a 381-bit square root per point) and, when check_subgroup != 0, the
 * r-torsion check run on the GPU.  A rejected point (x >= p, both flags, no such point, outside the
 * subgroup) returns DP_E_ARG naming the first bad index and leaves the context uninitialised.     */
int dp_init_compressed(dp_ctx *ctx, const void *bases48, size_t n_bases, uint64_t domain_size, uint64_t quot_domain_size,
                       int check_subgroup);
/* bases [start, start + n) back as raw G1Affine (104 B each; identity = (0, 1, true))             */
int dp_get_bases(dp_ctx *ctx, uint64_t start, size_t n, void *out104);

/* ---- PlonkSlave.varMsm (src/worker.rs:159-185) -----------------------------------------------
 * out = sum_{k < min(end-start, n_scalars)} scalars[k] * bases[start + k]
 * (VariableBaseMSM::multi_scalar_mul(&bases[start..end], &scalars) truncates to the shorter).
 * scalars: canonical BigInteger256.  out: 144 B raw G1Projective, normalised (Z = 1) or identity. */
int dp_msm(dp_ctx *ctx, uint64_t start, uint64_t end, const void *scalars, size_t n_scalars, void *out);

/* Several varMsm requests issued together (the dispatcher joins the commitments of a round,
 * dispatcher2.rs:316-321, 526-532): scalars of request k+1 are copied in under the kernels of
 * request k, and the narrow tail kernels of k overlap the wide head kernels of k+1.  Arrays of
 * n_jobs entries, each with the semantics of dp_msm. */
int dp_msm_batch(dp_ctx *ctx, size_t n_jobs, const uint64_t *starts, const uint64_t *ends, const void *const *scalars,
                 const size_t *n_scalars, void *const *outs);

/* varMsm answered asynchronously (the Rust worker returns a Promise, as it already does for
 * fft2Prepare, src/worker.rs:293): dp_msm_submit queues the copy-in and the kernels behind whatever
 * the context is doing and returns; dp_msm_collect blocks until THAT job is done and writes the
 * 144-byte G1Projective.  `id` is the caller's (unique among pending jobs; at most 64 pending).
 * `scalars` must stay valid until the copy-in has run (dp_msm_collect, or dp_sync).  Lets
 * commitments share the GPU with transforms whose time goes into PCIe transfers.                 */
int dp_msm_submit(dp_ctx *ctx, uint64_t id, uint64_t start, uint64_t end, const void *scalars, size_t n_scalars);
int dp_msm_collect(dp_ctx *ctx, uint64_t id, void *out144);

/* ---- commit_polynomial (src/worker.rs:117-123) -----------------------------------------------
 * Fr::into_repr on every coefficient, zero-pad to bases.len(), MSM over all bases.
 * coeffs: n raw Fr (Montgomery), n <= n_bases. */
int dp_commit(dp_ctx *ctx, const void *coeffs, size_t n, void *out);

/* ---- PlonkSlave.fftInit (src/worker.rs:187-233) ----------------------------------------------
 * Opens task `id` (an open task with the same id is replaced, as `fft_tasks.insert` does).
 * workloads[w] is worker w's row / column range; n_workloads must equal n_workers and the ranges
 * must tile [0,r) x [0,c) in equal power-of-two blocks. */
int dp_fft_init(dp_ctx *ctx, uint64_t id, const dp_fft_workload *workloads, size_t n_workloads, int is_quot,
                int is_inv, int is_coset);

/* ---- PlonkSlave.fft1 (src/worker.rs:235-278) -------------------------------------------------
 * Hands in local row `i` (global row i + row_start).  len is normally c; a shorter row is zero-extended
 * and a longer one cut to c, which is what `c_domain.fft_in_place(&mut v)` does to `v` in the reference
 * (it resizes to the domain size) - a dispatcher may therefore leave out the zero tail of a padded
 * polynomial's rows.  The row transform (fft1_helper, worker.rs:66-94) runs on the device, at the latest
 * in dp_fft2_prepare. */
int dp_fft1(dp_ctx *ctx, uint64_t id, uint64_t i, const void *row, size_t len);
/* n_rows consecutive local rows in one call (rows = n_rows * c Fr, row-major) */
int dp_fft1_rows(dp_ctx *ctx, uint64_t id, uint64_t i_first, uint64_t n_rows, const void *rows);
/* the same for rows that are all cut after their first row_len <= c entries (rows = n_rows * row_len Fr,
 * compact): only the non-zero part of a padded polynomial's rows crosses PCIe, and the row kernel neither
 * reads the zero tail nor spends multiplications on it */
int dp_fft1_rows_short(dp_ctx *ctx, uint64_t id, uint64_t i_first, uint64_t n_rows, const void *rows, size_t row_len);

/* ---- PlonkSlave.fft2Prepare + PlonkPeer.fftExchange (src/worker.rs:280-345, 412-438) ---------
 * Finishes the row phase and moves every (rows_p x cols_q) block to its owner.  n_workers == 1:
 * purely local.  n_workers > 1: either peers were attached with dp_peer_attach (the blocks are
 * written straight into the owners' memory over NVLink by the row kernel), or the caller drives
 * the split API below around its own all-to-all (NCCL / torch.distributed). */
int dp_fft2_prepare(dp_ctx *ctx, uint64_t id);

/* split exchange: after _begin, *send_dev / *recv_dev are DEVICE pointers to n_workers blocks of
 * *block_elems Fr each (block q of send = my rows x q's columns, row-major = the payload of
 * fftExchange, worker.rs:327-330; block p of recv = p's rows x my columns).  The caller performs
 * the all-to-all (block q of send -> rank q, into block `me` of its recv) and calls _end. */
int dp_fft_exchange_begin(dp_ctx *ctx, uint64_t id, void **send_dev, void **recv_dev, uint64_t *block_elems);
int dp_fft_exchange_end(dp_ctx *ctx, uint64_t id);
/* Stream-ordered form of the same step, for a host that issues the exchange with its own NCCL
 * communicator: dp_fft_exchange_begin_async returns without waiting for the row phase; the two buffers
 * are then valid only for work enqueued on the context's compute stream (dp_compute_stream gives the
 * cudaStream_t), i.e. ncclGroupStart(); ncclSend / ncclRecv(..., stream) x W; ncclGroupEnd(); followed
 * by dp_fft_exchange_end.  Nothing blocks the host: copy-in, row kernels, all-to-all and column
 * kernels of consecutive tasks pipeline as they do for a single worker.                             */
int dp_fft_exchange_begin_async(dp_ctx *ctx, uint64_t id, void **send_dev, void **recv_dev, uint64_t *block_elems);
int dp_compute_stream(dp_ctx *ctx, void **stream);

/* ---- PlonkSlave.fft2 (src/worker.rs:347-381) -------------------------------------------------
 * Column transforms (fft2_helper, worker.rs:96-115); writes the local columns back to back
 * (n_cols * r Fr, column k at out + k*r*32 = the k-th `Data` of the reply) and drops the task. */
int dp_fft2(dp_ctx *ctx, uint64_t id, void *out, size_t out_bytes);

/* ---- whole-domain transform ------------------------------------------------------------------
 * Radix2EvaluationDomain::{fft,ifft,coset_fft,coset_ifft} on one device (round1's
 * ifft_in_place, worker.rs:398; the dispatcher-local coset_ifft, dispatcher2.rs:507).
 * data: n Fr in, 2^log_n Fr out (n <= 2^log_n, zero-padded like ark's resize); capacity of the
 * buffer must be 2^log_n elements. */
int dp_ntt(dp_ctx *ctx, void *data, size_t n, uint32_t log_n, int is_inv, int is_coset);

/* ---- PlonkSlave.round1 (src/worker.rs:383-408) -----------------------------------------------
 * evals (n Fr) -> ifft -> wire = (b0 + b1*X) * (X^n - 1) + poly -> commitment.  The reference
 * draws b0,b1 from the worker's ThreadRng (a CSPRNG): pass two secret, uniformly random Fr from the
 * host's own RNG in `blind` (2 raw Fr; also what makes the call reproducible in tests), or NULL to let
 * the library draw them from the operating system's entropy pool (getrandom(2), rejection-sampled
 * below r; DP_E_STATE if that source is unavailable).  The blinded polynomial stays resident
 * (state.wire, worker.rs:58) and can be read back with dp_get_wire. */
int dp_round1(dp_ctx *ctx, const void *evals, size_t n, const void *blind, void *out);
int dp_get_wire(dp_ctx *ctx, void *out, size_t out_bytes, size_t *n_coeffs);

/* ---- "next" row: the round-2 permutation grand product -----------------------
 * What the dispatcher computes serially with one field division per row (src/dispatcher2.rs:
 * 329-345): z[0] = 1, z[j+1] = z[j] * prod_i (w_i[j] + gamma + beta*id_i[j])
 *                                     / prod_i (w_i[j] + gamma + beta*sigma_i[j]),  j < n-1.
 * wires[i][j] = circuit.witness[wire_variables[i][j]], id_perm[i][j] = extended_id_permutation[i*n+j],
 * sigma_perm[i][j] = extended_id_permutation[perm_i*n+perm_j]; all [num_wire_types][n] raw Fr, as are
 * beta, gamma (one Fr each) and out (n Fr).  DP_E_ARG when a denominator is zero. */
int dp_perm_product(dp_ctx *ctx, const void *wires, const void *id_perm, const void *sigma_perm, size_t num_wire_types, size_t n,
                    const void *beta, const void *gamma, void *out);
/* same with device-resident inputs and output (beta, gamma stay host pointers to one Fr each) */
int dp_perm_product_dev(dp_ctx *ctx, const void *wires_dev, const void *id_dev, const void *sigma_dev, size_t num_wire_types, size_t n,
                        const void *beta, const void *gamma, void *out_dev);

/* ---- "next" row: rounds 3-5 of Prover::prove on polynomials resident on the worker
 *
 * The reference declares round3 / round4 / round5 RPCs (hello_world.capnp:26-44) but never implements
 * them: the dispatcher pulls every polynomial back and does this arithmetic serially
 * (src/dispatcher2.rs:363-690).  These entries are the bodies those RPCs would call.  The plain
 * entries take host buffers (copied in and out); the *_dev entries take device pointers for the
 * polynomial-sized arrays (challenges and points stay host-side, 32 B each).  All values raw
 * Montgomery Fr; results byte-identical to the reference's sequential code.                        */

typedef struct dp_quotient_args {
    const void *selectors[13]; /* q_lc[4], q_mul[2], q_hash[4], q_o, q_c, q_ecc (dispatcher2.rs:437-450) */
    const void *sigmas[5];
    const void *wires[5];
    const void *perm;      /* permutation product polynomial z                                          */
    const void *pub_input;
    /* ^ 25 arrays of quot_domain_size Fr: coset evaluations over the quotient domain (lines 381-432)  */
    const void *k;         /* vk.k, 5 Fr, host                                                          */
    const void *alpha, *beta, *gamma; /* 1 Fr each, host                                                */
} dp_quotient_args;

/* Round 3, src/dispatcher2.rs:434-504: out[i] = 1/Z_H(x_i) * (gate(i) + alpha * permutation(i)) +
 * alpha^2/n * (z[i] - 1)/(x_i - 1), x_i = g * omega_m^i, over the quotient domain given to dp_init
 * (GATE_WIDTH 4, five wire types); the input of the final coset iFFT (line 507, dp_ntt).          */
int dp_quotient_evals(dp_ctx *ctx, const dp_quotient_args *host_arrays, void *out);
int dp_quotient_evals_dev(dp_ctx *ctx, const dp_quotient_args *dev_arrays, void *out_dev);

/* Round 4, DensePolynomial::evaluate (src/dispatcher2.rs:535-548): out32 = sum_j coeffs[j] * point^j */
int dp_poly_eval(dp_ctx *ctx, const void *coeffs, size_t n, const void *point, void *out32);
int dp_poly_eval_dev(dp_ctx *ctx, const void *coeffs_dev, size_t n, const void *point, void *out32);

/* Round 5, the folds of src/dispatcher2.rs:566-649 (lin_poly, r_quot, batch_poly):
 * out[j] = sum_{i<k} coeffs[i] * polys[i][j], polys[i] zero-extended past lens[i]; k <= 32.
 * polys / lens / coeffs are host arrays; polys[i] points to host (plain) or device (_dev) memory.  */
int dp_poly_lincomb(dp_ctx *ctx, const void *const *polys, const size_t *lens, const void *coeffs, size_t k, void *out, size_t out_len);
int dp_poly_lincomb_dev(dp_ctx *ctx, const void *const *polys_dev, const size_t *lens, const void *coeffs, size_t k, void *out_dev,
                        size_t out_len);

/* Round 5, the opening witnesses (src/dispatcher2.rs:651-666, 672-688): the n-1 coefficients of
 * p(X) / (X - point) into out, and the remainder p(point) into rem32 when it is not NULL.
 * out_dev must not overlap coeffs_dev (DP_E_ARG): the division is not done in place.              */
int dp_poly_div_linear(dp_ctx *ctx, const void *coeffs, size_t n, const void *point, void *out, void *rem32);
int dp_poly_div_linear_dev(dp_ctx *ctx, const void *coeffs_dev, size_t n, const void *point, void *out_dev, void *rem32);

/* ---- worker-resident polynomials ---------------------------------------------------------------
 * `state.wire` of the reference (src/worker.rs:58,400-405) generalised: named device buffers the
 * *_dev entries above, dp_ntt_dev, dp_perm_product_dev and dp_commit_dev work on, so a polynomial
 * crosses PCIe once (or never: outputs of one step are inputs of the next).  A Rust worker has no
 * device allocator of its own; this is it.                                                          */
/* create / overwrite polynomial `poly_id`: `capacity` Fr on the device (>= n; e.g. the quotient domain
 * size for a polynomial that will be transformed in place), coefficients [0, n) copied from the host,
 * the rest zero.  Re-putting with the same capacity reuses the buffer.                              */
int dp_poly_put(dp_ctx *ctx, uint64_t poly_id, const void *coeffs, size_t n, size_t capacity);
int dp_poly_ptr(dp_ctx *ctx, uint64_t poly_id, void **dev, size_t *capacity); /* the device address         */
int dp_poly_get(dp_ctx *ctx, uint64_t poly_id, size_t offset, size_t n, void *out);    /* copy back        */
int dp_poly_free(dp_ctx *ctx, uint64_t poly_id);
/* commit_polynomial (src/worker.rs:117-123) of n coefficients already on the device -> 144 B        */
int dp_commit_dev(dp_ctx *ctx, const void *coeffs_dev, size_t n, void *out144);

/* ---- peer transport for n_workers > 1 ---------------------------------------------------------
 * Exchange arena shared between the GPUs of one box through CUDA IPC: every rank exports a
 * handle, the ranks swap them out of band (torch.distributed / the capnp control plane) and attach
 * the others'.  Afterwards dp_fft2_prepare stores the row-phase output straight into the owners'
 * arenas over NVLink and returns when its stores are complete; the caller provides the barrier across
 * ranks between fft2Prepare and fft2 - the dispatcher's join over the fft2Prepare replies is one.
 * Slots: the arena holds as many receive matrices ([r][c/n_workers] Fr of the LARGER domain given to dp_init) as fit
 * into arena_bytes, at least two are required; exchange number k uses slot k mod n_slots on every rank, so every rank
 * must issue its exchanges in the same order (as for a collective) and at most n_slots tasks per context may sit
 * between dp_fft2_prepare and dp_fft2: one more dp_fft2_prepare returns DP_E_STATE (nothing is consumed: call dp_fft2
 * on an earlier task, then retry - or use dp_fft_exchange_begin/_end, which has per-task buffers and no limit).
 * The slot sequence only advances on success.  A worker behind the reference dispatcher (up to 26 transforms in
 * flight, dispatcher2.rs:382-414) creates 32 slots; two are enough for one transform at a time.
 * arena_bytes >= n_slots * (r * c / n_workers) * 32 for the largest domain; the same on every rank. */
#define DP_IPC_HANDLE_BYTES 64
int dp_peer_arena_create(dp_ctx *ctx, uint64_t arena_bytes, void *handle_out /* DP_IPC_HANDLE_BYTES */);
int dp_peer_attach(dp_ctx *ctx, uint64_t peer, const void *handle /* DP_IPC_HANDLE_BYTES */);
int dp_peer_ready(const dp_ctx *ctx); /* 1 when the arena exists and every peer is attached */

/* ---- instrumentation --------------------------------------------------------------------------*/
/* device-side time (ms, CUDA events on the context stream) and kernel launches of the last call */
int dp_last_timing(const dp_ctx *ctx, float *kernel_ms, uint64_t *launches);
/* total kernel launches since dp_create */
uint64_t dp_launch_count(const dp_ctx *ctx);
/* block until everything queued on the context stream has finished */
int dp_sync(dp_ctx *ctx);

/* device-side ms of the three phases of the last MSM on ctx: digit sort (count/scan/scatter/tasks),
 * bucket accumulation (the dominant kernel), bucket reduction + window combine + normalise */
int dp_last_msm_breakdown(const dp_ctx *ctx, float *sort_ms, float *accumulate_ms, float *reduce_ms);

/* what dp_init's MSM tuning found: one MSM over the context's own window-multiple table with the plain pipeline
 * (XYZZ chunks) and one with 2 batched-affine tree levels in front of it (1, 2 and 3 with DP_MSM_TUNE=2; *affine_ms = the best);
 * *equal = all results were the same 144 bytes (1), some differed (0: the plain pipeline is kept), or the tuning did not
 * run (-1: DP_MSM_TUNE=0, small SRS, or DP_MSM_AFFINE set); *levels = what MSMs of this context use (0 = plain).  Replaces nothing in the reference: ark-ec has one algorithm. */
int dp_msm_tuning(const dp_ctx *ctx, float *plain_ms, float *affine_ms, int *levels, int *equal);
/* the same experiment in full: ms of one MSM with 0 (plain), 1, 2 and 3 tree levels (0 = not measured) */
int dp_msm_tuning_all(const dp_ctx *ctx, float ms_by_levels[4]);

/* synthetic SRS: n distinct points k_i*G (k_i = SplitMix64(seed, i)), raw 104-byte G1Affine each,
 * computed on the device and written to `out` (host memory, or device memory of the same GPU); feeds
 * dp_init in benches and tests */
int dp_debug_gen_bases(dp_ctx *ctx, uint64_t seed, size_t n, void *out);

/* test hook: lower the pass-planning limits (sub-transform sizes 2^k handled by one kernel pass;
 * defaults 11 / 9) and/or steer the MSM (0 = automatic: precomputed window multiples when the SRS is
 * large enough; 1 = per-window bucket sets with automatic width; c >= 2 = per-window sets of width c)
 * so that small inputs exercise the multi-pass NTT plans and every MSM geometry. */
int dp_debug_set_limits(dp_ctx *ctx, uint32_t max_contig_log_k, uint32_t max_strided_log_k, int msm_window_bits);
/* test hook: the single-worker three-pass transform plan is used for domains of at least
 * 2^min_log_n points (default 20); 0 switches it off (the 2-D row / column plan is used everywhere). */
int dp_debug_set_three_pass(dp_ctx *ctx, uint32_t min_log_n);

/* device-resident variants used by bench.py to time the kernels with inputs already in HBM.
 * All pointers are DEVICE pointers owned by the caller (e.g. torch tensors). */
int dp_msm_dev(dp_ctx *ctx, uint64_t start, uint64_t end, const void *scalars_dev, size_t n_scalars, void *out_dev);
/* several MSMs issued together (the commitments of one prover round, join_all in dispatcher2.rs:
 * 316-321, 526-532): the narrow tail kernels of MSM k overlap the wide head kernels of MSM k+1.
 * Arrays of n_jobs entries; same semantics per job as dp_msm_dev. */
int dp_msm_dev_batch(dp_ctx *ctx, size_t n_jobs, const uint64_t *starts, const uint64_t *ends, const void *const *scalars_dev,
                     const size_t *n_scalars, void *const *outs_dev);
int dp_ntt_dev(dp_ctx *ctx, void *data_dev, uint32_t log_n, int is_inv, int is_coset);
/* the same on a buffer whose entries from n_valid on are ZERO (a resident coefficient vector shorter than the
 * domain: n coefficients evaluated on the 8n-point coset, dispatcher2.rs:386-388): the forward transform neither
 * reads nor multiplies the zero tail.  wait = 0 returns once the kernels are queued (dp_sync waits).        */
int dp_ntt_dev_padded(dp_ctx *ctx, void *data_dev, size_t n_valid, uint32_t log_n, int is_inv, int is_coset, int wait);
/* full 2-D pipeline of one worker on device-resident rows: rows_dev = my rows (n_rows*c Fr),
 * cols_dev receives my columns (n_cols*r Fr); n_workers must be 1 or peers attached. */
int dp_fft_dev(dp_ctx *ctx, const void *rows_dev, void *cols_dev, int is_quot, int is_inv, int is_coset);
/* Promise about the rows given to the dp_fft_dev* entries of one domain (is_quot) from now on: every row is
 * zero from column valid_cols on - the shape of n coefficients on the 8n-point quotient domain, where
 * valid_cols = c/8 (dispatcher2.rs:386-388).  Forward transforms then neither read nor multiply the zero
 * tail (same effect as short rows through dp_fft1); inverse transforms ignore the promise.  0 = no promise. */
int dp_fft_dev_hint_valid_cols(dp_ctx *ctx, int is_quot, uint64_t valid_cols);
/* the same split around the caller's all-to-all for n_workers > 1 (one transform in flight per ctx):
 * _rows runs the row phase on my row block (rows_dev: r/W x c Fr) and returns the exchange buffers
 * exactly like dp_fft_exchange_begin; after the all-to-all _cols runs the column phase into
 * cols_dev (c/W x r Fr).  With n_workers == 1 recv == send and no exchange is needed. */
int dp_fft_dev_rows(dp_ctx *ctx, const void *rows_dev, int is_quot, int is_inv, int is_coset, void **send_dev,
                    void **recv_dev, uint64_t *block_elems);
int dp_fft_dev_cols(dp_ctx *ctx, void *cols_dev);
/* fused variant of _rows for attached peers: the row kernel stores into the owners' arenas; the
 * caller then only needs a barrier across ranks before dp_fft_dev_cols */
int dp_fft_dev_rows_p2p(dp_ctx *ctx, const void *rows_dev, int is_quot, int is_inv, int is_coset);
/* the whole distributed transform of one worker without host involvement: row kernels (stores into
 * the owners' arenas), a device-side barrier kernel (system-scope arrival counters in the arenas,
 * over NVLink), column kernels, queued back to back on the context stream; returns when done.
 * Every rank must call it for the same transforms in the same order (it blocks like a collective). */
int dp_fft_dev_p2p(dp_ctx *ctx, const void *rows_dev, void *cols_dev, int is_quot, int is_inv, int is_coset);
/* the same, returning as soon as the kernels are queued (rows_dev / cols_dev must stay valid until dp_sync,
 * which also reports a barrier time-out as DP_E_COMM).  Any number of transforms may be queued back to back:
 * the two arena slots are recycled in stream order behind the device-side barriers.                      */
int dp_fft_dev_p2p_async(dp_ctx *ctx, const void *rows_dev, void *cols_dev, int is_quot, int is_inv, int is_coset);

#ifdef __cplusplus
}
#endif
#endif /* DPLONK_H */
