/*
 * proxb200.h -- C ABI of libproxb200.so: the B200-native proximal-gradient hot path.
 *
 * The reference (ProximalAlgorithms.jl v0.7.0, pure Julia) has NO FFI boundary of its own; its plug-in surface
 * is method extension on array/functional types.  The entry points below are exactly what a
 * `ccall` shim specialising the reference's iterators on a device-vector type binds; every function cites the
 * reference arithmetic it replaces (paths relative to the reference tree, abbreviated:
 *   FFB = src/algorithms/fast_forward_backward.jl, FB = src/algorithms/forward_backward.jl,
 *   FBT = src/utilities/fb_tools.jl, BM = benchmark/benchmarks.jl).
 * The reference-side binding a maintainer would add is shown in INTEGRATION.md and shipped in
 * proximalalgorithms.jl_b200/julia/B200Prox.jl.
 *
 * Conventions
 *   - plain C: pointers, sizes, doubles.  No torch / C++ types cross this boundary.
 *   - every function returns 0 on success, a PB_E* code otherwise; pb_last_error() gives the message (thread local).
 *   - device pointers are raw CUDA device addresses on the context's device (from pb_malloc, cudaMalloc or
 *     tensor.data_ptr()); `dtype` selects float / double for every vector argument of the call.
 *   - kernels are enqueued on the context's stream and return immediately.  Reduction results are left in the context's
 *     device scalar block (PB_NSCALARS doubles); pb_read_scalars() copies it to the host and synchronises -- that is the
 *     one host sync per iteration the reference's driver loop needs (src/ProximalAlgorithms.jl:116-120).
 *   - scalar parameters (gamma, beta, lambda...) are passed as double and narrowed to `dtype` on entry; callers that
 *     keep their scalars in R = real(eltype(x0)) (as the reference does) lose nothing in the widening.
 *   - one host thread per context.  The library never calls back into the host language.
 */
#ifndef PROXB200_H
#define PROXB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB_VERSION_STRING "0.1.0"

/* ---- status codes -------------------------------------------------------------------------------------------- */
enum {
  PB_OK = 0,
  PB_EINVAL = 1,   /* bad argument (null pointer, negative size, unknown dtype/prox kind, misaligned group...) */
  PB_ECUDA = 2,    /* a CUDA runtime call or kernel launch failed; message carries cudaGetErrorString */
  PB_ENOMEM = 3,   /* host or device allocation failed */
  PB_EUNSUPPORTED = 4
};

/* ---- element types --------------------------------------------------------------------------------------------- */
enum { PB_F32 = 0, PB_F64 = 1 };

/* ---- proximable terms (ProximalOperators.jl semantics; call sites FFB:80,141  FB:72,118  FBT:49) ---------------- */
enum {
  PB_PROX_ZERO = 0,    /* ProximalCore.Zero: identity prox, value 0                                   */
  PB_PROX_L1 = 1,      /* NormL1(lambda): soft threshold, value lambda*||z||_1   (BM:52,60)               */
  PB_PROX_BOX = 2,     /* IndBox(lo,hi): clamp, value 0  (test/problems/test_nonconvex_qp.jl:19,33)       */
  PB_PROX_SCALE = 3,   /* z = s*y with a caller-supplied factor: phase 2 of IndBallL2 (s = min(1, r/||y||)) */
  PB_PROX_L21 = 4,     /* NormL21(lambda, dim=1) on contiguous groups of `group` elements                  */
  PB_PROX_SQRL2 = 5,   /* Translate(SqrNormL2(lambda), -b): f = lambda/2*||x - b||^2, b = v0 or 0
                          (test/problems/test_lasso_small.jl:38); prox z = (y - b)/(1 + gamma*lambda) + b           */
  PB_PROX_BALL = 6     /* IndBallL2(r), p0 = r: z = y if ||y|| <= r else y*r/||y||, value 0.  pb_fb_step / pb_ffb_step / pb_solve only, one
                          GPU: the step enqueues a reduction pass (||x - gamma*grad||^2 -> AUX3 slot) and the fused kernel forms the scale
                          factor r/||y|| from that slot ON THE DEVICE -- no host round trip between the two phases.  Row-sharded iterates
                          need the norm combined across ranks first: use pb_forward + exchange + PB_PROX_SCALE (PB_EUNSUPPORTED here) */
};

typedef struct pb_prox {
  int32_t kind;        /* PB_PROX_*                                                                        */
  int32_t group;       /* PB_PROX_L21: group length (elements); otherwise ignored                          */
  double p0;           /* L1/L21/SQRL2: lambda;  BOX: lo (used when v0 == NULL);  SCALE: s                  */
  double p1;           /* BOX: hi (used when v1 == NULL)                                                   */
  const void* v0;      /* BOX: optional per-element lower bounds (device, same dtype as the vectors);
                          SQRL2: optional translation vector b                                              */
  const void* v1;      /* BOX: optional per-element upper bounds                                            */
} pb_prox;

/* ---- device scalar block layout --------------------------------------------------------------------------------
 * Sums are accumulated as double-double (hi, lo) with error-free transformations so that the rounded result does not
 * depend on grid size or on how a vector is sharded across GPUs (hard part 1).  A consumer adds the
 * pairs of all shards in double-double and rounds once: value = hi + lo.
 */
enum {
  PB_S_GSUM = 0,      /* [0],[1]  hi,lo of  sum|z_i| (L1),  sum_g scal_g*||y_g|| (L21), 0 otherwise; g(z) = lambda*that */
  PB_S_RESSQ = 2,     /* [2],[3]  sum (x_i - z_i)^2                 -> norm(res)^2 of FBT:4                         */
  PB_S_GDR = 4,       /* [4],[5]  sum grad_i*(x_i - z_i)            -> dot(grad_f_x, res) of FBT:4                  */
  PB_S_RESINF = 6,    /* [6]      max |x_i - z_i| (NaN propagates)  -> norm(res, Inf) of FB:125-126, FFB:147-152    */
  PB_S_AUX = 8,       /* [8],[9]  kernel-specific sum: ||y||^2 (pb_forward), ||v||^2 (pb_nrm2sq), <a,b> (pb_dot),
                                  ||A x - b||^2 (pb_lsq_*_residual), ||x - b||^2 (pb_sqdist)                        */
  PB_S_AUXINF = 10,   /* [10]     kernel-specific max: max|v_i| (pb_norm_inf)                                       */
  PB_S_AUX2 = 12,     /* [12],[13] <s,y> of pb_lbfgs_update (own slots: the update can be enqueued behind a fused step   */
  PB_S_AUX3 = 14,     /* [14],[15] <y,y> of pb_lbfgs_update  and an f evaluation and read back in ONE exchange)          */
  PB_NSCALARS = 16
};

typedef struct pb_ctx pb_ctx;

/* ---- library / context ----------------------------------------------------------------------------------------- */
const char* pb_version(void);
const char* pb_last_error(void);
int pb_device_count(int* count);

/* Create a context on CUDA device `device`.  borrow_stream != 0: `stream` is an existing cudaStream_t the context
 * borrows (e.g. torch's current stream; NULL is the legacy default stream).  borrow_stream == 0: `stream` is ignored
 * and the context creates and owns a non-blocking stream. */
int pb_ctx_create(int device, void* stream, int borrow_stream, pb_ctx** out);
int pb_ctx_destroy(pb_ctx* ctx);
/* Make the context's device the calling thread's current CUDA device.  Kernels are launched on the context's stream, which belongs
 * to that device: a host thread that drives several contexts on different GPUs (or a fresh thread, which starts on device 0) calls
 * this before using a context.  pb_solve / pb_panoc_solve do it themselves (and restore the previous device). */
int pb_ctx_make_current(pb_ctx* ctx);
void* pb_ctx_stream(pb_ctx* ctx);
int pb_ctx_sync(pb_ctx* ctx);
/* Device address of the PB_NSCALARS-double scalar block (so a host framework can all-gather it in place). */
double* pb_ctx_scalars_dev(pb_ctx* ctx);
/* Redirect the scalar block to caller-owned device memory (PB_NSCALARS doubles, e.g. a framework tensor that is the
 * send buffer of the per-iteration all-gather).  NULL restores the context's own block. */
int pb_ctx_set_scalars_dev(pb_ctx* ctx, double* dev);
/* Tuning knobs of the streaming kernels (results never depend on them; tests sweep them to prove it). */
enum {
  PB_OPT_CTAS_PER_SM = 0,   /* CTAs per SM of the grid-stride kernels; 0 = per-kernel default                      */
  PB_OPT_STREAM_HINTS = 1,  /* ld/st cache-streaming hints: -1 = auto (on when the working set exceeds L2), 0, 1   */
  PB_OPT_UNROLL = 2,        /* 16-byte packs in flight per thread and input stream: 0 = default, or 1, 2, 4, 8       */
  PB_OPT_STEP_IMPL = 3,     /* fused step implementation: 0 = default, 1 = register (LDG) pipeline, 2 = TMA bulk-copy
                               shared-memory ring                                                                 */
  PB_OPT_FUSED_EXCHANGE = 4,/* 1: pb_fb_step / pb_ffb_step also perform the per-iteration exchange (pb_xchg_*) in their
                               last CTA, so the iteration needs no collective launch, memcpy or stream sync       */
  PB_OPT_MULTI_ITER = 6,    /* pb_solve, fixed-stepsize FFB with an element-wise gradient source (LinearFunction, SquaredDistance):
                               0 = auto (ONE persistent kernel loops over the iterations, csrc/step_multi.cu; off when contexts
                               of one process share a GPU), -1 = never (one launch per iteration), 1 = always.  Same results */
  PB_OPT_LSQ_FISTA = 9,     /* pb_solve, fixed-stepsize FFB on a block-diagonal least-squares term: 0 = auto (ONE sweep of A per iteration:
                               gradient, fused step and the next residual's partial products from the same tiles, csrc/lsq_fista.cu, when A
                               exceeds 64 MB), -1 = never (residual + gradient + step kernels), 1 = whenever the shape allows.  Same bits   */
  PB_OPT_LSQ_FUSED = 8,     /* pb_lsq_blockdiag_value_and_gradient: 0 = auto (ONE persistent kernel that reads every block of A from HBM
                               once and its second sweep from L2 when A exceeds L2 and a block fits it, csrc/lsq_fused.cu), -1 = never
                               (residual kernel + gradient kernel), k = 1..8: always, keeping k blocks between the sweeps. Same bits */
  PB_OPT_GEMV_SCALAR = 7,   /* 1: r = A x - b with the thread-per-row kernel (4-byte loads) instead of 16-byte row packs; same bits     */
  PB_OPT_PERSISTENT = 5     /* pb_solve on cache-resident dense least squares (m*n*sizeof <= 8 MB): 0 = auto (whole solve
                               in one persistent cooperative kernel, csrc/persist.cu), -1 = never (one kernel per
                               operation), 1..32 = persistent on at most that many CTAs.  Results do not depend on it */
};
int pb_ctx_set_option(pb_ctx* ctx, int option, int value);
/* Number of kernels this context has launched since creation (bench.py reports it as gpu_launches). */
int64_t pb_ctx_launch_count(pb_ctx* ctx);

/* ---- C1: per-iteration exchange of the scalar block, by the GPU itself --------------------------------------------
 * Replaces "NCCL all-gather + cudaMemcpy + stream synchronise" per iteration: each rank's reducing kernel pushes its
 * PB_NSCALARS doubles to every peer over NVLink (peer stores into cudaIpc-mapped buffers), waits for the peers' rows and
 * copies all rows into mapped pinned host memory; the host polls a flag there.  With world == 1 it is a zero-copy
 * read-back.  Set-up: every rank calls pb_xchg_init (gets a PB_IPC_HANDLE_BYTES handle), the host framework all-gathers
 * the handles (any transport), every rank calls pb_xchg_connect.  All ranks must issue the same sequence of exchanges. */
#define PB_IPC_HANDLE_BYTES 64
#define PB_MAX_WORLD 8
int pb_xchg_init(pb_ctx* ctx, int rank, int world, void* handle_out);
int pb_xchg_connect(pb_ctx* ctx, const void* all_handles /* world x PB_IPC_HANDLE_BYTES, rank order */);
/* Single-process form (one host process drives all GPUs, one host thread per context): the n contexts
 * belong to this process, ctxs[r] initialised with pb_xchg_init(ctxs[r], r, n, NULL); peers are reached by peer access. */
int pb_xchg_connect_local(pb_ctx** ctxs, int n);
int pb_xchg_shutdown(pb_ctx* ctx);
/* Enqueue a stand-alone exchange of the current scalar block (for reads that do not follow a fused step). */
int pb_exchange(pb_ctx* ctx);
/* Wait for the most recent exchange (launching one if none is pending) and return the world x PB_NSCALARS rows in rank
 * order.  Polls pinned memory: no CUDA call on the fast path. */
int pb_exchange_wait(pb_ctx* ctx, double* rows_out, double timeout_s);

/* ---- memory (lets a host without a CUDA binding own device vectors: Julia `B200Vector`) -------------------------- */
int pb_malloc(pb_ctx* ctx, size_t bytes, void** dptr);
int pb_free(pb_ctx* ctx, void* dptr);
int pb_host_alloc(size_t bytes, void** hptr);                /* pinned host memory */
int pb_host_free(void* hptr);
int pb_upload(pb_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);    /* async on the ctx stream */
int pb_download(pb_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);  /* synchronises */
int pb_copy(pb_ctx* ctx, void* dst_dev, const void* src_dev, size_t bytes);       /* `copy`, FFB:74, FB:66 */
int pb_memset_zero(pb_ctx* ctx, void* dst_dev, size_t bytes);                     /* `zero(x)` */
/* out[i] = scale * u(offset + i), u in [-1, 1) a counter-based (splitmix64) function of the GLOBAL index: a row shard holds
 * exactly the values of the unsharded vector (bench.py: identical synthetic data for every number of GPUs and for the CPU
 * port, oracle/c/fb_port.c: port_fill). */
int pb_fill_counter(pb_ctx* ctx, int dtype, int64_t n, int64_t offset, uint64_t seed, double scale, void* out);
/* Copy the device scalar block to `out` (PB_NSCALARS doubles) and synchronise the stream. */
int pb_read_scalars(pb_ctx* ctx, double* out);

/* ---- K1: fused forward-backward step ------------------------------------------------------------------------------
 * One pass:  y = x - gamma*grad (FB:117, FFB:140, FBT:48);  z = prox_{gamma g}(y) (FB:118, FFB:141, FBT:49);
 * res = x - z (FB:120, FFB:142, FBT:50);  reductions GSUM, RESSQ, GDR (f_model, FBT:3-5) and RESINF (stop rule).
 * `y` and `res` may be NULL (not materialised).  mul and add are rounded separately (no FMA contraction), as Julia's
 * broadcast does.  Algorithmic traffic: 3 vectors (read x, grad; write z). */
int pb_fb_step(pb_ctx* ctx, int dtype, int64_t n, const void* x, const void* grad, double gamma, const pb_prox* g,
               void* y, void* z, void* res);

/* ---- K2: K1 plus the next iteration's extrapolation in the same pass ---------------------------------------------
 * x_next = z + beta*(z - z_prev) (FFB:135), where z_prev is the previous forward-backward point.  x_next must not
 * alias x.  Algorithmic traffic: 5 vectors (read x, grad, z_prev; write z, x_next). */
int pb_ffb_step(pb_ctx* ctx, int dtype, int64_t n, const void* x, const void* grad, const void* z_prev, double gamma,
                double beta, const pb_prox* g, void* y, void* z, void* res, void* x_next);

/* ---- K3: standalone prox (init `prox(g, y, gamma)` FFB:80, FB:72; user-driven splittings) ------------------------- */
int pb_prox_apply(pb_ctx* ctx, int dtype, int64_t n, const void* y, double gamma, const pb_prox* g, void* z);
/* y = x - gamma*grad and AUX = ||y||^2 : phase 1 of IndBallL2 (needs the global norm before scaling). */
int pb_forward(pb_ctx* ctx, int dtype, int64_t n, const void* x, const void* grad, double gamma, void* y);

/* ---- K6: vector utilities (unfused path, lazy state fields) ------------------------------------------------------- */
int pb_extrapolate(pb_ctx* ctx, int dtype, int64_t n, const void* z, const void* z_prev, double beta, void* x); /* FFB:135 */
int pb_residual(pb_ctx* ctx, int dtype, int64_t n, const void* x, const void* z, const void* grad_or_null,
                void* res);                             /* res = x - z with RESSQ, RESINF (+GDR when grad given) */
int pb_add_scalar(pb_ctx* ctx, int dtype, int64_t n, const void* x, double c, void* out);   /* x .+ 1, FBT:9 */
int pb_sub(pb_ctx* ctx, int dtype, int64_t n, const void* a, const void* b, void* out);     /* a - b;  AUX = ||a-b||^2 (FBT:11) */
int pb_nrm2sq(pb_ctx* ctx, int dtype, int64_t n, const void* v);                             /* AUX = sum v^2, AUXINF = max|v| */
int pb_dot(pb_ctx* ctx, int dtype, int64_t n, const void* a, const void* b);                 /* AUX = sum a*b */

/* ---- K4: least-squares smooth term  f(x) = 0.5*||A x - b||^2  (BM:11-17) ------------------------------------------
 * dense: A column-major m x n with leading dimension lda (Julia Matrix).  residual: r = A x - b, AUX = ||r||^2.
 * gradient: grad = A' r.  `b` may be NULL (r = A x: partial product of a column shard). */
int pb_lsq_dense_residual(pb_ctx* ctx, int dtype, int64_t m, int64_t n, const void* A, int64_t lda, const void* x,
                          const void* b, void* r);
int pb_lsq_dense_gradient(pb_ctx* ctx, int dtype, int64_t m, int64_t n, const void* A, int64_t lda, const void* r,
                          void* grad);
/* Column shard of a dense m x n_global matrix on a rank of the device exchange (pb_xchg_init / pb_xchg_connect[_local]); SURVEY.md
 * section 8e "Dense-A gradient under this partition".  A: this rank's columns [col_offset, col_offset + n_local), x: its slice.
 * ONE kernel combines the rank's chunk partials, pushes them to every peer over NVLink and folds ALL chunks in global chunk order:
 * r = A x - b and AUX = ||r||^2 are replicated on every rank and bit-identical to the single-GPU pb_lsq_dense_residual on the whole
 * matrix.  Shard boundaries must be multiples of pb_lsq_dense_chunk_cols(dtype, m, n_global) (the last shard takes the remainder; empty
 * shards are allowed).  flags & 1: AUX is written on rank 0 only (0 elsewhere) -- for callers that sum AUX over the ranks.
 * grad = A_p' r is then the local pb_lsq_dense_gradient.  Every rank must make the same sequence of sharded calls. */
int pb_lsq_dense_residual_sharded(pb_ctx* ctx, int dtype, int64_t m, int64_t n_local, const void* A, int64_t lda, const void* x,
                                  const void* b, void* r, int64_t n_global, int64_t col_offset, int flags);
int64_t pb_lsq_dense_chunk_cols(int dtype, int64_t m, int64_t n);
/* block-diagonal ("implicit A via batched GEMV", BASELINE.json configs[1]): nblk blocks, block k is a column-major
 * mb x nb matrix at A + k*mb*nb; x has nblk*nb entries, r/b have nblk*mb. */
int pb_lsq_blockdiag_residual(pb_ctx* ctx, int dtype, int64_t nblk, int64_t mb, int64_t nb, const void* A,
                              const void* x, const void* b, void* r);
int pb_lsq_blockdiag_gradient(pb_ctx* ctx, int dtype, int64_t nblk, int64_t mb, int64_t nb, const void* A,
                              const void* r, void* grad);
/* value AND gradient of the block-diagonal term in one call: r = A x - b, AUX = ||r||^2, grad = A' r, bit-identical to
 * pb_lsq_blockdiag_residual + pb_lsq_blockdiag_gradient.  For matrices beyond L2 whose blocks fit it, one persistent kernel streams every
 * block from HBM once (chunk partials), assembles r_k, and re-reads the block for A_k' r_k while it is still L2 resident
 * (csrc/lsq_fused.cu): ~1 sweep of HBM traffic per value_and_gradient instead of 2 (BM:11-17 per block). */
int pb_lsq_blockdiag_value_and_gradient(pb_ctx* ctx, int dtype, int64_t nblk, int64_t mb, int64_t nb, const void* A, const void* x,
                                        const void* b, void* r, void* grad);
/* SquaredDistance (BM:19-28): grad = x - b, AUX = ||x - b||^2. */
int pb_sqdist(pb_ctx* ctx, int dtype, int64_t n, const void* x, const void* b, void* grad);

/* ---- K7: L-BFGS directions and the vector algebra of PANOC's line search ------------------------------------------
 * (src/accel/lbfgs.jl, src/algorithms/panoc.jl; abbreviated LB, PN below).  Single GPU: the two-loop recursion keeps its
 * coefficients on the device (no host round trip between its 2m+1 launches), which needs un-sharded dot products. */
int pb_lincomb2(pb_ctx* ctx, int dtype, int64_t n, double a, const void* x, double b, const void* y,
                void* out);                               /* out = a.*x .+ b.*y  (PN:183-184, :214-215, :234-237) */
int pb_scale(pb_ctx* ctx, int dtype, int64_t n, double s, const void* x, void* out);   /* out = s.*x  (PN:116,120) */

typedef struct pb_lbfgs pb_lbfgs;          /* LBFGSOperator{M} (LB:5-28): ring of M (+1 spare) pairs on the device */
int pb_lbfgs_create(pb_ctx* ctx, int dtype, int64_t n, int mem, pb_lbfgs** out);        /* initialize, LB:102-104 */
int pb_lbfgs_destroy(pb_lbfgs* op);
int pb_lbfgs_reset(pb_lbfgs* op);                                                       /* reset!, LB:53-56      */
int pb_lbfgs_info(const pb_lbfgs* op, int* currmem, int* curridx, double* H);
int pb_lbfgs_pair(const pb_lbfgs* op, int pos, void** s, void** y, double* ys);        /* ring position 1..M    */
/* update!(L, s, y) (LB:30-51) fused with PANOC's differences (PN:125-126): s = a - a_prev, y = b - b_prev (a NULL
 * "prev" means s = a / y = b) are written into the ring's spare slot in one pass with AUX2 = <s,y>, AUX3 = <y,y>.
 * Asynchronous.  The caller reads the two sums with its per-iteration exchange and calls pb_lbfgs_commit, which does
 * the `if ys > 0` bookkeeping (LB:34-48) in the element type. */
int pb_lbfgs_update(pb_ctx* ctx, pb_lbfgs* op, const void* a, const void* a_prev, const void* b, const void* b_prev);
int pb_lbfgs_commit(pb_lbfgs* op, double ys, double yty, int* accepted);
/* mul!(d, L, v) (LB:66-95) followed by d .*= scale (PN:116 uses -1), and optionally x_d = x + d (PN:183) in the last
 * launch (x, x_d both NULL to skip).  d may alias v.  Asynchronous: 2*currmem + 2 launches, no host read-back. */
int pb_lbfgs_apply(pb_ctx* ctx, pb_lbfgs* op, const void* v, double scale, void* d, const void* x, void* x_d);

/* ---- K8: fused Douglas-Rachford iteration (src/algorithms/douglas_rachford.jl:54-63) -----------------------------
 * y = prox_{gamma f}(x); r = 2y - x; z = prox_{gamma g}(r); res = y - z; x_out = x - res in ONE pass, with RESINF =
 * norm(res, Inf) (stop rule, :65-69).  f and g must be element-wise kinds (ZERO, L1, BOX, SQRL2);
 * PB_EUNSUPPORTED otherwise.  x_out may alias x.  y, r, z, res are optional outputs (NULL = not materialised).
 * Algorithmic traffic: 2 vectors (read x, write x_out), +1 for the data vector of SQRL2. */
int pb_dr_step(pb_ctx* ctx, int dtype, int64_t n, const void* x, double gamma, const pb_prox* f, const pb_prox* g,
               void* x_out, void* y, void* r, void* z, void* res);

/* prox!(out, convex_conjugate(h), v, gamma) by the Moreau identity out = v - gamma*prox_{h/gamma}(v/gamma) (ProximalCore;
 * call site src/algorithms/primal_dual.jl:195) for the element-wise kinds; Zero* = IndZero gives out = 0. */
int pb_conj_prox(pb_ctx* ctx, int dtype, int64_t n, const void* v, double gamma, const pb_prox* h, void* out);

/* ---- K10: Douglas-Rachford iteration of anisotropic TV denoising in consensus form (BASELINE.json configs[4]) -------
 * minimize 0.5*||u - b||^2 + lambda*TV(u) split into five terms (data, even/odd horizontal pairs, even/odd vertical
 * pairs) on five stacked copies x[5][H][W] (row-major images); one call = one whole iteration of
 * douglas_rachford.jl:54-63 with F = separable sum of the five proxes and G = consensus (average), in ONE pass:
 * reads x (and b), writes x_out (must not alias x), RESINF = max|res| over the five copies; y[5][H][W] and the consensus
 * image z[H][W] are optional outputs.  TV is not in the reference: the splitting is ours (csrc/tv_kernels.cu,
 * oracle/tv_oracle.py; parity unpinned).  Row shards: this call owns global rows [row0, row0 + H) of an Hglob-row image;
 * halo_prev / halo_next point to the ONE neighbouring row (of the copy whose vertical pair straddles the shard boundary)
 * inside the neighbour's x buffer, typically peer memory opened with pb_ipc_open (read over NVLink), NULL at the border. */
int pb_dr_tv_step(pb_ctx* ctx, int dtype, int64_t H, int64_t W, const void* x, const void* b, double gamma, double lambda,
                  void* x_out, void* y, void* z, int64_t row0, int64_t Hglob, const void* halo_prev,
                  const void* halo_next);
/* cudaIpc plumbing for buffers other ranks read directly (pb_malloc'ed memory only; handles are PB_IPC_HANDLE_BYTES). */
int pb_ipc_export(pb_ctx* ctx, void* dptr, void* handle_out);
int pb_ipc_open(pb_ctx* ctx, const void* handle, void** dptr);
int pb_ipc_close(pb_ctx* ctx, void* dptr);

/* ---- K11: forward differences of an H x W row-major image and the adjoint (anisotropic TV by Chambolle-Pock: the `L` of
 * src/algorithms/primal_dual.jl with h = lambda*||.||_1).  forward: out[0][i][j] = u[i][j+1] - u[i][j] (0 in the last column),
 * out[1][i][j] = u[i+1][j] - u[i][j] (0 in the last row); adjoint: the exact transpose.  out must not alias the input. */
int pb_fd2d_forward(pb_ctx* ctx, int dtype, int64_t H, int64_t W, const void* u, void* out /* 2*H*W */);
int pb_fd2d_adjoint(pb_ctx* ctx, int dtype, int64_t H, int64_t W, const void* pq /* 2*H*W */, void* out /* H*W */);

/* ---- K9: prox of the dense least-squares term (DouglasRachford's `f = LeastSquares(A, b)`, test_lasso_small.jl:39,205-214;
 * benchmark/benchmarks.jl:87-93).  ProximalOperators' LeastSquaresDirect restated: q = lambda*A'b + x/gamma; tall A:
 * y = (lambda*A'A + I/gamma)^-1 q; wide A: y = gamma*(q - lambda*A'((lambda*AA' + I/gamma)^-1 (A q))).  A (column-major m x n,
 * lda = m) and b are borrowed device arrays that must outlive the operator; min(m, n) <= 4096.  The factorisation is redone
 * (and the call synchronises) whenever gamma changes; otherwise apply is asynchronous and leaves AUX = ||A y - b||^2. */
typedef struct pb_lsqprox pb_lsqprox;
int pb_lsq_prox_create(pb_ctx* ctx, int dtype, int64_t m, int64_t n, const void* A, const void* b, double lambda,
                       pb_lsqprox** out);
int pb_lsq_prox_destroy(pb_lsqprox* op);
int pb_lsq_prox_apply(pb_ctx* ctx, pb_lsqprox* op, const void* x, double gamma, void* y);

/* ---- native driver loop ---------------------------------------------------------------------------------------------
 * The reference's IterativeAlgorithm loop (src/ProximalAlgorithms.jl:114-123) around ForwardBackward / FastForwardBackward
 * (forward_backward.jl:65-123, fast_forward_backward.jl:73-145) with the line search of fb_tools.jl:24-63, for the built-in
 * terms, executed inside the library: same kernel sequence and the same scalar arithmetic in R = real(eltype(x0)) as a
 * host-language driver, without its per-iteration interpreter cost.  Scalars are read through the device exchange when one
 * is attached to the context (any world size), else by memcpy. */
enum { PB_F_LSQ_DENSE = 0, PB_F_LSQ_BLOCKDIAG = 1, PB_F_SQDIST = 2, PB_F_LINEAR = 3 };
enum { PB_ALG_FB = 0, PB_ALG_FFB = 1 };
enum { PB_SEQ_ADAPTIVE = 0, PB_SEQ_FIXED = 1, PB_SEQ_SIMPLE = 2, PB_SEQ_CONSTANT = 3 };   /* src/accel/nesterov.jl */

typedef struct pb_smooth {
  int32_t kind;          /* PB_F_*                                                                              */
  int32_t pad;
  int64_t m, n, lda;     /* dense: A is m x n column-major                                                      */
  int64_t nblk, mb, nb;  /* block-diagonal: nblk column-major mb x nb blocks.  Dense COLUMN SHARD on a rank of the
                          * device exchange (world > 1): n = local columns, nb = n_global, nblk = col_offset            */
  const void* A;         /* device matrix (dense / block-diagonal)                                              */
  const void* b;         /* device: right-hand side (least squares), b (SquaredDistance) or c (LinearFunction)   */
  void* r;               /* device scratch for the residual: m or nblk*mb entries                               */
} pb_smooth;

typedef struct pb_solve_opts {
  int32_t algorithm;     /* PB_ALG_*                                                                            */
  int32_t adaptive;      /* backtracking line search on/off (fast_forward_backward.jl:50-51)                    */
  int32_t sequence;      /* PB_SEQ_* extrapolation sequence of FFB                                              */
  int32_t profile;       /* 1: time the loop and every fused-step launch with CUDA events (result->loop_ms, ...)  */
  int64_t maxit;
  int64_t n_global;      /* length of the whole (unsharded) iterate; 0 = n                                      */
  double tol;            /* stop when norm(res, Inf)/gamma <= tol (negative: never)                             */
  double gamma;          /* stepsize; <= 0: estimate 1/L with fb_tools.jl:7-12 (requires adaptive)              */
  double mf;             /* convexity modulus seeding AdaptiveNesterovSequence                                  */
  double constant_beta;  /* value of PB_SEQ_CONSTANT                                                            */
  double minimum_gamma, reduce_gamma, increase_gamma;
  /* Optional pipelining of the fixed-stepsize FFB loop (all three non-NULL, device exchange attached): the kernel of
   * iteration k+1 is launched BEFORE the host has received the scalars of iteration k, so the host-side stop test, the
   * exchange latency and the launch overhead hide behind the next kernel.  Iteration k+1 only reads what iteration k
   * wrote and writes into these spare n-vectors, so when the stop test of iteration k fires the state of iteration k is
   * intact (x, grad, z, z_prev as reported in pb_solve_result) and one speculative launch is discarded.  Results
   * (iterates, scalars, iteration count) are identical to the unpipelined loop. */
  void *spare_x, *spare_z, *spare_grad;
} pb_solve_opts;

typedef struct pb_solve_result {
  int64_t iterations;    /* k of the reference's driver loop (the init state counts as 1)                       */
  int64_t backtracks;
  double gamma, f_x, g_z, res_inf;
  int32_t warned_small_gamma;   /* fb_tools.jl:59-61 would have warned                                         */
  int32_t persistent_ctas;      /* > 0: the solve ran as ONE persistent kernel on that many CTAs (csrc/persist.cu)    */
  void *x, *grad, *z, *z_prev;  /* which of the caller's buffers hold the final state fields (buffers are swapped) */
  double loop_ms;               /* opts->profile: CUDA-event time from the first kernel of init to the last kernel  */
  double step_kernel_ms;        /* opts->profile: summed CUDA-event time of the timed fused-step launches            */
  int64_t step_kernel_launches; /* opts->profile: how many launches that sum covers (at most 4096)                   */
  /* rank-combined reductions of the final state's fused step, rounded once from their double-double sums: norm(res)^2,
   * dot(grad_f_x, res) (the f_model terms, fb_tools.jl:3-5) and sum|z| (L1) / sum_g scal_g*||y_g|| (L21).  Identical
   * for every sharding of the iterate -- bench.py prints them as its cross-N parity fingerprint. */
  double res_sq, gdr, gsum;
  int32_t multi_iter_kernel;    /* 1: the iterations ran inside ONE persistent kernel (csrc/step_multi.cu); step_kernel_launches
                                   then counts iterations and step_kernel_ms is that kernel's duration                        */
  int32_t pad;
} pb_solve_result;

/* x holds copy(x0) on entry.  grad, z, scratch: n-vectors.  z_prev, x_next: FFB only.  grad_z: adaptive FB only. */
int pb_solve(pb_ctx* ctx, int dtype, int64_t n, const pb_smooth* f, const pb_prox* g, const pb_solve_opts* opts, void* x,
             void* grad, void* z, void* z_prev, void* x_next, void* grad_z, void* scratch, pb_solve_result* result);

/* ---- native PANOC driver (src/algorithms/panoc.jl:88-259 around the kernels above; twin of the Python host panoc.py) -------
 * f is a built-in smooth term acting on A x (A = NULL: the identity; else dense column-major Am x n), g a single-pass prox kind,
 * directions L-BFGS with memory lbfgs_mem (0: NoAcceleration).  Work vectors are allocated and released by the call; the solution
 * state.z is copied to z_out.  Single GPU. */
typedef struct pb_panoc_opts {
  int64_t maxit;
  double tol;              /* stop when norm(res, Inf)/gamma <= tol (negative: never)                                */
  double alpha, beta;      /* panoc.jl:45-46 (0.95, 0.5)                                                           */
  double gamma;            /* stepsize; <= 0: alpha / lower_bound_smoothness_constant (requires adaptive)          */
  double minimum_gamma;
  int32_t adaptive, max_backtracks, lbfgs_mem, quadratic;   /* quadratic: ProximalCore.is_generalized_quadratic(f)  */
  int64_t Am, An;
  const void* A;
} pb_panoc_opts;

typedef struct pb_panoc_result {
  int64_t iterations, gamma_backtracks, tau_backtracks;
  double gamma, f_Ax, g_z, res_inf, tau;
  int32_t warned_small_gamma, pad;
} pb_panoc_result;

int pb_panoc_solve(pb_ctx* ctx, int dtype, int64_t n, const pb_smooth* f, const pb_prox* g, const pb_panoc_opts* opts,
                   const void* x0, void* z_out, pb_panoc_result* result);

/* Diagnostic of the persistent solver (csrc/persist.cu): clock64() cycles CTA 0 spent per phase in the last pb_solve that ran
 * persistently with opts->profile = 1.  out8: 0 A*v chunk partials, 1 grid barrier + scalar fold, 2 r = sum of partials - b and
 * ||r||^2, 3 A'r, 4 fused step, 5 everything else, 6-7 unused. */
int pb_persist_phase_cycles(pb_ctx* ctx, int64_t* out8);

/* ---- host-buffer convenience (the "plugin call with HOST buffers"): upload x, grad, z_prev, run K2, download z, x_next
 * and the scalar block.  All host pointers; temporary device buffers are cached in the context. */
int pb_ffb_step_host(pb_ctx* ctx, int dtype, int64_t n, const void* x, const void* grad, const void* z_prev,
                     double gamma, double beta, const pb_prox* g, void* z, void* x_next, double* scalars);

#ifdef __cplusplus
}
#endif
#endif /* PROXB200_H */
