// solve.cu -- native driver loop: the reference's IterativeAlgorithm loop (src/ProximalAlgorithms.jl:114-123) around the
// ForwardBackward / FastForwardBackward iterations (forward_backward.jl:65-123, fast_forward_backward.jl:73-145) and the
// line search (fb_tools.jl:24-63), for the built-in smooth / proximable terms, entirely inside the library.
//
// Why it exists: the iteration is a handful of kernel launches plus ~40 scalar flops; driven from an interpreted host the
// per-iteration host cost (tens of microseconds) is the bottleneck on small problems and at 8 GPUs.  The reference's
// driver is compiled Julia; this is its native equivalent.  It issues the SAME kernel sequence and the SAME scalar
// arithmetic in R = real(eltype(x0)) as the Python host (algorithms.py), so both produce identical iterates and
// iteration counts (tests/test_gpu_solvers.py::test_native_driver_equals_python_host).
//
// Host-side scalar code only; every vector operation is one of the library's kernels.
#include <math.h>
#include <string.h>

#include <limits>

#include "common.cuh"
#include "solve_scalar.h"

bool pb_bd_fista_eligible(const pb_ctx* ctx, int dtype, const pb_smooth* f, const pb_prox* g, const void* x, const void* z_prev, const void* grad,
                          const void* z, const void* x_next);   // lsq_fista.cu
int pb_bd_fista_step(pb_ctx* ctx, int dtype, const pb_smooth* f, const pb_prox* g, double gamma, double beta, const void* x, const void* z_prev,
                     void* grad, void* z, void* x_next);
bool pb_multi_eligible(const pb_ctx* ctx, int dtype, int64_t n, const pb_smooth* f, const pb_prox* g, const pb_solve_opts* o, const void* x_next);   // step_multi.cu
int pb_multi_run(pb_ctx* ctx, int dtype, int64_t n, const pb_smooth* f, const pb_prox* g, const pb_solve_opts* o, void* const X[3],
                 void* const Z[3], int64_t* k_out, double comb[4], float* kernel_ms);
bool pb_persist_eligible(const pb_ctx* ctx, int dtype, const pb_smooth* f, const pb_prox* g, const pb_solve_opts* o);   // persist.cu
int pb_persist_solve(pb_ctx* ctx, int dtype, int64_t n, const pb_smooth* f, const pb_prox* g, const pb_solve_opts* o, void* x,
                     void* grad, void* z, void* z_prev, pb_solve_result* out);

namespace {

struct Comb {          // rank-combined view of one exchange (mirror of host.py: Scalars)
  double gsum, res_sq, gdr, aux, res_inf;
  double local_aux;    // this rank's own AUX pair (dense least squares: r is replicated, not summed)
};

inline void two_sum(double a, double b, double& s, double& e) {
  s = a + b;
  const double bb = s - a;
  e = (a - (s - bb)) + (b - bb);
}
inline void dd_add(double& hi, double& lo, double bh, double bl) {
  double s, e;
  two_sum(hi, bh, s, e);
  e += lo + bl;
  const double h = s + e;
  lo = e - (h - s);
  hi = h;
}
inline double fold(const double* rows, int world, int slot) {
  double hi = 0.0, lo = 0.0;
  for (int p = 0; p < world; ++p) dd_add(hi, lo, rows[p * PB_NSCALARS + slot], rows[p * PB_NSCALARS + slot + 1]);
  return hi + lo;
}
inline double fold_max(const double* rows, int world, int slot) {
  double m = 0.0;
  for (int p = 0; p < world; ++p) {
    const double v = rows[p * PB_NSCALARS + slot];
    if (v != v) return v;
    if (v > m) m = v;
  }
  return m;
}

// The one host synchronisation of an iteration: device exchange when attached (any world size), else memcpy read-back.
int read_comb(pb_ctx* ctx, Comb* c) {
  double rows[PB_MAX_RANKS * PB_NSCALARS];
  int world = 1, rank = 0;
  if (ctx->xchg_world > 0 && ctx->xchg_connected) {
    world = ctx->xchg_world;
    rank = ctx->xchg_rank;
    int rc = pb_exchange_wait(ctx, rows, 30.0);
    if (rc != PB_OK) return rc;
  } else {
    int rc = pb_read_scalars(ctx, rows);
    if (rc != PB_OK) return rc;
  }
  c->gsum = fold(rows, world, PB_S_GSUM);
  c->res_sq = fold(rows, world, PB_S_RESSQ);
  c->gdr = fold(rows, world, PB_S_GDR);
  c->aux = fold(rows, world, PB_S_AUX);
  c->res_inf = fold_max(rows, world, PB_S_RESINF);
  c->local_aux = rows[rank * PB_NSCALARS + PB_S_AUX] + rows[rank * PB_NSCALARS + PB_S_AUX + 1];
  return PB_OK;
}

// Same, for one specific exchange of the pipelined loop (sequence number recorded right after the publishing launch).
int read_comb_seq(pb_ctx* ctx, unsigned int seq, Comb* c) {
  double rows[PB_MAX_RANKS * PB_NSCALARS];
  const int world = ctx->xchg_world, rank = ctx->xchg_rank;
  int rc = pb_xchg_wait_seq(ctx, seq, rows, 30.0);
  if (rc != PB_OK) return rc;
  c->gsum = fold(rows, world, PB_S_GSUM);
  c->res_sq = fold(rows, world, PB_S_RESSQ);
  c->gdr = fold(rows, world, PB_S_GDR);
  c->aux = fold(rows, world, PB_S_AUX);
  c->res_inf = fold_max(rows, world, PB_S_RESINF);
  c->local_aux = rows[rank * PB_NSCALARS + PB_S_AUX] + rows[rank * PB_NSCALARS + PB_S_AUX + 1];
  return PB_OK;
}

template <typename R>
struct Solver {
  pb_ctx* ctx;
  int dtype;
  int64_t n;
  const pb_smooth* f;
  const pb_prox* g;
  const pb_solve_opts* o;
  void *x, *grad, *z, *z_prev, *x_next, *grad_z, *scratch;
  // state scalars
  R gamma, f_x, g_z;
  Comb sc;
  int64_t backtracks;
  int warned;
  bool lazy_value;                       // LinearFunction with a fixed stepsize: f(x) is never consumed inside the loop
  // optional profiling (opts->profile): CUDA events around the whole loop and around every fused-step launch
  static const int kMaxStepEvents = 4096;
  cudaEvent_t ev_loop[2];
  cudaEvent_t* ev_step;
  int n_step_events, n_created;
  bool profile;

  static R sq_half(double sum_sq) { return pb_sq_half<R>(sum_sq); }   // norm(v)^2/2, sqrt-then-square (benchmark/benchmarks.jl:16)
  R f_value(const Comb& c) const {
    switch (f->kind) {
      case PB_F_LSQ_DENSE: return sq_half(c.local_aux);
      case PB_F_LINEAR: return (R)c.aux;
      default: return sq_half(c.aux);
    }
  }
  R g_value(const Comb& c) const {
    if (g->kind == PB_PROX_L1 || g->kind == PB_PROX_L21) return (R)g->p0 * (R)c.gsum;
    return R(0);
  }
  static R f_model(R fx, double gdr, double res_sq, R Lc) { return pb_f_model<R>(fx, gdr, res_sq, Lc); }   // fb_tools.jl:3-5

  int residual(const void* v) {     // f's residual pass at v; the value is Deferred in the AUX slot
    switch (f->kind) {
      case PB_F_LSQ_DENSE:
        // column shard (device exchange, world > 1): combine + NVLink all-gather + fold in one kernel; r and AUX replicated
        if (ctx->xchg_world > 1)
          return pb_lsq_dense_residual_sharded(ctx, dtype, f->m, f->n, f->A, f->lda, v, f->b, f->r, f->nb, f->nblk, 0);
        return pb_lsq_dense_residual(ctx, dtype, f->m, f->n, f->A, f->lda, v, f->b, f->r);
      case PB_F_LSQ_BLOCKDIAG: return pb_lsq_blockdiag_residual(ctx, dtype, f->nblk, f->mb, f->nb, f->A, v, f->b, f->r);
      default: return PB_EUNSUPPORTED;
    }
  }
  int eval_f(const void* v, void* grad_out) {     // value_and_gradient(f, v) -> grad_out, value Deferred
    int rc;
    switch (f->kind) {
      case PB_F_LSQ_DENSE:
        if ((rc = residual(v))) return rc;
        return pb_lsq_dense_gradient(ctx, dtype, f->m, f->n, f->A, f->lda, f->r, grad_out);
      case PB_F_LSQ_BLOCKDIAG:
        return pb_lsq_blockdiag_value_and_gradient(ctx, dtype, f->nblk, f->mb, f->nb, f->A, v, f->b, f->r, grad_out);
      case PB_F_SQDIST:
        return pb_sqdist(ctx, dtype, n, v, f->b, grad_out);
      case PB_F_LINEAR:
        if (grad_out != f->b && (rc = pb_copy(ctx, grad_out, f->b, (size_t)n * (dtype == PB_F32 ? 4 : 8)))) return rc;
        if (lazy_value) return PB_OK;      // value computed once for the final state (see run())
        return pb_dot(ctx, dtype, n, f->b, v);
      default:
        pb_set_error("pb_solve: unknown smooth term %d", f->kind);
        return PB_EINVAL;
    }
  }
  int eval_f_value(const void* v) {      // f(v) only (the FFB line search discards the gradient, fast_forward_backward.jl:112-128)
    if (f->kind == PB_F_LSQ_DENSE || f->kind == PB_F_LSQ_BLOCKDIAG) return residual(v);
    return eval_f(v, scratch);
  }
  int step(const void* xin, const void* gr, void* zout, bool extrap, R beta) { return step_to(xin, gr, z_prev, zout, x_next, extrap, beta); }
  int step_to(const void* xin, const void* gr, const void* zp, void* zout, void* xnext_out, bool extrap, R beta) {
    // events are created on first use (the adaptive paths launch more steps than maxit: one per backtrack)
    bool timed = profile && n_step_events < kMaxStepEvents;
    if (timed && n_step_events == n_created) {
      if (cudaEventCreate(&ev_step[2 * n_created]) == cudaSuccess && cudaEventCreate(&ev_step[2 * n_created + 1]) == cudaSuccess) {
        ++n_created;
      } else {
        cudaGetLastError();
        timed = profile = false;          // stop timing rather than record on an invalid handle
      }
    }
    if (timed) cudaEventRecord(ev_step[2 * n_step_events], ctx->stream);
    const int rc = extrap ? pb_ffb_step(ctx, dtype, n, xin, gr, zp, (double)gamma, (double)beta, g, nullptr, zout, nullptr, xnext_out)
                          : pb_fb_step(ctx, dtype, n, xin, gr, (double)gamma, g, nullptr, zout, nullptr);
    if (timed) {
      cudaEventRecord(ev_step[2 * n_step_events + 1], ctx->stream);
      ++n_step_events;
    }
    return rc;
  }
  bool stop() const {                    // norm(res, Inf)/gamma <= tol
    const R rn = (R)sc.res_inf;
    return (double)(rn / gamma) <= o->tol;
  }

  // fb_tools.jl:24-63 (A = nothing, Az aliased to z).  want_grad: FB keeps grad f(z) in grad_z.
  int backtrack(bool want_grad, R* f_z_out) {
    const R eps = std::numeric_limits<R>::epsilon();
    int rc;
    R f_upp = f_model(f_x, sc.gdr, sc.res_sq, R(1) / gamma);
    if ((rc = want_grad ? eval_f(z, grad_z) : eval_f_value(z))) return rc;
    Comb c;
    if ((rc = read_comb(ctx, &c))) return rc;
    R f_z = f_value(c);
    R tol = R(10) * eps * (R(1) + (R)fabs((double)f_z));
    while (f_z > f_upp + tol && gamma >= (R)o->minimum_gamma) {
      gamma = gamma * (R)o->reduce_gamma;
      if ((rc = step(x, grad, z, false, R(0)))) return rc;
      if ((rc = read_comb(ctx, &sc))) return rc;
      g_z = g_value(sc);
      f_upp = f_model(f_x, sc.gdr, sc.res_sq, R(1) / gamma);
      if ((rc = want_grad ? eval_f(z, grad_z) : eval_f_value(z))) return rc;
      if ((rc = read_comb(ctx, &c))) return rc;
      f_z = f_value(c);
      tol = R(10) * eps * (R(1) + (R)fabs((double)f_z));
      ++backtracks;
    }
    if (gamma < (R)o->minimum_gamma) warned = 1;     // fb_tools.jl:59-61 (the host prints the warning)
    *f_z_out = f_z;
    return PB_OK;
  }

  int run(pb_solve_result* out) {
    int rc;
    const bool fast = o->algorithm == PB_ALG_FFB;
    const bool adaptive = o->adaptive != 0;
    backtracks = 0;
    warned = 0;
    lazy_value = f->kind == PB_F_LINEAR && !adaptive && o->gamma > 0;
    profile = o->profile != 0;
    ev_step = nullptr;
    n_step_events = n_created = 0;
    const bool profile_req = profile;
    if (profile) {
      ev_step = new cudaEvent_t[2 * kMaxStepEvents];
      cudaEventCreate(&ev_loop[0]);
      cudaEventCreate(&ev_loop[1]);
      cudaEventRecord(ev_loop[0], ctx->stream);
    }
    rc = run_loop(out);
    if (profile_req) {
      if (rc == PB_OK && profile) {
        cudaEventSynchronize(ev_loop[1]);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev_loop[0], ev_loop[1]);
        out->loop_ms = ms;
        double tot = 0.0;
        for (int i = 0; i < n_step_events; ++i) {
          if (cudaEventElapsedTime(&ms, ev_step[2 * i], ev_step[2 * i + 1]) == cudaSuccess) tot += ms;
        }
        out->step_kernel_ms = tot;
        out->step_kernel_launches = n_step_events;
      }
      for (int i = 0; i < 2 * n_created; ++i) cudaEventDestroy(ev_step[i]);
      cudaEventDestroy(ev_loop[0]);
      cudaEventDestroy(ev_loop[1]);
      delete[] ev_step;
    }
    return rc;
  }

  int run_loop(pb_solve_result* out) {
    int rc;
    const bool fast = o->algorithm == PB_ALG_FFB;
    const bool adaptive = o->adaptive != 0;
    // ---- init: forward_backward.jl:65-84 / fast_forward_backward.jl:73-97 ----
    if ((rc = eval_f(x, grad))) return rc;
    bool fx_pending = true;
    if (o->gamma <= 0) {                 // fb_tools.jl:7-12 with A = I
      if ((rc = read_comb(ctx, &sc))) return rc;
      f_x = f_value(sc);
      fx_pending = false;
      if ((rc = pb_add_scalar(ctx, dtype, n, x, 1.0, scratch))) return rc;
      if ((rc = eval_f(scratch, z))) return rc;                      // z is free at this point: holds grad f(x + 1)
      if ((rc = pb_sub(ctx, dtype, n, z, grad, z))) return rc;
      Comb c2;
      if ((rc = read_comb(ctx, &c2))) return rc;
      const int64_t n_glob = o->n_global > 0 ? o->n_global : n;
      const R lower = (R)sqrt(c2.aux) / (R)sqrt((double)n_glob);
      gamma = R(1) / lower;
    } else {
      gamma = (R)o->gamma;
    }
    Nesterov<R> seq;
    seq.init(o->sequence, (R)o->mf, (R)o->constant_beta);
    R beta_next = R(0);
    // element-wise gradient source + fixed stepsize: one persistent kernel loops over the iterations (step_multi.cu), same results
    if (fast && !adaptive && pb_multi_eligible(ctx, dtype, n, f, g, o, x_next)) {
      rc = run_ffb_multi(out);
      if (rc != PB_EUNSUPPORTED) return rc;
    }
    // block-diagonal least squares + fixed stepsize: ONE sweep of A per iteration (lsq_fista.cu), same results
    if (fast && !adaptive && x_next && n > 0 && pb_bd_fista_eligible(ctx, dtype, f, g, x, z_prev, grad, z, x_next)) return run_ffb_bd_fista(out, seq);
    if (fast) {
      if ((rc = pb_copy(ctx, z_prev, x, (size_t)n * (dtype == PB_F32 ? 4 : 8)))) return rc;     // z_prev = copy(x)
      if (!adaptive) {
        beta_next = seq.next(gamma);
        rc = step(x, grad, z, true, beta_next);
      } else {
        rc = step(x, grad, z, false, R(0));
      }
    } else {
      rc = step(x, grad, z, false, R(0));
    }
    if (rc) return rc;
    const bool pipelined = fast && !adaptive && o->spare_x && o->spare_z && o->spare_grad && ctx->xchg_world > 0 &&
                           ctx->xchg_connected && ctx->xchg_fused && n > 0;
    if (pipelined) return run_ffb_pipelined(out, seq, beta_next);
    if ((rc = read_comb(ctx, &sc))) return rc;
    if (fx_pending) f_x = f_value(sc);
    g_z = g_value(sc);
    // ---- driver loop: src/ProximalAlgorithms.jl:114-123 ----
    int64_t k = 1;
    for (;; ++k) {
      if (k >= o->maxit || stop()) break;
      if (!fast) {                       // forward_backward.jl:86-123
        if (adaptive) {
          gamma = gamma * (R)o->increase_gamma;
          R f_z;
          if ((rc = backtrack(true, &f_z))) return rc;
          f_x = f_z;
          void* t = x; x = z; z = t;
          t = grad; grad = grad_z; grad_z = t;
          if ((rc = step(x, grad, z, false, R(0)))) return rc;
          if ((rc = read_comb(ctx, &sc))) return rc;
        } else {
          void* t = x; x = z; z = t;
          if ((rc = eval_f(x, grad))) return rc;
          if ((rc = step(x, grad, z, false, R(0)))) return rc;
          if ((rc = read_comb(ctx, &sc))) return rc;
          f_x = f_value(sc);
        }
        g_z = g_value(sc);
      } else {                           // fast_forward_backward.jl:106-145
        if (adaptive) {
          gamma = gamma * (R)o->increase_gamma;
          R f_z;
          if ((rc = backtrack(false, &f_z))) return rc;
          const R beta = seq.next(gamma);
          if ((rc = pb_extrapolate(ctx, dtype, n, z, z_prev, (double)beta, x))) return rc;
          void* t = z_prev; z_prev = z; z = t;
          if ((rc = eval_f(x, grad))) return rc;
          if ((rc = step(x, grad, z, false, R(0)))) return rc;
        } else {
          gamma = (R)o->gamma > 0 ? (R)o->gamma : gamma;
          void* t = x; x = x_next; x_next = t;        // :135, computed by the previous fused pass
          t = z_prev; z_prev = z; z = t;              // :136
          if ((rc = eval_f(x, grad))) return rc;
          beta_next = seq.next(gamma);
          if ((rc = step(x, grad, z, true, beta_next))) return rc;
        }
        if ((rc = read_comb(ctx, &sc))) return rc;
        f_x = f_value(sc);
        g_z = g_value(sc);
      }
    }
    return finish(out, k);
  }

  // Fixed-stepsize FFB with one iteration of look-ahead (pb_solve_opts: spare_x / spare_z / spare_grad).  Invariant at the top of
  // the loop: the kernel of iteration k has been launched and will publish exchange `seq_k`; X[ix] = x_k, X[ixn] = x_{k+1}
  // (written by that kernel), Z[izp] = z_{k-1}, Z[iz] = z_k (ditto), G[ig] = grad f(x_k).  The body of iteration k+1 reads
  // x_{k+1}, z_k and writes the third x / z buffer and the other gradient buffer, so nothing of iteration k is overwritten
  // before its scalars have been examined.
  int run_ffb_pipelined(pb_solve_result* out, Nesterov<R>& seq, R beta_next) {
    int rc;
    void* X[3] = {x, x_next, o->spare_x};
    void* Z[3] = {z_prev, z, o->spare_z};
    void* G[2] = {grad, (f->kind == PB_F_LINEAR && grad == f->b) ? grad : o->spare_grad};
    int ix = 0, ixn = 1, izp = 0, iz = 1, ig = 0;
    unsigned int seq_k = ctx->xchg_seq;               // published by the init step launched by the caller
    (void)beta_next;
    // from here on the step kernels leave the fold + exchange to a 1-CTA kernel on the side stream (step_kernels.cu:
    // launch_step_deferred), so the main stream goes straight from one step kernel to the next
    if ((rc = pb_step_defer(ctx, 1))) return rc;
    int64_t k = 1;
    for (;;) {
      const bool spec = k < o->maxit;
      Nesterov<R> seq_saved = seq;
      unsigned int seq_next = 0;
      const int ixf = 3 - ix - ixn, izf = 3 - izp - iz, igo = 1 - ig;
      if (spec) {                                      // fast_forward_backward.jl:130-142 for iteration k+1, launched ahead
        if ((rc = eval_f_to(X[ixn], G[igo]))) {
          pb_step_defer(ctx, 0);
          return rc;
        }
        const R beta = seq.next(gamma);
        if ((rc = step_to(X[ixn], G[igo], Z[iz], Z[izf], X[ixf], true, beta))) {
          pb_step_defer(ctx, 0);
          return rc;
        }
        seq_next = ctx->xchg_seq;
      }
      if ((rc = read_comb_seq(ctx, seq_k, &sc))) {
        pb_step_defer(ctx, 0);
        return rc;
      }
      f_x = f_value(sc);
      g_z = g_value(sc);
      if (k >= o->maxit || stop()) {                   // src/ProximalAlgorithms.jl:117
        seq = seq_saved;
        break;
      }
      // commit the speculation: iteration k+1 becomes the current one
      ix = ixn;
      ixn = ixf;
      izp = iz;
      iz = izf;
      ig = igo;
      seq_k = seq_next;
      ++k;
    }
    ctx->xchg_pending = 0;                             // whatever was published last has been consumed or is discarded
    if ((rc = pb_step_defer(ctx, 0))) return rc;       // main stream waits for the outstanding folds
    x = X[ix];
    x_next = X[ixn];
    z = Z[iz];
    z_prev = Z[izp];
    grad = G[ig];
    return finish(out, k);
  }

  int eval_f_to(const void* v, void* grad_out) { return eval_f(v, grad_out); }

  // Fixed-stepsize FFB on a block-diagonal least-squares term with one sweep of A per iteration (csrc/lsq_fista.cu): the kernel of iteration
  // k computes grad f(x_k) from r_k = A x_k - b, the fused step, and the partial products of A x_{k+1}; its combine leaves r_{k+1} in f->r and
  // ||r_{k+1}||^2 in AUX, i.e. the value of the NEXT point, which is shifted here.  On entry eval_f(x, grad) of the init has left r_1 and AUX.
  int run_ffb_bd_fista(pb_solve_result* out, Nesterov<R>& seq) {
    int rc;
    if ((rc = pb_copy(ctx, z_prev, x, (size_t)n * (dtype == PB_F32 ? 4 : 8)))) return rc;     // z_prev = copy(x)
    Comb c0;
    if ((rc = read_comb(ctx, &c0))) return rc;
    R f_cur = f_value(c0);
    // one iteration of look-ahead when the caller gave spare vectors: without it the GPU idles for a host round trip after every sweep
    const bool via_xchg = ctx->xchg_world > 0 && ctx->xchg_connected;
    if (o->spare_x && o->spare_z && o->spare_grad && pb_aligned16(o->spare_x) && pb_aligned16(o->spare_z) && pb_aligned16(o->spare_grad) &&
        (!via_xchg || ctx->xchg_fused))
      return run_ffb_bd_fista_ahead(out, seq, f_cur, via_xchg);
    int64_t k = 1;
    for (;;) {
      const R beta = seq.next(gamma);
      if ((rc = pb_bd_fista_step(ctx, dtype, f, g, (double)gamma, (double)beta, x, z_prev, grad, z, x_next))) return rc;
      if ((rc = read_comb(ctx, &sc))) return rc;
      f_x = f_cur;
      g_z = g_value(sc);
      if (k >= o->maxit || stop()) break;                       // src/ProximalAlgorithms.jl:117
      f_cur = f_value(sc);                                      // AUX = ||A x_{k+1} - b||^2
      void* t = x; x = x_next; x_next = t;                      // fast_forward_backward.jl:135, computed by the fused pass
      t = z_prev; z_prev = z; z = t;                            // :136
      ++k;
    }
    return finish(out, k);
  }

  // The same loop with the sweep of iteration k+1 launched before the scalars of iteration k are examined (the discard rule of
  // run_ffb_pipelined: iteration k+1 writes the third x / z buffer and the other gradient buffer, so the state of a stopping iteration k
  // is intact; f->r is scratch).  The scalars of an iteration reach the host either through the exchange (sequence numbers, parity
  // double-buffered landing zone) or, without one, through a 128-byte copy into one of two pinned slots followed by an event.
  struct AheadMark {
    unsigned int seq;
    int slot;
  };
  int ahead_mark(bool via_xchg, int slot, AheadMark* m) {
    m->slot = slot;
    m->seq = ctx->xchg_seq;
    if (via_xchg) return PB_OK;
    PB_CHECK_CUDA(cudaMemcpyAsync(ctx->ahead_host + (size_t)slot * PB_NSCALARS, ctx->scalars_dev, PB_NSCALARS * sizeof(double), cudaMemcpyDeviceToHost,
                                  ctx->stream));
    PB_CHECK_CUDA(cudaEventRecord(ctx->ahead_ev[slot], ctx->stream));
    return PB_OK;
  }
  int ahead_read(bool via_xchg, const AheadMark& m, Comb* c) {
    if (via_xchg) return read_comb_seq(ctx, m.seq, c);
    PB_CHECK_CUDA(cudaEventSynchronize(ctx->ahead_ev[m.slot]));
    const double* rows = ctx->ahead_host + (size_t)m.slot * PB_NSCALARS;
    c->gsum = fold(rows, 1, PB_S_GSUM);
    c->res_sq = fold(rows, 1, PB_S_RESSQ);
    c->gdr = fold(rows, 1, PB_S_GDR);
    c->aux = fold(rows, 1, PB_S_AUX);
    c->res_inf = fold_max(rows, 1, PB_S_RESINF);
    c->local_aux = rows[PB_S_AUX] + rows[PB_S_AUX + 1];
    return PB_OK;
  }
  int run_ffb_bd_fista_ahead(pb_solve_result* out, Nesterov<R>& seq, R f_cur, bool via_xchg) {
    int rc;
    if (!via_xchg && !ctx->ahead_host) {
      void* h = nullptr;
      PB_CHECK_CUDA(cudaMallocHost(&h, 2 * PB_NSCALARS * sizeof(double)));
      PB_CHECK_CUDA(cudaEventCreateWithFlags(&ctx->ahead_ev[0], cudaEventDisableTiming));
      PB_CHECK_CUDA(cudaEventCreateWithFlags(&ctx->ahead_ev[1], cudaEventDisableTiming));
      ctx->ahead_host = static_cast<double*>(h);
    }
    void* X[3] = {x, x_next, o->spare_x};
    void* Z[3] = {z_prev, z, o->spare_z};
    void* G[2] = {grad, o->spare_grad};
    int ix = 0, ixn = 1, izp = 0, iz = 1, ig = 0;
    AheadMark mk, mk_next;
    mk_next.seq = 0;
    mk_next.slot = 0;
    {
      const R beta = seq.next(gamma);
      if ((rc = pb_bd_fista_step(ctx, dtype, f, g, (double)gamma, (double)beta, X[ix], Z[izp], G[ig], Z[iz], X[ixn]))) return rc;
      if ((rc = ahead_mark(via_xchg, 1, &mk))) return rc;
    }
    int64_t k = 1;
    for (;;) {
      const bool spec = k < o->maxit;
      const Nesterov<R> seq_saved = seq;
      const int ixf = 3 - ix - ixn, izf = 3 - izp - iz, igo = 1 - ig;
      if (spec) {                                       // fast_forward_backward.jl:130-142 for iteration k+1, launched ahead
        const R beta = seq.next(gamma);
        if ((rc = pb_bd_fista_step(ctx, dtype, f, g, (double)gamma, (double)beta, X[ixn], Z[iz], G[igo], Z[izf], X[ixf]))) return rc;
        if ((rc = ahead_mark(via_xchg, (int)((k + 1) & 1), &mk_next))) return rc;
      }
      if ((rc = ahead_read(via_xchg, mk, &sc))) return rc;
      f_x = f_cur;
      g_z = g_value(sc);
      if (k >= o->maxit || stop()) {                    // src/ProximalAlgorithms.jl:117
        seq = seq_saved;
        break;
      }
      f_cur = f_value(sc);                              // AUX = ||A x_{k+1} - b||^2
      ix = ixn;
      ixn = ixf;
      izp = iz;
      iz = izf;
      ig = igo;
      mk = mk_next;
      ++k;
    }
    if (via_xchg) ctx->xchg_pending = 0;                // whatever was published last has been consumed or is discarded
    x = X[ix];
    x_next = X[ixn];
    z = Z[iz];
    z_prev = Z[izp];
    grad = G[ig];
    return finish(out, k);
  }

  // Fixed-stepsize FFB, f = <c, .> or SquaredDistance: iterations 1 .. k inside ONE kernel (csrc/step_multi.cu).  The rings are laid
  // out so that iteration 1 reads the caller's x (= copy(x0)) and z_prev (= copy(x)): X[1] = x, Z[0] = z_prev.
  int run_ffb_multi(pb_solve_result* out) {
    int rc;
    if ((rc = pb_copy(ctx, z_prev, x, (size_t)n * (dtype == PB_F32 ? 4 : 8)))) return rc;     // z_prev = copy(x)
    void* X[3] = {o->spare_x, x, x_next};
    void* Z[3] = {z_prev, z, o->spare_z};
    int64_t k = 0;
    double comb[4];
    float ms = 0.f;
    rc = pb_multi_run(ctx, dtype, n, f, g, o, X, Z, &k, comb, profile ? &ms : nullptr);
    if (rc != PB_OK) return rc;
    x = X[k % 3];
    x_next = X[(k + 1) % 3];
    z = Z[k % 3];
    z_prev = Z[(k + 2) % 3];
    sc.gsum = comb[0];
    sc.res_sq = comb[1];
    sc.gdr = comb[2];
    sc.res_inf = comb[3];
    sc.aux = sc.local_aux = 0.0;
    g_z = g_value(sc);
    if (f->kind == PB_F_SQDIST) {          // grad f and f of the final state (the loop itself never consumes f with a fixed stepsize)
      if ((rc = eval_f(x, grad))) return rc;
      Comb c;
      if ((rc = read_comb(ctx, &c))) return rc;
      f_x = f_value(c);
    }
    rc = finish(out, k);                   // LinearFunction: f(x) = <c, x> of the final state (lazy_value)
    out->multi_iter_kernel = 1;
    if (profile) {
      out->loop_ms = ms;
      out->step_kernel_ms = ms;
      out->step_kernel_launches = k;       // the one launch covers k iterations: step_kernel_ms / launches = time per iteration
      profile = false;                     // run() must not overwrite these with its own events
    }
    return rc;
  }

  int finish(pb_solve_result* out, int64_t k) {
    int rc;
    if (ev_step) cudaEventRecord(ev_loop[1], ctx->stream);     // the K iterations end here
    if (lazy_value) {                    // f(x) of the final state (LinearFunction: <c, x>)
      if ((rc = pb_dot(ctx, dtype, n, f->b, x))) return rc;
      Comb c;
      if ((rc = read_comb(ctx, &c))) return rc;
      f_x = f_value(c);
    }
    out->iterations = k;
    out->gamma = (double)gamma;
    out->f_x = (double)f_x;
    out->g_z = (double)g_z;
    out->res_inf = sc.res_inf;
    out->res_sq = sc.res_sq;
    out->gdr = sc.gdr;
    out->gsum = sc.gsum;
    out->backtracks = backtracks;
    out->warned_small_gamma = warned;
    out->x = x;
    out->grad = grad;
    out->z = z;
    out->z_prev = z_prev;
    return PB_OK;
  }
};

}  // namespace

extern "C" int pb_solve(pb_ctx* ctx, int dtype, int64_t n, const pb_smooth* f, const pb_prox* g, const pb_solve_opts* o,
                        void* x, void* grad, void* z, void* z_prev, void* x_next, void* grad_z, void* scratch,
                        pb_solve_result* out) {
  PB_REQUIRE(ctx != nullptr && f != nullptr && g != nullptr && o != nullptr && out != nullptr, "null argument");
  PB_REQUIRE(dtype == PB_F32 || dtype == PB_F64, "dtype must be PB_F32 or PB_F64");
  PB_REQUIRE(n >= 0 && o->maxit >= 1, "need n >= 0 and maxit >= 1");
  PB_REQUIRE(x && grad && z && scratch, "null state vector");
  PB_REQUIRE(o->algorithm == PB_ALG_FB || o->algorithm == PB_ALG_FFB, "unknown algorithm");
  PB_REQUIRE(o->algorithm != PB_ALG_FFB || (z_prev && (o->adaptive || x_next)), "FFB needs z_prev (and x_next when the stepsize is fixed)");
  PB_REQUIRE(o->algorithm != PB_ALG_FB || !o->adaptive || grad_z, "adaptive FB needs grad_z");
  PB_REQUIRE(o->gamma > 0 || o->adaptive, "a fixed stepsize needs gamma > 0");
  PB_REQUIRE(g->kind == PB_PROX_ZERO || g->kind == PB_PROX_L1 || g->kind == PB_PROX_BOX || g->kind == PB_PROX_L21 ||
                 (g->kind == PB_PROX_BALL && ctx->xchg_world <= 1),
             "pb_solve supports Zero, NormL1, IndBox, NormL21 and (on one GPU) IndBallL2");
  PB_REQUIRE(f->kind != PB_F_LSQ_DENSE || ctx->xchg_world <= 1 || (f->nb >= f->n && f->nblk >= 0 && f->nblk + f->n <= f->nb),
             "a column-sharded dense least-squares term needs nb = n_global and nblk = col_offset (proxb200.h, pb_smooth)");
  memset(out, 0, sizeof(*out));
  PbDeviceGuard dev_guard(ctx);
  // cache-resident dense least squares: the whole loop runs on the device in one persistent kernel (persist.cu), same results
  if (pb_persist_eligible(ctx, dtype, f, g, o) && n == f->n && n > 0) return pb_persist_solve(ctx, dtype, n, f, g, o, x, grad, z, z_prev, out);
  if (dtype == PB_F32) {
    Solver<float> s{ctx, dtype, n, f, g, o, x, grad, z, z_prev, x_next, grad_z, scratch};
    return s.run(out);
  }
  Solver<double> s{ctx, dtype, n, f, g, o, x, grad, z, z_prev, x_next, grad_z, scratch};
  return s.run(out);
}
