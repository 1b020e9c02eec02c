// panoc_solve.cu -- native driver loop for PANOC (src/algorithms/panoc.jl:88-112 init, :138-255 step, :257-259 stop / solution)
// with L-BFGS or no acceleration (src/accel/lbfgs.jl, noaccel.jl) and the stepsize backtracking of src/utilities/fb_tools.jl:24-63,
// for the built-in smooth terms (pb_smooth) composed with A = I or a dense matrix, and the single-pass proximable terms.
//
// It is a line-by-line native twin of proximalalgorithms.jl_b200/panoc.py: the same kernel sequence, the same buffer renames instead
// of copies, the same scalar arithmetic in R = real(eltype(x0)) (including Julia's Float64 promotion of `0.5 / gamma`,
// panoc.jl:196-198), one host synchronisation per accepted iteration.  What it removes is the interpreter cost per launch
// (10-15 us x ~21 launches per iteration), which dominates on small problems.
// STATUS: staged for round 2 -- compiles, mirrors the hardware-verified Python host, not yet run on a GPU.
#include <math.h>
#include <string.h>

#include <limits>

#include "common.cuh"

namespace {

struct Rows {
  double gsum, res_sq, gdr, aux, aux2, aux3, res_inf, local_aux;
};

inline void two_sum(double a, double b, double& s, double& e) {
  s = a + b;
  const double bb = s - a;
  e = (a - (s - bb)) + (b - bb);
}
inline double fold(const double* rows, int world, int slot) {
  double hi = 0.0, lo = 0.0;
  for (int p = 0; p < world; ++p) {
    double s, e;
    two_sum(hi, rows[p * PB_NSCALARS + slot], s, e);
    e += lo + rows[p * PB_NSCALARS + slot + 1];
    const double h = s + e;
    lo = e - (h - s);
    hi = h;
  }
  return hi + lo;
}

int read_rows(pb_ctx* ctx, Rows* c) {
  double rows[PB_MAX_RANKS * PB_NSCALARS];
  int world = 1, rank = 0, rc;
  if (ctx->xchg_world > 0 && ctx->xchg_connected) {
    world = ctx->xchg_world;
    rank = ctx->xchg_rank;
    rc = pb_exchange_wait(ctx, rows, 30.0);
  } else {
    rc = pb_read_scalars(ctx, rows);
  }
  if (rc != PB_OK) return rc;
  c->gsum = fold(rows, world, PB_S_GSUM);
  c->res_sq = fold(rows, world, PB_S_RESSQ);
  c->gdr = fold(rows, world, PB_S_GDR);
  c->aux = fold(rows, world, PB_S_AUX);
  c->aux2 = fold(rows, world, PB_S_AUX2);
  c->aux3 = fold(rows, world, PB_S_AUX3);
  double m = 0.0;
  for (int p = 0; p < world; ++p) {
    const double v = rows[p * PB_NSCALARS + PB_S_RESINF];
    if (v != v) {
      m = v;
      break;
    }
    if (v > m) m = v;
  }
  c->res_inf = m;
  c->local_aux = rows[rank * PB_NSCALARS + PB_S_AUX] + rows[rank * PB_NSCALARS + PB_S_AUX + 1];
  return PB_OK;
}

template <typename R>
struct Panoc {
  pb_ctx* ctx;
  int dtype;
  int64_t n, m;            // primal dimension, dimension of A x (m == n and the twins alias when A = I)
  const pb_smooth* f;
  const pb_prox* g;
  const pb_panoc_opts* o;
  bool ident;
  size_t es;
  // n-space vectors
  void *X[3], *RES[2], *Z[2], *d, *G[2], *At_grad_f_Az, *scratch_n;
  // m-space vectors (aliases of the n-space ones when ident)
  void *AX[2], *GM[2], *Ad, *Az_buf, *grad_f_Az;
  // roles (pointers into the pools)
  void *x, *x_prev, *x_d, *x_spare, *res, *res_prev, *z, *z_curr;
  void *Ax, *Ax_d, *grad_f_Ax, *grad_f_Ax_d, *At_grad_f_Ax, *At_grad_f_Ax_d, *Az;
  pb_lbfgs* H;
  // scalars
  R gamma, f_Ax, f_Ax_d, g_z, tau, alpha, beta;
  Rows sc;
  int64_t gamma_backtracks, tau_backtracks;
  int warned;
  void* owned[24];
  int nowned;

  static R sq_half(double sum_sq) {
    const R nr = (R)sqrt(sum_sq);
    return (nr * nr) / R(2);
  }
  R f_value(const Rows& c) const {
    switch (f->kind) {
      case PB_F_LSQ_DENSE: return sq_half(c.local_aux);
      case PB_F_LINEAR: return (R)c.aux;
      default: return sq_half(c.aux);
    }
  }
  R g_value(const Rows& c) const {
    if (g->kind == PB_PROX_L1 || g->kind == PB_PROX_L21) return (R)g->p0 * (R)c.gsum;
    return R(0);
  }
  static R f_model(R fx, double gdr, double res_sq, R Lc) {     // fb_tools.jl:3-5
    const R nr = (R)sqrt(res_sq);
    return (fx - (R)gdr) + (Lc / R(2)) * (nr * nr);
  }
  R Lc() const { return alpha / gamma; }
  R fbe() const { return f_model(f_Ax, sc.gdr, sc.res_sq, Lc()) + g_z; }     // panoc.jl:85-86, :202

  int alloc(void** p, int64_t len) {
    const int rc = pb_malloc(ctx, (size_t)(len > 0 ? len : 1) * es, p);
    if (rc == PB_OK) owned[nowned++] = *p;
    return rc;
  }
  void release() {
    cudaStreamSynchronize(ctx->stream);
    for (int k = 0; k < nowned; ++k) pb_free(ctx, owned[k]);
    nowned = 0;
    if (H) pb_lbfgs_destroy(H);
    H = nullptr;
  }

  // value_and_gradient(f, v) for v in m-space: gradient into grad_out, value Deferred in the AUX slot
  int eval_f(const void* v, void* grad_out) {
    int rc;
    switch (f->kind) {
      case PB_F_LSQ_DENSE:
        if ((rc = pb_lsq_dense_residual(ctx, dtype, f->m, f->n, f->A, f->lda, v, f->b, f->r))) return rc;
        return pb_lsq_dense_gradient(ctx, dtype, f->m, f->n, f->A, f->lda, f->r, grad_out);
      case PB_F_LSQ_BLOCKDIAG:
        if ((rc = pb_lsq_blockdiag_residual(ctx, dtype, f->nblk, f->mb, f->nb, f->A, v, f->b, f->r))) return rc;
        return pb_lsq_blockdiag_gradient(ctx, dtype, f->nblk, f->mb, f->nb, f->A, f->r, grad_out);
      case PB_F_SQDIST:
        return pb_sqdist(ctx, dtype, m, v, f->b, grad_out);
      case PB_F_LINEAR:
        if ((rc = pb_copy(ctx, grad_out, f->b, (size_t)m * es))) return rc;
        return pb_dot(ctx, dtype, m, f->b, v);
      default:
        pb_set_error("pb_panoc_solve: unknown smooth term %d", f->kind);
        return PB_EINVAL;
    }
  }
  int mulA(void* out, const void* v) {      // out = A v
    return pb_lsq_dense_residual(ctx, dtype, o->Am, o->An, o->A, o->Am, v, nullptr, out);
  }
  int mulAt(void* out, const void* v) {     // out = A' v
    return pb_lsq_dense_gradient(ctx, dtype, o->Am, o->An, o->A, o->Am, v, out);
  }
  int lincomb(double a, const void* u, double b, const void* v, void* out, int64_t len) {
    return pb_lincomb2(ctx, dtype, len, a, u, b, v, out);
  }
  int step_kernel() {                        // y, z, res from x and At_grad_f_Ax (panoc.jl:199-201, :246-248)
    return pb_fb_step(ctx, dtype, n, x, At_grad_f_Ax, (double)gamma, g, nullptr, z, res);
  }
  int read_step(bool f_pending) {
    int rc;
    if ((rc = read_rows(ctx, &sc))) return rc;
    if (f_pending) f_Ax = f_value(sc);
    g_z = g_value(sc);
    return PB_OK;
  }
  bool stop() const { return (double)((R)sc.res_inf / gamma) <= o->tol; }     // panoc.jl:257-258

  // ---- init, panoc.jl:88-112 ----
  int init(const void* x0) {
    int rc;
    nowned = 0;
    H = nullptr;
    gamma_backtracks = tau_backtracks = 0;
    warned = 0;
    alpha = (R)o->alpha;
    beta = (R)o->beta;
    for (int k = 0; k < 3; ++k)
      if ((rc = alloc(&X[k], n))) return rc;
    for (int k = 0; k < 2; ++k) {
      if ((rc = alloc(&RES[k], n))) return rc;
      if ((rc = alloc(&Z[k], n))) return rc;
      if ((rc = alloc(&G[k], n))) return rc;
    }
    if ((rc = alloc(&d, n))) return rc;
    if ((rc = alloc(&At_grad_f_Az, n))) return rc;
    if ((rc = alloc(&scratch_n, n))) return rc;
    if (ident) {
      AX[0] = AX[1] = nullptr;
      GM[0] = G[0];
      GM[1] = G[1];
      Ad = d;
      Az_buf = nullptr;
      grad_f_Az = At_grad_f_Az;
    } else {
      for (int k = 0; k < 2; ++k) {
        if ((rc = alloc(&AX[k], m))) return rc;
        if ((rc = alloc(&GM[k], m))) return rc;
      }
      if ((rc = alloc(&Ad, m))) return rc;
      if ((rc = alloc(&Az_buf, m))) return rc;
      if ((rc = alloc(&grad_f_Az, m))) return rc;
    }
    x = X[0];
    x_prev = X[1];
    x_d = X[2];
    x_spare = nullptr;
    if ((rc = pb_copy(ctx, x, x0, (size_t)n * es))) return rc;                      // :89
    if (ident) {
      Ax = x;
    } else {
      Ax = AX[0];
      if ((rc = mulA(Ax, x))) return rc;                                             // :90
    }
    grad_f_Ax = GM[0];
    if ((rc = eval_f(Ax, grad_f_Ax))) return rc;                                     // :91
    bool f_pending = true;
    if (o->gamma <= 0) {                                                             // :92-95, fb_tools.jl:7-12
      if ((rc = read_rows(ctx, &sc))) return rc;
      f_Ax = f_value(sc);
      f_pending = false;
      void* xeps = scratch_n;
      if ((rc = pb_add_scalar(ctx, dtype, n, x, 1.0, xeps))) return rc;
      void* Axeps = xeps;
      void* geps = GM[1];
      if (!ident) {
        Axeps = Ad;                                                                  // free at this point
        if ((rc = mulA(Axeps, xeps))) return rc;
      }
      if ((rc = eval_f(Axeps, geps))) return rc;
      if ((rc = pb_sub(ctx, dtype, m, geps, grad_f_Ax, geps))) return rc;
      if (!ident) {
        if ((rc = mulAt(xeps, geps))) return rc;
        if ((rc = pb_nrm2sq(ctx, dtype, n, xeps))) return rc;
      }
      Rows c2;
      if ((rc = read_rows(ctx, &c2))) return rc;
      const R lower = (R)sqrt(c2.aux) / (R)sqrt((double)n);
      gamma = alpha / lower;
    } else {
      gamma = (R)o->gamma;
    }
    if (ident) {
      At_grad_f_Ax = grad_f_Ax;
    } else {
      At_grad_f_Ax = G[0];
      if ((rc = mulAt(At_grad_f_Ax, grad_f_Ax))) return rc;                          // :96
    }
    z = Z[0];
    z_curr = Z[1];
    res = RES[0];
    res_prev = RES[1];
    if ((rc = step_kernel())) return rc;                                             // :97-98, :109
    if ((rc = read_step(f_pending))) return rc;
    if (o->lbfgs_mem > 0 && (rc = pb_lbfgs_create(ctx, dtype, n, o->lbfgs_mem, &H))) return rc;   // :110
    tau = R(0);
    At_grad_f_Ax_d = G[1];
    grad_f_Ax_d = GM[1];
    Ax_d = ident ? x_d : AX[1];
    Az = ident ? z_curr : Az_buf;
    f_Ax_d = R(0);
    return PB_OK;
  }

  // fb_tools.jl:24-63 as called at panoc.jl:143-159
  int backtrack_gamma(R* f_Az_out, R* f_upp_out) {
    const R eps = std::numeric_limits<R>::epsilon();
    int rc;
    R f_upp = f_model(f_Ax, sc.gdr, sc.res_sq, Lc());
    auto f_at_z = [&](R* out) -> int {
      int r2;
      if (ident) {
        Az = z;
      } else if ((r2 = mulA(Az, z))) {
        return r2;
      }
      if ((r2 = eval_f(Az, grad_f_Az))) return r2;
      Rows c;
      if ((r2 = read_rows(ctx, &c))) return r2;
      *out = f_value(c);
      return PB_OK;
    };
    R f_Az;
    if ((rc = f_at_z(&f_Az))) return rc;
    R tol = R(10) * eps * (R(1) + (R)fabs((double)f_Az));
    while (f_Az > f_upp + tol && gamma >= (R)o->minimum_gamma) {
      gamma = gamma * R(0.5);
      if ((rc = step_kernel())) return rc;
      if ((rc = read_rows(ctx, &sc))) return rc;
      g_z = g_value(sc);
      f_upp = f_model(f_Ax, sc.gdr, sc.res_sq, Lc());
      if ((rc = f_at_z(&f_Az))) return rc;
      tol = R(10) * eps * (R(1) + (R)fabs((double)f_Az));
      ++gamma_backtracks;
    }
    if (gamma < (R)o->minimum_gamma) warned = 1;
    *f_Az_out = f_Az;
    *f_upp_out = f_upp;
    return PB_OK;
  }

  static void* other(void* const pool[2], const void* p) { return pool[0] == p ? pool[1] : pool[0]; }

  // ---- step, panoc.jl:138-255 (statement order of panoc.py: step) ----
  int step() {
    int rc;
    const R inf = std::numeric_limits<R>::infinity();
    R f_Az = inf, a = inf, b = inf, c = inf, f_upp;
    if (o->adaptive) {                                                               // :141-164
      const R gamma_prev = gamma;
      if ((rc = backtrack_gamma(&f_Az, &f_upp))) return rc;
      if (gamma != gamma_prev && H) pb_lbfgs_reset(H);
    } else {
      f_upp = f_model(f_Ax, sc.gdr, sc.res_sq, Lc());                                // :166
    }
    const R FBE_x = f_upp + g_z;                                                     // :170
    const double res_sq_x = sc.res_sq;
    // direction and x_d = x + d (:173, :183); x_prev <- x (:176) is a rename
    void* xd_buf = nullptr;
    void* spare = nullptr;
    for (int k = 0; k < 3; ++k)
      if (X[k] != x && !xd_buf) xd_buf = X[k];
    for (int k = 0; k < 3; ++k)
      if (X[k] != x && X[k] != xd_buf) spare = X[k];
    if (H) {
      if ((rc = pb_lbfgs_apply(ctx, H, res, -1.0, d, x, xd_buf))) return rc;         // :114-117 + :183
    } else {
      if ((rc = pb_scale(ctx, dtype, n, -1.0, res, d))) return rc;                   // :119-120
      if ((rc = lincomb(1.0, x, 1.0, d, xd_buf, n))) return rc;
    }
    x_prev = x;
    x_d = xd_buf;
    x_spare = spare;
    tau = R(1);                                                                      // :180
    if (ident) {
      Ax_d = x_d;
      At_grad_f_Ax_d = G[0];
      grad_f_Ax_d = At_grad_f_Ax_d;
      if ((rc = eval_f(x_d, At_grad_f_Ax_d))) return rc;                             // :185-187
    } else {
      if ((rc = mulA(Ad, d))) return rc;                                             // :181
      Ax_d = Ax;                                                                     // in place over Ax's buffer
      if ((rc = lincomb(1.0, Ax, 1.0, Ad, Ax_d, m))) return rc;                      // :184
      grad_f_Ax_d = GM[0];
      if ((rc = eval_f(Ax_d, grad_f_Ax_d))) return rc;                               // :185-186
      At_grad_f_Ax_d = G[0];
      if ((rc = mulAt(At_grad_f_Ax_d, grad_f_Ax_d))) return rc;                      // :187
    }
    // :189-194 -- renames instead of copies
    x = x_d;
    Ax = Ax_d;
    grad_f_Ax = grad_f_Ax_d;
    At_grad_f_Ax = At_grad_f_Ax_d;
    {
      void* t = z_curr;
      z_curr = z;
      z = t;
    }
    if (ident) Az = z_curr;
    {
      void* t = res_prev;
      res_prev = res;
      res = t;                                                                       // :177
    }
    // :196-198 -- `0.5 / gamma` is a Float64 expression in Julia: sigma and threshold are Float64 also for R = Float32
    const double sigma = (double)beta * (0.5 / (double)gamma) * (double)(R)(R(1) - alpha);
    const R tolF = R(10) * std::numeric_limits<R>::epsilon() * (R(1) + (R)fabs((double)FBE_x));
    const R nr = (R)sqrt(res_sq_x);
    const double threshold = (double)FBE_x - sigma * (double)(R)(nr * nr) + (double)tolF;
    if ((rc = step_kernel())) return rc;                                             // :199-201
    if (H && (rc = pb_lbfgs_update(ctx, H, x, x_prev, res, res_prev))) return rc;    // speculative :252
    if ((rc = read_step(true))) return rc;
    f_Ax_d = f_Ax;                                                                   // :187, :194
    R FBE_new = fbe();                                                               // :202
    bool moved = false;
    for (int k = 1; k <= o->max_backtracks; ++k) {                                   // :204-250
      if ((double)FBE_new <= threshold) break;
      moved = true;
      if (f_Az == inf && !ident && (rc = mulA(Az, z_curr))) return rc;               // :209-211
      tau = k >= o->max_backtracks ? R(0) : tau / R(2);                              // :213
      const R one_m = R(1) - tau;
      if (x == x_d) {                                                                // un-alias before overwriting x
        x = x_spare;
        if (ident) Ax = x;
      }
      if ((rc = lincomb((double)tau, x_d, (double)one_m, z_curr, x, n))) return rc;  // :214
      if (!ident) {
        if (Ax == Ax_d) Ax = other(AX, Ax_d);
        if ((rc = lincomb((double)tau, Ax_d, (double)one_m, Az, Ax, m))) return rc;  // :215
      }
      if (At_grad_f_Ax == At_grad_f_Ax_d) {
        At_grad_f_Ax = other(G, At_grad_f_Ax_d);
        if (ident) grad_f_Ax = At_grad_f_Ax;
      }
      if (!ident && grad_f_Ax == grad_f_Ax_d) grad_f_Ax = other(GM, grad_f_Ax_d);
      bool f_pending = false;
      if (o->quadratic) {                                                            // :217-237
        if (f_Az == inf) {
          if ((rc = eval_f(Az, grad_f_Az))) return rc;
          Rows cz;
          if ((rc = read_rows(ctx, &cz))) return rc;
          f_Az = f_value(cz);
        }
        if (c == inf) {
          if (!ident && (rc = mulAt(At_grad_f_Az, grad_f_Az))) return rc;
          c = f_Az;
          Rows cd;
          if ((rc = pb_dot(ctx, dtype, m, Ax_d, grad_f_Az))) return rc;
          if ((rc = read_rows(ctx, &cd))) return rc;
          const R d1 = (R)cd.aux;
          if ((rc = pb_dot(ctx, dtype, m, Az, grad_f_Az))) return rc;
          if ((rc = read_rows(ctx, &cd))) return rc;
          const R d2 = (R)cd.aux;
          b = d1 - d2;
          a = (f_Ax_d - b) - c;
        }
        f_Ax = (a * (tau * tau) + b * tau) + c;
        if (!ident && (rc = lincomb((double)tau, grad_f_Ax_d, (double)one_m, grad_f_Az, grad_f_Ax, m))) return rc;
        if ((rc = lincomb((double)tau, At_grad_f_Ax_d, (double)one_m, At_grad_f_Az, At_grad_f_Ax, n))) return rc;
      } else {                                                                       // :238-244
        if ((rc = eval_f(Ax, grad_f_Ax))) return rc;
        f_pending = true;
        if (!ident && (rc = mulAt(At_grad_f_Ax, grad_f_Ax))) return rc;
      }
      if ((rc = step_kernel())) return rc;                                           // :246-248
      if ((rc = read_step(f_pending))) return rc;
      FBE_new = fbe();                                                               // :249
      ++tau_backtracks;
    }
    if (H) {                                                                         // :252, :123-128
      if (moved) {
        if ((rc = pb_lbfgs_update(ctx, H, x, x_prev, res, res_prev))) return rc;
        Rows cu;
        if ((rc = read_rows(ctx, &cu))) return rc;
        sc.aux2 = cu.aux2;
        sc.aux3 = cu.aux3;
      }
      if ((rc = pb_lbfgs_commit(H, sc.aux2, sc.aux3, nullptr))) return rc;
    }
    return PB_OK;
  }

  int run(const void* x0, void* z_out, pb_panoc_result* out) {
    int rc = init(x0);
    int64_t k = 1;
    if (rc == PB_OK) {
      for (;; ++k) {                                                                 // src/ProximalAlgorithms.jl:114-123
        if (k >= o->maxit || stop()) break;
        if ((rc = step())) break;
      }
    }
    if (rc == PB_OK) rc = pb_copy(ctx, z_out, z, (size_t)n * es);                    // default_solution = state.z (:259)
    if (rc == PB_OK) {
      out->iterations = k;
      out->gamma_backtracks = gamma_backtracks;
      out->tau_backtracks = tau_backtracks;
      out->gamma = (double)gamma;
      out->f_Ax = (double)f_Ax;
      out->g_z = (double)g_z;
      out->res_inf = sc.res_inf;
      out->tau = (double)tau;
      out->warned_small_gamma = warned;
    }
    release();
    return rc;
  }
};

}  // namespace

extern "C" int pb_panoc_solve(pb_ctx* ctx, int dtype, int64_t n, const pb_smooth* f, const pb_prox* g, const pb_panoc_opts* o,
                              const void* x0, void* z_out, pb_panoc_result* out) {
  PB_REQUIRE(ctx != nullptr && f != nullptr && g != nullptr && o != nullptr && out != nullptr, "null argument");
  PB_REQUIRE(dtype == PB_F32 || dtype == PB_F64, "dtype must be PB_F32 or PB_F64");
  PB_REQUIRE(n >= 0 && o->maxit >= 1, "need n >= 0 and maxit >= 1");
  PB_REQUIRE(n == 0 || (x0 && z_out), "null vector");
  PB_REQUIRE(o->gamma > 0 || o->adaptive, "a fixed stepsize needs gamma > 0");
  PB_REQUIRE(o->lbfgs_mem >= 0 && o->lbfgs_mem <= 32 && o->max_backtracks >= 0, "bad L-BFGS memory / max_backtracks");
  PB_REQUIRE(g->kind == PB_PROX_ZERO || g->kind == PB_PROX_L1 || g->kind == PB_PROX_BOX || g->kind == PB_PROX_L21,
             "pb_panoc_solve supports the single-pass prox kinds (Zero, NormL1, IndBox, NormL21)");
  PB_REQUIRE(ctx->xchg_world <= 1, "pb_panoc_solve is single-GPU (the L-BFGS recursion needs un-sharded dot products)");
  PB_REQUIRE(o->A == nullptr || (o->Am > 0 && o->An == n), "A must be Am x n");
  memset(out, 0, sizeof(*out));
  PbDeviceGuard dev_guard(ctx);
  const int64_t m = o->A ? o->Am : n;
  if (dtype == PB_F32) {
    Panoc<float> s;
    memset(&s, 0, sizeof(s));
    s.ctx = ctx, s.dtype = dtype, s.n = n, s.m = m, s.f = f, s.g = g, s.o = o, s.ident = o->A == nullptr, s.es = 4;
    return s.run(x0, z_out, out);
  }
  Panoc<double> s;
  memset(&s, 0, sizeof(s));
  s.ctx = ctx, s.dtype = dtype, s.n = n, s.m = m, s.f = f, s.g = g, s.o = o, s.ident = o->A == nullptr, s.es = 8;
  return s.run(x0, z_out, out);
}
