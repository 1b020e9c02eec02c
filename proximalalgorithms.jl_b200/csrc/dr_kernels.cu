// dr_kernels.cu -- K8: fused Douglas-Rachford iteration (src/algorithms/douglas_rachford.jl:54-63) for element-wise prox pairs.
//
//   y = prox_{gamma f}(x);  r = 2y - x;  z = prox_{gamma g}(r);  res = y - z;  x <- x - res;   stop on norm(res, Inf)/gamma
//
// The reference makes five passes over six vectors (prox!, broadcast, prox!, broadcast, broadcast, norm: 13 vector
// reads/writes per element).  Here the whole iteration is ONE pass: read x (+ the data vector b of a translated
// quadratic), write x; y, r, z, res are only materialised on request (lazy state fields / the final solution), and
// norm(res, Inf) comes out of the same pass.  Algorithmic traffic: 2 vectors per iteration (3 with b).
// HBM-bound streaming kernel, no tensor-core formulation exists.
#include "step_common.cuh"

struct DrProx {
  int kind;            // PB_PROX_ZERO / L1 / BOX / SQRL2
  double a, b;         // L1: gl = gamma*lambda.  BOX: lo, hi.  SQRL2: den = 1 + gamma*lambda (all in the element type)
  const void* v0;      // BOX: per-element lo;  SQRL2: translation vector b (f(x) = lambda/2 ||x - b||^2), may be NULL
  const void* v1;      // BOX: per-element hi
};

struct DrParams {
  const void* x;
  void* x_out;
  void *y, *r, *z, *res;   // optional outputs
  int64_t n;
  DrProx f, g;
  PbWorkspace* ws;
  double* outs;
  XchgParams xchg;
};

// v0e / v1e: the element of the term's per-element vectors (BOX bounds, SQRL2 translation) when those vectors exist
// K >= 0: the kind is a compile-time constant (no per-element dispatch); K < 0: read it from the descriptor
template <typename T, int K>
__device__ __forceinline__ T dr_prox(const DrProx& p, T v, T v0e, T v1e) {
  const int kind = K >= 0 ? K : p.kind;
  switch (kind) {
    case PB_PROX_L1:
      return prox_elem<T, PB_PROX_L1>(v, (T)p.a, T(0));
    case PB_PROX_BOX:
      return prox_elem<T, PB_PROX_BOX>(v, p.v0 ? v0e : (T)p.a, p.v1 ? v1e : (T)p.b);
    case PB_PROX_SQRL2:
      // Translate(SqrNormL2(lambda), -b): w = v - b; w / (1 + gamma*lambda); + b   (three separately rounded operations)
      if (p.v0) return add_rn(sub_rn(v, v0e) / (T)p.a, v0e);
      return v / (T)p.a;
    default:
      return v;
  }
}

template <typename T, int VEC, int UNROLL, int FK, int GK>
__global__ void __launch_bounds__(PB_BLOCK) k_dr_step(DrParams p) {
  constexpr int64_t TILE = (int64_t)PB_BLOCK * VEC * UNROLL;
  const T* __restrict__ x = static_cast<const T*>(p.x);
  const T* __restrict__ f0 = static_cast<const T*>(p.f.v0);
  const T* __restrict__ f1 = static_cast<const T*>(p.f.v1);
  const T* __restrict__ g0 = static_cast<const T*>(p.g.v0);
  const T* __restrict__ g1 = static_cast<const T*>(p.g.v1);
  T* __restrict__ xo = static_cast<T*>(p.x_out);
  T* __restrict__ yo = static_cast<T*>(p.y);
  T* __restrict__ ro = static_cast<T*>(p.r);
  T* __restrict__ zo = static_cast<T*>(p.z);
  T* __restrict__ so = static_cast<T*>(p.res);
  // The only reduction is the stop norm max|res| (NaN-propagating), kept in the element type: the pass moves 8-12 bytes per
  // element, so per-element conversions to double (ncu: XU / ADU pipes) would make it instruction-bound.
  T mx = T(0);
  auto elem = [&](T xv, T f0e, T f1e, T g0e, T g1e, T& y, T& r, T& z, T& res, T& xn) {
    y = dr_prox<T, FK>(p.f, xv, f0e, f1e);           // :58
    r = sub_rn(mul_rn(T(2), y), xv);                 // :59
    z = dr_prox<T, GK>(p.g, r, g0e, g1e);            // :60
    res = sub_rn(y, z);                              // :61
    xn = sub_rn(xv, res);                            // :62
    const T ar = fabs(res);
    mx = (ar > mx || ar != ar) ? ar : mx;
  };
  auto do_pack = [&](int64_t i, const Pack<T, VEC>& xv, const Pack<T, VEC>& f0v) {
    Pack<T, VEC> f1v, g0v, g1v, y, r, z, res, xn;
    if (f1) f1v = ld_pack<T, VEC, false>(f1 + i);
    if (g0) g0v = ld_pack<T, VEC, false>(g0 + i);
    if (g1) g1v = ld_pack<T, VEC, false>(g1 + i);
#pragma unroll
    for (int e = 0; e < VEC; ++e)
      elem(xv.v[e], f0 ? f0v.v[e] : T(0), f1 ? f1v.v[e] : T(0), g0 ? g0v.v[e] : T(0), g1 ? g1v.v[e] : T(0), y.v[e], r.v[e],
           z.v[e], res.v[e], xn.v[e]);
    st_pack<T, VEC, true>(xo + i, xn);
    if (yo) st_pack<T, VEC, true>(yo + i, y);
    if (ro) st_pack<T, VEC, true>(ro + i, r);
    if (zo) st_pack<T, VEC, true>(zo + i, z);
    if (so) st_pack<T, VEC, true>(so + i, res);
  };
  // same balanced schedule as k_step: full rounds of one TILE per CTA (UNROLL packs per stream in flight per thread), then
  // the remaining tiles at pack granularity
  const int64_t ntiles = p.n / TILE;
  const int64_t rounds = ntiles / gridDim.x;
  for (int64_t rd_ = 0; rd_ < rounds; ++rd_) {
    const int64_t base = (rd_ * gridDim.x + blockIdx.x) * TILE + (int64_t)threadIdx.x * VEC;
    Pack<T, VEC> xv[UNROLL], fv[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int64_t i = base + (int64_t)u * PB_BLOCK * VEC;
      xv[u] = ld_pack<T, VEC, true>(x + i);
      if (f0) fv[u] = ld_pack<T, VEC, true>(f0 + i);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) do_pack(base + (int64_t)u * PB_BLOCK * VEC, xv[u], fv[u]);
  }
  const int64_t rem_start = rounds * gridDim.x * TILE;
  const int64_t rem_packs = (p.n - rem_start) / VEC;
  for (int64_t q = (int64_t)blockIdx.x * PB_BLOCK + threadIdx.x; q < rem_packs; q += (int64_t)gridDim.x * PB_BLOCK) {
    const int64_t i = rem_start + q * VEC;
    Pack<T, VEC> xq = ld_pack<T, VEC, true>(x + i), fq;
    if (f0) fq = ld_pack<T, VEC, true>(f0 + i);
    do_pack(i, xq, fq);
  }
  for (int64_t i = rem_start + rem_packs * VEC + (int64_t)blockIdx.x * PB_BLOCK + threadIdx.x; i < p.n;
       i += (int64_t)gridDim.x * PB_BLOCK) {
    T y, r, z, res, xn;
    elem(x[i], f0 ? f0[i] : T(0), f1 ? f1[i] : T(0), g0 ? g0[i] : T(0), g1 ? g1[i] : T(0), y, r, z, res, xn);
    xo[i] = xn;
    if (yo) yo[i] = y;
    if (ro) ro[i] = r;
    if (zo) zo[i] = z;
    if (so) so[i] = res;
  }
  Acc<0, 1> acc;
  acc.clear();
  acc.m[0] = (double)mx;
  OutMap map;
  map.sum_slot[0] = map.sum_slot[1] = map.sum_slot[2] = map.sum_slot[3] = -1;
  map.max_slot[0] = PB_S_RESINF;
  map.max_slot[1] = -1;
  grid_reduce<0, 1, PB_BLOCK>(acc, p.ws, p.outs, map, &p.xchg);
}

// prox parameters in the element type, combined on the host with one rounding each (like the package does)
template <typename T>
static int dr_fill(DrProx* d, const pb_prox* g, double gamma) {
  d->kind = g->kind;
  d->a = d->b = 0.0;
  d->v0 = d->v1 = nullptr;
  switch (g->kind) {
    case PB_PROX_ZERO:
      return PB_OK;
    case PB_PROX_L1:
      d->a = (double)mul_rn_host((T)gamma, (T)g->p0);
      return PB_OK;
    case PB_PROX_BOX:
      d->a = g->p0;
      d->b = g->p1;
      d->v0 = g->v0;
      d->v1 = g->v1;
      return PB_OK;
    case PB_PROX_SQRL2: {
      const T gl = mul_rn_host((T)gamma, (T)g->p0);
      volatile T den = T(1) + gl;
      d->a = (double)den;
      d->v0 = g->v0;
      return PB_OK;
    }
    default:
      pb_set_error("pb_dr_step: prox kind %d is not element-wise (use the unfused sequence)", g->kind);
      return PB_EUNSUPPORTED;
  }
}

extern "C" int pb_dr_step(pb_ctx* ctx, int dtype, int64_t n, const void* x, double gamma, const pb_prox* f, const pb_prox* g,
                          void* x_out, void* y, void* r, void* z, void* res) {
  PB_REQUIRE(ctx != nullptr, "null context");
  PB_REQUIRE(dtype == PB_F32 || dtype == PB_F64, "dtype must be PB_F32 or PB_F64");
  PB_REQUIRE(n >= 0, "negative length");
  PB_REQUIRE(f != nullptr && g != nullptr, "null prox descriptor");
  PB_REQUIRE(n == 0 || (x && x_out), "null vector");
  DrParams p;
  p.x = x;
  p.x_out = x_out;
  p.y = y;
  p.r = r;
  p.z = z;
  p.res = res;
  p.n = n;
  int rc = dtype == PB_F32 ? dr_fill<float>(&p.f, f, gamma) : dr_fill<double>(&p.f, f, gamma);
  if (rc != PB_OK) return rc;
  rc = dtype == PB_F32 ? dr_fill<float>(&p.g, g, gamma) : dr_fill<double>(&p.g, g, gamma);
  if (rc != PB_OK) return rc;
  p.ws = ctx->ws;
  p.outs = ctx->scalars_dev;
  pb_xchg_next(ctx, &p.xchg, ctx->xchg_fused != 0 && n > 0);
  const void* ptrs[] = {x, x_out, y, r, z, res, p.f.v0, p.f.v1, p.g.v0, p.g.v1};
  bool vec_ok = true;
  for (const void* q : ptrs) vec_ok = vec_ok && (!q || pb_aligned16(q));
  // grid = SMs x co-resident CTAs (a partial second wave of a grid-stride kernel serialises, see step_kernels.cu)
  auto launch = [&](auto kern, int64_t work_per_cta) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, PB_BLOCK, 0) != cudaSuccess || occ < 1) occ = 1;
    int per_sm = ctx->ctas_per_sm > 0 ? ctx->ctas_per_sm : 8;
    if (per_sm > occ) per_sm = occ;
    int64_t g_ = (int64_t)ctx->sm_count * per_sm;
    const int64_t need = (n + work_per_cta - 1) / work_per_cta;
    if (g_ > need) g_ = need;
    if (g_ < 1) g_ = 1;
    if (g_ > PB_MAX_CTAS) g_ = PB_MAX_CTAS;
    kern<<<(unsigned)g_, PB_BLOCK, 0, ctx->stream>>>(p);
  };
  // packs in flight per thread and stream: PB_OPT_UNROLL (1, 2, 4); default from the sweep in profiles/r01_next_rows.md.
  // The prox kinds are compile-time parameters of the vectorised kernels (16 pairs); misaligned views take a generic kernel.
  const int unroll = ctx->unroll == 1 || ctx->unroll == 2 || ctx->unroll == 4 ? ctx->unroll : 2;
  const int fk = p.f.kind, gk = p.g.kind;
#define PB_DR_U(T, VEC, FK, GK)                                                                  \
  do {                                                                                           \
    if (unroll == 4)                                                                             \
      launch(k_dr_step<T, VEC, 4, FK, GK>, (int64_t)PB_BLOCK * VEC * 4);                         \
    else if (unroll == 2)                                                                        \
      launch(k_dr_step<T, VEC, 2, FK, GK>, (int64_t)PB_BLOCK * VEC * 2);                         \
    else                                                                                         \
      launch(k_dr_step<T, VEC, 1, FK, GK>, (int64_t)PB_BLOCK * VEC);                             \
  } while (0)
#define PB_DR_G(T, VEC, FK)                                                                      \
  do {                                                                                           \
    switch (gk) {                                                                                \
      case PB_PROX_ZERO: PB_DR_U(T, VEC, FK, PB_PROX_ZERO); break;                               \
      case PB_PROX_L1: PB_DR_U(T, VEC, FK, PB_PROX_L1); break;                                   \
      case PB_PROX_BOX: PB_DR_U(T, VEC, FK, PB_PROX_BOX); break;                                 \
      default: PB_DR_U(T, VEC, FK, PB_PROX_SQRL2); break;                                        \
    }                                                                                            \
  } while (0)
#define PB_DR_F(T, VEC)                                                                          \
  do {                                                                                           \
    switch (fk) {                                                                                \
      case PB_PROX_ZERO: PB_DR_G(T, VEC, PB_PROX_ZERO); break;                                   \
      case PB_PROX_L1: PB_DR_G(T, VEC, PB_PROX_L1); break;                                       \
      case PB_PROX_BOX: PB_DR_G(T, VEC, PB_PROX_BOX); break;                                     \
      default: PB_DR_G(T, VEC, PB_PROX_SQRL2); break;                                            \
    }                                                                                            \
  } while (0)
  if (dtype == PB_F32) {
    if (vec_ok)
      PB_DR_F(float, 4);
    else
      launch(k_dr_step<float, 1, 4, -1, -1>, (int64_t)PB_BLOCK * 4);
  } else {
    if (vec_ok)
      PB_DR_F(double, 2);
    else
      launch(k_dr_step<double, 1, 4, -1, -1>, (int64_t)PB_BLOCK * 4);
  }
#undef PB_DR_F
#undef PB_DR_G
#undef PB_DR_U
  PB_LAUNCH_CHECK(ctx);
  return PB_OK;
}

// standalone prox of the translated quadratic (pb_prox_apply dispatches here): z = (y - b)/(1 + gamma*lambda) + b,
// GSUM = sum ((y - b)/(1 + gamma*lambda))^2, so that f(z) = lambda/2 * GSUM (ProximalOperators SqrNormL2 / Translate).
template <typename T>
__global__ void __launch_bounds__(PB_BLOCK) k_prox_sqrl2(const T* __restrict__ y, const T* __restrict__ b, T* __restrict__ z,
                                                         int64_t n, double den_d, PbWorkspace* ws, double* outs) {
  constexpr bool COMP = sizeof(T) == 8;
  const T den = (T)den_d;
  Acc<1, 0> acc;
  acc.clear();
  for (int64_t i = (int64_t)blockIdx.x * PB_BLOCK + threadIdx.x; i < n; i += (int64_t)gridDim.x * PB_BLOCK) {
    const T w = b ? sub_rn(y[i], b[i]) / den : y[i] / den;
    z[i] = b ? add_rn(w, b[i]) : w;
    if (COMP)
      dd_add_prod(acc.s[0], (double)w, (double)w);
    else
      acc.s[0].hi = __fma_rn((double)w, (double)w, acc.s[0].hi);
  }
  OutMap map;
  map.sum_slot[0] = PB_S_GSUM;
  map.sum_slot[1] = map.sum_slot[2] = map.sum_slot[3] = -1;
  map.max_slot[0] = map.max_slot[1] = -1;
  grid_reduce<1, 0, PB_BLOCK>(acc, ws, outs, map);
}

int pb_prox_sqrl2_apply(pb_ctx* ctx, int dtype, int64_t n, const void* y, double gamma, const pb_prox* g, void* z) {
  PB_REQUIRE(ctx != nullptr, "null context");
  PB_REQUIRE(dtype == PB_F32 || dtype == PB_F64, "dtype must be PB_F32 or PB_F64");
  PB_REQUIRE(n >= 0, "negative length");
  PB_REQUIRE(n == 0 || (y && z), "null vector");
  DrProx d;
  const int rc = dtype == PB_F32 ? dr_fill<float>(&d, g, gamma) : dr_fill<double>(&d, g, gamma);
  if (rc != PB_OK) return rc;
  const int grid = pb_stream_grid(ctx, PB_BLOCK * 4, n, 4);
  if (dtype == PB_F32)
    k_prox_sqrl2<float><<<grid, PB_BLOCK, 0, ctx->stream>>>((const float*)y, (const float*)d.v0, (float*)z, n, d.a, ctx->ws,
                                                             ctx->scalars_dev);
  else
    k_prox_sqrl2<double><<<grid, PB_BLOCK, 0, ctx->stream>>>((const double*)y, (const double*)d.v0, (double*)z, n, d.a, ctx->ws,
                                                              ctx->scalars_dev);
  PB_LAUNCH_CHECK(ctx);
  return PB_OK;
}

// ---- prox of a convex conjugate by the Moreau identity (ProximalCore ConvexConjugate; primal_dual.jl:195) -------------------
// out = v - gamma * prox_{h/gamma}(v/gamma) for the element-wise kinds; Zero* = IndZero -> out = 0.  One pass (2 vectors).
template <typename T>
__global__ void __launch_bounds__(PB_BLOCK) k_conj_prox(const T* __restrict__ v, T* __restrict__ out, int64_t n, DrProx d,
                                                        double gamma_d, int zero_conj) {
  const T gamma = (T)gamma_d;
  const T* __restrict__ v0 = static_cast<const T*>(d.v0);
  const T* __restrict__ v1 = static_cast<const T*>(d.v1);
  for (int64_t i = (int64_t)blockIdx.x * PB_BLOCK + threadIdx.x; i < n; i += (int64_t)gridDim.x * PB_BLOCK) {
    if (zero_conj) {
      out[i] = T(0);
      continue;
    }
    const T vv = v[i];
    const T p = dr_prox<T, -1>(d, vv / gamma, v0 ? v0[i] : T(0), v1 ? v1[i] : T(0));
    out[i] = sub_rn(vv, mul_rn(gamma, p));
  }
}

extern "C" int pb_conj_prox(pb_ctx* ctx, int dtype, int64_t n, const void* v, double gamma, const pb_prox* h, void* out) {
  PB_REQUIRE(ctx != nullptr, "null context");
  PB_REQUIRE(dtype == PB_F32 || dtype == PB_F64, "dtype must be PB_F32 or PB_F64");
  PB_REQUIRE(n >= 0, "negative length");
  PB_REQUIRE(h != nullptr, "null prox descriptor");
  PB_REQUIRE(n == 0 || (v && out), "null vector");
  PB_REQUIRE(gamma > 0, "gamma must be positive");
  DrProx d;
  int rc;
  // the inner prox runs with parameter 1/gamma, formed in the element type
  if (dtype == PB_F32) {
    volatile float ig = 1.0f / (float)gamma;
    rc = dr_fill<float>(&d, h, (double)ig);
  } else {
    volatile double ig = 1.0 / gamma;
    rc = dr_fill<double>(&d, h, ig);
  }
  if (rc != PB_OK) return rc;
  const int grid = pb_stream_grid(ctx, PB_BLOCK * 4, n, 4);
  const int zc = h->kind == PB_PROX_ZERO;
  if (dtype == PB_F32)
    k_conj_prox<float><<<grid, PB_BLOCK, 0, ctx->stream>>>((const float*)v, (float*)out, n, d, gamma, zc);
  else
    k_conj_prox<double><<<grid, PB_BLOCK, 0, ctx->stream>>>((const double*)v, (double*)out, n, d, gamma, zc);
  PB_LAUNCH_CHECK(ctx);
  return PB_OK;
}
