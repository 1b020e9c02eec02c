// qn_kernels.cu -- K7: device-resident L-BFGS (src/accel/lbfgs.jl) and the vector algebra of PANOC's line search
// (src/algorithms/panoc.jl:114-128, :183-184, :214-215, :228-237).
//
// B200 design notes
//   * The two-loop recursion (lbfgs.jl:66-95) is a chain of 2m+1 launches with NO host round trip: every launch fuses
//     "d <- d -/+ c*u" with the dot product the NEXT step needs, and the coefficient c = dot/ys (alpha of loop 1,
//     alpha - beta of loop 2) is formed on the device, by every thread, from the double-double dot the previous launch
//     left in the scalar block.  alpha_i stays in device memory between the two loops.  The unfused reference does
//     2m dots + 2m axpys + 2 scalings = 5m+4 vector passes plus 2m host-visible scalars; this chain does 4 passes per
//     launch (read d, u, w; write d) and the last launch also emits x_d = x + d (panoc.jl:183).
//   * update! (lbfgs.jl:30-51) is one pass: s = x - x_prev and y = res - res_prev (panoc.jl:125-126) are written straight
//     into the ring's spare slot while <s,y> and <y,y> are reduced; the host only commits the slot if <s,y> > 0.
//     (The reference: 2 subtractions, 2 copies into L.s/L.y, dot, 2 copies into the ring, dot = 18 vector passes; here 6.)
//   * All kernels are single coalesced HBM passes with 16-byte packs; products are rounded separately from sums exactly
//     like Julia's broadcasts (no FMA contraction); dots are exact for float data and double-double for double data.
#include <new>
#include <string.h>

#include "common.cuh"

// ---------------------------------------------------------------------------------------------------------------
// out = a.*x .+ b.*y   (panoc.jl:183-184 with a = b = 1, :214-215 and :234-237 with a = tau, b = 1 - tau)
// out = s.*x           (panoc.jl:116 `d .*= -1`, :120 `.-res`)
// ---------------------------------------------------------------------------------------------------------------
template <typename T, int VEC, bool TWO>
__global__ void __launch_bounds__(PB_BLOCK) k_lincomb(const T* __restrict__ x, const T* __restrict__ y, T* __restrict__ out,
                                                      int64_t n, double a_d, double b_d) {
  const T a = (T)a_d, b = (T)b_d;
  const int64_t npacks = n / VEC;
  for (int64_t q = (int64_t)blockIdx.x * PB_BLOCK + threadIdx.x; q < npacks; q += (int64_t)gridDim.x * PB_BLOCK) {
    const int64_t i = q * VEC;
    Pack<T, VEC> xv = ld_pack<T, VEC, false>(x + i), yv, o;
    if constexpr (TWO) yv = ld_pack<T, VEC, false>(y + i);
#pragma unroll
    for (int e = 0; e < VEC; ++e) o.v[e] = TWO ? add_rn(mul_rn(a, xv.v[e]), mul_rn(b, yv.v[e])) : mul_rn(a, xv.v[e]);
    st_pack<T, VEC, false>(out + i, o);
  }
  for (int64_t i = npacks * VEC + (int64_t)blockIdx.x * PB_BLOCK + threadIdx.x; i < n; i += (int64_t)gridDim.x * PB_BLOCK)
    out[i] = TWO ? add_rn(mul_rn(a, x[i]), mul_rn(b, y[i])) : mul_rn(a, x[i]);
}

template <typename T, bool TWO>
static int launch_lincomb(pb_ctx* ctx, int64_t n, double a, const void* x, double b, const void* y, void* out) {
  constexpr int VEC = 16 / sizeof(T);
  const bool vec_ok = pb_aligned16(x) && (!TWO || pb_aligned16(y)) && pb_aligned16(out);
  if (vec_ok) {
    const int grid = pb_stream_grid(ctx, (int64_t)PB_BLOCK * VEC * 4, n, 4);
    k_lincomb<T, VEC, TWO><<<grid, PB_BLOCK, 0, ctx->stream>>>((const T*)x, (const T*)y, (T*)out, n, a, b);
  } else {
    const int grid = pb_stream_grid(ctx, (int64_t)PB_BLOCK * 4, n, 4);
    k_lincomb<T, 1, TWO><<<grid, PB_BLOCK, 0, ctx->stream>>>((const T*)x, (const T*)y, (T*)out, n, a, b);
  }
  PB_LAUNCH_CHECK(ctx);
  return PB_OK;
}

extern "C" int pb_lincomb2(pb_ctx* ctx, int dtype, int64_t n, double a, const void* x, double b, const void* y, void* out) {
  PB_REQUIRE(ctx != nullptr, "null context");
  PB_REQUIRE(dtype == PB_F32 || dtype == PB_F64, "dtype must be PB_F32 or PB_F64");
  PB_REQUIRE(n >= 0, "negative length");
  PB_REQUIRE(n == 0 || (x && y && out), "null vector");
  return dtype == PB_F32 ? launch_lincomb<float, true>(ctx, n, a, x, b, y, out)
                         : launch_lincomb<double, true>(ctx, n, a, x, b, y, out);
}

extern "C" int pb_scale(pb_ctx* ctx, int dtype, int64_t n, double s, const void* x, void* out) {
  PB_REQUIRE(ctx != nullptr, "null context");
  PB_REQUIRE(dtype == PB_F32 || dtype == PB_F64, "dtype must be PB_F32 or PB_F64");
  PB_REQUIRE(n >= 0, "negative length");
  PB_REQUIRE(n == 0 || (x && out), "null vector");
  return dtype == PB_F32 ? launch_lincomb<float, false>(ctx, n, s, x, 0.0, nullptr, out)
                         : launch_lincomb<double, false>(ctx, n, s, x, 0.0, nullptr, out);
}

// ---------------------------------------------------------------------------------------------------------------
// L-BFGS operator
// ---------------------------------------------------------------------------------------------------------------
#define PB_LBFGS_MAX_MEM 32

struct pb_lbfgs {
  pb_ctx* ctx;
  int dtype;
  int64_t n;
  int mem;                               // M of LBFGS(M)
  int currmem, curridx;                  // lbfgs.jl:6-7 (curridx is 1-based as in the reference, 0 = empty)
  int slot_of[PB_LBFGS_MAX_MEM + 1];     // logical ring position (1..M) -> physical slot (0..M)
  int spare;                             // physical slot the next update writes to
  int pending;                           // an update kernel has filled `spare` and waits for pb_lbfgs_commit
  void* store;                           // 2*(M+1) vectors: s slots then y slots
  size_t stride;                         // bytes between consecutive slots (16-byte aligned)
  double ys[PB_LBFGS_MAX_MEM + 1];       // <s,y> per physical slot, held in the element type (lbfgs.jl:11)
  double H;                              // lbfgs.jl:13
  double* alpha_dev;                     // lbfgs.jl:12, lives on the device between loop 1 and loop 2
};

static inline void* lb_s(const pb_lbfgs* L, int slot) { return (char*)L->store + (size_t)slot * L->stride; }
static inline void* lb_y(const pb_lbfgs* L, int slot) { return (char*)L->store + (size_t)(L->mem + 1 + slot) * L->stride; }

extern "C" int pb_lbfgs_create(pb_ctx* ctx, int dtype, int64_t n, int mem, pb_lbfgs** out) {
  PB_REQUIRE(ctx != nullptr && out != nullptr, "null argument");
  PB_REQUIRE(dtype == PB_F32 || dtype == PB_F64, "dtype must be PB_F32 or PB_F64");
  PB_REQUIRE(n >= 0, "negative length");
  PB_REQUIRE(mem >= 1 && mem <= PB_LBFGS_MAX_MEM, "memory must be in 1..32");
  pb_lbfgs* L = new (std::nothrow) pb_lbfgs();
  if (!L) {
    pb_set_error("pb_lbfgs_create: out of host memory");
    return PB_ENOMEM;
  }
  L->ctx = ctx;
  L->dtype = dtype;
  L->n = n;
  L->mem = mem;
  L->currmem = L->curridx = 0;
  L->spare = 0;
  L->pending = 0;
  L->H = 1.0;
  const size_t es = dtype == PB_F32 ? 4 : 8;
  L->stride = (((size_t)n * es + 255) / 256) * 256;
  if (L->stride == 0) L->stride = 256;
  L->store = nullptr;
  L->alpha_dev = nullptr;
  cudaError_t e = cudaSetDevice(ctx->device);
  if (e == cudaSuccess) e = cudaMalloc(&L->store, L->stride * 2 * (size_t)(mem + 1));
  if (e == cudaSuccess) e = cudaMalloc((void**)&L->alpha_dev, sizeof(double) * (PB_LBFGS_MAX_MEM + 1));
  if (e == cudaSuccess) e = cudaMemsetAsync(L->store, 0, L->stride * 2 * (size_t)(mem + 1), ctx->stream);   // zero(x), lbfgs.jl:17-18
  if (e == cudaSuccess) e = cudaMemsetAsync(L->alpha_dev, 0, sizeof(double) * (PB_LBFGS_MAX_MEM + 1), ctx->stream);
  if (e != cudaSuccess) {
    pb_set_error("pb_lbfgs_create: %s", cudaGetErrorString(e));
    if (L->store) cudaFree(L->store);
    if (L->alpha_dev) cudaFree(L->alpha_dev);
    delete L;
    return e == cudaErrorMemoryAllocation ? PB_ENOMEM : PB_ECUDA;
  }
  for (int k = 0; k <= PB_LBFGS_MAX_MEM; ++k) {
    L->slot_of[k] = -1;
    L->ys[k] = 0.0;
  }
  *out = L;
  return PB_OK;
}

extern "C" int pb_lbfgs_destroy(pb_lbfgs* L) {
  if (!L) return PB_OK;
  cudaStreamSynchronize(L->ctx->stream);
  if (L->store) cudaFree(L->store);
  if (L->alpha_dev) cudaFree(L->alpha_dev);
  delete L;
  return PB_OK;
}

// reset! (lbfgs.jl:53-56)
extern "C" int pb_lbfgs_reset(pb_lbfgs* L) {
  PB_REQUIRE(L != nullptr, "null operator");
  L->currmem = L->curridx = 0;
  L->H = 1.0;
  L->pending = 0;
  return PB_OK;
}

extern "C" int pb_lbfgs_info(const pb_lbfgs* L, int* currmem, int* curridx, double* H) {
  PB_REQUIRE(L != nullptr, "null operator");
  if (currmem) *currmem = L->currmem;
  if (curridx) *curridx = L->curridx;
  if (H) *H = L->H;
  return PB_OK;
}

// Device address of the stored pair at logical ring position `pos` (1..M); NULL pointers are skipped.
extern "C" int pb_lbfgs_pair(const pb_lbfgs* L, int pos, void** s, void** y, double* ys) {
  PB_REQUIRE(L != nullptr, "null operator");
  PB_REQUIRE(pos >= 1 && pos <= L->mem && L->slot_of[pos] >= 0, "empty ring position");
  if (s) *s = lb_s(L, L->slot_of[pos]);
  if (y) *y = lb_y(L, L->slot_of[pos]);
  if (ys) *ys = L->ys[L->slot_of[pos]];
  return PB_OK;
}

// ---- update: s = a - a_prev, y = b - b_prev into the spare slot; AUX2 = <s,y>, AUX3 = <y,y> --------------------------
template <typename T, int VEC>
__global__ void __launch_bounds__(PB_BLOCK) k_lbfgs_update(const T* __restrict__ a, const T* __restrict__ ap,
                                                           const T* __restrict__ b, const T* __restrict__ bp,
                                                           T* __restrict__ so, T* __restrict__ yo, int64_t n,
                                                           PbWorkspace* ws, double* outs) {
  constexpr bool COMP = sizeof(T) == 8;
  Acc<2, 0> acc;
  acc.clear();
  auto elem = [&](T av, T apv, T bv, T bpv, T& s, T& y) {
    s = ap ? sub_rn(av, apv) : av;
    y = bp ? sub_rn(bv, bpv) : bv;
    const double sd = (double)s, yd = (double)y;
    if (COMP) {
      dd_add_prod(acc.s[0], sd, yd);
      dd_add_prod(acc.s[1], yd, yd);
    } else {
      acc.s[0].hi = __fma_rn(sd, yd, acc.s[0].hi);
      acc.s[1].hi = __fma_rn(yd, yd, acc.s[1].hi);
    }
  };
  const int64_t npacks = n / VEC;
  for (int64_t q = (int64_t)blockIdx.x * PB_BLOCK + threadIdx.x; q < npacks; q += (int64_t)gridDim.x * PB_BLOCK) {
    const int64_t i = q * VEC;
    Pack<T, VEC> av = ld_pack<T, VEC, false>(a + i), bv = ld_pack<T, VEC, false>(b + i), apv, bpv, s, y;
    if (ap) apv = ld_pack<T, VEC, false>(ap + i);
    if (bp) bpv = ld_pack<T, VEC, false>(bp + i);
#pragma unroll
    for (int e = 0; e < VEC; ++e) elem(av.v[e], ap ? apv.v[e] : T(0), bv.v[e], bp ? bpv.v[e] : T(0), s.v[e], y.v[e]);
    st_pack<T, VEC, false>(so + i, s);
    st_pack<T, VEC, false>(yo + i, y);
  }
  for (int64_t i = npacks * VEC + (int64_t)blockIdx.x * PB_BLOCK + threadIdx.x; i < n; i += (int64_t)gridDim.x * PB_BLOCK) {
    T s, y;
    elem(a[i], ap ? ap[i] : T(0), b[i], bp ? bp[i] : T(0), s, y);
    so[i] = s;
    yo[i] = y;
  }
  OutMap map;
  map.sum_slot[0] = PB_S_AUX2;
  map.sum_slot[1] = PB_S_AUX3;
  map.sum_slot[2] = map.sum_slot[3] = -1;
  map.max_slot[0] = map.max_slot[1] = -1;
  grid_reduce<2, 0, PB_BLOCK>(acc, ws, outs, map);
}

extern "C" int pb_lbfgs_update(pb_ctx* ctx, pb_lbfgs* L, const void* a, const void* a_prev, const void* b,
                               const void* b_prev) {
  PB_REQUIRE(ctx != nullptr && L != nullptr, "null argument");
  PB_REQUIRE(L->ctx == ctx, "operator belongs to another context");
  PB_REQUIRE(L->n == 0 || (a && b), "null vector");
  void* so = lb_s(L, L->spare);
  void* yo = lb_y(L, L->spare);
  const int64_t n = L->n;
  const bool vec_ok = pb_aligned16(a) && pb_aligned16(b) && (!a_prev || pb_aligned16(a_prev)) && (!b_prev || pb_aligned16(b_prev));
  if (L->dtype == PB_F32) {
    if (vec_ok) {
      const int grid = pb_stream_grid(ctx, PB_BLOCK * 4 * 4, n, 4);
      k_lbfgs_update<float, 4><<<grid, PB_BLOCK, 0, ctx->stream>>>((const float*)a, (const float*)a_prev, (const float*)b,
                                                                    (const float*)b_prev, (float*)so, (float*)yo, n, ctx->ws,
                                                                    ctx->scalars_dev);
    } else {
      const int grid = pb_stream_grid(ctx, PB_BLOCK * 4, n, 4);
      k_lbfgs_update<float, 1><<<grid, PB_BLOCK, 0, ctx->stream>>>((const float*)a, (const float*)a_prev, (const float*)b,
                                                                    (const float*)b_prev, (float*)so, (float*)yo, n, ctx->ws,
                                                                    ctx->scalars_dev);
    }
  } else {
    if (vec_ok) {
      const int grid = pb_stream_grid(ctx, PB_BLOCK * 2 * 4, n, 4);
      k_lbfgs_update<double, 2><<<grid, PB_BLOCK, 0, ctx->stream>>>((const double*)a, (const double*)a_prev, (const double*)b,
                                                                     (const double*)b_prev, (double*)so, (double*)yo, n,
                                                                     ctx->ws, ctx->scalars_dev);
    } else {
      const int grid = pb_stream_grid(ctx, PB_BLOCK * 4, n, 4);
      k_lbfgs_update<double, 1><<<grid, PB_BLOCK, 0, ctx->stream>>>((const double*)a, (const double*)a_prev, (const double*)b,
                                                                     (const double*)b_prev, (double*)so, (double*)yo, n,
                                                                     ctx->ws, ctx->scalars_dev);
    }
  }
  PB_LAUNCH_CHECK(ctx);
  L->pending = 1;
  return PB_OK;
}

// Host half of update! (lbfgs.jl:33-48): `ys`, `yty` are the (rank-combined) sums the update kernel reduced, as read by
// the caller's per-iteration exchange.  Scalars are kept in the element type.  *accepted = 1 if the pair entered the ring.
extern "C" int pb_lbfgs_commit(pb_lbfgs* L, double ys, double yty, int* accepted) {
  PB_REQUIRE(L != nullptr, "null operator");
  PB_REQUIRE(L->pending, "no update is pending");
  L->pending = 0;
  int acc = 0;
  if (L->dtype == PB_F32) {
    const float ysf = (float)ys, ytyf = (float)yty;
    if (ysf > 0.0f) {
      acc = 1;
      L->ys[L->spare] = (double)ysf;
      L->H = (double)(ysf / ytyf);
    }
  } else if (ys > 0.0) {
    acc = 1;
    L->ys[L->spare] = ys;
    L->H = ys / yty;
  }
  if (acc) {
    L->curridx += 1;                                  // lbfgs.jl:35-38
    if (L->curridx > L->mem) L->curridx = 1;
    L->currmem += 1;                                  // lbfgs.jl:39-42
    if (L->currmem > L->mem) L->currmem = L->mem;
    const int evicted = L->slot_of[L->curridx];       // the pair this position held (if any) becomes the spare
    L->slot_of[L->curridx] = L->spare;
    if (evicted >= 0) {
      L->spare = evicted;
    } else {
      int used[PB_LBFGS_MAX_MEM + 1] = {0};
      for (int k = 1; k <= L->mem; ++k)
        if (L->slot_of[k] >= 0) used[L->slot_of[k]] = 1;
      for (int k = 0; k <= L->mem; ++k)
        if (!used[k]) {
          L->spare = k;
          break;
        }
    }
  }
  if (accepted) *accepted = acc;
  return PB_OK;
}

// ---- two-loop recursion ----------------------------------------------------------------------------------------------
struct QnParams {
  const void* d_in;    // current d (v for the first launch)
  const void* u;       // y_i (loop 1) or s_i (loop 2); unused in mode 2
  const void* w;       // vector of the NEXT dot product, or NULL
  void* d_out;
  const void* x;       // optional: xd = x + d_out (last launch only)
  void* xd;
  int64_t n;
  double ys;           // <s_i, y_i> of the pair being applied
  double post1, post2; // multipliers applied after the update (H after loop 1; the caller's scale at the very end)
  int mode;            // 0: alpha = dot/ys, store, d -= alpha*u;  1: beta = dot/ys, d += (alpha - beta)*u;  2: d = d_in
  int slot;
  double* alpha_dev;
  PbWorkspace* ws;
  double* outs;
  const double* dot_src;   // scalar block the coefficient's dot product is read from: `outs` on one GPU, the rank-combined copy
                           // (pb_ctx::chain_dev) when the vectors are row shards
  XchgParams xchg;         // world > 0: the dot this launch reduces is exchanged and folded over the ranks inside the kernel
};

template <typename T, int VEC>
__global__ void __launch_bounds__(PB_BLOCK) k_qn(QnParams p) {
  constexpr bool COMP = sizeof(T) == 8;
  const T* __restrict__ d = static_cast<const T*>(p.d_in);
  const T* __restrict__ u = static_cast<const T*>(p.u);
  const T* __restrict__ w = static_cast<const T*>(p.w);
  const T* __restrict__ x = static_cast<const T*>(p.x);
  T* __restrict__ out = static_cast<T*>(p.d_out);
  T* __restrict__ xd = static_cast<T*>(p.xd);
  // coefficient from the dot product the previous launch left in the scalar block (kernel boundary = visibility)
  T c = T(0);
  if (p.mode != 2) {
    const volatile double* o = p.dot_src;
    const T dotv = (T)(o[PB_S_AUX] + o[PB_S_AUX + 1]);          // real(dot(.,.)) rounded to R
    const T q = dotv / (T)p.ys;                                  // lbfgs.jl:77, :92
    if (p.mode == 0) {
      c = q;
      if (blockIdx.x == 0 && threadIdx.x == 0) p.alpha_dev[p.slot] = (double)q;
    } else {
      c = sub_rn((T)(*(const volatile double*)(p.alpha_dev + p.slot)), q);   // alphas[idx] - beta, lbfgs.jl:93
    }
  }
  const T post1 = (T)p.post1, post2 = (T)p.post2;
  const int mode = p.mode;
  Acc<1, 0> acc;
  acc.clear();
  auto elem = [&](T dv, T uv, T wv, T xv, T& o_, T& xo_) {
    T t = mode == 0 ? sub_rn(dv, mul_rn(c, uv)) : (mode == 1 ? add_rn(dv, mul_rn(c, uv)) : dv);
    t = mul_rn(mul_rn(t, post1), post2);
    o_ = t;
    if (w) {
      if (COMP)
        dd_add_prod(acc.s[0], (double)wv, (double)t);
      else
        acc.s[0].hi = __fma_rn((double)wv, (double)t, acc.s[0].hi);
    }
    if (x) xo_ = add_rn(xv, t);
  };
  const int64_t npacks = p.n / VEC;
  for (int64_t q = (int64_t)blockIdx.x * PB_BLOCK + threadIdx.x; q < npacks; q += (int64_t)gridDim.x * PB_BLOCK) {
    const int64_t i = q * VEC;
    Pack<T, VEC> dv = ld_pack<T, VEC, false>(d + i), uv, wv, xv, o_, xo_;
    if (mode != 2) uv = ld_pack<T, VEC, false>(u + i);
    if (w) wv = ld_pack<T, VEC, false>(w + i);
    if (x) xv = ld_pack<T, VEC, false>(x + i);
#pragma unroll
    for (int e = 0; e < VEC; ++e)
      elem(dv.v[e], mode != 2 ? uv.v[e] : T(0), w ? wv.v[e] : T(0), x ? xv.v[e] : T(0), o_.v[e], xo_.v[e]);
    st_pack<T, VEC, false>(out + i, o_);
    if (x) st_pack<T, VEC, false>(xd + i, xo_);
  }
  for (int64_t i = npacks * VEC + (int64_t)blockIdx.x * PB_BLOCK + threadIdx.x; i < p.n; i += (int64_t)gridDim.x * PB_BLOCK) {
    T o_, xo_;
    elem(d[i], mode != 2 ? u[i] : T(0), w ? w[i] : T(0), x ? x[i] : T(0), o_, xo_);
    out[i] = o_;
    if (x) xd[i] = xo_;
  }
  if (w) {
    OutMap map;
    map.sum_slot[0] = PB_S_AUX;
    map.sum_slot[1] = map.sum_slot[2] = map.sum_slot[3] = -1;
    map.max_slot[0] = map.max_slot[1] = -1;
    grid_reduce<1, 0, PB_BLOCK>(acc, p.ws, p.outs, map, &p.xchg);
  }
}

static int launch_qn(pb_ctx* ctx, int dtype, QnParams p) {
  // row shards (device exchange attached, world > 1): every dot of the chain is summed over the ranks inside the kernel
  const bool sharded = ctx->xchg_world > 1 && ctx->xchg_connected;
  memset(&p.xchg, 0, sizeof(p.xchg));
  p.dot_src = sharded ? ctx->chain_dev : p.outs;
  if (sharded && p.w) {
    pb_xchg_next(ctx, &p.xchg, true);
    p.xchg.gather_out = ctx->chain_dev;
    ctx->xchg_pending = 0;              // nothing of this exchange reaches the host: a later read publishes the block afresh
  }
  const bool vec_ok = pb_aligned16(p.d_in) && pb_aligned16(p.d_out) && (!p.u || pb_aligned16(p.u)) &&
                      (!p.w || pb_aligned16(p.w)) && (!p.x || (pb_aligned16(p.x) && pb_aligned16(p.xd)));
  if (dtype == PB_F32) {
    if (vec_ok)
      k_qn<float, 4><<<pb_stream_grid(ctx, PB_BLOCK * 4 * 4, p.n, 4), PB_BLOCK, 0, ctx->stream>>>(p);
    else
      k_qn<float, 1><<<pb_stream_grid(ctx, PB_BLOCK * 4, p.n, 4), PB_BLOCK, 0, ctx->stream>>>(p);
  } else {
    if (vec_ok)
      k_qn<double, 2><<<pb_stream_grid(ctx, PB_BLOCK * 2 * 4, p.n, 4), PB_BLOCK, 0, ctx->stream>>>(p);
    else
      k_qn<double, 1><<<pb_stream_grid(ctx, PB_BLOCK * 4, p.n, 4), PB_BLOCK, 0, ctx->stream>>>(p);
  }
  PB_LAUNCH_CHECK(ctx);
  return PB_OK;
}

// d = scale * (H * v)   (mul!, lbfgs.jl:66-72, followed by PANOC's `d .*= -1` when scale = -1, panoc.jl:114-117);
// optionally x_d = x + d (panoc.jl:183) in the last launch.  d may alias v.  Asynchronous: no host read-back.
extern "C" int pb_lbfgs_apply(pb_ctx* ctx, pb_lbfgs* L, const void* v, double scale, void* d, const void* x, void* x_d) {
  PB_REQUIRE(ctx != nullptr && L != nullptr, "null argument");
  PB_REQUIRE(L->ctx == ctx, "operator belongs to another context");
  PB_REQUIRE(L->n == 0 || (v && d), "null vector");
  PB_REQUIRE((x == nullptr) == (x_d == nullptr), "x and x_d go together");
  PB_REQUIRE(!L->pending, "an update is pending: call pb_lbfgs_commit first");
  const int m = L->currmem, M = L->mem;
  QnParams p;
  p.n = L->n;
  p.alpha_dev = L->alpha_dev;
  p.ws = ctx->ws;
  p.outs = ctx->scalars_dev;
  p.x = nullptr;
  p.xd = nullptr;
  if (m == 0) {                                       // d = (v * H) * scale
    p.d_in = v;
    p.u = p.w = nullptr;
    p.d_out = d;
    p.ys = 1.0;
    p.post1 = L->H;
    p.post2 = scale;
    p.mode = 2;
    p.slot = 0;
    p.x = x;
    p.xd = x_d;
    return launch_qn(ctx, L->dtype, p);
  }
  // visiting order of loop 1: curridx, curridx-1, ... (wrapping at 0 -> M), lbfgs.jl:74-84
  int order[PB_LBFGS_MAX_MEM];
  int idx = L->curridx;
  for (int k = 0; k < m; ++k) {
    order[k] = L->slot_of[idx];
    idx -= 1;
    if (idx == 0) idx = M;
  }
  // first dot: <s_{i1}, v>
  int rc;
  if (ctx->xchg_world > 1 && ctx->xchg_connected) {
    // row shards: the dot must be summed over the ranks on the device -> a copy link of the chain (d = v) carries it
    p.d_in = v;
    p.u = nullptr;
    p.w = lb_s(L, order[0]);
    p.d_out = d;
    p.ys = 1.0;
    p.post1 = p.post2 = 1.0;
    p.mode = 2;
    p.slot = 0;
    rc = launch_qn(ctx, L->dtype, p);
  } else {
    rc = pb_dot(ctx, L->dtype, L->n, lb_s(L, order[0]), v);
  }
  if (rc != PB_OK) return rc;
  for (int k = 0; k < m; ++k) {                       // loop 1
    const int sl = order[k];
    p.d_in = k == 0 ? v : d;
    p.u = lb_y(L, sl);
    p.d_out = d;
    p.ys = L->ys[sl];
    p.mode = 0;
    p.slot = sl;
    p.post2 = 1.0;
    if (k + 1 < m) {
      p.w = lb_s(L, order[k + 1]);
      p.post1 = 1.0;
    } else {                                          // `d .*= L.H` (lbfgs.jl:69) and the first dot of loop 2
      p.w = lb_y(L, sl);
      p.post1 = L->H;
    }
    rc = launch_qn(ctx, L->dtype, p);
    if (rc != PB_OK) return rc;
  }
  for (int k = m - 1; k >= 0; --k) {                  // loop 2 visits the pairs in the opposite order, lbfgs.jl:86-95
    const int sl = order[k];
    p.d_in = d;
    p.u = lb_s(L, sl);
    p.d_out = d;
    p.ys = L->ys[sl];
    p.mode = 1;
    p.slot = sl;
    p.post1 = 1.0;
    if (k > 0) {
      p.w = lb_y(L, order[k - 1]);
      p.post2 = 1.0;
    } else {
      p.w = nullptr;
      p.post2 = scale;
      p.x = x;
      p.xd = x_d;
    }
    rc = launch_qn(ctx, L->dtype, p);
    if (rc != PB_OK) return rc;
  }
  return PB_OK;
}
