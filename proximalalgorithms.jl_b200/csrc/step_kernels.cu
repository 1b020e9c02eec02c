// step_kernels.cu -- K1/K2 fused forward-backward step, K3 standalone prox, K6 vector utilities.
//
// Every kernel is a single coalesced HBM pass: 16-byte packed loads (float4 / double2), UNROLL independent packs per
// input array in flight per thread, grid sized as a multiple of the SM count, per-thread double(-double) accumulators,
// warp-shuffle -> shared -> last-CTA reduction (common.cuh).  All of them are bandwidth bound (1-3 flop/byte), so no
// tensor-core path exists for this file.
#include <string.h>

#include "step_common.cuh"

template <typename T, int PROX, bool EXTRAP, int VEC, int UNROLL, bool HINT>
__global__ void __launch_bounds__(PB_BLOCK) k_step(StepParams p) {
  constexpr bool COMP = sizeof(T) == 8;
  constexpr bool VECP = PROX == PB_PROX_BOX || PROX == PB_PROX_SQRL2;   // prox kinds with an optional per-element vector in lo_v
  constexpr int64_t TILE = (int64_t)PB_BLOCK * VEC * UNROLL;
  const T* __restrict__ x = static_cast<const T*>(p.x);
  const T* __restrict__ g = static_cast<const T*>(p.grad);
  const T* __restrict__ zp = static_cast<const T*>(p.z_prev);
  const T* __restrict__ lov = static_cast<const T*>(p.lo_v);
  const T* __restrict__ hiv = static_cast<const T*>(p.hi_v);
  T* __restrict__ yo = static_cast<T*>(p.y);
  T* __restrict__ zo = static_cast<T*>(p.z);
  T* __restrict__ ro = static_cast<T*>(p.res);
  T* __restrict__ xo = static_cast<T*>(p.x_next);
  const T gamma = (T)p.gamma, beta = (T)p.beta;
  const T pa = PROX == PB_PROX_BALL ? ball_scale<T>(p.out, (T)p.a) : (T)p.a, pb = (T)p.b;
  const int64_t n = p.n;
  const int64_t ntiles = n / TILE;

  Acc<3, 1> acc, pk;
  acc.clear();
  pk.clear();

  // one 16-byte pack: compute, store, accumulate
  auto do_pack = [&](int64_t i, const Pack<T, VEC>& xq, const Pack<T, VEC>& gq, const Pack<T, VEC>& zq) {
    Pack<T, VEC> lo, hi, yv, zn, rv, xn;
    if (VECP && lov) lo = ld_pack<T, VEC, false>(lov + i);
    if (PROX == PB_PROX_BOX && hiv) hi = ld_pack<T, VEC, false>(hiv + i);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const T l = (VECP && lov) ? lo.v[e] : pa;
      const T h = (PROX == PB_PROX_BOX && hiv) ? hi.v[e] : pb;
      StepElem<T, PROX, EXTRAP>::template run<COMP>(xq.v[e], gq.v[e], EXTRAP ? zq.v[e] : T(0), l, h, gamma, beta, yv.v[e],
                                                    zn.v[e], rv.v[e], xn.v[e], COMP ? acc : pk, lov != nullptr);
    }
    if constexpr (!COMP) fold_pack<PROX>(acc, pk);
    st_pack<T, VEC, HINT>(zo + i, zn);
    if constexpr (EXTRAP) st_pack<T, VEC, HINT>(xo + i, xn);
    if (yo) st_pack<T, VEC, HINT>(yo + i, yv);
    if (ro) st_pack<T, VEC, HINT>(ro + i, rv);
  };

  // Balanced schedule: `rounds` full rounds in which EVERY CTA streams one TILE (UNROLL packs per thread in flight), then
  // the remaining < gridDim.x tiles are shared by all CTAs at pack granularity, so no CTA works a whole tile longer than
  // the others (at 1.25e7 elements per GPU -- n = 1e8 on 8 GPUs -- that imbalance was 0.7 of 10.3 tiles).
  const int64_t rounds = ntiles / gridDim.x;
  for (int64_t r = 0; r < rounds; ++r) {
    const int64_t base = (r * gridDim.x + blockIdx.x) * TILE + (int64_t)threadIdx.x * VEC;
    Pack<T, VEC> xv[UNROLL], gv[UNROLL], zv[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int64_t i = base + (int64_t)u * PB_BLOCK * VEC;
      xv[u] = ld_pack<T, VEC, HINT>(x + i);
      gv[u] = ld_pack<T, VEC, HINT>(g + i);
      if constexpr (EXTRAP) zv[u] = ld_pack<T, VEC, HINT>(zp + i);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) do_pack(base + (int64_t)u * PB_BLOCK * VEC, xv[u], gv[u], zv[u]);
  }
  const int64_t rem_start = rounds * gridDim.x * TILE;
  const int64_t rem_packs = (n - rem_start) / VEC;
  for (int64_t q = (int64_t)blockIdx.x * PB_BLOCK + threadIdx.x; q < rem_packs; q += (int64_t)gridDim.x * PB_BLOCK) {
    const int64_t i = rem_start + q * VEC;
    Pack<T, VEC> xq = ld_pack<T, VEC, HINT>(x + i), gq = ld_pack<T, VEC, HINT>(g + i), zq;
    if constexpr (EXTRAP) zq = ld_pack<T, VEC, HINT>(zp + i);
    do_pack(i, xq, gq, zq);
  }
  // ragged tail (< VEC elements), element-wise
  for (int64_t i = rem_start + rem_packs * VEC + (int64_t)blockIdx.x * PB_BLOCK + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * PB_BLOCK) {
    const T l = (VECP && lov) ? lov[i] : pa;
    const T h = (PROX == PB_PROX_BOX && hiv) ? hiv[i] : pb;
    T yv, zn, rv, xn;
    StepElem<T, PROX, EXTRAP>::template run<COMP>(x[i], g[i], EXTRAP ? zp[i] : T(0), l, h, gamma, beta, yv, zn, rv, xn,
                                                  COMP ? acc : pk, lov != nullptr);
    if constexpr (!COMP) fold_pack<PROX>(acc, pk);
    zo[i] = zn;
    if constexpr (EXTRAP) xo[i] = xn;
    if (yo) yo[i] = yv;
    if (ro) ro[i] = rv;
  }

  OutMap map;
  map.sum_slot[0] = PB_S_GSUM;
  map.sum_slot[1] = PB_S_RESSQ;
  map.sum_slot[2] = PB_S_GDR;
  map.sum_slot[3] = -1;
  map.max_slot[0] = PB_S_RESINF;
  map.max_slot[1] = -1;
  // deferred form: freeze the scalar block as the iteration left it (stream order: everything this iteration enqueued before the
  // step has finished, nothing of the next iteration has started) -- the fold kernel runs concurrently with the next iteration
  if (p.defer && blockIdx.x == 0 && threadIdx.x < PB_NSCALARS) p.ws->snap[threadIdx.x] = __ldcg(p.out + threadIdx.x);
  grid_reduce<3, 1, PB_BLOCK>(acc, p.ws, p.out, map, &p.xchg, p.defer != 0);
}

// Deferred form of the step's epilogue: fold the partials of `nctas` CTAs, write the scalar block, exchange.  <<<1, PB_BLOCK>>>
// on the context's side stream, overlapping the NEXT step kernel on the main stream.
__global__ void __launch_bounds__(PB_BLOCK) k_step_fold(PbWorkspace* ws, unsigned int nctas, XchgParams xp) {
  double* out = ws->blk;
  if (threadIdx.x < PB_NSCALARS) out[threadIdx.x] = ws->snap[threadIdx.x];
  __syncthreads();
  Acc<3, 1> acc;
  OutMap map;
  map.sum_slot[0] = PB_S_GSUM;
  map.sum_slot[1] = PB_S_RESSQ;
  map.sum_slot[2] = PB_S_GDR;
  map.sum_slot[3] = -1;
  map.max_slot[0] = PB_S_RESINF;
  map.max_slot[1] = -1;
  fold_partials<3, 1, PB_BLOCK>(acc, ws, nctas, out, map, &xp, false);
}

// ----------------------------------------------------------------------------------------------------------------
// NormL21 step: one warp per contiguous group.  y is recomputed in the second sweep (group data is L1 resident).
// group norm: float data -> exact double sum of squares; double data -> double-double.
// ----------------------------------------------------------------------------------------------------------------
template <typename T, bool EXTRAP>
__global__ void __launch_bounds__(PB_BLOCK) k_step_l21(StepParams p, int group) {
  constexpr bool COMP = sizeof(T) == 8;
  const T* __restrict__ x = static_cast<const T*>(p.x);
  const T* __restrict__ g = static_cast<const T*>(p.grad);
  const T* __restrict__ zp = static_cast<const T*>(p.z_prev);
  T* __restrict__ yo = static_cast<T*>(p.y);
  T* __restrict__ zo = static_cast<T*>(p.z);
  T* __restrict__ ro = static_cast<T*>(p.res);
  T* __restrict__ xo = static_cast<T*>(p.x_next);
  const T gamma = (T)p.gamma, beta = (T)p.beta, gl = (T)p.a;
  const int lane = threadIdx.x & 31;
  const int64_t ngroups = p.n / group;
  const int64_t warp0 = ((int64_t)blockIdx.x * PB_BLOCK + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * PB_BLOCK) >> 5;

  Acc<3, 1> acc;
  acc.clear();
  for (int64_t gi = warp0; gi < ngroups; gi += nwarps) {
    const int64_t base = gi * group;
    dd ss;
    ss.hi = ss.lo = 0.0;
    for (int e = lane; e < group; e += 32) {
      const T yv = sub_rn(x[base + e], mul_rn(gamma, g[base + e]));
      if (COMP)
        dd_add_prod(ss, (double)yv, (double)yv);
      else
        ss.hi = __fma_rn((double)yv, (double)yv, ss.hi);
    }
    if constexpr (COMP) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        dd o;
        o.hi = __shfl_xor_sync(0xffffffffu, ss.hi, off);
        o.lo = __shfl_xor_sync(0xffffffffu, ss.lo, off);
        ss = dd_sum(ss, o);
      }
    } else {
      // float data: the lane sums are sums of exact double products; a plain double shuffle tree keeps ~1e-16 relative accuracy on a
      // value that is rounded to float right after (the double-double tree cost ~70 FP64 instructions per group and per lane: half of
      // the FP64 pipe at HBM speed)
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) ss.hi = __dadd_rn(ss.hi, __shfl_xor_sync(0xffffffffu, ss.hi, off));
    }
    // nslice = sqrt(sum) in the element type; scal = 1 - gl/nslice, clamped at 0 (NaN propagates like the package)
    const T ns = (T)sqrt(ss.hi + ss.lo);
    T scal = sub_rn(T(1), gl / ns);
    scal = (scal <= T(0)) ? T(0) : scal;
    if (lane == 0) {
      const double contrib = (double)mul_rn(scal, ns);
      if (COMP)
        dd_add(acc.s[0], contrib);
      else
        acc.s[0].hi += contrib;
    }
    for (int e = lane; e < group; e += 32) {
      const T xv = x[base + e], gv = g[base + e];
      const T yv = sub_rn(xv, mul_rn(gamma, gv));
      const T zv = mul_rn(scal, yv);
      const T rv = sub_rn(xv, zv);
      zo[base + e] = zv;
      if (yo) yo[base + e] = yv;
      if (ro) ro[base + e] = rv;
      if constexpr (EXTRAP) xo[base + e] = add_rn(zv, mul_rn(beta, sub_rn(zv, zp[base + e])));
      const double rd = (double)rv, gd = (double)gv;
      if (COMP) {
        dd_add_prod(acc.s[1], rd, rd);
        dd_add_prod(acc.s[2], gd, rd);
      } else {
        acc.s[1].hi = __fma_rn(rd, rd, acc.s[1].hi);
        acc.s[2].hi = __fma_rn(gd, rd, acc.s[2].hi);
      }
      acc.m[0] = nanmax(acc.m[0], fabs(rd));
    }
  }
  OutMap map;
  map.sum_slot[0] = PB_S_GSUM;
  map.sum_slot[1] = PB_S_RESSQ;
  map.sum_slot[2] = PB_S_GDR;
  map.sum_slot[3] = -1;
  map.max_slot[0] = PB_S_RESINF;
  map.max_slot[1] = -1;
  grid_reduce<3, 1, PB_BLOCK>(acc, p.ws, p.out, map, &p.xchg);
}

// NormL21 step, vectorised single pass: group = KP * 32 * VEC elements, one warp per group, each lane keeps its KP packs of
// x, grad (and z_prev) in registers, so every vector is read exactly once (the generic kernel above re-reads x and grad
// from L1 in its second sweep and uses scalar loads).
template <typename T, bool EXTRAP, int KP>
__global__ void __launch_bounds__(PB_BLOCK) k_step_l21_vec(StepParams p) {
  constexpr bool COMP = sizeof(T) == 8;
  constexpr int VEC = 16 / sizeof(T);
  constexpr int GROUP = KP * 32 * VEC;
  const T* __restrict__ x = static_cast<const T*>(p.x);
  const T* __restrict__ g = static_cast<const T*>(p.grad);
  const T* __restrict__ zp = static_cast<const T*>(p.z_prev);
  T* __restrict__ yo = static_cast<T*>(p.y);
  T* __restrict__ zo = static_cast<T*>(p.z);
  T* __restrict__ ro = static_cast<T*>(p.res);
  T* __restrict__ xo = static_cast<T*>(p.x_next);
  const T gamma = (T)p.gamma, beta = (T)p.beta, gl = (T)p.a;
  const int lane = threadIdx.x & 31;
  const int64_t ngroups = p.n / GROUP;
  const int64_t warp0 = ((int64_t)blockIdx.x * PB_BLOCK + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * PB_BLOCK) >> 5;
  Acc<3, 1> acc;
  acc.clear();
  // software pipeline: the loads of the warp's NEXT group are issued before the current group is reduced, so the
  // shuffle / sqrt / divide latency chain of one group overlaps the memory latency of the next
  Pack<T, VEC> nx[KP], ng[KP], nz[KP];
  auto load_group = [&](int64_t gi_) {
    const int64_t b_ = gi_ * GROUP + (int64_t)lane * VEC;
#pragma unroll
    for (int q = 0; q < KP; ++q) {
      nx[q] = ld_pack<T, VEC, true>(x + b_ + q * 32 * VEC);
      ng[q] = ld_pack<T, VEC, true>(g + b_ + q * 32 * VEC);
      if constexpr (EXTRAP) nz[q] = ld_pack<T, VEC, true>(zp + b_ + q * 32 * VEC);
    }
  };
  if (warp0 < ngroups) load_group(warp0);
  for (int64_t gi = warp0; gi < ngroups; gi += nwarps) {
    const int64_t base = gi * GROUP + (int64_t)lane * VEC;
    Pack<T, VEC> xv[KP], gv[KP], zv[KP], yv[KP];
#pragma unroll
    for (int q = 0; q < KP; ++q) {
      xv[q] = nx[q];
      gv[q] = ng[q];
      if constexpr (EXTRAP) zv[q] = nz[q];
    }
    if (gi + nwarps < ngroups) load_group(gi + nwarps);
    dd ss;
    ss.hi = ss.lo = 0.0;
#pragma unroll
    for (int q = 0; q < KP; ++q)
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        yv[q].v[e] = sub_rn(xv[q].v[e], mul_rn(gamma, gv[q].v[e]));
        if (COMP)
          dd_add_prod(ss, (double)yv[q].v[e], (double)yv[q].v[e]);
        else
          ss.hi = __fma_rn((double)yv[q].v[e], (double)yv[q].v[e], ss.hi);
      }
    if constexpr (COMP) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        dd o;
        o.hi = __shfl_xor_sync(0xffffffffu, ss.hi, off);
        o.lo = __shfl_xor_sync(0xffffffffu, ss.lo, off);
        ss = dd_sum(ss, o);
      }
    } else {
      // float data: the lane sums are sums of exact double products; a plain double shuffle tree keeps ~1e-16 relative accuracy on a
      // value that is rounded to float right after (the double-double tree cost ~70 FP64 instructions per group and per lane: half of
      // the FP64 pipe at HBM speed)
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) ss.hi = __dadd_rn(ss.hi, __shfl_xor_sync(0xffffffffu, ss.hi, off));
    }
    const T ns = (T)sqrt(ss.hi + ss.lo);
    T scal = sub_rn(T(1), gl / ns);
    scal = (scal <= T(0)) ? T(0) : scal;
    if (lane == 0) {
      const double contrib = (double)mul_rn(scal, ns);
      if (COMP)
        dd_add(acc.s[0], contrib);
      else
        acc.s[0].hi += contrib;
    }
#pragma unroll
    for (int q = 0; q < KP; ++q) {
      Pack<T, VEC> zn, rv, xn;
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        zn.v[e] = mul_rn(scal, yv[q].v[e]);
        rv.v[e] = sub_rn(xv[q].v[e], zn.v[e]);
        if constexpr (EXTRAP) xn.v[e] = add_rn(zn.v[e], mul_rn(beta, sub_rn(zn.v[e], zv[q].v[e])));
        const double rd = (double)rv.v[e], gd = (double)gv[q].v[e];
        if (COMP) {
          dd_add_prod(acc.s[1], rd, rd);
          dd_add_prod(acc.s[2], gd, rd);
        } else {
          acc.s[1].hi = __fma_rn(rd, rd, acc.s[1].hi);
          acc.s[2].hi = __fma_rn(gd, rd, acc.s[2].hi);
        }
        acc.m[0] = nanmax(acc.m[0], fabs(rd));
      }
      const int64_t i = base + q * 32 * VEC;
      st_pack<T, VEC, true>(zo + i, zn);
      if constexpr (EXTRAP) st_pack<T, VEC, true>(xo + i, xn);
      if (yo) st_pack<T, VEC, true>(yo + i, yv[q]);
      if (ro) st_pack<T, VEC, true>(ro + i, rv);
    }
  }
  OutMap map;
  map.sum_slot[0] = PB_S_GSUM;
  map.sum_slot[1] = PB_S_RESSQ;
  map.sum_slot[2] = PB_S_GDR;
  map.sum_slot[3] = -1;
  map.max_slot[0] = PB_S_RESINF;
  map.max_slot[1] = -1;
  grid_reduce<3, 1, PB_BLOCK>(acc, p.ws, p.out, map, &p.xchg);
}

// ----------------------------------------------------------------------------------------------------------------
// generic element-wise kernel with up to three inputs, one output and one sum / one max reduction (K3, K6)
// ----------------------------------------------------------------------------------------------------------------
enum {
  OP_PROX_L1 = 0,   // out = prox_l1(a; gl = s0);                    sum = |out|
  OP_PROX_BOX,      // out = clamp(a, s0 | b[], s1 | c[]);
  OP_PROX_SCALE,    // out = s0 > 1 ? a : s0*a
  OP_COPY,          // out = a
  OP_FORWARD,       // out = a - s0*b;                                 sum = out^2
  OP_EXTRAP,        // out = a + s0*(a - b)
  OP_RESIDUAL,      // out = a - b;   sums: (a-b)^2, c*(a-b) (c optional), max |a-b|  -> handled by k_residual
  OP_ADD_SCALAR,    // out = a + s0
  OP_SUB,           // out = a - b;                                    sum = out^2
  OP_NRM2SQ,        // (no out)                                        sum = a^2,  max = |a|
  OP_DOT,           // (no out)                                        sum = a*b
  OP_FORWARD_NRM    // (no out)                                        sum = (a - s0*b)^2   (phase 1 of IndBallL2 inside the step)
};

struct EwParams {
  const void* a;
  const void* b;
  const void* c;
  void* out;
  int64_t n;
  double s0, s1;
  PbWorkspace* ws;
  double* outs;
  int sum_slot, max_slot;
};

template <typename T, int OP, bool COMP>
__device__ __forceinline__ T ew_elem(T a, T b, T c, T s0, T s1, bool hb, bool hc, Acc<1, 1>& acc) {
  T o = a;
  double sum_term = 0.0;
  bool have_sum = false, have_prod = false;
  double pa = 0.0, pb = 0.0;
  if constexpr (OP == OP_PROX_L1) {
    o = prox_elem<T, PB_PROX_L1>(a, s0, s1);
    sum_term = fabs((double)o);
    have_sum = true;
  } else if constexpr (OP == OP_PROX_BOX) {
    o = prox_elem<T, PB_PROX_BOX>(a, hb ? b : s0, hc ? c : s1);
  } else if constexpr (OP == OP_PROX_SCALE) {
    o = prox_elem<T, PB_PROX_SCALE>(a, s0, s1);
  } else if constexpr (OP == OP_FORWARD) {
    o = sub_rn(a, mul_rn(s0, b));
    pa = pb = (double)o;
    have_prod = true;
  } else if constexpr (OP == OP_FORWARD_NRM) {
    o = sub_rn(a, mul_rn(s0, b));
    pa = pb = (double)o;
    have_prod = true;
  } else if constexpr (OP == OP_EXTRAP) {
    o = add_rn(a, mul_rn(s0, sub_rn(a, b)));
  } else if constexpr (OP == OP_ADD_SCALAR) {
    o = add_rn(a, s0);
  } else if constexpr (OP == OP_SUB) {
    o = sub_rn(a, b);
    pa = pb = (double)o;
    have_prod = true;
  } else if constexpr (OP == OP_NRM2SQ) {
    pa = pb = (double)a;
    have_prod = true;
    acc.m[0] = nanmax(acc.m[0], fabs((double)a));
  } else if constexpr (OP == OP_DOT) {
    pa = (double)a;
    pb = (double)b;
    have_prod = true;
  }
  if (have_sum) {
    if (COMP)
      dd_add(acc.s[0], sum_term);
    else
      acc.s[0].hi += sum_term;
  }
  if (have_prod) {
    if (COMP)
      dd_add_prod(acc.s[0], pa, pb);
    else
      acc.s[0].hi = __fma_rn(pa, pb, acc.s[0].hi);
  }
  return o;
}

template <typename T, int OP, int VEC, int UNROLL>
__global__ void __launch_bounds__(PB_BLOCK) k_ew(EwParams p) {
  constexpr bool COMP = sizeof(T) == 8;
  constexpr bool USES_B = (OP == OP_FORWARD || OP == OP_EXTRAP || OP == OP_SUB || OP == OP_DOT || OP == OP_FORWARD_NRM);
  constexpr bool HAS_OUT = !(OP == OP_NRM2SQ || OP == OP_DOT || OP == OP_FORWARD_NRM);
  constexpr int64_t TILE = (int64_t)PB_BLOCK * VEC * UNROLL;
  const T* __restrict__ a = static_cast<const T*>(p.a);
  const T* __restrict__ b = static_cast<const T*>(p.b);
  const T* __restrict__ c = static_cast<const T*>(p.c);
  T* __restrict__ out = static_cast<T*>(p.out);
  const bool hb = (OP == OP_PROX_BOX) && b != nullptr;
  const bool hc = (OP == OP_PROX_BOX) && c != nullptr;
  const T s0 = (T)p.s0, s1 = (T)p.s1;
  const int64_t n = p.n, ntiles = n / TILE;
  Acc<1, 1> acc;
  acc.clear();
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t base = tile * TILE + (int64_t)threadIdx.x * VEC;
    Pack<T, VEC> av[UNROLL], bv[UNROLL], cv[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int64_t i = base + (int64_t)u * PB_BLOCK * VEC;
      av[u] = ld_pack<T, VEC, false>(a + i);
      if (USES_B || hb) bv[u] = ld_pack<T, VEC, false>(b + i);
      if (hc) cv[u] = ld_pack<T, VEC, false>(c + i);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int64_t i = base + (int64_t)u * PB_BLOCK * VEC;
      Pack<T, VEC> o;
#pragma unroll
      for (int e = 0; e < VEC; ++e)
        o.v[e] = ew_elem<T, OP, COMP>(av[u].v[e], (USES_B || hb) ? bv[u].v[e] : T(0), hc ? cv[u].v[e] : T(0), s0, s1, hb,
                                      hc, acc);
      if constexpr (HAS_OUT) st_pack<T, VEC, false>(out + i, o);
    }
  }
  for (int64_t i = ntiles * TILE + (int64_t)blockIdx.x * PB_BLOCK + threadIdx.x; i < n; i += (int64_t)gridDim.x * PB_BLOCK) {
    const T o = ew_elem<T, OP, COMP>(a[i], (USES_B || hb) ? b[i] : T(0), hc ? c[i] : T(0), s0, s1, hb, hc, acc);
    if constexpr (HAS_OUT) out[i] = o;
  }
  if (p.sum_slot >= 0 || p.max_slot >= 0) {
    OutMap map;
    map.sum_slot[0] = p.sum_slot;
    map.sum_slot[1] = map.sum_slot[2] = map.sum_slot[3] = -1;
    map.max_slot[0] = p.max_slot;
    map.max_slot[1] = -1;
    grid_reduce<1, 1, PB_BLOCK>(acc, p.ws, p.outs, map);
  }
}

// res = x - z with RESSQ, RESINF and optionally GDR
template <typename T>
__global__ void __launch_bounds__(PB_BLOCK) k_residual(const T* __restrict__ x, const T* __restrict__ z,
                                                       const T* __restrict__ g, T* __restrict__ res, int64_t n,
                                                       PbWorkspace* ws, double* outs) {
  constexpr bool COMP = sizeof(T) == 8;
  Acc<2, 1> acc;
  acc.clear();
  for (int64_t i = (int64_t)blockIdx.x * PB_BLOCK + threadIdx.x; i < n; i += (int64_t)gridDim.x * PB_BLOCK) {
    const T r = sub_rn(x[i], z[i]);
    if (res) res[i] = r;
    const double rd = (double)r;
    if (COMP) {
      dd_add_prod(acc.s[0], rd, rd);
      if (g) dd_add_prod(acc.s[1], (double)g[i], rd);
    } else {
      acc.s[0].hi = __fma_rn(rd, rd, acc.s[0].hi);
      if (g) acc.s[1].hi = __fma_rn((double)g[i], rd, acc.s[1].hi);
    }
    acc.m[0] = nanmax(acc.m[0], fabs(rd));
  }
  OutMap map;
  map.sum_slot[0] = PB_S_RESSQ;
  map.sum_slot[1] = g ? PB_S_GDR : -1;
  map.sum_slot[2] = map.sum_slot[3] = -1;
  map.max_slot[0] = PB_S_RESINF;
  map.max_slot[1] = -1;
  grid_reduce<2, 1, PB_BLOCK>(acc, ws, outs, map);
}

// NormL21 standalone prox, one warp per group
template <typename T>
__global__ void __launch_bounds__(PB_BLOCK) k_prox_l21(const T* __restrict__ y, T* __restrict__ z, int64_t n, int group,
                                                       double gl_d, PbWorkspace* ws, double* outs) {
  constexpr bool COMP = sizeof(T) == 8;
  const T gl = (T)gl_d;
  const int lane = threadIdx.x & 31;
  const int64_t ngroups = n / group;
  const int64_t warp0 = ((int64_t)blockIdx.x * PB_BLOCK + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * PB_BLOCK) >> 5;
  Acc<1, 1> acc;
  acc.clear();
  for (int64_t gi = warp0; gi < ngroups; gi += nwarps) {
    const int64_t base = gi * group;
    dd ss;
    ss.hi = ss.lo = 0.0;
    for (int e = lane; e < group; e += 32) {
      const double yv = (double)y[base + e];
      if (COMP)
        dd_add_prod(ss, yv, yv);
      else
        ss.hi = __fma_rn(yv, yv, ss.hi);
    }
    if constexpr (COMP) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        dd o;
        o.hi = __shfl_xor_sync(0xffffffffu, ss.hi, off);
        o.lo = __shfl_xor_sync(0xffffffffu, ss.lo, off);
        ss = dd_sum(ss, o);
      }
    } else {
      // float data: the lane sums are sums of exact double products; a plain double shuffle tree keeps ~1e-16 relative accuracy on a
      // value that is rounded to float right after (the double-double tree cost ~70 FP64 instructions per group and per lane: half of
      // the FP64 pipe at HBM speed)
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) ss.hi = __dadd_rn(ss.hi, __shfl_xor_sync(0xffffffffu, ss.hi, off));
    }
    const T ns = (T)sqrt(ss.hi + ss.lo);
    T scal = sub_rn(T(1), gl / ns);
    scal = (scal <= T(0)) ? T(0) : scal;
    if (lane == 0) {
      const double contrib = (double)mul_rn(scal, ns);
      if (COMP)
        dd_add(acc.s[0], contrib);
      else
        acc.s[0].hi += contrib;
    }
    for (int e = lane; e < group; e += 32) z[base + e] = mul_rn(scal, y[base + e]);
  }
  OutMap map;
  map.sum_slot[0] = PB_S_GSUM;
  map.sum_slot[1] = map.sum_slot[2] = map.sum_slot[3] = -1;
  map.max_slot[0] = map.max_slot[1] = -1;
  grid_reduce<1, 1, PB_BLOCK>(acc, ws, outs, map);
}

// ----------------------------------------------------------------------------------------------------------------
// host-side dispatch
// ----------------------------------------------------------------------------------------------------------------
static bool use_hints(const pb_ctx* ctx, int64_t n, size_t elt, int nvec) {
  if (ctx->stream_hints >= 0) return ctx->stream_hints != 0;
  return (size_t)n * elt * nvec > ctx->l2_bytes;  // working set cannot live in L2: stream through it
}

// grid = SMs x CTAs-per-SM, never more CTAs than are co-resident: a grid-stride kernel with a partial second wave would
// serialise (measured: 592 CTAs at 3 resident/SM ran 4 % slower than 296).
template <typename K>
static int resident_grid(pb_ctx* ctx, K kern, int* occ_cache, int64_t work_per_cta, int64_t n, int default_per_sm) {
  if (*occ_cache == 0) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, PB_BLOCK, 0) != cudaSuccess || occ < 1) occ = 1;
    *occ_cache = occ;
  }
  int per_sm = ctx->ctas_per_sm > 0 ? ctx->ctas_per_sm : default_per_sm;
  if (per_sm > *occ_cache) per_sm = *occ_cache;
  int64_t g = (int64_t)ctx->sm_count * per_sm;
  int64_t need = (n + work_per_cta - 1) / work_per_cta;
  if (need < 1) need = 1;
  if (g > need) g = need;
  if (g > PB_MAX_CTAS) g = PB_MAX_CTAS;
  return (int)g;
}

template <typename T, int PROX, bool EXTRAP, int UNROLL>
static int launch_step_u(pb_ctx* ctx, const StepParams& p, bool vec_ok, bool hint) {
  constexpr int VEC = 16 / sizeof(T);
  // occupancy is queried on the context's device: one cache row per device ordinal
  static int occ_dev[PB_MAX_DEVICES][3] = {};
  int* occ = occ_dev[ctx->device < PB_MAX_DEVICES ? ctx->device : 0];
  if (vec_ok) {
    if (hint) {
      auto kern = k_step<T, PROX, EXTRAP, VEC, UNROLL, true>;
      const int grid = resident_grid(ctx, kern, &occ[0], (int64_t)PB_BLOCK * VEC * UNROLL, p.n, 2);
      ctx->last_step_grid = grid;
      kern<<<grid, PB_BLOCK, 0, ctx->stream>>>(p);
    } else {
      auto kern = k_step<T, PROX, EXTRAP, VEC, UNROLL, false>;
      const int grid = resident_grid(ctx, kern, &occ[1], (int64_t)PB_BLOCK * VEC * UNROLL, p.n, 2);
      ctx->last_step_grid = grid;
      kern<<<grid, PB_BLOCK, 0, ctx->stream>>>(p);
    }
  } else {
    auto kern = k_step<T, PROX, EXTRAP, 1, UNROLL, false>;
    const int grid = resident_grid(ctx, kern, &occ[2], (int64_t)PB_BLOCK * UNROLL, p.n, 4);
    ctx->last_step_grid = grid;
    kern<<<grid, PB_BLOCK, 0, ctx->stream>>>(p);
  }
  PB_LAUNCH_CHECK(ctx);
  return PB_OK;
}

template <typename T, int PROX, bool EXTRAP>
static int launch_step_t(pb_ctx* ctx, const StepParams& p, bool vec_ok) {
  // defaults from the interleaved A/B on B200 (profiles/r01_tune_step.md): L1 -> 4 packs in flight with streaming hints,
  // box -> 2 packs without hints; PB_OPT_UNROLL / PB_OPT_STREAM_HINTS override.
  bool hint = use_hints(ctx, p.n, sizeof(T), EXTRAP ? 5 : 3);
  int unroll = ctx->unroll;
  if constexpr (PROX == PB_PROX_BOX) {
    if (ctx->stream_hints < 0) hint = false;
    if (unroll == 0) unroll = 2;
  }
  // the unroll sweep exists for the two headline prox kinds only (keeps the binary small)
  if constexpr (PROX == PB_PROX_L1 || PROX == PB_PROX_BOX) {
    switch (unroll) {
      case 1: return launch_step_u<T, PROX, EXTRAP, 1>(ctx, p, vec_ok, hint);
      case 2: return launch_step_u<T, PROX, EXTRAP, 2>(ctx, p, vec_ok, hint);
      case 8: return launch_step_u<T, PROX, EXTRAP, 8>(ctx, p, vec_ok, hint);
      default: break;
    }
  }
  return launch_step_u<T, PROX, EXTRAP, 4>(ctx, p, vec_ok, hint);
}

template <typename T, int PROX, bool EXTRAP>
static int launch_step_t(pb_ctx* ctx, const StepParams& p, bool vec_ok);

// Split form (pb_ctx::defer_fold, set by the pipelined driver loop): the step kernel leaves its per-CTA partials in one of two
// alternating workspaces; k_step_fold on the side stream folds them, writes the scalar block and does the exchange while the
// main stream is free to run the next step.  Stream dependencies: fold k waits for step k (event), step k+2 waits for fold k
// (it reuses the workspace).  The exchange sequence numbers are issued in step order, so the host protocol is unchanged.
template <typename T, bool EXTRAP>
static int launch_step_deferred(pb_ctx* ctx, StepParams p, const pb_prox* g, bool vec_ok) {
  const int par = (int)(ctx->defer_count & 1u);
  if (ctx->defer_count >= 2) PB_CHECK_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_fold[par], 0));
  p.ws = ctx->ws_defer[par];
  p.defer = 1;
  p.xchg.world = 0;
  int rc;
  switch (g->kind) {
    case PB_PROX_ZERO: rc = launch_step_t<T, PB_PROX_ZERO, EXTRAP>(ctx, p, vec_ok); break;
    case PB_PROX_L1: rc = launch_step_t<T, PB_PROX_L1, EXTRAP>(ctx, p, vec_ok); break;
    case PB_PROX_BOX: rc = launch_step_t<T, PB_PROX_BOX, EXTRAP>(ctx, p, vec_ok); break;
    case PB_PROX_SQRL2: rc = launch_step_t<T, PB_PROX_SQRL2, EXTRAP>(ctx, p, vec_ok); break;
    case PB_PROX_BALL: rc = launch_step_t<T, PB_PROX_BALL, EXTRAP>(ctx, p, vec_ok); break;
    default: rc = launch_step_t<T, PB_PROX_SCALE, EXTRAP>(ctx, p, vec_ok); break;
  }
  if (rc != PB_OK) return rc;
  PB_CHECK_CUDA(cudaEventRecord(ctx->ev_main[par], ctx->stream));
  PB_CHECK_CUDA(cudaStreamWaitEvent(ctx->side_stream, ctx->ev_main[par], 0));
  XchgParams xp;
  pb_xchg_next(ctx, &xp, true);
  k_step_fold<<<1, PB_BLOCK, 0, ctx->side_stream>>>(ctx->ws_defer[par], (unsigned int)ctx->last_step_grid, xp);
  PB_LAUNCH_CHECK(ctx);
  PB_CHECK_CUDA(cudaEventRecord(ctx->ev_fold[par], ctx->side_stream));
  ctx->defer_count += 1;
  return PB_OK;
}

template <typename T, bool EXTRAP>
static int launch_step_prox(pb_ctx* ctx, StepParams p, const pb_prox* g, bool vec_ok) {
  const T gamma = (T)p.gamma;
  // prox parameters in the element type
  switch (g->kind) {
    case PB_PROX_ZERO:
      break;
    case PB_PROX_L1:
    case PB_PROX_L21:
      p.a = (double)mul_rn_host(gamma, (T)g->p0);   // gl = gamma*lambda, one rounding in R like the package
      break;
    case PB_PROX_BOX:
      p.a = g->p0;
      p.b = g->p1;
      p.lo_v = g->v0;
      p.hi_v = g->v1;
      if ((p.lo_v && !pb_aligned16(p.lo_v)) || (p.hi_v && !pb_aligned16(p.hi_v))) vec_ok = false;
      break;
    case PB_PROX_SCALE:
    case PB_PROX_BALL:             // p.a = r; the kernel forms r / ||y|| from the AUX3 slot (step_common launched the norm pass)
      p.a = g->p0;
      break;
    case PB_PROX_SQRL2: {           // den = 1 + gamma*lambda in the element type (two roundings, as pb_dr_step / pb_prox_apply)
      const T gl = mul_rn_host(gamma, (T)g->p0);
      volatile T den = T(1) + gl;
      p.b = (double)den;
      p.lo_v = g->v0;
      if (p.lo_v && !pb_aligned16(p.lo_v)) vec_ok = false;
      break;
    }
    default:
      pb_set_error("unknown prox kind %d", g->kind);
      return PB_EINVAL;
  }
  if (g->kind == PB_PROX_L21) {
    PB_REQUIRE(g->group > 0 && p.n % g->group == 0, "NormL21 group must divide n");
    const int64_t ngroups = p.n / g->group;
    const int grid = pb_stream_grid(ctx, PB_BLOCK / 32, ngroups, 8);
    constexpr int VEC = 16 / sizeof(T);
    const int kp = (g->group % (32 * VEC) == 0) ? g->group / (32 * VEC) : 0;
    if (vec_ok && kp == 1)
      k_step_l21_vec<T, EXTRAP, 1><<<grid, PB_BLOCK, 0, ctx->stream>>>(p);
    else if (vec_ok && kp == 2)
      k_step_l21_vec<T, EXTRAP, 2><<<grid, PB_BLOCK, 0, ctx->stream>>>(p);
    else if (vec_ok && kp == 4)
      k_step_l21_vec<T, EXTRAP, 4><<<grid, PB_BLOCK, 0, ctx->stream>>>(p);
    else
      k_step_l21<T, EXTRAP><<<grid, PB_BLOCK, 0, ctx->stream>>>(p, g->group);
    PB_LAUNCH_CHECK(ctx);
    return PB_OK;
  }
  // measured default: K2 (5 streams) -> register pipeline, K1 (3 streams) -> TMA bulk-copy ring
  const int impl = ctx->step_impl != 0 ? ctx->step_impl : (EXTRAP ? 1 : 2);
  if (p.defer && impl == 1) return launch_step_deferred<T, EXTRAP>(ctx, p, g, vec_ok);
  p.defer = 0;
  if (impl == 2 && vec_ok && p.n >= (int64_t)1 << 16) {
    const int rc = pb_launch_step_tma(ctx, sizeof(T) == 4 ? PB_F32 : PB_F64, g->kind, EXTRAP, p);
    if (rc != PB_EUNSUPPORTED) return rc;
  }
  switch (g->kind) {
    case PB_PROX_ZERO:
      return launch_step_t<T, PB_PROX_ZERO, EXTRAP>(ctx, p, vec_ok);
    case PB_PROX_L1:
      return launch_step_t<T, PB_PROX_L1, EXTRAP>(ctx, p, vec_ok);
    case PB_PROX_BOX:
      return launch_step_t<T, PB_PROX_BOX, EXTRAP>(ctx, p, vec_ok);
    case PB_PROX_SQRL2:
      return launch_step_t<T, PB_PROX_SQRL2, EXTRAP>(ctx, p, vec_ok);
    case PB_PROX_BALL:
      return launch_step_t<T, PB_PROX_BALL, EXTRAP>(ctx, p, vec_ok);
    default:
      return launch_step_t<T, PB_PROX_SCALE, EXTRAP>(ctx, p, vec_ok);
  }
}

template <int OP>
static int launch_ew(pb_ctx* ctx, int dtype, int64_t n, const void* a, const void* b, const void* c, void* out, double s0, double s1,
                     int sum_slot, int max_slot);

static int step_common(pb_ctx* ctx, int dtype, int64_t n, const void* x, const void* grad, const void* z_prev,
                       double gamma, double beta, const pb_prox* g, void* y, void* z, void* res, void* x_next,
                       bool extrap) {
  PB_REQUIRE(ctx != nullptr, "null context");
  PB_REQUIRE(dtype == PB_F32 || dtype == PB_F64, "dtype must be PB_F32 or PB_F64");
  PB_REQUIRE(n >= 0, "negative length");
  PB_REQUIRE(g != nullptr, "null prox descriptor");
  PB_REQUIRE(n == 0 || (x && grad && z), "null vector");
  PB_REQUIRE(!extrap || n == 0 || (z_prev && x_next), "null z_prev / x_next");
  PB_REQUIRE(!extrap || n == 0 || x_next != x, "x_next must not alias x");
  if (g->kind == PB_PROX_BALL && n > 0) {
    // IndBallL2, phase 1: ||x - gamma*grad||^2 -> AUX3 (no output vector; phase 2 recomputes y with the same two roundings)
    if (ctx->xchg_world > 1) {
      pb_set_error("PB_PROX_BALL needs the norm of y combined across ranks: use pb_forward + exchange + PB_PROX_SCALE on row shards");
      return PB_EUNSUPPORTED;
    }
    const int rc1 = launch_ew<OP_FORWARD_NRM>(ctx, dtype, n, x, grad, nullptr, nullptr, dtype == PB_F32 ? (double)(float)gamma : gamma, 0, PB_S_AUX3, -1);
    if (rc1 != PB_OK) return rc1;
  }
  StepParams p;
  p.x = x;
  p.grad = grad;
  p.z_prev = z_prev;
  p.y = y;
  p.z = z;
  p.res = res;
  p.x_next = x_next;
  p.lo_v = p.hi_v = nullptr;
  p.n = n;
  p.gamma = gamma;
  p.beta = beta;
  p.a = p.b = 0.0;
  p.ws = ctx->ws;
  p.out = ctx->scalars_dev;
  // deferred fold: only the register-pipeline kernel of the single-pass prox kinds has the split form
  const int impl_ = ctx->step_impl != 0 ? ctx->step_impl : (extrap ? 1 : 2);
  p.defer = (ctx->defer_fold && ctx->xchg_fused && n > 0 && impl_ == 1 && g->kind != PB_PROX_L21) ? 1 : 0;
  if (p.defer)
    memset(&p.xchg, 0, sizeof(p.xchg));
  else
    pb_xchg_next(ctx, &p.xchg, ctx->xchg_fused != 0 && n > 0);
  bool vec_ok = pb_aligned16(x) && pb_aligned16(grad) && pb_aligned16(z) && (!y || pb_aligned16(y)) &&
                (!res || pb_aligned16(res)) && (!extrap || (pb_aligned16(z_prev) && pb_aligned16(x_next)));
  if (dtype == PB_F32)
    return extrap ? launch_step_prox<float, true>(ctx, p, g, vec_ok) : launch_step_prox<float, false>(ctx, p, g, vec_ok);
  return extrap ? launch_step_prox<double, true>(ctx, p, g, vec_ok) : launch_step_prox<double, false>(ctx, p, g, vec_ok);
}

// Switch the split (deferred-fold) form of the fused step on / off (internal: used by the pipelined loop of solve.cu).  Turning it
// off makes the main stream wait for the last fold, so the scalar block and the workspaces are quiescent afterwards.
// Side stream, the two alternating workspaces and the chaining events of the split form; all or nothing.
static int defer_setup(pb_ctx* ctx) {
  cudaStream_t side = nullptr;
  PbWorkspace* ws[2] = {nullptr, nullptr};
  cudaEvent_t em[2] = {nullptr, nullptr}, ef[2] = {nullptr, nullptr};
  cudaError_t e = cudaSetDevice(ctx->device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking);
  for (int k = 0; k < 2 && e == cudaSuccess; ++k) {
    e = cudaMalloc((void**)&ws[k], sizeof(PbWorkspace));
    if (e == cudaSuccess) e = cudaMemset(ws[k], 0, sizeof(PbWorkspace));
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&em[k], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ef[k], cudaEventDisableTiming);
  }
  if (e != cudaSuccess) {
    for (int k = 0; k < 2; ++k) {
      if (ws[k]) cudaFree(ws[k]);
      if (em[k]) cudaEventDestroy(em[k]);
      if (ef[k]) cudaEventDestroy(ef[k]);
    }
    if (side) cudaStreamDestroy(side);
    pb_set_error("pb_step_defer: %s", cudaGetErrorString(e));
    return e == cudaErrorMemoryAllocation ? PB_ENOMEM : PB_ECUDA;
  }
  for (int k = 0; k < 2; ++k) {
    ctx->ws_defer[k] = ws[k];
    ctx->ev_main[k] = em[k];
    ctx->ev_fold[k] = ef[k];
  }
  ctx->side_stream = side;
  return PB_OK;
}

int pb_step_defer(pb_ctx* ctx, int on) {
  if (on) {
    if (!ctx->side_stream) {
      const int rc = defer_setup(ctx);
      if (rc != PB_OK) return rc;
    }
    ctx->defer_count = 0;
    ctx->defer_fold = 1;
    return PB_OK;
  }
  if (ctx->defer_fold) {
    ctx->defer_fold = 0;
    for (int k = 0; k < 2 && k < (int)ctx->defer_count; ++k) PB_CHECK_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_fold[k], 0));
  }
  return PB_OK;
}

extern "C" int pb_fb_step(pb_ctx* ctx, int dtype, int64_t n, const void* x, const void* grad, double gamma,
                          const pb_prox* g, void* y, void* z, void* res) {
  return step_common(ctx, dtype, n, x, grad, nullptr, gamma, 0.0, g, y, z, res, nullptr, false);
}

extern "C" int pb_ffb_step(pb_ctx* ctx, int dtype, int64_t n, const void* x, const void* grad, const void* z_prev,
                           double gamma, double beta, const pb_prox* g, void* y, void* z, void* res, void* x_next) {
  return step_common(ctx, dtype, n, x, grad, z_prev, gamma, beta, g, y, z, res, x_next, true);
}

// ---- element-wise utilities ------------------------------------------------------------------------------------
template <typename T, int OP>
static int launch_ew_t(pb_ctx* ctx, const EwParams& p) {
  constexpr int VEC = 16 / sizeof(T);
  constexpr int UNROLL = 4;
  const bool vec_ok = pb_aligned16(p.a) && (!p.b || pb_aligned16(p.b)) && (!p.c || pb_aligned16(p.c)) &&
                      (!p.out || pb_aligned16(p.out));
  if (vec_ok) {
    const int grid = pb_stream_grid(ctx, (int64_t)PB_BLOCK * VEC * UNROLL, p.n, 4);
    k_ew<T, OP, VEC, UNROLL><<<grid, PB_BLOCK, 0, ctx->stream>>>(p);
  } else {
    const int grid = pb_stream_grid(ctx, (int64_t)PB_BLOCK * UNROLL, p.n, 4);
    k_ew<T, OP, 1, UNROLL><<<grid, PB_BLOCK, 0, ctx->stream>>>(p);
  }
  PB_LAUNCH_CHECK(ctx);
  return PB_OK;
}

template <int OP>
static int launch_ew(pb_ctx* ctx, int dtype, int64_t n, const void* a, const void* b, const void* c, void* out,
                     double s0, double s1, int sum_slot, int max_slot) {
  PB_REQUIRE(ctx != nullptr, "null context");
  PB_REQUIRE(dtype == PB_F32 || dtype == PB_F64, "dtype must be PB_F32 or PB_F64");
  PB_REQUIRE(n >= 0, "negative length");
  PB_REQUIRE(n == 0 || a != nullptr, "null vector");
  EwParams p;
  p.a = a;
  p.b = b;
  p.c = c;
  p.out = out;
  p.n = n;
  p.s0 = s0;
  p.s1 = s1;
  p.ws = ctx->ws;
  p.outs = ctx->scalars_dev;
  p.sum_slot = sum_slot;
  p.max_slot = max_slot;
  return dtype == PB_F32 ? launch_ew_t<float, OP>(ctx, p) : launch_ew_t<double, OP>(ctx, p);
}

int pb_prox_sqrl2_apply(pb_ctx* ctx, int dtype, int64_t n, const void* y, double gamma, const pb_prox* g, void* z);  // dr_kernels.cu

extern "C" int pb_prox_apply(pb_ctx* ctx, int dtype, int64_t n, const void* y, double gamma, const pb_prox* g, void* z) {
  PB_REQUIRE(g != nullptr, "null prox descriptor");
  PB_REQUIRE(n == 0 || z != nullptr, "null output");
  switch (g->kind) {
    case PB_PROX_SQRL2:
      return pb_prox_sqrl2_apply(ctx, dtype, n, y, gamma, g, z);
    case PB_PROX_ZERO:
      return launch_ew<OP_COPY>(ctx, dtype, n, y, nullptr, nullptr, z, 0, 0, -1, -1);
    case PB_PROX_L1: {
      const double gl = dtype == PB_F32 ? (double)mul_rn_host((float)gamma, (float)g->p0) : mul_rn_host(gamma, g->p0);
      return launch_ew<OP_PROX_L1>(ctx, dtype, n, y, nullptr, nullptr, z, gl, 0, PB_S_GSUM, -1);
    }
    case PB_PROX_BOX:
      return launch_ew<OP_PROX_BOX>(ctx, dtype, n, y, g->v0, g->v1, z, g->p0, g->p1, -1, -1);
    case PB_PROX_SCALE:
      return launch_ew<OP_PROX_SCALE>(ctx, dtype, n, y, nullptr, nullptr, z, g->p0, 0, -1, -1);
    case PB_PROX_L21: {
      PB_REQUIRE(ctx != nullptr, "null context");
      PB_REQUIRE(dtype == PB_F32 || dtype == PB_F64, "dtype must be PB_F32 or PB_F64");
      PB_REQUIRE(g->group > 0 && n % g->group == 0, "NormL21 group must divide n");
      const int grid = pb_stream_grid(ctx, PB_BLOCK / 32, n / g->group, 8);
      if (dtype == PB_F32)
        k_prox_l21<float><<<grid, PB_BLOCK, 0, ctx->stream>>>((const float*)y, (float*)z, n, g->group,
                                                               (double)mul_rn_host((float)gamma, (float)g->p0), ctx->ws,
                                                               ctx->scalars_dev);
      else
        k_prox_l21<double><<<grid, PB_BLOCK, 0, ctx->stream>>>((const double*)y, (double*)z, n, g->group,
                                                                mul_rn_host(gamma, g->p0), ctx->ws, ctx->scalars_dev);
      PB_LAUNCH_CHECK(ctx);
      return PB_OK;
    }
    default:
      pb_set_error("unknown prox kind %d", g->kind);
      return PB_EINVAL;
  }
}

extern "C" int pb_forward(pb_ctx* ctx, int dtype, int64_t n, const void* x, const void* grad, double gamma, void* y) {
  PB_REQUIRE(n == 0 || (grad && y), "null vector");
  return launch_ew<OP_FORWARD>(ctx, dtype, n, x, grad, nullptr, y, gamma, 0, PB_S_AUX, -1);
}

extern "C" int pb_extrapolate(pb_ctx* ctx, int dtype, int64_t n, const void* z, const void* z_prev, double beta, void* x) {
  PB_REQUIRE(n == 0 || (z_prev && x), "null vector");
  return launch_ew<OP_EXTRAP>(ctx, dtype, n, z, z_prev, nullptr, x, beta, 0, -1, -1);
}

extern "C" int pb_add_scalar(pb_ctx* ctx, int dtype, int64_t n, const void* x, double c, void* out) {
  PB_REQUIRE(n == 0 || out, "null vector");
  return launch_ew<OP_ADD_SCALAR>(ctx, dtype, n, x, nullptr, nullptr, out, c, 0, -1, -1);
}

extern "C" int pb_sub(pb_ctx* ctx, int dtype, int64_t n, const void* a, const void* b, void* out) {
  PB_REQUIRE(n == 0 || (b && out), "null vector");
  return launch_ew<OP_SUB>(ctx, dtype, n, a, b, nullptr, out, 0, 0, PB_S_AUX, -1);
}

extern "C" int pb_nrm2sq(pb_ctx* ctx, int dtype, int64_t n, const void* v) {
  return launch_ew<OP_NRM2SQ>(ctx, dtype, n, v, nullptr, nullptr, nullptr, 0, 0, PB_S_AUX, PB_S_AUXINF);
}

extern "C" int pb_dot(pb_ctx* ctx, int dtype, int64_t n, const void* a, const void* b) {
  PB_REQUIRE(n == 0 || b, "null vector");
  return launch_ew<OP_DOT>(ctx, dtype, n, a, b, nullptr, nullptr, 0, 0, PB_S_AUX, -1);
}

extern "C" int pb_residual(pb_ctx* ctx, int dtype, int64_t n, const void* x, const void* z, const void* grad, void* res) {
  PB_REQUIRE(ctx != nullptr, "null context");
  PB_REQUIRE(dtype == PB_F32 || dtype == PB_F64, "dtype must be PB_F32 or PB_F64");
  PB_REQUIRE(n >= 0, "negative length");
  PB_REQUIRE(n == 0 || (x && z), "null vector");
  const int grid = pb_stream_grid(ctx, PB_BLOCK * 4, n, 4);
  if (dtype == PB_F32)
    k_residual<float><<<grid, PB_BLOCK, 0, ctx->stream>>>((const float*)x, (const float*)z, (const float*)grad,
                                                           (float*)res, n, ctx->ws, ctx->scalars_dev);
  else
    k_residual<double><<<grid, PB_BLOCK, 0, ctx->stream>>>((const double*)x, (const double*)z, (const double*)grad,
                                                            (double*)res, n, ctx->ws, ctx->scalars_dev);
  PB_LAUNCH_CHECK(ctx);
  return PB_OK;
}

// ---- host-buffer entry point -------------------------------------------------------------------------------------
extern "C" int pb_ffb_step_host(pb_ctx* ctx, int dtype, int64_t n, const void* x, const void* grad, const void* z_prev,
                                double gamma, double beta, const pb_prox* g, void* z, void* x_next, double* scalars) {
  PB_REQUIRE(ctx != nullptr, "null context");
  PB_REQUIRE(dtype == PB_F32 || dtype == PB_F64, "dtype must be PB_F32 or PB_F64");
  PB_REQUIRE(n >= 0, "negative length");
  PB_REQUIRE(n == 0 || (x && grad && z_prev && z && x_next), "null host vector");
  const size_t bytes = (size_t)n * (dtype == PB_F32 ? 4 : 8);
  if (ctx->hbuf_bytes < bytes) {
    for (int k = 0; k < 5; ++k) {
      if (ctx->hbuf[k]) cudaFree(ctx->hbuf[k]);
      ctx->hbuf[k] = nullptr;
    }
    ctx->hbuf_bytes = 0;
    for (int k = 0; k < 5; ++k) PB_CHECK_CUDA(cudaMalloc(&ctx->hbuf[k], bytes ? bytes : 16));
    ctx->hbuf_bytes = bytes;
  }
  PB_CHECK_CUDA(cudaMemcpyAsync(ctx->hbuf[0], x, bytes, cudaMemcpyHostToDevice, ctx->stream));
  PB_CHECK_CUDA(cudaMemcpyAsync(ctx->hbuf[1], grad, bytes, cudaMemcpyHostToDevice, ctx->stream));
  PB_CHECK_CUDA(cudaMemcpyAsync(ctx->hbuf[2], z_prev, bytes, cudaMemcpyHostToDevice, ctx->stream));
  int rc = pb_ffb_step(ctx, dtype, n, ctx->hbuf[0], ctx->hbuf[1], ctx->hbuf[2], gamma, beta, g, nullptr, ctx->hbuf[3],
                       nullptr, ctx->hbuf[4]);
  if (rc != PB_OK) return rc;
  PB_CHECK_CUDA(cudaMemcpyAsync(z, ctx->hbuf[3], bytes, cudaMemcpyDeviceToHost, ctx->stream));
  PB_CHECK_CUDA(cudaMemcpyAsync(x_next, ctx->hbuf[4], bytes, cudaMemcpyDeviceToHost, ctx->stream));
  if (scalars) return pb_read_scalars(ctx, scalars);
  PB_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
  return PB_OK;
}
