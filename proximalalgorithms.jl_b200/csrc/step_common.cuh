// step_common.cuh -- per-element arithmetic of the fused forward-backward step, shared by the register-pipeline kernel
// (step_kernels.cu) and the TMA bulk-copy ring kernel (step_tma.cu).
#pragma once
#include "common.cuh"

// ----------------------------------------------------------------------------------------------------------------
// element-wise prox (ProximalOperators.jl semantics, restated from the package's published algorithm)
// ----------------------------------------------------------------------------------------------------------------
template <typename T, int PROX>
__device__ __forceinline__ T prox_elem(T y, T a, T b) {
  if constexpr (PROX == PB_PROX_L1) {
    // z = y + (y <= -gl ? gl : (y >= gl ? -gl : -y)),  a = gl = gamma*lambda
    T sel = (y <= -a) ? a : ((y >= a) ? -a : -y);
    return add_rn(y, sel);
  } else if constexpr (PROX == PB_PROX_BOX) {
    return (y < a) ? a : ((y > b) ? b : y);
  } else if constexpr (PROX == PB_PROX_SCALE || PROX == PB_PROX_BALL) {
    return (a > T(1)) ? y : mul_rn(a, y);
  } else {
    return y;
  }
}

// IndBallL2 on one GPU: scale factor r / ||y|| in the element type from the (hi, lo) pair the reduction pass left in the AUX3 slot --
// the arithmetic of functions.py: IndBallL2.scale_factor (sqrt of the rounded double sum, cast, one division in R)
template <typename T>
__device__ __forceinline__ T ball_scale(const double* out, T r) {
  const double ysq = __dadd_rn(__ldcg(out + PB_S_AUX3), __ldcg(out + PB_S_AUX3 + 1));
  const T ny = (T)sqrt(ysq);
  return r / ny;
}

struct StepParams {
  const void* x;
  const void* grad;
  const void* z_prev;
  void* y;
  void* z;
  void* res;
  void* x_next;
  const void* lo_v;
  const void* hi_v;
  int64_t n;
  double gamma, beta, a, b;  // a, b: prox parameters already combined on the host in the element type
  PbWorkspace* ws;
  double* out;
  XchgParams xchg;   // fused per-iteration exchange (world == 0: off)
  int defer;         // 1: write per-CTA partials only; pb_step_fold_launch folds and exchanges on the side stream
};

template <typename T, int PROX, bool EXTRAP>
struct StepElem {
  // processes one element, updates accumulators; returns z and (optionally) writes y, res, x_next through references
  template <bool COMP>
  // PB_PROX_SQRL2 (Translate(SqrNormL2(lambda), -b)): lo = b_i (when has_b), hi = den = 1 + gamma*lambda;
  // w = (y - b)/den, z = w + b (three separately rounded operations, as csrc/dr_kernels.cu: dr_prox); GSUM = sum w^2
  static __device__ __forceinline__ void run(T x, T g, T zp, T lo, T hi, T gamma, T beta, T& y, T& z, T& r, T& xn,
                                             Acc<3, 1>& acc, bool has_b = false) {
    y = sub_rn(x, mul_rn(gamma, g));
    T w = T(0);
    if constexpr (PROX == PB_PROX_SQRL2) {
      w = has_b ? sub_rn(y, lo) / hi : y / hi;
      z = has_b ? add_rn(w, lo) : w;
    } else {
      z = prox_elem<T, PROX>(y, lo, hi);
    }
    r = sub_rn(x, z);
    if constexpr (EXTRAP) xn = add_rn(z, mul_rn(beta, sub_rn(z, zp)));
    const double rd = (double)r, gd = (double)g;
    if constexpr (COMP) {
      if constexpr (PROX == PB_PROX_L1) dd_add(acc.s[0], fabs((double)z));
      if constexpr (PROX == PB_PROX_SQRL2) dd_add_prod(acc.s[0], (double)w, (double)w);
      dd_add_prod(acc.s[1], rd, rd);
      dd_add_prod(acc.s[2], gd, rd);
    } else {
      // float data: the products are exact in double.  `acc` is then the accumulator of ONE 16-byte pack (plain double sums in
      // element order); the caller folds it into the thread's double-double accumulator with fold_pack() after each pack
      if constexpr (PROX == PB_PROX_L1) acc.s[0].hi += fabs((double)z);
      if constexpr (PROX == PB_PROX_SQRL2) acc.s[0].hi = __fma_rn((double)w, (double)w, acc.s[0].hi);
      acc.s[1].hi = __fma_rn(rd, rd, acc.s[1].hi);
      acc.s[2].hi = __fma_rn(gd, rd, acc.s[2].hi);
    }
    acc.m[0] = nanmax(acc.m[0], fabs(rd));
  }
};

// float data: add the sums of one pack (computed in element order, see StepElem::run) to the thread's double-double accumulator and
// reset the pack accumulator.  A pack is the same four elements whichever thread, CTA, unroll factor, step implementation or GPU
// processes it (row shards are 32-element aligned), and the double-double accumulation of the pack sums is exact to ~1e-32, so the
// ROUNDED totals are independent of grid size and of the number of shards -- for float data too (bench.py's `parity` fingerprint).
template <int PROX>
__device__ __forceinline__ void fold_pack(Acc<3, 1>& acc, Acc<3, 1>& pk) {
  if constexpr (PROX == PB_PROX_L1 || PROX == PB_PROX_SQRL2) dd_add(acc.s[0], pk.s[0].hi);
  dd_add(acc.s[1], pk.s[1].hi);
  dd_add(acc.s[2], pk.s[2].hi);
  acc.m[0] = nanmax(acc.m[0], pk.m[0]);
  pk.s[0].hi = pk.s[1].hi = pk.s[2].hi = 0.0;
  pk.m[0] = 0.0;
}

// TMA-ring implementation (step_tma.cu).  Returns PB_EUNSUPPORTED when the configuration is not covered (caller falls
// back to the register pipeline).
int pb_launch_step_tma(pb_ctx* ctx, int dtype, int prox_kind, bool extrap, const StepParams& p);
