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
  } else if constexpr (PROX == PB_PROX_SCALE) {
    return (a > T(1)) ? y : mul_rn(a, y);
  } else {
    return y;
  }
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
  static __device__ __forceinline__ void run(T x, T g, T zp, T lo, T hi, T gamma, T beta, T& y, T& z, T& r, T& xn,
                                             Acc<3, 1>& acc) {
    y = sub_rn(x, mul_rn(gamma, g));
    z = prox_elem<T, PROX>(y, lo, hi);
    r = sub_rn(x, z);
    if constexpr (EXTRAP) xn = add_rn(z, mul_rn(beta, sub_rn(z, zp)));
    const double rd = (double)r, gd = (double)g;
    if constexpr (COMP) {
      if constexpr (PROX == PB_PROX_L1) dd_add(acc.s[0], fabs((double)z));
      dd_add_prod(acc.s[1], rd, rd);
      dd_add_prod(acc.s[2], gd, rd);
    } else {
      // float data: the products are exact in double; plain double accumulation per thread, double-double across threads
      if constexpr (PROX == PB_PROX_L1) acc.s[0].hi += fabs((double)z);
      acc.s[1].hi = __fma_rn(rd, rd, acc.s[1].hi);
      acc.s[2].hi = __fma_rn(gd, rd, acc.s[2].hi);
    }
    acc.m[0] = nanmax(acc.m[0], fabs(rd));
  }
};


// TMA-ring implementation (step_tma.cu).  Returns PB_EUNSUPPORTED when the configuration is not covered (caller falls
// back to the register pipeline).
int pb_launch_step_tma(pb_ctx* ctx, int dtype, int prox_kind, bool extrap, const StepParams& p);
