// solve_scalar.h -- the scalar side of the forward-backward iterations, in R = real(eltype(x0)), shared by the host driver
// loop (solve.cu) and the on-device driver loop (persist.cu) so that both take bit-identical decisions.
//   Nesterov<R>        src/accel/nesterov.jl:14-17, :36, :51-54, :89-103
//   pb_f_model         src/utilities/fb_tools.jl:3-5   (norm(res)^2 is sqrt-then-square in R)
//   pb_sq_half         benchmark/benchmarks.jl:16      (norm(r)^2 / 2)
// Every operation must round separately (no FMA contraction): host code is compiled with -ffp-contract=off, the device
// translation unit that includes this header with -fmad=false (build.py).
#pragma once
#include <math.h>

#include "proxb200.h"

#if defined(__CUDACC__)
#define PB_HD __host__ __device__ __forceinline__
#else
#define PB_HD inline
#endif

// square root IN R: sqrtf for float, sqrt for double (both correctly rounded on host and device)
PB_HD float pb_rs(float v) { return sqrtf(v); }
PB_HD double pb_rs(double v) { return sqrt(v); }

template <typename R>
PB_HD R pb_sq_half(double sum_sq) {
  const R nr = (R)sqrt(sum_sq);
  return (nr * nr) / R(2);
}

template <typename R>
PB_HD R pb_f_model(R fx, double gdr, double res_sq, R Lc) {
  const R nr = (R)sqrt(res_sq);
  return (fx - (R)gdr) + (Lc / R(2)) * (nr * nr);
}

template <typename R>
struct Nesterov {
  int kind;
  R m, stepsize, theta;   // adaptive
  R t;                    // fixed
  long k;                 // simple
  R constant;
  PB_HD void init(int kind_, R mf, R constant_beta) {
    kind = kind_;
    m = mf;
    stepsize = R(-1);
    theta = R(-1);
    t = R(1);
    k = 1;
    constant = constant_beta;
  }
  PB_HD R next(R gamma) {
    switch (kind) {
      case PB_SEQ_FIXED: {             // nesterov.jl:14-17
        const R t_next = (R(1) + pb_rs(R(1) + R(4) * (t * t))) / R(2);
        const R beta = (t - R(1)) / t_next;
        t = t_next;
        return beta;
      }
      case PB_SEQ_SIMPLE: {            // nesterov.jl:36
        const R beta = R(k - 1) / R(k + 2);
        ++k;
        return beta;
      }
      case PB_SEQ_CONSTANT:            // nesterov.jl:51-54
        return constant;
      default: {                       // AdaptiveNesterovSequence, nesterov.jl:89-103
        if (stepsize < 0) {
          stepsize = gamma;
          theta = m > 0 ? pb_rs(m * gamma) : R(1);
        }
        const R th2 = theta * theta;
        const R b = th2 / stepsize - m;
        const R delta = b * b + (R(4) * th2) / (stepsize * gamma);
        const R theta_n = (gamma * (pb_rs(delta) - b)) / R(2);
        const R beta = ((gamma * theta) * (R(1) - theta)) / (stepsize * theta_n + gamma * th2);
        stepsize = gamma;
        theta = theta_n;
        return beta;
      }
    }
  }
};
