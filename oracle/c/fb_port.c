/* CPU ORACLE (C port) -- TEST / BASELINE INFRASTRUCTURE ONLY, never on the product path.
 *
 * Plain-C restatement of the reference's UNFUSED per-iteration op sequence, one loop per Julia broadcast, used
 *   (a) as the timed "restated CPU baseline" of bench.py (`cpu_baseline` leg and `--impl reference`), and
 *   (b) as a second, independent checker of the numpy oracle (tests/test_oracle_port.py).
 * The reference is pure Julia and cannot be compiled or run in this image (no julia), so this is a PORT, not the
 * reference itself: bench.py labels it kind = "port".
 *
 * Pass structure follows src/algorithms/fast_forward_backward.jl:134-142 + the stop norm of :147-152:
 *     x .= z .+ beta .* (z .- z_prev)            (:135)   3 vector passes
 *     swap(z_prev, z)                            (:136)   pointer swap (done by the caller)
 *     grad_f_x .= grad                           (:139)   2   (the gradient is supplied as a buffer: "fused step only")
 *     y .= x .- gamma .* grad_f_x                (:140)   3
 *     g_z = prox!(z, g, y, gamma)                (:141)   2   NormL1 / IndBox (ProximalOperators.jl semantics)
 *     res .= x .- z                              (:142)   3
 *     norm(res, Inf)                             (:152)   1
 * Each loop is an OpenMP `parallel for` (static schedule): Julia's broadcast is single threaded, so giving the port all
 * host threads makes this baseline FASTER than the reference would be; single-thread timing is obtained with
 * OMP_NUM_THREADS=1 (the reference's own benchmark convention, benchmark/runbenchmarks.jl:46-47).
 * Compile WITHOUT -ffast-math / FMA contraction (-ffp-contract=off) so roundings match the numpy oracle bit for bit.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int port_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* torch.distributed.run exports OMP_NUM_THREADS=1 to its workers: the baseline sets its thread count explicitly so that it uses
 * the same host cores whether bench.py is started directly or under torchrun. */
void port_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

enum { PORT_PROX_L1 = 1, PORT_PROX_BOX = 2 };

#define DEFINE_PORT(T, SUF, FABS)                                                                                    \
  /* one unfused FISTA iteration; returns norm(res, Inf); *g_z receives g(z).  x is overwritten (extrapolated point). */ \
  double port_ffb_iteration_##SUF(int64_t n, T* x, const T* z_in, const T* z_prev, const T* grad_src, T* grad_f_x,   \
                                  T* y, T* z_out, T* res, T gamma, T beta, int prox, T p0, T p1, double* g_z) {      \
    int64_t i;                                                                                                       \
    _Pragma("omp parallel for schedule(static)") for (i = 0; i < n; ++i) x[i] = z_in[i] + beta * (z_in[i] - z_prev[i]); \
    _Pragma("omp parallel for schedule(static)") for (i = 0; i < n; ++i) grad_f_x[i] = grad_src[i];                   \
    _Pragma("omp parallel for schedule(static)") for (i = 0; i < n; ++i) y[i] = x[i] - gamma * grad_f_x[i];          \
    double gs = 0.0;                                                                                                 \
    if (prox == PORT_PROX_L1) {                                                                                      \
      const T gl = gamma * p0;                                                                                       \
      _Pragma("omp parallel for schedule(static) reduction(+ : gs)") for (i = 0; i < n; ++i) {                        \
        const T v = y[i];                                                                                            \
        const T zz = v + (v <= -gl ? gl : (v >= gl ? -gl : -v));                                                     \
        z_out[i] = zz;                                                                                               \
        gs += (double)FABS(zz);                                                                                      \
      }                                                                                                              \
      *g_z = (double)p0 * gs;                                                                                        \
    } else {                                                                                                         \
      _Pragma("omp parallel for schedule(static)") for (i = 0; i < n; ++i) {                                         \
        const T v = y[i];                                                                                            \
        z_out[i] = v < p0 ? p0 : (v > p1 ? p1 : v);                                                                  \
      }                                                                                                              \
      *g_z = 0.0;                                                                                                    \
    }                                                                                                                \
    _Pragma("omp parallel for schedule(static)") for (i = 0; i < n; ++i) res[i] = x[i] - z_out[i];                   \
    double m = 0.0;                                                                                                  \
    _Pragma("omp parallel for schedule(static) reduction(max : m)") for (i = 0; i < n; ++i) {                         \
      const double a = (double)FABS(res[i]);                                                                         \
      if (a > m) m = a;                                                                                              \
    }                                                                                                                \
    return m;                                                                                                        \
  }                                                                                                                  \
  /* one unfused forward-backward iteration with fixed stepsize (forward_backward.jl:111-120): caller swaps x,z */    \
  double port_fb_iteration_##SUF(int64_t n, const T* x, const T* grad_src, T* grad_f_x, T* y, T* z_out, T* res,       \
                                 T gamma, int prox, T p0, T p1, double* g_z) {                                       \
    int64_t i;                                                                                                       \
    _Pragma("omp parallel for schedule(static)") for (i = 0; i < n; ++i) grad_f_x[i] = grad_src[i];                   \
    _Pragma("omp parallel for schedule(static)") for (i = 0; i < n; ++i) y[i] = x[i] - gamma * grad_f_x[i];          \
    double gs = 0.0;                                                                                                 \
    if (prox == PORT_PROX_L1) {                                                                                      \
      const T gl = gamma * p0;                                                                                       \
      _Pragma("omp parallel for schedule(static) reduction(+ : gs)") for (i = 0; i < n; ++i) {                        \
        const T v = y[i];                                                                                            \
        const T zz = v + (v <= -gl ? gl : (v >= gl ? -gl : -v));                                                     \
        z_out[i] = zz;                                                                                               \
        gs += (double)FABS(zz);                                                                                      \
      }                                                                                                              \
      *g_z = (double)p0 * gs;                                                                                        \
    } else {                                                                                                         \
      _Pragma("omp parallel for schedule(static)") for (i = 0; i < n; ++i) {                                         \
        const T v = y[i];                                                                                            \
        z_out[i] = v < p0 ? p0 : (v > p1 ? p1 : v);                                                                  \
      }                                                                                                              \
      *g_z = 0.0;                                                                                                    \
    }                                                                                                                \
    _Pragma("omp parallel for schedule(static)") for (i = 0; i < n; ++i) res[i] = x[i] - z_out[i];                   \
    double m = 0.0;                                                                                                  \
    _Pragma("omp parallel for schedule(static) reduction(max : m)") for (i = 0; i < n; ++i) {                         \
      const double a = (double)FABS(res[i]);                                                                         \
      if (a > m) m = a;                                                                                              \
    }                                                                                                                \
    return m;                                                                                                        \
  }                                                                                                                  \
  /* parallel first-touch initialisation so pages are spread over NUMA nodes like a threaded run would have them */   \
  void port_fill_##SUF(int64_t n, T* v, uint64_t seed, T scale) {                                                     \
    int64_t i;                                                                                                       \
    _Pragma("omp parallel for schedule(static)") for (i = 0; i < n; ++i) {                                           \
      uint64_t s = (uint64_t)i * 0x9E3779B97F4A7C15ull + seed;                                                       \
      s ^= s >> 30; s *= 0xBF58476D1CE4E5B9ull; s ^= s >> 27; s *= 0x94D049BB133111EBull; s ^= s >> 31;              \
      v[i] = scale * (T)(((double)(s >> 11) * (1.0 / 9007199254740992.0)) * 2.0 - 1.0);                              \
    }                                                                                                                \
  }

DEFINE_PORT(float, f32, fabsf)
DEFINE_PORT(double, f64, fabs)
