// persist.cu -- the WHOLE ForwardBackward / FastForwardBackward solve in ONE cooperative kernel launch, for dense
// least-squares problems whose matrix stays cache resident (the reference's own benchmark suite: benchmark/benchmarks.jl:30-61
// runs on 5x10, 50x100 and 500x1000 Float64 fixtures; BASELINE.json configs[0] is 200x500).
//
// Why: driven from the host, such an iteration is ~8 kernel launches and 2 host synchronisations (20-35 us) around a few
// hundred nanoseconds of arithmetic -- the B200 loses to one CPU core.  Here the reference's driver loop
// (src/ProximalAlgorithms.jl:114-123), the iteration (forward_backward.jl:65-123, fast_forward_backward.jl:73-145), the line
// search (fb_tools.jl:24-63), the stepsize estimate (fb_tools.jl:7-12) and the Nesterov sequences (src/accel/nesterov.jl) all
// run ON the device: zero host round trips per iteration, one launch per solve.
//
// Layout: G <= 32 co-resident CTAs (cooperative launch) x 512 threads.  The columns of A are dealt to the CTAs in whole
// column chunks (the chunking of lsq_order.h); a CTA keeps ITS slices of x, grad, z, z_prev, ... in shared memory for the
// whole solve and reads its column slice of A through L1 (read-only, so it stays there / in L2: 4 MB at most).  What crosses
// CTAs per f-evaluation is the m-vector of chunk partials of A v (global memory, L2) and per step the 7 reduction partials;
// both are published before ONE grid barrier: the step's reductions travel together with the partial product the NEXT
// operation needs (A z for the line search / next gradient, A x_next for fixed-stepsize FISTA), so an iteration costs one
// barrier (two for adaptive FISTA, whose extrapolated point is only known after the line search).  Every CTA then folds the
// partials itself, in the same fixed order, and takes the same decisions from bit-identical scalars (R arithmetic shared
// with the host loop through solve_scalar.h; this file is compiled with -fmad=false).
//
// Parity: the products restate the summation orders of lsq_kernels.cu (selected by the same rule, lsq_order.h), the step is
// the same StepElem arithmetic as K1/K2, the reductions are double-double: iterates, scalars and iteration counts are
// bit-identical to the multi-kernel path (tests/test_gpu_persist.py), hence the same counts as the oracle on the fixtures.
#include <string.h>

#include "lsq_order.h"
#include "solve_scalar.h"
#include "step_common.cuh"

#define PS_THREADS 512
#define PS_MAX_CTAS 32
#define PS_NSLOT 6
#define PS_SCAL 12           // doubles per CTA in the reduction exchange: 3 (hi, lo) pairs, 1 max, pad, (hi, lo) of this CTA's share of ||r||^2

struct PersistParams {
  const void* A;
  const void* b;
  int64_t m, n, lda;
  void *x, *grad, *z, *z_prev;          // global n-vectors: x holds copy(x0) on entry; all four receive the final state
  // prox
  int prox_kind;
  double p0, p1;
  const void* lo_v;
  const void* hi_v;
  // options (pb_solve_opts)
  int algorithm, adaptive, sequence;
  int64_t maxit, n_global;
  double tol, gamma, mf, constant_beta, minimum_gamma, reduce_gamma, increase_gamma;
  // summation orders (lsq_order.h)
  PbLsqOrder ord;
  // cross-CTA workspace (global): chunk partials [2][nchunk][m], reduction partials [2][G][PS_SCAL], barrier counter
  void* partial;
  void* r_glob;                          // [m]: r = A v - b assembled from the CTAs' row shares (G > 1)
  double* scal;
  unsigned int* bar;                     // one arrival flag per CTA (epoch number), 32 bytes apart
  int ns_max;                            // slice capacity (elements) of one shared-memory slot
  int shared_xchg;                       // 1: one CTA, the cross-CTA buffers live in shared memory
  int a_in_smem;                         // 1: every CTA copies its column slice of A into shared memory at start
  int timing;                            // 1: CTA 0 accumulates clock64() per phase into `cycles`
  long long* cycles;                     // [PS_NPHASE]
  pb_solve_result* result;               // device copy, written by CTA 0
};

enum { PS_PH_GEMV_N = 0, PS_PH_BARRIER, PS_PH_COMBINE, PS_PH_GEMV_T, PS_PH_STEP, PS_PH_OTHER, PS_NPHASE = 8 };
#define PS_FLAG_STRIDE 8                 // unsigned ints between two CTAs' arrival flags

// ---------------------------------------------------------------------------------------------------------------------
// grid barrier: monotone counter, release/acquire at gpu scope.  All CTAs are co-resident (cooperative launch).
// ---------------------------------------------------------------------------------------------------------------------
// Arrival: every CTA owns one flag; it stores the epoch number there (release) once its data for this epoch is written.  Waiting:
// lane c of warp 0 polls CTA c's flag (acquire), so the G flags are watched in parallel and nothing serialises on one address.
__device__ __forceinline__ void ps_arrive(unsigned int* bar, unsigned int epoch) {
  __syncthreads();                       // this CTA's partials are written
  if (gridDim.x > 1 && threadIdx.x == 0) {   // release is cumulative: it also orders the other threads' writes seen through bar.sync
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(bar + (size_t)blockIdx.x * PS_FLAG_STRIDE), "r"(epoch) : "memory");
  }
}
__device__ __forceinline__ void ps_wait_warp0(const unsigned int* bar, unsigned int epoch) {   // called by warp 0 only
  if (gridDim.x > 1 && threadIdx.x < gridDim.x) {
    const unsigned int* f = bar + (size_t)threadIdx.x * PS_FLAG_STRIDE;
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
    } while ((int)(v - epoch) < 0);
  }
  __syncwarp();
}

// cross-CTA data goes through L2 (L1 is not coherent); with ONE CTA the same buffers live in shared memory
template <typename T>
__device__ __forceinline__ void ps_st(T* p, T v, bool one) {
  if (one)
    *p = v;
  else
    __stcg(p, v);
}
template <typename T>
__device__ __forceinline__ T ps_ld(const T* p, bool one) {
  return one ? *p : __ldcg(p);
}

// warp_reduce of common.cuh with the level loop ROLLED: each double-double level is ~100 instructions per sum, and the fully unrolled
// trees made every phase routine 20-40 KB of SASS (instruction-cache misses dominated the small problems)
template <int NSUM, int NMAX>
__device__ __forceinline__ void ps_warp_reduce(Acc<NSUM, NMAX>& a) {
#pragma unroll 1
  for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
    for (int k = 0; k < NSUM; ++k) {
      dd o;
      o.hi = shfl_down_d(a.s[k].hi, off);
      o.lo = shfl_down_d(a.s[k].lo, off);
      a.s[k] = dd_sum(a.s[k], o);
    }
#pragma unroll
    for (int k = 0; k < NMAX; ++k) a.m[k] = nanmax(a.m[k], shfl_down_d(a.m[k], off));
  }
}

// Block reduction restricted to the warps that hold data (`active` threads, tid < active): the double-double shuffle trees are
// FP64-pipe bound, and a slice of 32 columns keeps 8 of the 512 threads busy.  Result valid in thread 0; `red`: 16 x 8 doubles.
template <int NSUM, int NMAX>
__device__ __forceinline__ void ps_block_reduce(Acc<NSUM, NMAX>& a, int active, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int naw = (active + 31) >> 5;
  if (naw < 1) naw = 1;
  if (warp < naw) {
    ps_warp_reduce<NSUM, NMAX>(a);
    if (lane == 0 && naw > 1) {
#pragma unroll
      for (int k = 0; k < NSUM; ++k) {
        red[warp * 8 + 2 * k] = a.s[k].hi;
        red[warp * 8 + 2 * k + 1] = a.s[k].lo;
      }
#pragma unroll
      for (int k = 0; k < NMAX; ++k) red[warp * 8 + 6 + k] = a.m[k];
    }
  }
  if (naw > 1) {
    __syncthreads();
    if (warp == 0) {
#pragma unroll
      for (int k = 0; k < NSUM; ++k) {
        a.s[k].hi = lane < naw ? red[lane * 8 + 2 * k] : 0.0;
        a.s[k].lo = lane < naw ? red[lane * 8 + 2 * k + 1] : 0.0;
      }
#pragma unroll
      for (int k = 0; k < NMAX; ++k) a.m[k] = lane < naw ? red[lane * 8 + 6 + k] : 0.0;
      ps_warp_reduce<NSUM, NMAX>(a);
    }
  }
}

// The phase routines below are __noinline__ on purpose: the driver loop calls them from many places, and fully inlined the kernel
// was ~1 MB of SASS -- every iteration then ran out of the instruction cache (measured: 1-2k cycles of fetch stalls per phase).
// ---------------------------------------------------------------------------------------------------------------------
// r = A v - b, phase 1: partial[ch][i] = sum_{j in chunk ch} A[i, j] v[j] for the chunks this CTA owns.  v: shared-memory slice,
// element j lives at v[j - j0].  Two summation orders, exactly those of k_gemv_n_partial / k_gemv_n_sub (lsq_kernels.cu).
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
__device__ __noinline__ void ps_gemv_n_rows(const PersistParams& p, const T* __restrict__ A, int64_t lda, const T* __restrict__ v, int64_t j0, int ch0,
                               int ch1, T* partial, T* shp, bool one) {
  // row-per-thread order: 4 column lanes, each a sequential FMA chain over its columns j = c0 + cl, c0 + cl + 4, ...; the
  // lanes are then added ((l0 + l1) + l2) + l3.  512 threads = 128 rows x 4 lanes, four row tiles in flight per thread.
  const int tid = threadIdx.x, rl = tid & 127, cl = tid >> 7;
  const int64_t m = p.m;
  for (int ch = ch0; ch < ch1; ++ch) {
    const int64_t c0 = (int64_t)ch * p.ord.chunk_cols;
    int64_t c1 = c0 + p.ord.chunk_cols;
    if (c1 > p.n) c1 = p.n;
    for (int64_t row0 = 0; row0 < m; row0 += 512) {
      T acc[4] = {T(0), T(0), T(0), T(0)};
      int64_t j = c0 + cl;
      for (; j + 4 < c1; j += 8) {              // two columns of this lane per trip: 8 independent loads in flight
        const T x0 = v[j - j0], x1 = v[j + 4 - j0];
        T a0[4], a1[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int64_t row = row0 + t * 128 + rl;
          a0[t] = row < m ? A[row + j * lda] : T(0);
          a1[t] = row < m ? A[row + (j + 4) * lda] : T(0);
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          acc[t] = fma(a0[t], x0, acc[t]);
          acc[t] = fma(a1[t], x1, acc[t]);
        }
      }
      for (; j < c1; j += 4) {
        const T x0 = v[j - j0];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int64_t row = row0 + t * 128 + rl;
          if (row < m) acc[t] = fma(A[row + j * lda], x0, acc[t]);
        }
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) shp[cl * 512 + t * 128 + rl] = acc[t];
      __syncthreads();
      const int64_t row = row0 + tid;
      if (row < m) {
        T s = shp[tid];
        s += shp[512 + tid];
        s += shp[1024 + tid];
        s += shp[1536 + tid];
        ps_st(partial + (int64_t)ch * m + row, s, one);
      }
      __syncthreads();
    }
  }
}

template <typename T>
__device__ __noinline__ void ps_gemv_n_sub(const PersistParams& p, const T* __restrict__ A, int64_t lda, const T* __restrict__ v, int64_t j0, int ch0,
                              int ch1, T* partial, T* shp, bool one) {
  // short-column order (m < 64): LPC lanes share a column, each owning up to KP 16-byte packs of it; a "virtual CTA" of 256
  // threads (8 warps) works one chunk exactly like k_gemv_n_sub does, two virtual CTAs side by side.
  constexpr int VEC = 16 / sizeof(T);
  const int lpc = p.ord.n_lpc, kp = p.ord.n_kp;
  const int vt = threadIdx.x & 255, vc = threadIdx.x >> 8;
  const int lane = vt & 31, warp = vt >> 5;
  const int cpw = 32 / lpc, sub = lane % lpc, colw = lane / lpc;
  const int64_t stride = (int64_t)8 * cpw;
  const int64_t m = p.m, npk = m / VEC;
  const int W = lpc * kp * VEC;                      // <= 64
  T* sh = shp + vc * (8 * 64);
  for (int chb = ch0; chb < ch1; chb += 2) {
    const int ch = chb + vc;
    const bool active = ch < ch1;
    if (active) {
      const int64_t c0 = (int64_t)ch * p.ord.chunk_cols;
      int64_t c1 = c0 + p.ord.chunk_cols;
      if (c1 > p.n) c1 = p.n;
      T acc[4][VEC];
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[q][e] = T(0);
      for (int64_t jb = c0 + (int64_t)warp * cpw + colw; jb < c1; jb += 4 * stride) {
#pragma unroll
        for (int s_ = 0; s_ < 4; ++s_) {
          const int64_t j = jb + s_ * stride;
          const bool live = j < c1;
          const T xv = live ? v[j - j0] : T(0);
          const T* __restrict__ a = A + (live ? j : c0) * lda;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int64_t pk = sub + (int64_t)q * lpc;
            if (q < kp && pk < npk) {
#pragma unroll
              for (int e = 0; e < VEC; ++e) acc[q][e] = fma(a[pk * VEC + e], xv, acc[q][e]);
            }
          }
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          if (q < kp) {                                // uniform across the warp
            T w = acc[q][e];
            for (int off = 16; off >= lpc; off >>= 1) w += __shfl_xor_sync(0xffffffffu, w, off);
            if (colw == 0) sh[warp * W + (q * lpc + sub) * VEC + e] = w;
          }
        }
    }
    __syncthreads();
    if (active) {
      for (int64_t i = vt; i < m; i += 256) {
        T s = sh[i];
#pragma unroll
        for (int w = 1; w < 8; ++w) s += sh[w * W + i];
        ps_st(partial + (int64_t)ch * m + i, s, one);
      }
    }
    __syncthreads();
  }
}

// phase 2: r_i = (sum over chunks, in chunk order) - b_i for rows [i0, i1); the dd sum of r_i^2 over those rows ends up in thread 0's
// `acc`.  One CTA: all rows, r into shared memory.  Several CTAs: each assembles ITS share of the rows (the partials of all chunks
// of a row are 8-byte loads from L2) and publishes it; after a second barrier everyone copies the m values of r.
template <typename T>
__device__ __noinline__ void ps_combine_rows(const PersistParams& p, const T* partial, T* r_out, bool r_shared, int64_t i0, int64_t i1, bool one,
                                Acc<1, 1>& acc, double* red) {
  constexpr bool COMP = sizeof(T) == 8;
  const T* __restrict__ b = static_cast<const T*>(p.b);
  const int64_t m = p.m;
  const int nchunk = (int)p.ord.nchunk;
  acc.clear();
  for (int64_t i = i0 + threadIdx.x; i < i1; i += PS_THREADS) {
    T s = ps_ld(partial + i, one);
    int c = 1;
    for (; c + 7 < nchunk; c += 8) {            // 8 independent loads in flight, added in chunk order
      T t[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) t[u] = ps_ld(partial + (int64_t)(c + u) * m + i, one);
#pragma unroll
      for (int u = 0; u < 8; ++u) s += t[u];
    }
    for (; c < nchunk; ++c) s += ps_ld(partial + (int64_t)c * m + i, one);
    const T rv = b ? sub_rn(s, __ldg(b + i)) : s;
    ps_st(r_out + i, rv, r_shared);
    if (COMP)
      dd_add_prod(acc.s[0], (double)rv, (double)rv);
    else
      acc.s[0].hi = __fma_rn((double)rv, (double)rv, acc.s[0].hi);
  }
  const int64_t rows = i1 - i0;
  ps_block_reduce<1, 1>(acc, (int)(rows < PS_THREADS ? rows : PS_THREADS), red);
}

// grad_j = sum_i A[i, j] r_i for this CTA's columns, in the order of k_gemv_t (warp per column) or k_gemv_t_sub (LPC lanes per column)
template <typename T>
__device__ __noinline__ void ps_gemv_t(const PersistParams& p, const T* __restrict__ A, int64_t lda, const T* __restrict__ r_sh, T* g, int64_t j0, int64_t j1) {
  constexpr int VEC = 16 / sizeof(T);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t m = p.m;
  if (!p.ord.t_sub) {
    for (int64_t j = j0 + warp; j < j1; j += PS_THREADS / 32) {
      const T* __restrict__ a = A + j * lda;
      T acc = T(0);
      int64_t i = lane;
      for (; i + 96 < m; i += 128) {
        const T a0 = a[i], a1 = a[i + 32], a2 = a[i + 64], a3 = a[i + 96];
        acc = fma(a0, r_sh[i], acc);
        acc = fma(a1, r_sh[i + 32], acc);
        acc = fma(a2, r_sh[i + 64], acc);
        acc = fma(a3, r_sh[i + 96], acc);
      }
      for (; i < m; i += 32) acc = fma(a[i], r_sh[i], acc);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
      if (lane == 0) g[j - j0] = acc;
    }
  } else {
    const int lpc = p.ord.t_lpc, kp = p.ord.t_kp;
    const int cpw = 32 / lpc, sub = lane % lpc, colw = lane / lpc;
    const int64_t npk = m / VEC;
    for (int64_t jb = j0 + (int64_t)warp * cpw; jb < j1; jb += (int64_t)(PS_THREADS / 32) * cpw) {
      const int64_t j = jb + colw;
      const bool live = j < j1;
      const T* __restrict__ a = A + (live ? j : j0) * lda;
      T acc = T(0);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int64_t pk = sub + (int64_t)q * lpc;
        if (q < kp && pk < npk) {
#pragma unroll
          for (int e = 0; e < VEC; ++e) acc = fma(a[pk * VEC + e], r_sh[pk * VEC + e], acc);
        }
      }
      for (int off = lpc >> 1; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
      if (live && sub == 0) g[j - j0] = acc;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// the fused forward-backward step on this CTA's slice (same StepElem arithmetic as K1/K2) + publication of its reduction partials
// ---------------------------------------------------------------------------------------------------------------------
template <typename T, int PROX, bool EXTRAP>
__device__ __noinline__ void ps_step_t(const PersistParams& p, const T* x, const T* g, const T* zp, T gamma, T beta, T* z, T* xn, int64_t j0,
                          int ns, double* scal_out, double* red) {
  constexpr bool COMP = sizeof(T) == 8;
  const T* __restrict__ lov = static_cast<const T*>(p.lo_v);
  const T* __restrict__ hiv = static_cast<const T*>(p.hi_v);
  T pa = T(0), pb = T(0);
  if (PROX == PB_PROX_L1) pa = mul_rn(gamma, (T)p.p0);          // gamma*lambda: one rounding in R, like the package
  if (PROX == PB_PROX_BOX) {
    pa = (T)p.p0;
    pb = (T)p.p1;
  }
  constexpr int VEC = 16 / sizeof(T);
  Acc<3, 1> acc, pk;
  acc.clear();
  pk.clear();
  // one 16-byte pack per thread and trip (slices start at multiples of 32 elements, so these are the packs of K1/K2): float sums
  // are grouped per pack exactly like there (fold_pack), the last < VEC elements of the vector are single-element groups
  for (int q = threadIdx.x; q * VEC < ns; q += PS_THREADS) {
    const bool whole = j0 + (int64_t)q * VEC + VEC <= p.n;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const int e = q * VEC + k;
      if (e < ns) {
        const T l = (PROX == PB_PROX_BOX && lov) ? __ldg(lov + j0 + e) : pa;
        const T h = (PROX == PB_PROX_BOX && hiv) ? __ldg(hiv + j0 + e) : pb;
        T yv, zn, rv, xv;
        StepElem<T, PROX, EXTRAP>::template run<COMP>(x[e], g[e], EXTRAP ? zp[e] : T(0), l, h, gamma, beta, yv, zn, rv, xv,
                                                      COMP ? acc : pk);
        z[e] = zn;
        if constexpr (EXTRAP) xn[e] = xv;
        if constexpr (!COMP) {
          if (!whole) fold_pack<PROX>(acc, pk);
        }
      }
    }
    if constexpr (!COMP) {
      if (whole) fold_pack<PROX>(acc, pk);
    }
  }
  const int packs = (ns + VEC - 1) / VEC;
  ps_block_reduce<3, 1>(acc, packs < PS_THREADS ? packs : PS_THREADS, red);
  if (threadIdx.x == 0) {
    const bool one = p.shared_xchg != 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      ps_st(scal_out + 2 * k, acc.s[k].hi, one);
      ps_st(scal_out + 2 * k + 1, acc.s[k].lo, one);
    }
    ps_st(scal_out + 6, acc.m[0], one);
  }
}

template <typename T>
__device__ void ps_step(const PersistParams& p, const T* x, const T* g, const T* zp, T gamma, T beta, bool extrap, T* z, T* xn,
                        int64_t j0, int ns, double* scal_out, double* red) {
  switch (p.prox_kind) {
    case PB_PROX_L1:
      if (extrap) ps_step_t<T, PB_PROX_L1, true>(p, x, g, zp, gamma, beta, z, xn, j0, ns, scal_out, red);
      else ps_step_t<T, PB_PROX_L1, false>(p, x, g, zp, gamma, beta, z, xn, j0, ns, scal_out, red);
      break;
    case PB_PROX_BOX:
      if (extrap) ps_step_t<T, PB_PROX_BOX, true>(p, x, g, zp, gamma, beta, z, xn, j0, ns, scal_out, red);
      else ps_step_t<T, PB_PROX_BOX, false>(p, x, g, zp, gamma, beta, z, xn, j0, ns, scal_out, red);
      break;
    default:
      if (extrap) ps_step_t<T, PB_PROX_ZERO, true>(p, x, g, zp, gamma, beta, z, xn, j0, ns, scal_out, red);
      else ps_step_t<T, PB_PROX_ZERO, false>(p, x, g, zp, gamma, beta, z, xn, j0, ns, scal_out, red);
      break;
  }
}

struct PsComb {
  double gsum, res_sq, gdr, res_inf, sumsq;
};

// The grid barrier and the fold of the G CTAs' reduction partials in one: warp 0 waits for the G arrival flags (lane c watches CTA c),
// lane c then loads CTA c's partials, a fixed shuffle tree folds them, and the rounded scalars are broadcast through shared memory.
// want_scal = false: barrier only (phases that publish just a partial product).
__device__ __noinline__ PsComb ps_wait_fold(const PersistParams& p, unsigned int epoch, const double* scal, bool want_scal, bool one, double* bc) {
  if (threadIdx.x < 32) {
    ps_wait_warp0(p.bar, epoch);
    if (want_scal) {
      Acc<4, 1> a;
      a.clear();
      if (threadIdx.x < gridDim.x) {
        const double* s = scal + (size_t)threadIdx.x * PS_SCAL;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          a.s[k].hi = ps_ld(s + 2 * k, one);
          a.s[k].lo = ps_ld(s + 2 * k + 1, one);
        }
        a.m[0] = ps_ld(s + 6, one);
        a.s[3].hi = ps_ld(s + 8, one);
        a.s[3].lo = ps_ld(s + 9, one);
      }
      if (gridDim.x > 1) ps_warp_reduce<4, 1>(a);
      if (threadIdx.x == 0) {
        bc[0] = a.s[0].hi + a.s[0].lo;
        bc[1] = a.s[1].hi + a.s[1].lo;
        bc[2] = a.s[2].hi + a.s[2].lo;
        bc[3] = a.m[0];
        bc[4] = a.s[3].hi + a.s[3].lo;
      }
    }
  }
  __syncthreads();
  PsComb c;
  c.gsum = bc[0];
  c.res_sq = bc[1];
  c.gdr = bc[2];
  c.res_inf = bc[3];
  c.sumsq = bc[4];
  __syncthreads();                       // everyone has read the broadcast before its next writer runs
  return c;
}

// ---------------------------------------------------------------------------------------------------------------------
// the solve
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(PS_THREADS, 1) k_persist_solve(PersistParams p) {
  typedef T R;
  constexpr bool COMP = sizeof(T) == 8;
  extern __shared__ __align__(16) unsigned char ps_smem[];
  __shared__ double bc[8];
  __shared__ double red[16 * 8];
  __shared__ double scal_sh[2 * PS_SCAL];
  __shared__ long long cyc[PS_NPHASE];
  T* slot[PS_NSLOT];
  for (int k = 0; k < PS_NSLOT; ++k) slot[k] = reinterpret_cast<T*>(ps_smem) + (size_t)k * p.ns_max;
  T* r_sh = reinterpret_cast<T*>(ps_smem) + (size_t)PS_NSLOT * p.ns_max;
  const int64_t m_pad = (p.m + 3) & ~(int64_t)3;
  T* shp = r_sh + m_pad;                              // 2048 elements: lane partials of the residual orders
  T* dyn_next = shp + 2048;                           // one-CTA exchange buffers and / or this CTA's slice of A follow

  const int G = gridDim.x, cta = blockIdx.x, tid = threadIdx.x;
  const int nchunk = (int)p.ord.nchunk;
  const int base = nchunk / G, extra = nchunk % G;
  const int ch0 = cta * base + (cta < extra ? cta : extra);
  const int ch1 = ch0 + base + (cta < extra ? 1 : 0);
  const int64_t j0 = (int64_t)ch0 * p.ord.chunk_cols;
  int64_t j1 = (int64_t)ch1 * p.ord.chunk_cols;
  if (j1 > p.n) j1 = p.n;
  if (j1 < j0) j1 = j0;
  const int ns = (int)(j1 - j0);
  const int64_t m = p.m;
  // Cross-CTA buffers: L2 (global) when several CTAs cooperate, shared memory when one does.  Every product round is
  // [publish chunk partials (+ the step's reduction partials)] -> barrier A -> [each CTA assembles its share of the rows of r and its
  // share of ||r||^2] -> barrier B + fold -> [everyone copies r].  Barrier B orders the next round's writes of `partial` after this
  // round's reads and barrier A the next writes of `r_glob` after this round's copies, so both are single buffered; the reduction
  // partials are read after B while a fast CTA may already publish the next ones: double buffered by round parity.
  const bool one = p.shared_xchg != 0;
  T* partial = static_cast<T*>(p.partial);
  T* r_glob = static_cast<T*>(p.r_glob);
  double* scal_buf[2] = {p.scal, p.scal + (size_t)PS_MAX_CTAS * PS_SCAL};
  if (one) {
    partial = dyn_next;
    dyn_next += (size_t)nchunk * m;
    scal_buf[0] = scal_sh;
    scal_buf[1] = scal_sh + PS_SCAL;
  }
  // this CTA's column slice of A: shared memory when it fits (read ~3 times per iteration for thousands of iterations), else L1/L2
  const T* Aeff = static_cast<const T*>(p.A);
  int64_t lda_eff = p.lda;
  if (p.a_in_smem) {
    const T* __restrict__ Ag = static_cast<const T*>(p.A);
    T* As = dyn_next;
    for (int64_t idx = tid; idx < (int64_t)ns * m; idx += PS_THREADS) {
      const int64_t jj = idx / m, ii = idx - jj * m;
      As[idx] = __ldg(Ag + ii + (j0 + jj) * p.lda);
    }
    Aeff = As - j0 * m;
    lda_eff = m;
  }
  const int64_t rows_base = m / G, rows_extra = m % G;
  const int64_t i0 = cta * rows_base + (cta < rows_extra ? cta : rows_extra);
  const int64_t i1 = i0 + rows_base + (cta < rows_extra ? 1 : 0);
  unsigned int epoch = 0, gen = 0;
  long long t_last = 0;
  if (p.timing && cta == 0 && tid == 0) {
    for (int k = 0; k < PS_NPHASE; ++k) cyc[k] = 0;
    t_last = clock64();
  }
  auto lap = [&](int phase) {                 // CTA 0, thread 0: cycles since the previous lap go to `phase`
    if (p.timing && cta == 0 && tid == 0) {
      const long long t = clock64();
      cyc[phase] += t - t_last;
      t_last = t;
    }
  };

  // slots (pointers are swapped exactly where the reference swaps its vectors)
  T *X = slot[0], *GR = slot[1], *Z = slot[2], *ZP = slot[3], *W1 = slot[4], *W2 = slot[5];
  // W1: x_next (fixed FFB) / grad_f_z (adaptive FB);  W2: scratch (x + 1 of the stepsize estimate)

  {
    const T* __restrict__ xg = static_cast<const T*>(p.x);
    for (int e = tid; e < ns; e += PS_THREADS) X[e] = xg[j0 + e];
  }
  __syncthreads();

  auto publish_Av = [&](const T* v) {          // chunk partials of A v into the buffer of the NEXT barrier
    lap(PS_PH_OTHER);
    if (p.ord.n_sub)
      ps_gemv_n_sub<T>(p, Aeff, lda_eff, v, j0, ch0, ch1, partial, shp, one);
    else
      ps_gemv_n_rows<T>(p, Aeff, lda_eff, v, j0, ch0, ch1, partial, shp, one);
    lap(PS_PH_GEMV_N);
  };
  auto scal_slot = [&]() { return scal_buf[gen & 1] + (size_t)cta * PS_SCAL; };
  // grid barrier (+ fold of the reduction partials of the current round)
  auto sync_fold = [&](bool want_scal) {
    lap(PS_PH_OTHER);
    epoch += 1;
    ps_arrive(p.bar, epoch);
    const PsComb c = ps_wait_fold(p, epoch, scal_buf[gen & 1], want_scal, one, bc);
    lap(PS_PH_BARRIER);
    return c;
  };
  // Completes a product round started by publish_Av (the step's reduction partials, if any, are already in this round's slots):
  // leaves r = A v - b in shared memory, returns ||r||^2 and (with_scal) the folded reductions of the step.
  auto round_Av = [&](bool with_scal, PsComb& sc_out) {
    double aux_out;
    Acc<1, 1> a;
    if (G == 1) {
      __syncthreads();
      ps_combine_rows<T>(p, partial, r_sh, true, 0, m, one, a, red);
      if (tid == 0) {
        double* so = scal_slot();
        ps_st(so + 8, a.s[0].hi, one);
        ps_st(so + 9, a.s[0].lo, one);
      }
      lap(PS_PH_COMBINE);
      const PsComb c = sync_fold(true);
      if (with_scal) sc_out = c;
      aux_out = c.sumsq;
    } else {
      (void)sync_fold(false);                    // barrier A: every chunk partial is in L2
      ps_combine_rows<T>(p, partial, r_glob, false, i0, i1, false, a, red);
      if (tid == 0) {
        double* so = scal_slot();
        __stcg(so + 8, a.s[0].hi);
        __stcg(so + 9, a.s[0].lo);
      }
      lap(PS_PH_COMBINE);
      const PsComb c = sync_fold(true);          // barrier B: every row share of r is in L2
      if (with_scal) sc_out = c;
      aux_out = c.sumsq;
      for (int64_t i = tid; i < m; i += PS_THREADS) r_sh[i] = __ldcg(r_glob + i);
      __syncthreads();
      lap(PS_PH_COMBINE);
    }
    gen += 1;
    return aux_out;
  };
  auto gemv_t = [&](T* gout) {
    lap(PS_PH_OTHER);
    ps_gemv_t<T>(p, Aeff, lda_eff, r_sh, gout, j0, j1);
    __syncthreads();
    lap(PS_PH_GEMV_T);
  };
  auto step = [&](const T* xs, const T* gs, const T* zps, R gam, R bet, bool extrap, T* zs, T* xns) {
    lap(PS_PH_OTHER);
    ps_step<T>(p, xs, gs, zps, gam, bet, extrap, zs, xns, j0, ns, scal_slot(), red);
    __syncthreads();
    lap(PS_PH_STEP);
  };
  auto f_value = [&](double aux) { return pb_sq_half<R>(aux); };
  auto g_value = [&](const PsComb& c) { return p.prox_kind == PB_PROX_L1 ? (R)p.p0 * (R)c.gsum : R(0); };

  const bool fast = p.algorithm == PB_ALG_FFB;
  const bool adaptive = p.adaptive != 0;
  const R min_gamma = (R)p.minimum_gamma, red_gamma = (R)p.reduce_gamma, inc_gamma = (R)p.increase_gamma;
  R gamma, f_x = R(0), g_z = R(0);
  double aux, aux_n;
  int64_t backtracks = 0;
  int warned = 0;

  // ---- init: forward_backward.jl:65-84 / fast_forward_backward.jl:73-97 ----
  PsComb sc, sc_dummy;
  publish_Av(X);
  aux = round_Av(false, sc_dummy);              // r = A x - b
  gemv_t(GR);                                   // grad f(x)
  bool fx_pending = true;
  if (p.gamma <= 0) {                           // fb_tools.jl:7-12 with A = I
    f_x = f_value(aux);
    fx_pending = false;
    for (int e = tid; e < ns; e += PS_THREADS) W2[e] = add_rn(X[e], T(1));
    __syncthreads();
    publish_Av(W2);
    (void)round_Av(false, sc_dummy);
    gemv_t(Z);                                  // z is free at this point: holds grad f(x + 1)
    Acc<1, 1> a;
    a.clear();
    for (int e = tid; e < ns; e += PS_THREADS) {
      const T d = sub_rn(Z[e], GR[e]);
      Z[e] = d;
      if (COMP)
        dd_add_prod(a.s[0], (double)d, (double)d);
      else
        a.s[0].hi = __fma_rn((double)d, (double)d, a.s[0].hi);
    }
    ps_block_reduce<1, 1>(a, ns < PS_THREADS ? ns : PS_THREADS, red);
    if (tid == 0) {
      double* so = scal_slot();
      ps_st(so + 0, a.s[0].hi, one);
      ps_st(so + 1, a.s[0].lo, one);
      for (int k = 2; k < 10; ++k) ps_st(so + k, 0.0, one);
    }
    const PsComb c2 = sync_fold(true);
    gen += 1;
    const int64_t n_glob = p.n_global > 0 ? p.n_global : p.n;
    const R lower = (R)sqrt(c2.gsum) / (R)sqrt((double)n_glob);
    gamma = R(1) / lower;
  } else {
    gamma = (R)p.gamma;
  }
  Nesterov<R> seq;
  seq.init(p.sequence, (R)p.mf, (R)p.constant_beta);
  R beta_next = R(0);
  __syncthreads();
  if (fast) {
    for (int e = tid; e < ns; e += PS_THREADS) ZP[e] = X[e];            // z_prev = copy(x)
    __syncthreads();
  }
  const bool fused_extrap = fast && !adaptive;
  if (fused_extrap) beta_next = seq.next(gamma);
  step(X, GR, ZP, gamma, beta_next, fused_extrap, Z, W1);
  publish_Av(fused_extrap ? W1 : Z);            // the product the next operation needs rides on the same barriers
  aux_n = round_Av(true, sc);
  if (fx_pending) f_x = f_value(aux);
  g_z = g_value(sc);

  // fb_tools.jl:24-63 (A = nothing, Az aliased to z).  On entry r_sh = A z - b and aux_n = ||r||^2 are already known.
  auto backtrack = [&](bool want_grad, R& f_z_out) {
    const R eps = sizeof(T) == 4 ? (R)1.1920928955078125e-07 : (R)2.220446049250313e-16;
    R f_upp = pb_f_model<R>(f_x, sc.gdr, sc.res_sq, R(1) / gamma);
    if (want_grad) gemv_t(W1);                  // grad f(z)
    R f_z = f_value(aux_n);
    R tol = R(10) * eps * (R(1) + (R)fabs((double)f_z));
    while (f_z > f_upp + tol && gamma >= min_gamma) {
      gamma = gamma * red_gamma;
      step(X, GR, ZP, gamma, R(0), false, Z, W1);
      publish_Av(Z);
      aux_n = round_Av(true, sc);
      g_z = g_value(sc);
      f_upp = pb_f_model<R>(f_x, sc.gdr, sc.res_sq, R(1) / gamma);
      if (want_grad) gemv_t(W1);
      f_z = f_value(aux_n);
      tol = R(10) * eps * (R(1) + (R)fabs((double)f_z));
      ++backtracks;
    }
    if (gamma < min_gamma) warned = 1;
    f_z_out = f_z;
  };
  auto stop = [&]() { return (double)((R)sc.res_inf / gamma) <= p.tol; };

  // ---- driver loop: src/ProximalAlgorithms.jl:114-123 ----
  int64_t k = 1;
  for (;; ++k) {
    if (k >= p.maxit || stop()) break;
    if (!fast) {                                  // forward_backward.jl:86-123
      if (adaptive) {
        gamma = gamma * inc_gamma;
        R f_z;
        backtrack(true, f_z);
        f_x = f_z;
        T* t = X; X = Z; Z = t;
        t = GR; GR = W1; W1 = t;
      } else {
        T* t = X; X = Z; Z = t;
        aux = aux_n;                              // r = A x - b was published with the previous step
        gemv_t(GR);
      }
      step(X, GR, ZP, gamma, R(0), false, Z, W1);
      publish_Av(Z);
      aux_n = round_Av(true, sc);
      if (!adaptive) f_x = f_value(aux);
      g_z = g_value(sc);
    } else {                                      // fast_forward_backward.jl:106-145
      if (adaptive) {
        gamma = gamma * inc_gamma;
        R f_z;
        backtrack(false, f_z);
        const R beta = seq.next(gamma);
        for (int e = tid; e < ns; e += PS_THREADS) X[e] = add_rn(Z[e], mul_rn(beta, sub_rn(Z[e], ZP[e])));   // :135
        T* t = ZP; ZP = Z; Z = t;                 // :136
        __syncthreads();
        publish_Av(X);
        aux = round_Av(false, sc_dummy);
        gemv_t(GR);
        step(X, GR, ZP, gamma, R(0), false, Z, W1);
        publish_Av(Z);
      } else {
        gamma = (R)p.gamma > 0 ? (R)p.gamma : gamma;
        T* t = X; X = W1; W1 = t;                 // :135, computed by the previous fused pass
        t = ZP; ZP = Z; Z = t;                    // :136
        aux = aux_n;
        gemv_t(GR);
        beta_next = seq.next(gamma);
        step(X, GR, ZP, gamma, beta_next, true, Z, W1);
        publish_Av(W1);
      }
      aux_n = round_Av(true, sc);
      f_x = f_value(aux);
      g_z = g_value(sc);
    }
  }

  // ---- final state ----
  {
    T* xg = static_cast<T*>(p.x);
    T* gg = static_cast<T*>(p.grad);
    T* zg = static_cast<T*>(p.z);
    T* zpg = static_cast<T*>(p.z_prev);
    for (int e = tid; e < ns; e += PS_THREADS) {
      xg[j0 + e] = X[e];
      gg[j0 + e] = GR[e];
      zg[j0 + e] = Z[e];
      if (fast && zpg) zpg[j0 + e] = ZP[e];
    }
  }
  if (cta == 0 && tid == 0) {
    pb_solve_result* out = p.result;
    out->iterations = k;
    out->backtracks = backtracks;
    out->gamma = (double)gamma;
    out->f_x = (double)f_x;
    out->g_z = (double)g_z;
    out->res_inf = sc.res_inf;
    out->res_sq = sc.res_sq;
    out->gdr = sc.gdr;
    out->gsum = sc.gsum;
    out->warned_small_gamma = warned;
    if (p.timing) {
      lap(PS_PH_OTHER);
      for (int q = 0; q < PS_NPHASE; ++q) p.cycles[q] = cyc[q];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
struct PersistPlan {
  int ctas;
  int ns_max;
  int shared_xchg, a_in_smem;
  size_t smem;
  PbLsqOrder ord;
};

template <typename T>
static bool persist_plan(const pb_ctx* ctx, const pb_smooth* f, int want_ctas, PersistPlan* plan) {
  const int64_t m = f->m, n = f->n;
  if (m < 1 || n < 1 || m > 4096) return false;
  const double bytes = (double)m * (double)n * sizeof(T);
  if (bytes > 8.0 * 1024 * 1024) return false;          // the matrix must stay L2 resident for thousands of iterations
  plan->ord = pb_lsq_order(sizeof(T), 1, m, n, f->lda, 0, f->A, f->r);
  const int nchunk = (int)plan->ord.nchunk;
  int G = want_ctas > 0 ? want_ctas : (bytes <= 96.0 * 1024 ? 1 : PS_MAX_CTAS);
  if (G > nchunk) G = nchunk;
  if (G > PS_MAX_CTAS) G = PS_MAX_CTAS;
  if (G > ctx->sm_count) G = ctx->sm_count;
  if (G < 1) G = 1;
  const int64_t ns_max = ((int64_t)(nchunk + G - 1) / G) * plan->ord.chunk_cols;
  const int64_t m_pad = (m + 3) & ~(int64_t)3;
  const size_t limit = 200 * 1024;
  size_t smem = ((size_t)PS_NSLOT * ns_max + m_pad + 2048) * sizeof(T);
  if (smem > limit) return false;
  // one CTA: the chunk partials [nchunk][m] live in shared memory too (no trip through L2)
  plan->shared_xchg = 0;
  if (G == 1 && smem + (size_t)nchunk * m * sizeof(T) <= limit) {
    plan->shared_xchg = 1;
    smem += (size_t)nchunk * m * sizeof(T);
  }
  // the CTA's column slice of A in shared memory when it fits
  plan->a_in_smem = 0;
  if (smem + (size_t)ns_max * m * sizeof(T) <= limit) {
    plan->a_in_smem = 1;
    smem += (size_t)ns_max * m * sizeof(T);
  }
  plan->ctas = G;
  plan->ns_max = (int)ns_max;
  plan->smem = smem;
  return true;
}

bool pb_persist_eligible(const pb_ctx* ctx, int dtype, const pb_smooth* f, const pb_prox* g, const pb_solve_opts* o) {
  if (ctx->persist_mode < 0) return false;
  if (f->kind != PB_F_LSQ_DENSE || ctx->xchg_world > 1) return false;
  if (g->kind != PB_PROX_ZERO && g->kind != PB_PROX_L1 && g->kind != PB_PROX_BOX) return false;
  if (o->gamma <= 0 && !o->adaptive) return false;
  PersistPlan plan;
  return dtype == PB_F32 ? persist_plan<float>(ctx, f, ctx->persist_mode, &plan) : persist_plan<double>(ctx, f, ctx->persist_mode, &plan);
}

template <typename T>
static int persist_run(pb_ctx* ctx, int64_t n, const pb_smooth* f, const pb_prox* g, const pb_solve_opts* o, void* x, void* grad,
                       void* z, void* z_prev, pb_solve_result* out) {
  PersistPlan plan;
  if (!persist_plan<T>(ctx, f, ctx->persist_mode, &plan)) {
    pb_set_error("pb_solve: problem is not eligible for the persistent kernel");
    return PB_EUNSUPPORTED;
  }
  PB_CHECK_CUDA(cudaSetDevice(ctx->device));
  // workspace: [2][nchunk][m] chunk partials + [2][32][8] reduction partials + barrier counter + result
  const size_t part_bytes = (size_t)plan.ord.nchunk * f->m * sizeof(T);
  const size_t rglob_off = (part_bytes + 255) & ~(size_t)255;
  const size_t scal_off = (rglob_off + (size_t)f->m * sizeof(T) + 255) & ~(size_t)255;
  const size_t scal_bytes = (size_t)2 * PS_MAX_CTAS * PS_SCAL * sizeof(double);
  const size_t bar_off = scal_off + scal_bytes;
  const size_t res_off = bar_off + (size_t)PS_MAX_CTAS * PS_FLAG_STRIDE * sizeof(unsigned int);
  const size_t cyc_off = (res_off + sizeof(pb_solve_result) + 15) & ~(size_t)15;
  const size_t total = cyc_off + PS_NPHASE * sizeof(long long);
  int rc = pb_ensure_scratch(ctx, total);
  if (rc != PB_OK) return rc;
  unsigned char* ws = static_cast<unsigned char*>(ctx->scratch);
  PB_CHECK_CUDA(cudaMemsetAsync(ws + scal_off, 0, total - scal_off, ctx->stream));
  PersistParams p;
  memset(&p, 0, sizeof(p));
  p.A = f->A;
  p.b = f->b;
  p.m = f->m;
  p.n = f->n;
  p.lda = f->lda;
  p.x = x;
  p.grad = grad;
  p.z = z;
  p.z_prev = z_prev;
  p.prox_kind = g->kind;
  p.p0 = g->p0;
  p.p1 = g->p1;
  p.lo_v = g->kind == PB_PROX_BOX ? g->v0 : nullptr;
  p.hi_v = g->kind == PB_PROX_BOX ? g->v1 : nullptr;
  p.algorithm = o->algorithm;
  p.adaptive = o->adaptive;
  p.sequence = o->sequence;
  p.maxit = o->maxit;
  p.n_global = o->n_global > 0 ? o->n_global : n;
  p.tol = o->tol;
  p.gamma = o->gamma;
  p.mf = o->mf;
  p.constant_beta = o->constant_beta;
  p.minimum_gamma = o->minimum_gamma;
  p.reduce_gamma = o->reduce_gamma;
  p.increase_gamma = o->increase_gamma;
  p.ord = plan.ord;
  p.partial = ws;
  p.r_glob = ws + rglob_off;
  p.shared_xchg = plan.shared_xchg;
  p.a_in_smem = plan.a_in_smem;
  p.scal = reinterpret_cast<double*>(ws + scal_off);
  p.bar = reinterpret_cast<unsigned int*>(ws + bar_off);
  p.ns_max = plan.ns_max;
  p.timing = o->profile ? 1 : 0;
  p.cycles = reinterpret_cast<long long*>(ws + cyc_off);
  p.result = reinterpret_cast<pb_solve_result*>(ws + res_off);
  auto kern = k_persist_solve<T>;
  static bool attr_done[PB_MAX_DEVICES][2] = {};
  const int dev = ctx->device < PB_MAX_DEVICES ? ctx->device : 0;
  if (!attr_done[dev][sizeof(T) == 8] || ctx->device >= PB_MAX_DEVICES) {
    PB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done[dev][sizeof(T) == 8] = true;
  }
  cudaEvent_t ev[2] = {nullptr, nullptr};
  if (o->profile) {
    PB_CHECK_CUDA(cudaEventCreate(&ev[0]));
    PB_CHECK_CUDA(cudaEventCreate(&ev[1]));
    PB_CHECK_CUDA(cudaEventRecord(ev[0], ctx->stream));
  }
  void* args[] = {&p};
  PB_CHECK_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3(plan.ctas), dim3(PS_THREADS), args, plan.smem, ctx->stream));
  ctx->launches++;
  if (o->profile) PB_CHECK_CUDA(cudaEventRecord(ev[1], ctx->stream));
  pb_solve_result host;
  PB_CHECK_CUDA(cudaMemcpyAsync(&host, p.result, sizeof(host), cudaMemcpyDeviceToHost, ctx->stream));
  if (o->profile) PB_CHECK_CUDA(cudaMemcpyAsync(ctx->persist_cycles, p.cycles, PS_NPHASE * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
  PB_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = host;
  out->x = x;
  out->grad = grad;
  out->z = z;
  out->z_prev = z_prev;
  out->persistent_ctas = plan.ctas;
  if (o->profile) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev[0], ev[1]);
    out->loop_ms = ms;
    cudaEventDestroy(ev[0]);
    cudaEventDestroy(ev[1]);
  }
  return PB_OK;
}

int pb_persist_solve(pb_ctx* ctx, int dtype, int64_t n, const pb_smooth* f, const pb_prox* g, const pb_solve_opts* o, void* x,
                     void* grad, void* z, void* z_prev, pb_solve_result* out) {
  if (dtype == PB_F32) return persist_run<float>(ctx, n, f, g, o, x, grad, z, z_prev, out);
  return persist_run<double>(ctx, n, f, g, o, x, grad, z, z_prev, out);
}

// Diagnostic: clock64() cycles CTA 0 spent per phase in the last persistent solve run with opts->profile = 1
// (0 A*v partials, 1 grid barrier + scalar fold, 2 r = sum of partials - b and ||r||^2, 3 A'r, 4 fused step, 5 everything else).
extern "C" int pb_persist_phase_cycles(pb_ctx* ctx, int64_t* out8) {
  PB_REQUIRE(ctx != nullptr && out8 != nullptr, "null argument");
  for (int k = 0; k < PS_NPHASE; ++k) out8[k] = (int64_t)ctx->persist_cycles[k];
  return PB_OK;
}
