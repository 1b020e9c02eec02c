// common.cuh -- shared device helpers for libproxb200 (sm_100a).
//   * context definition
//   * double-double accumulation (error-free transformations) and the deterministic block -> grid reduction
//   * 16-byte packed loads/stores with optional streaming cache hints
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "proxb200.h"
#include "xchg.cuh"

#define PB_BLOCK 256            // threads per CTA of every streaming kernel
#define PB_MAX_CTAS 4096        // upper bound on grid size of reducing kernels (workspace rows)
#define PB_MAX_SUMS 4           // double-double sums a kernel may reduce
#define PB_MAX_MAXS 2           // max-reductions a kernel may reduce
#define PB_MAX_DEVICES 64       // device ordinals with their own per-device caches (function attributes, occupancy)

// Reduction workspace in device memory: a ticket counter and per-CTA partials laid out [quantity][cta].
struct PbWorkspace {
  unsigned int ticket;
  unsigned int pad[63];
  double sum_hi[PB_MAX_SUMS][PB_MAX_CTAS];
  double sum_lo[PB_MAX_SUMS][PB_MAX_CTAS];
  double mx[PB_MAX_MAXS][PB_MAX_CTAS];
  // deferred fold only: snapshot of the scalar block taken by the step kernel (slots other kernels of the iteration wrote, e.g.
  // the f value in AUX) and the private block the fold kernel completes and exchanges
  double snap[PB_NSCALARS];
  double blk[PB_NSCALARS];
};

struct pb_ctx {
  int device;
  cudaStream_t stream;
  bool owns_stream;
  int sm_count;
  size_t l2_bytes;
  int ctas_per_sm;     // 0 = per-kernel default
  int stream_hints;    // -1 auto, 0 off, 1 on
  int unroll;          // 0 = default
  int step_impl;       // 0 = default, 1 = register pipeline, 2 = TMA bulk ring
  int persist_mode;    // PB_OPT_PERSISTENT: 0 auto, -1 never, k > 0 at most k CTAs
  long long persist_cycles[8];   // per-phase clock64() totals of the last profiled persistent solve
  int lsq_fista;       // PB_OPT_LSQ_FISTA: 0 auto, -1 never, 1 always (when the shape allows)
  int lsq_fused;       // PB_OPT_LSQ_FUSED: 0 auto, -1 never, k > 0 always with k blocks kept between the two sweeps
  int gemv_scalar;     // PB_OPT_GEMV_SCALAR: 1 = thread-per-row residual kernel (4-byte loads) instead of 16-byte row packs
  int multi_mode;      // PB_OPT_MULTI_ITER: 0 auto, -1 never, 1 force (even when contexts share a device)
  void* multi_ws;      // workspace of the persistent multi-iteration step kernel (step_multi.cu)
  double* chain_dev;   // PB_NSCALARS doubles: rank-combined dots of a device-side chain (sharded L-BFGS recursion, qn_kernels.cu)
  double* scalars_dev;   // active scalar block (own or caller supplied)
  double* scalars_own;
  double* scalars_host;  // pinned mirror
  double* ahead_host;    // pinned, 2 x PB_NSCALARS: scalar blocks of two iterations in flight (look-ahead loops without the exchange)
  cudaEvent_t ahead_ev[2];
  PbWorkspace* ws;
  int64_t launches;
  // scratch for the dense / block-diagonal products (partial sums of column chunks)
  void* scratch;
  size_t scratch_bytes;
  // cached device buffers of pb_ffb_step_host
  void* hbuf[5];
  size_t hbuf_bytes;
  // C1 exchange (xchg.cu): own IPC-exported buffer, peer mappings, mapped pinned landing zone for the host
  unsigned long long* xchg_own;
  unsigned long long* xchg_peer[PB_MAX_RANKS];
  unsigned long long* xchg_host_words;      // host pointer (pinned, mapped): [parity][PB_MAX_RANKS][32] words
  unsigned long long* xchg_host_words_dev;  // its device alias
  unsigned int xchg_seq;                    // last sequence number issued
  unsigned int xchg_vseq;                   // last sequence number of the VECTOR exchange (C2, lsq_kernels.cu)
  int xchg_rank, xchg_world;                // world == 0: not initialised
  int xchg_connected;
  int xchg_local;                           // peers are contexts of this process (pb_xchg_connect_local): nothing to cudaIpcClose
  int xchg_shared_device;                   // some peer context lives on the same GPU: kernels that need the whole GPU to be
                                            // co-resident while they wait for a peer (step_multi.cu) are not used
  int xchg_fused;                           // K1/K2 push in-kernel
  int xchg_pending;                         // a launched kernel will publish xchg_seq
  int64_t xchg_pending_launch;              // value of `launches` right after that kernel's launch
  // deferred fold (pipelined driver loop, solve.cu): the fused step only writes its per-CTA partials; a 1-CTA kernel on a
  // side stream folds them, writes the scalar block and performs the exchange WHILE the next step kernel already runs
  int defer_fold;                           // 1: pb_fb_step / pb_ffb_step use the split form when the kernel supports it
  PbWorkspace* ws_defer[2];                 // partial workspaces, alternated by step parity
  cudaStream_t side_stream;
  cudaEvent_t ev_main[2], ev_fold[2];
  unsigned int defer_count;                 // steps issued in deferred form since defer_fold was switched on
  int last_step_grid;                       // grid size of the most recent k_step launch (the fold kernel needs it)
};

// Fill the kernel-side parameters for the next in-kernel exchange (advances the sequence number); world = 0 if disabled.
void pb_xchg_next(pb_ctx* ctx, XchgParams* xp, bool want);
void pb_xchg_vec_next(pb_ctx* ctx, XchgVecParams* xv);
// Wait for one specific exchange (by sequence number) in the pinned landing zone; rows_out: world x PB_NSCALARS doubles.
int pb_xchg_wait_seq(pb_ctx* ctx, unsigned int seq, double* rows_out, double timeout_s);
int pb_step_defer(pb_ctx* ctx, int on);

// Make the context's device current for the calling thread for the lifetime of the guard (a process that drives several GPUs has
// one host thread per context; a fresh thread starts on device 0).  Restores the previous device on exit.
struct PbDeviceGuard {
  int prev;
  bool switched;
  explicit PbDeviceGuard(const pb_ctx* ctx) : prev(-1), switched(false) {
    if (ctx && cudaGetDevice(&prev) == cudaSuccess && prev != ctx->device) switched = cudaSetDevice(ctx->device) == cudaSuccess;
  }
  ~PbDeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};

void pb_set_error(const char* fmt, ...);
int pb_ensure_scratch(pb_ctx* ctx, size_t bytes);

#define PB_CHECK_CUDA(call)                                                                    \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      pb_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));      \
      return PB_ECUDA;                                                                         \
    }                                                                                          \
  } while (0)

#define PB_REQUIRE(cond, msg)                                  \
  do {                                                         \
    if (!(cond)) {                                             \
      pb_set_error("%s: %s", __func__, msg);                   \
      return PB_EINVAL;                                        \
    }                                                          \
  } while (0)

#define PB_LAUNCH_CHECK(ctx)                                   \
  do {                                                         \
    (ctx)->launches++;                                         \
    PB_CHECK_CUDA(cudaGetLastError());                         \
  } while (0)

// ---------------------------------------------------------------------------------------------------------------
// rounding-exact scalar arithmetic: Julia broadcasts `x .- gamma .* g` round the product and the difference
// separately (no muladd in the reference), so the intrinsics below are used to forbid FMA contraction.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }

// ---------------------------------------------------------------------------------------------------------------
// double-double accumulator
// ---------------------------------------------------------------------------------------------------------------
struct dd {
  double hi, lo;
};

// a += v  (Knuth two-sum into hi, error into lo)
__device__ __forceinline__ void dd_add(dd& a, double v) {
  double s = __dadd_rn(a.hi, v);
  double bb = __dsub_rn(s, a.hi);
  double e = __dadd_rn(__dsub_rn(a.hi, __dsub_rn(s, bb)), __dsub_rn(v, bb));
  a.hi = s;
  a.lo = __dadd_rn(a.lo, e);
}
// a += p*q exactly (two-prod via FMA)
__device__ __forceinline__ void dd_add_prod(dd& a, double p, double q) {
  double pr = __dmul_rn(p, q);
  double er = __fma_rn(p, q, -pr);
  dd_add(a, pr);
  a.lo = __dadd_rn(a.lo, er);
}
// a + b for two double-doubles, renormalised
__device__ __forceinline__ dd dd_sum(const dd& a, const dd& b) {
  double s = __dadd_rn(a.hi, b.hi);
  double bb = __dsub_rn(s, a.hi);
  double e = __dadd_rn(__dsub_rn(a.hi, __dsub_rn(s, bb)), __dsub_rn(b.hi, bb));
  e = __dadd_rn(e, __dadd_rn(a.lo, b.lo));
  dd r;
  r.hi = __dadd_rn(s, e);                       // fast two-sum renormalisation
  r.lo = __dsub_rn(e, __dsub_rn(r.hi, s));
  return r;
}

// NaN-propagating max of non-negative values (Julia's norm(., Inf) returns NaN if any entry is NaN)
__device__ __forceinline__ double nanmax(double m, double a) { return (a > m || a != a) ? a : m; }

template <int NSUM, int NMAX>
struct Acc {
  dd s[NSUM > 0 ? NSUM : 1];
  double m[NMAX > 0 ? NMAX : 1];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int k = 0; k < (NSUM > 0 ? NSUM : 1); ++k) s[k].hi = s[k].lo = 0.0;
#pragma unroll
    for (int k = 0; k < (NMAX > 0 ? NMAX : 1); ++k) m[k] = 0.0;
  }
};

struct OutMap {
  int sum_slot[PB_MAX_SUMS];  // index into the scalar block of the hi word (lo = +1); -1 = discard
  int max_slot[PB_MAX_MAXS];
};

__device__ __forceinline__ double shfl_down_d(double v, int off) { return __shfl_down_sync(0xffffffffu, v, off); }

template <int NSUM, int NMAX>
__device__ __forceinline__ void warp_reduce(Acc<NSUM, NMAX>& a) {
#pragma unroll
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

// Reduce the per-thread accumulators of one CTA; result valid in thread 0.  BLOCK must be a multiple of 32, <= 1024.
template <int NSUM, int NMAX, int BLOCK>
__device__ __forceinline__ void block_reduce(Acc<NSUM, NMAX>& a) {
  constexpr int NW = BLOCK / 32;
  __shared__ double sh_hi[NSUM > 0 ? NSUM : 1][NW];
  __shared__ double sh_lo[NSUM > 0 ? NSUM : 1][NW];
  __shared__ double sh_mx[NMAX > 0 ? NMAX : 1][NW];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  warp_reduce<NSUM, NMAX>(a);
  __syncthreads();  // protects reuse of the shared arrays when called twice
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NSUM; ++k) {
      sh_hi[k][warp] = a.s[k].hi;
      sh_lo[k][warp] = a.s[k].lo;
    }
#pragma unroll
    for (int k = 0; k < NMAX; ++k) sh_mx[k][warp] = a.m[k];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < NSUM; ++k) {
      a.s[k].hi = lane < NW ? sh_hi[k][lane] : 0.0;
      a.s[k].lo = lane < NW ? sh_lo[k][lane] : 0.0;
    }
#pragma unroll
    for (int k = 0; k < NMAX; ++k) a.m[k] = lane < NW ? sh_mx[k][lane] : 0.0;
    warp_reduce<NSUM, NMAX>(a);
  }
}

// CTA partial -> workspace; the last CTA to arrive folds all partials in a fixed order (deterministic for a given
// grid, and -- thanks to the double-double arithmetic -- equal after rounding for any grid) and writes the scalar block.
// Fold the `nctas` per-CTA partials of `ws` in a fixed order, write the scalar block and (optionally) exchange it.  Executed by
// ONE CTA: the last CTA of the producing kernel, or the stand-alone fold kernel of the deferred form.
template <int NSUM, int NMAX, int BLOCK>
__device__ __forceinline__ void fold_partials(Acc<NSUM, NMAX>& a, PbWorkspace* ws, unsigned int nctas, double* out,
                                              const OutMap& map, const XchgParams* xp, bool reset_ticket) {
  a.clear();
  for (unsigned int c = threadIdx.x; c < nctas; c += BLOCK) {
#pragma unroll
    for (int k = 0; k < NSUM; ++k) {
      dd o;
      o.hi = __ldcg(&ws->sum_hi[k][c]);
      o.lo = __ldcg(&ws->sum_lo[k][c]);
      a.s[k] = dd_sum(a.s[k], o);
    }
#pragma unroll
    for (int k = 0; k < NMAX; ++k) a.m[k] = nanmax(a.m[k], __ldcg(&ws->mx[k][c]));
  }
  block_reduce<NSUM, NMAX, BLOCK>(a);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NSUM; ++k)
      if (map.sum_slot[k] >= 0) {
        out[map.sum_slot[k]] = a.s[k].hi;
        out[map.sum_slot[k] + 1] = a.s[k].lo;
      }
#pragma unroll
    for (int k = 0; k < NMAX; ++k)
      if (map.max_slot[k] >= 0) out[map.max_slot[k]] = a.m[k];
    if (reset_ticket) ws->ticket = 0;  // ready for the next launch on this stream
  }
  // fused C1: the CTA that produced the final scalars also exchanges them with the peers and hands them to the host -- or, for a
  // device-side chain (gather_out), folds the ranks' rows itself so that the next kernel finds the global sums
  if (xp != nullptr && xp->world > 0) {
    __syncthreads();
    if (xp->gather_out != nullptr) {
      __shared__ double rows_sh[PB_MAX_RANKS * PB_NSCALARS];
      const bool ok = xchg_push_gather(*xp, out, rows_sh);
      if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NSUM; ++k)
          if (map.sum_slot[k] >= 0) {
            double hi = 0.0, lo = 0.0;
            for (int r = 0; r < xp->world; ++r) {            // host-identical double-double fold in rank order (solve.cu: fold)
              const double bh = rows_sh[r * PB_NSCALARS + map.sum_slot[k]], bl = rows_sh[r * PB_NSCALARS + map.sum_slot[k] + 1];
              const double s = __dadd_rn(hi, bh);
              const double bb = __dsub_rn(s, hi);
              double e = __dadd_rn(__dsub_rn(hi, __dsub_rn(s, bb)), __dsub_rn(bh, bb));
              e = __dadd_rn(e, __dadd_rn(lo, bl));
              const double h = __dadd_rn(s, e);
              lo = __dsub_rn(e, __dsub_rn(h, s));
              hi = h;
            }
            xp->gather_out[map.sum_slot[k]] = ok ? hi : __longlong_as_double(0x7ff8000000000000ll);
            xp->gather_out[map.sum_slot[k] + 1] = lo;
          }
      }
    } else {
      xchg_push_wait(*xp, out);
    }
  }
}

template <int NSUM, int NMAX, int BLOCK>
__device__ __forceinline__ void grid_reduce(Acc<NSUM, NMAX>& a, PbWorkspace* ws, double* out, const OutMap& map,
                                            const XchgParams* xp = nullptr, bool defer = false) {
  __shared__ bool is_last;
  block_reduce<NSUM, NMAX, BLOCK>(a);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NSUM; ++k) {
      ws->sum_hi[k][blockIdx.x] = a.s[k].hi;
      ws->sum_lo[k][blockIdx.x] = a.s[k].lo;
    }
#pragma unroll
    for (int k = 0; k < NMAX; ++k) ws->mx[k][blockIdx.x] = a.m[k];
    if (!defer) {
      __threadfence();
      unsigned int t = atomicAdd(&ws->ticket, 1u);
      is_last = (t == gridDim.x - 1);
    }
  }
  if (defer) return;                    // deferred form: a separate 1-CTA kernel folds (kernel boundary = visibility)
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  fold_partials<NSUM, NMAX, BLOCK>(a, ws, gridDim.x, out, map, xp, true);
}

// ---------------------------------------------------------------------------------------------------------------
// packed 16-byte global memory access
// ---------------------------------------------------------------------------------------------------------------
template <typename T, int VEC>
struct alignas(sizeof(T) * VEC) Pack {
  T v[VEC];
};

template <typename T, int VEC, bool HINT>
__device__ __forceinline__ Pack<T, VEC> ld_pack(const T* p) {
  Pack<T, VEC> r;
  if constexpr (sizeof(T) * VEC == 16) {
    float4 t = HINT ? __ldcs(reinterpret_cast<const float4*>(p)) : __ldg(reinterpret_cast<const float4*>(p));
    *reinterpret_cast<float4*>(&r) = t;
  } else {
    static_assert(VEC == 1, "only 16-byte packs or scalars");
    r.v[0] = HINT ? __ldcs(p) : __ldg(p);
  }
  return r;
}

template <typename T, int VEC, bool HINT>
__device__ __forceinline__ void st_pack(T* p, const Pack<T, VEC>& r) {
  if constexpr (sizeof(T) * VEC == 16) {
    if (HINT)
      __stcs(reinterpret_cast<float4*>(p), *reinterpret_cast<const float4*>(&r));
    else
      *reinterpret_cast<float4*>(p) = *reinterpret_cast<const float4*>(&r);
  } else {
    if (HINT)
      __stcs(p, r.v[0]);
    else
      *p = r.v[0];
  }
}

// host-side single rounding product in the element type (gamma*lambda is formed in R by the reference)
static inline float mul_rn_host(float a, float b) {
  volatile float r = a * b;
  return r;
}
static inline double mul_rn_host(double a, double b) {
  volatile double r = a * b;
  return r;
}

static inline bool pb_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// default grid for a grid-stride streaming kernel
static inline int pb_stream_grid(const pb_ctx* ctx, int64_t work_items_per_cta, int64_t n, int default_ctas_per_sm) {
  int per_sm = ctx->ctas_per_sm > 0 ? ctx->ctas_per_sm : default_ctas_per_sm;
  int64_t g = (int64_t)ctx->sm_count * per_sm;
  int64_t need = (n + work_items_per_cta - 1) / work_items_per_cta;
  if (need < 1) need = 1;
  if (g > need) g = need;
  if (g > PB_MAX_CTAS) g = PB_MAX_CTAS;
  return (int)g;
}
