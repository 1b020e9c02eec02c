// step_multi.cu -- fixed-stepsize FastForwardBackward with an ELEMENT-WISE gradient source (f = <c, .>: gradient supplied as a
// buffer; f = SquaredDistance(b): grad = x - b) as ONE persistent kernel that loops over the iterations itself.
//
// Why: with such an f the whole iteration (fast_forward_backward.jl:130-142) is element-wise, so a CTA's slice of iteration k+1
// depends only on its own slice of iteration k.  One launch per iteration pays a launch ramp and a tail every ~40 us at n = 1e8 on
// 8 GPUs (SCALE_r01: 50.8 us per launch against 38.8 us of pure streaming, efficiency 0.69).  Here the streaming CTAs never stop:
// each walks its slice iteration after iteration; only what the driver loop (src/ProximalAlgorithms.jl:114-123) really needs per
// iteration leaves the slice -- the reduction partials.
//
// Roles.  CTAs 1..G-1 stream.  CTA 0 is the SERVICE CTA: for every iteration it waits until all streaming CTAs have delivered
// their partials (ticket), folds them in a fixed order (double-double), exchanges the scalar block with the other GPUs inside the
// kernel (xchg.cuh: NVLink peer stores, sequence-numbered), folds the ranks in rank order, evaluates the stop test
// `norm(res, Inf)/gamma <= tol` in R and publishes the decision.  Every iteration of the HBM pass still reads x, grad (or b), z_prev
// and writes z, x_next exactly once (20 B/element in Float32): nothing is fused ACROSS iterations.
//
// Look-ahead and the discard rule (as csrc/solve.cu: run_ffb_pipelined).  x and z live in rings of three buffers: iteration k reads
// X[k%3], Z[(k-1)%3] and writes Z[k%3], X[(k+1)%3].  A streaming CTA starts iteration k only when the decision on iteration k-2 is
// known and was "continue": iteration k-1 may then still be undecided (one speculative iteration), and if its stop test fires the
// state of iteration k-1 -- X[(k-1)%3], Z[(k-1)%3], Z[(k-2)%3] -- has not been touched by iteration k, which wrote Z[k%3] and
// X[(k+1)%3] = X[(k-2)%3], a buffer state k-1 no longer needs.  Results (iterates, scalars, iteration count) are bit-identical to
// the one-launch-per-iteration loop (tests/test_gpu_multi.py).
#include <string.h>

#include "solve_scalar.h"
#include "step_common.cuh"

#define SM_BLOCK 256
#define SM_UNROLL 4
#define SM_RING 4                      // iterations of reduction partials kept (a CTA is at most 2 iterations ahead of the service CTA)
#define SM_MAX_CTAS 1024

struct MultiWs {
  unsigned long long done_iter;        // last iteration whose stop test has been evaluated (monotone)
  unsigned long long stop_iter;        // 0, or the iteration at which the loop ends (stop test fired or maxit reached)
  unsigned int error;                  // 1: a peer did not publish within the device time-out
  unsigned int pad0[27];
  unsigned int ticket[SM_RING][32];    // [k % SM_RING][0]: streaming CTAs that have delivered iteration k (128 bytes apart)
  double sum_hi[SM_RING][3][SM_MAX_CTAS];
  double sum_lo[SM_RING][3][SM_MAX_CTAS];
  double mx[SM_RING][SM_MAX_CTAS];
  double final_block[PB_NSCALARS];     // rank-combined scalars of the final iteration: gsum, res_sq, gdr (rounded), res_inf
  long long iterations;
};

struct MultiParams {
  void* X[3];
  void* Z[3];
  const void* c;                       // LINEAR: the gradient;  SQDIST: b
  int fkind;
  int64_t n;
  int prox_kind;
  double p0, p1;
  const void* lo_v;
  const void* hi_v;
  int sequence;
  double gamma, mf, constant_beta, tol;
  int64_t maxit;
  int hint;                            // streaming cache hints (working set exceeds L2)
  MultiWs* ws;
  XchgParams xchg;                     // world == 0: single GPU without an attached exchange
  unsigned int seq0;                   // iteration k is exchange number seq_of(seq0, k)
};

__host__ __device__ __forceinline__ unsigned int sm_seq_of(unsigned int seq0, long long k) {
  // sequence numbers skip the reserved values 0 and PB_XCHG_ERROR_SEQ
  return (unsigned int)(((unsigned long long)seq0 + (unsigned long long)k - 1ull) % 0xFFFFFFFEull) + 1u;
}

__device__ __forceinline__ unsigned long long ld_acq_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_rel_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acq_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// host-identical double-double fold of the ranks' (hi, lo) pairs (solve.cu: fold / dd_add), thread 0 only
__device__ __forceinline__ void sm_dd_add(double& hi, double& lo, double bh, double bl) {
  const double s = __dadd_rn(hi, bh);
  const double bb = __dsub_rn(s, hi);
  double e = __dadd_rn(__dsub_rn(hi, __dsub_rn(s, bb)), __dsub_rn(bh, bb));
  e = __dadd_rn(e, __dadd_rn(lo, bl));
  const double h = __dadd_rn(s, e);
  lo = __dsub_rn(e, __dsub_rn(h, s));
  hi = h;
}

template <typename T, int PROX, bool SQDIST, bool HINT>
__device__ __forceinline__ void sm_stream_iteration(const MultiParams& p, long long k, T gamma, T beta, int sc, int nstream,
                                                    Acc<3, 1>& acc) {
  constexpr bool COMP = sizeof(T) == 8;
  constexpr int VEC = 16 / sizeof(T);
  const T* __restrict__ x = static_cast<const T*>(p.X[k % 3]);
  const T* __restrict__ zp = static_cast<const T*>(p.Z[(k + 2) % 3]);
  T* __restrict__ zo = static_cast<T*>(p.Z[k % 3]);
  T* __restrict__ xo = static_cast<T*>(p.X[(k + 1) % 3]);
  const T* __restrict__ cv = static_cast<const T*>(p.c);
  const T* __restrict__ lov = static_cast<const T*>(p.lo_v);
  const T* __restrict__ hiv = static_cast<const T*>(p.hi_v);
  T pa = T(0), pb = T(0);
  if (PROX == PB_PROX_L1) pa = mul_rn(gamma, (T)p.p0);
  if (PROX == PB_PROX_BOX) {
    pa = (T)p.p0;
    pb = (T)p.p1;
  }
  Acc<3, 1> pk;
  pk.clear();
  auto do_pack = [&](int64_t i, const Pack<T, VEC>& xq, const Pack<T, VEC>& cq, const Pack<T, VEC>& zq) {
    Pack<T, VEC> lo, hi, zn, xn;
    if (PROX == PB_PROX_BOX && lov) lo = ld_pack<T, VEC, false>(lov + i);
    if (PROX == PB_PROX_BOX && hiv) hi = ld_pack<T, VEC, false>(hiv + i);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const T l = (PROX == PB_PROX_BOX && lov) ? lo.v[e] : pa;
      const T h = (PROX == PB_PROX_BOX && hiv) ? hi.v[e] : pb;
      const T g = SQDIST ? sub_rn(xq.v[e], cq.v[e]) : cq.v[e];      // SquaredDistance: grad = x - b (benchmark/benchmarks.jl:25-28)
      T yv, rv;
      StepElem<T, PROX, true>::template run<COMP>(xq.v[e], g, zq.v[e], l, h, gamma, beta, yv, zn.v[e], rv, xn.v[e], COMP ? acc : pk);
    }
    if constexpr (!COMP) fold_pack<PROX>(acc, pk);
    st_pack<T, VEC, HINT>(zo + i, zn);
    st_pack<T, VEC, HINT>(xo + i, xn);
  };
  // The balanced schedule of k_step (step_kernels.cu), with a STATIC tile -> CTA map so that a CTA meets the same elements in every
  // iteration: `rounds` full rounds in which every streaming CTA takes one tile (SM_UNROLL packs per thread in flight), then the
  // remaining < nstream tiles shared at pack granularity, then the < VEC ragged elements.
  constexpr int64_t TILE = (int64_t)SM_BLOCK * VEC * SM_UNROLL;
  const int64_t ntiles = p.n / TILE, rounds = ntiles / nstream;
  for (int64_t r = 0; r < rounds; ++r) {
    const int64_t base = (r * nstream + sc) * TILE + (int64_t)threadIdx.x * VEC;
    Pack<T, VEC> xv[SM_UNROLL], cq[SM_UNROLL], zv[SM_UNROLL];
#pragma unroll
    for (int u = 0; u < SM_UNROLL; ++u) {
      const int64_t i = base + (int64_t)u * SM_BLOCK * VEC;
      xv[u] = ld_pack<T, VEC, HINT>(x + i);
      cq[u] = ld_pack<T, VEC, HINT>(cv + i);
      zv[u] = ld_pack<T, VEC, HINT>(zp + i);
    }
#pragma unroll
    for (int u = 0; u < SM_UNROLL; ++u) do_pack(base + (int64_t)u * SM_BLOCK * VEC, xv[u], cq[u], zv[u]);
  }
  const int64_t rem_start = rounds * nstream * TILE;
  const int64_t rem_packs = (p.n - rem_start) / VEC;
  for (int64_t q = (int64_t)sc * SM_BLOCK + threadIdx.x; q < rem_packs; q += (int64_t)nstream * SM_BLOCK) {
    const int64_t i = rem_start + q * VEC;
    const Pack<T, VEC> xq = ld_pack<T, VEC, HINT>(x + i), cq = ld_pack<T, VEC, HINT>(cv + i), zq = ld_pack<T, VEC, HINT>(zp + i);
    do_pack(i, xq, cq, zq);
  }
  for (int64_t i = rem_start + rem_packs * VEC + (int64_t)sc * SM_BLOCK + threadIdx.x; i < p.n; i += (int64_t)nstream * SM_BLOCK) {
    const T l = (PROX == PB_PROX_BOX && lov) ? lov[i] : pa;       // the last < VEC elements: single-element groups, like k_step
    const T h = (PROX == PB_PROX_BOX && hiv) ? hiv[i] : pb;
    const T g = SQDIST ? sub_rn(x[i], cv[i]) : cv[i];
    T yv, zn, rv, xn;
    StepElem<T, PROX, true>::template run<COMP>(x[i], g, zp[i], l, h, gamma, beta, yv, zn, rv, xn, COMP ? acc : pk);
    if constexpr (!COMP) fold_pack<PROX>(acc, pk);
    zo[i] = zn;
    xo[i] = xn;
  }
}

template <typename T, int PROX, bool SQDIST, bool HINT>
__global__ void __launch_bounds__(SM_BLOCK, 2) k_step_multi(MultiParams p) {
  typedef T R;
  __shared__ double rows_sh[PB_MAX_RANKS * PB_NSCALARS];
  __shared__ double blk_sh[PB_NSCALARS];
  __shared__ unsigned long long ctl[2];            // [0] stop_iter seen by this CTA, [1] unused
  __shared__ double beta_sh;
  MultiWs* ws = p.ws;
  const int nstream = (int)gridDim.x - 1;
  const int tid = threadIdx.x;
  const R gamma = (R)p.gamma;
  Nesterov<R> seq;
  seq.init(p.sequence, (R)p.mf, (R)p.constant_beta);

  if (blockIdx.x == 0) {
    // ---------------- service CTA ----------------
    for (long long k = 1;; ++k) {
      const int slot = (int)(k % SM_RING);
      if (tid == 0) {
        while (ld_acq_u32(&ws->ticket[slot][0]) < (unsigned int)nstream) {
        }
        ws->ticket[slot][0] = 0;                   // next used by iteration k + SM_RING, which cannot start before k + 2 is decided
      }
      __syncthreads();
      Acc<3, 1> a;
      OutMap map;
      map.sum_slot[0] = PB_S_GSUM;
      map.sum_slot[1] = PB_S_RESSQ;
      map.sum_slot[2] = PB_S_GDR;
      map.sum_slot[3] = -1;
      map.max_slot[0] = PB_S_RESINF;
      map.max_slot[1] = -1;
      a.clear();
      for (int c = tid; c < nstream; c += SM_BLOCK) {
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          dd o;
          o.hi = __ldcg(&ws->sum_hi[slot][q][c]);
          o.lo = __ldcg(&ws->sum_lo[slot][q][c]);
          a.s[q] = dd_sum(a.s[q], o);
        }
        a.m[0] = nanmax(a.m[0], __ldcg(&ws->mx[slot][c]));
      }
      block_reduce<3, 1, SM_BLOCK>(a);
      if (tid < PB_NSCALARS) blk_sh[tid] = 0.0;
      __syncthreads();
      if (tid == 0) {
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          blk_sh[map.sum_slot[q]] = a.s[q].hi;
          blk_sh[map.sum_slot[q] + 1] = a.s[q].lo;
        }
        blk_sh[PB_S_RESINF] = a.m[0];
      }
      __syncthreads();
      int world = 1;
      bool failed = false;
      if (p.xchg.world > 1) {                      // the per-iteration exchange, inside the kernel
        world = p.xchg.world;
        XchgParams xp = p.xchg;
        xp.seq = sm_seq_of(p.seq0, k);
        failed = !xchg_push_gather(xp, blk_sh, rows_sh);
      } else {
        if (tid < PB_NSCALARS) rows_sh[tid] = blk_sh[tid];
      }
      __syncthreads();
      if (tid == 0) {
        double f_[3];
        for (int q = 0; q < 3; ++q) {              // ranks folded in rank order, exactly like the host (solve.cu: fold)
          double hi = 0.0, lo = 0.0;
          for (int r = 0; r < world; ++r) sm_dd_add(hi, lo, rows_sh[r * PB_NSCALARS + 2 * q], rows_sh[r * PB_NSCALARS + 2 * q + 1]);
          f_[q] = hi + lo;
        }
        double mxv = 0.0;
        for (int r = 0; r < world; ++r) {
          const double v = rows_sh[r * PB_NSCALARS + PB_S_RESINF];
          if (v != v) {
            mxv = v;
            break;
          }
          if (v > mxv) mxv = v;
        }
        const R rn = (R)mxv;
        const bool stop = failed || k >= p.maxit || (double)(rn / gamma) <= p.tol;     // src/ProximalAlgorithms.jl:117
        if (stop) {
          ws->final_block[0] = f_[0];
          ws->final_block[1] = f_[1];
          ws->final_block[2] = f_[2];
          ws->final_block[3] = mxv;
          ws->iterations = k;
          if (failed) ws->error = 1;
          __threadfence();
          st_rel_u64(&ws->stop_iter, (unsigned long long)k);
        }
        st_rel_u64(&ws->done_iter, (unsigned long long)k);
        ctl[0] = stop ? 1ull : 0ull;
      }
      __syncthreads();
      if (ctl[0]) return;
      __syncthreads();
    }
  }

  // ---------------- streaming CTAs ----------------
  const int sc = (int)blockIdx.x - 1;
  for (long long k = 1; k <= p.maxit; ++k) {
    if (tid == 0) {
      unsigned long long stop_at = 0;
      if (k >= 3) {
        while (ld_acq_u64(&ws->done_iter) < (unsigned long long)(k - 2)) {
        }
        stop_at = ld_acq_u64(&ws->stop_iter);
      }
      ctl[0] = stop_at;
      beta_sh = (double)seq.next(gamma);           // beta of the extrapolation this pass fuses (fast_forward_backward.jl:134)
    }
    __syncthreads();
    const unsigned long long stop_at = ctl[0];
    const T beta = (T)beta_sh;
    if (stop_at != 0) return;                      // the loop ended at iteration stop_at <= k - 2
    Acc<3, 1> acc;
    acc.clear();
    sm_stream_iteration<T, PROX, SQDIST, HINT>(p, k, gamma, beta, sc, nstream, acc);
    block_reduce<3, 1, SM_BLOCK>(acc);
    if (tid == 0) {
      const int slot = (int)(k % SM_RING);
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        __stcg(&ws->sum_hi[slot][q][sc], acc.s[q].hi);
        __stcg(&ws->sum_lo[slot][q][sc], acc.s[q].lo);
      }
      __stcg(&ws->mx[slot][sc], acc.m[0]);
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(&ws->ticket[slot][0]) : "memory");
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
// The workspace is allocated when the context is created, never inside a solve: cudaMalloc may wait for the device to go idle, and on
// a device shared by several contexts a peer's persistent kernel may already be spinning there, waiting for THIS rank's kernel.
size_t pb_multi_ws_bytes() { return sizeof(MultiWs); }

bool pb_multi_eligible(const pb_ctx* ctx, int dtype, int64_t n, const pb_smooth* f, const pb_prox* g, const pb_solve_opts* o, const void* x_next) {
  (void)dtype;
  if (ctx->multi_mode < 0) return false;
  if (ctx->xchg_shared_device && ctx->multi_mode <= 0) return false;   // contexts sharing a GPU cannot all be co-resident
  if (o->algorithm != PB_ALG_FFB || o->adaptive || o->gamma <= 0) return false;
  if (f->kind != PB_F_LINEAR && f->kind != PB_F_SQDIST) return false;
  if (g->kind != PB_PROX_ZERO && g->kind != PB_PROX_L1 && g->kind != PB_PROX_BOX) return false;
  if (!o->spare_x || !o->spare_z || !x_next || n <= 0) return false;
  if (ctx->xchg_world > 1 && !ctx->xchg_connected) return false;
  return true;
}

template <typename T, int PROX, bool SQDIST>
static int multi_launch(pb_ctx* ctx, const MultiParams& p, int grid, bool hint) {
  void* args[] = {(void*)&p};
  if (hint)
    PB_CHECK_CUDA(cudaLaunchCooperativeKernel((const void*)k_step_multi<T, PROX, SQDIST, true>, dim3(grid), dim3(SM_BLOCK), args, 0, ctx->stream));
  else
    PB_CHECK_CUDA(cudaLaunchCooperativeKernel((const void*)k_step_multi<T, PROX, SQDIST, false>, dim3(grid), dim3(SM_BLOCK), args, 0, ctx->stream));
  ctx->launches++;
  return PB_OK;
}

template <typename T, bool SQDIST>
static int multi_launch_prox(pb_ctx* ctx, const MultiParams& p, int grid, bool hint) {
  switch (p.prox_kind) {
    case PB_PROX_L1: return multi_launch<T, PB_PROX_L1, SQDIST>(ctx, p, grid, hint);
    case PB_PROX_BOX: return multi_launch<T, PB_PROX_BOX, SQDIST>(ctx, p, grid, hint);
    default: return multi_launch<T, PB_PROX_ZERO, SQDIST>(ctx, p, grid, hint);
  }
}

// Runs iterations 1 .. k_final of the driver loop (the init step is iteration 1).  On entry x holds copy(x0) and z_prev = copy(x);
// on return *k_out is the iteration the loop ended at, idx_* tell which ring buffers hold its state, comb[4] = rank-combined
// gsum, res_sq, gdr (rounded from double-double) and res_inf.
int pb_multi_run(pb_ctx* ctx, int dtype, int64_t n, const pb_smooth* f, const pb_prox* g, const pb_solve_opts* o, void* const X[3],
                 void* const Z[3], int64_t* k_out, double comb[4], float* kernel_ms) {
  PB_CHECK_CUDA(cudaSetDevice(ctx->device));
  PB_REQUIRE(ctx->multi_ws != nullptr, "context has no multi-iteration workspace");
  PB_CHECK_CUDA(cudaMemsetAsync(ctx->multi_ws, 0, sizeof(MultiWs), ctx->stream));
  MultiParams p;
  memset(&p, 0, sizeof(p));
  for (int q = 0; q < 3; ++q) {
    p.X[q] = X[q];
    p.Z[q] = Z[q];
  }
  p.c = f->b;
  p.fkind = f->kind;
  p.n = n;
  p.prox_kind = g->kind;
  p.p0 = g->p0;
  p.p1 = g->p1;
  p.lo_v = g->kind == PB_PROX_BOX ? g->v0 : nullptr;
  p.hi_v = g->kind == PB_PROX_BOX ? g->v1 : nullptr;
  p.sequence = o->sequence;
  p.gamma = dtype == PB_F32 ? (double)(float)o->gamma : o->gamma;
  p.mf = o->mf;
  p.constant_beta = o->constant_beta;
  p.tol = o->tol;
  p.maxit = o->maxit;
  p.ws = static_cast<MultiWs*>(ctx->multi_ws);
  const size_t elt = dtype == PB_F32 ? 4 : 8;
  const bool hint = ctx->stream_hints >= 0 ? ctx->stream_hints != 0 : (size_t)n * elt * 5 > ctx->l2_bytes;
  bool vec_ok = pb_aligned16(f->b) && (!p.lo_v || pb_aligned16(p.lo_v)) && (!p.hi_v || pb_aligned16(p.hi_v));
  for (int q = 0; q < 3; ++q) vec_ok = vec_ok && pb_aligned16(X[q]) && pb_aligned16(Z[q]);
  if (!vec_ok) {
    pb_set_error("pb_multi_run: vectors must be 16-byte aligned");
    return PB_EUNSUPPORTED;
  }
  unsigned int last_seq = ctx->xchg_seq;
  if (ctx->xchg_world > 1) {
    XchgParams xp;
    memset(&xp, 0, sizeof(xp));
    for (int r = 0; r < ctx->xchg_world; ++r) xp.peer[r] = ctx->xchg_peer[r];
    xp.host_words = ctx->xchg_host_words_dev;
    xp.rank = ctx->xchg_rank;
    xp.world = ctx->xchg_world;
    p.xchg = xp;
    p.seq0 = ctx->xchg_seq + 1;
    if (p.seq0 == 0 || p.seq0 == PB_XCHG_ERROR_SEQ) p.seq0 = 1;
  }
  // grid: one service CTA + streaming CTAs, all co-resident (cooperative launch)
  int occ = 0;
  {
    cudaError_t e = dtype == PB_F32 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_step_multi<float, PB_PROX_L1, false, true>, SM_BLOCK, 0)
                                    : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_step_multi<double, PB_PROX_L1, false, true>, SM_BLOCK, 0);
    if (e != cudaSuccess || occ < 1) occ = 1;
  }
  int per_sm = ctx->ctas_per_sm > 0 ? ctx->ctas_per_sm : 2;
  if (per_sm > occ) per_sm = occ;
  int64_t grid = (int64_t)ctx->sm_count * per_sm;
  const int64_t need = (n / (16 / (int64_t)elt) + (int64_t)SM_BLOCK * SM_UNROLL - 1) / ((int64_t)SM_BLOCK * SM_UNROLL) + 1;
  if (grid > need) grid = need;
  if (grid > SM_MAX_CTAS) grid = SM_MAX_CTAS;
  if (grid < 2) grid = 2;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  if (kernel_ms) {
    PB_CHECK_CUDA(cudaEventCreate(&ev[0]));
    PB_CHECK_CUDA(cudaEventCreate(&ev[1]));
    PB_CHECK_CUDA(cudaEventRecord(ev[0], ctx->stream));
  }
  int rc;
  if (dtype == PB_F32)
    rc = f->kind == PB_F_SQDIST ? multi_launch_prox<float, true>(ctx, p, (int)grid, hint) : multi_launch_prox<float, false>(ctx, p, (int)grid, hint);
  else
    rc = f->kind == PB_F_SQDIST ? multi_launch_prox<double, true>(ctx, p, (int)grid, hint) : multi_launch_prox<double, false>(ctx, p, (int)grid, hint);
  if (rc != PB_OK) return rc;
  if (kernel_ms) PB_CHECK_CUDA(cudaEventRecord(ev[1], ctx->stream));
  MultiWs* hws = static_cast<MultiWs*>(ctx->multi_ws);
  struct {
    double fb[PB_NSCALARS];
    long long iterations;
  } tail;
  unsigned int err = 0;
  PB_CHECK_CUDA(cudaMemcpyAsync(&tail, &hws->final_block[0], sizeof(tail), cudaMemcpyDeviceToHost, ctx->stream));
  PB_CHECK_CUDA(cudaMemcpyAsync(&err, &hws->error, sizeof(err), cudaMemcpyDeviceToHost, ctx->stream));
  PB_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
  if (kernel_ms) {
    cudaEventElapsedTime(kernel_ms, ev[0], ev[1]);
    cudaEventDestroy(ev[0]);
    cudaEventDestroy(ev[1]);
  }
  if (ctx->xchg_world > 1) {
    last_seq = sm_seq_of(p.seq0, tail.iterations);
    ctx->xchg_seq = last_seq;
    ctx->xchg_pending = 0;
  }
  if (err) {
    pb_set_error("pb_solve (persistent step kernel): a peer did not publish its scalar block within the device time-out");
    return PB_ECUDA;
  }
  *k_out = tail.iterations;
  for (int q = 0; q < 4; ++q) comb[q] = tail.fb[q];
  return PB_OK;
}
