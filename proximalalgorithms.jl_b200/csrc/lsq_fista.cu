// lsq_fista.cu -- one fixed-stepsize FastForwardBackward iteration on a block-diagonal least-squares term with A read ONCE.
//
// Reference sequence per iteration (fast_forward_backward.jl:130-142 with benchmark/benchmarks.jl:11-17 as f):
//     res = A x - b;  f_x = norm(res)^2 / 2;  grad = A' res          two sweeps of A (4 GB each at BASELINE.json configs[1])
//     y = x - gamma grad;  z = prox(y);  res = x - z;  x+ = z + beta (z - z_prev)      (fused step K2, 5 vectors)
// With a FIXED stepsize the extrapolated point x+ is element-wise in grad, so the columns of A that produce grad_j also produce the
// contribution a_j x+_j to the NEXT iteration's residual A x+ - b.  One kernel therefore sweeps A once per iteration:
//
// Two kernels share that plan: k_bd_fista_ws (default: the three phases below run on their own warps, tiles flow through them over
// mbarriers) and k_bd_fista (phases one after the other behind CTA barriers; kept for row counts beyond the role layout of the first).
//
//   unit = (block k, column chunk c) -- the chunking of lsq_order.h --, one CTA, the chunk streamed through a TMA ring in tiles:
//     A:  grad_j = a_j' r_k                      for the tile's columns      (order of k_gemv_t_sub: LPC lanes per column)
//     B:  the fused step on those columns          -> z_j, x+_j, reductions    (StepElem, step_common.cuh; grad, z, x+ leave as 16-byte stores)
//     C:  acc[row pack][lane] += a_j * x+_j                                    (order of k_gemv_n_partial: 4 column lanes, sequential FMA chains)
//   then the chunk partial of A x+ and the unit's reduction partials are stored; k_bd_fista_combine assembles r+ = (partials in chunk
//   order) - b, ||r+||^2 and folds the step reductions of all units (double-double), with the optional in-kernel exchange.
//
// Every arithmetic order is that of the separate kernels, so iterates, scalars and iteration counts are bit-identical to
// residual + gradient + K2 (tests/test_gpu_lsq.py, tests/test_gpu_solvers.py); HBM traffic per iteration drops from 2 |A| + 5 n
// to |A| + 5 n elements.
#include <stdlib.h>
#include <string.h>

#include "lsq_order.h"
#include "step_common.cuh"
#include "tma.cuh"

#define LF_BLOCK 256
#define LF_STAGES 3

struct LfParams {
  const void* A;
  const void* r;            // A x - b of the CURRENT x (all blocks)
  const void* x;
  const void* z_prev;
  void* grad;
  void* z;
  void* x_next;
  void* partial;            // [nchunk][nblk][mb] chunk partials of A x_next
  double* unit_red;         // [units][8]: gsum (hi, lo), res_sq (hi, lo), gdr (hi, lo), res_inf
  int64_t nblk, mb, nb, chunk_cols;
  int nchunk, t_lpc, t_kp, tile_cols, stages, wa, wb, wait_ns;
  int prox_kind;
  double gamma, beta, pa, pb;       // prox parameters already combined in the element type (launch_step_prox convention)
};

template <typename T, int PROX>
__global__ void __launch_bounds__(LF_BLOCK, 2) k_bd_fista(LfParams p) {
  constexpr bool COMP = sizeof(T) == 8;
  constexpr int VEC = 16 / sizeof(T);
  extern __shared__ __align__(128) unsigned char lf_smem[];
  __shared__ uint64_t full[LF_STAGES];
  __shared__ uint64_t aux_full;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t mb = p.mb, nb = p.nb;
  const int npk = (int)(mb / VEC);
  const int TC = p.tile_cols;
  const uint32_t col_bytes = (uint32_t)(mb * sizeof(T));
  const int c = blockIdx.x, k = blockIdx.y;
  const int64_t c0 = (int64_t)c * p.chunk_cols;
  int64_t c1 = c0 + p.chunk_cols;
  if (c1 > nb) c1 = nb;
  const int ncols = (int)(c1 - c0);
  const int ntile = (ncols + TC - 1) / TC;
  const int64_t cc4 = (p.chunk_cols + 3) & ~(int64_t)3;

  T* ring = reinterpret_cast<T*>(lf_smem);                     // [LF_STAGES][TC][mb]
  T* xs = ring + (size_t)LF_STAGES * TC * mb;                  // x chunk
  T* zps = xs + cc4;                                           // z_prev chunk
  T* g_sm = zps + cc4;                                         // [TC] grad of the tile
  T* xn_sm = g_sm + TC;                                        // [TC] x_next of the tile
  Pack<T, VEC>* lp = reinterpret_cast<Pack<T, VEC>*>(xn_sm + TC);   // [4][npk] lane partials

  const T* __restrict__ A = static_cast<const T*>(p.A);
  const T* __restrict__ src = A + ((int64_t)k * nb + c0) * mb;
  const int64_t j0 = (int64_t)k * nb + c0;                     // first element of this unit in the n-vectors
  const T gamma = (T)p.gamma, beta = (T)p.beta, pa = (T)p.pa, pb = (T)p.pb;

  auto issue = [&](int t) {                                    // thread 0
    const int s = t % LF_STAGES;
    const int tc = ncols - t * TC < TC ? ncols - t * TC : TC;
    const uint32_t bytes = (uint32_t)tc * col_bytes;
    mbar_expect_tx(&full[s], bytes);
    bulk_g2s(ring + (size_t)s * TC * mb, src + (int64_t)t * TC * mb, bytes, &full[s]);
  };
  if (tid == 0) {
    for (int s = 0; s < LF_STAGES; ++s) mbar_init(&full[s], 1);
    mbar_init(&aux_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t vb = (uint32_t)ncols * (uint32_t)sizeof(T);
    mbar_expect_tx(&aux_full, 2 * vb);
    bulk_g2s(xs, static_cast<const T*>(p.x) + j0, vb, &aux_full);
    bulk_g2s(zps, static_cast<const T*>(p.z_prev) + j0, vb, &aux_full);
    const int pre = ntile < LF_STAGES ? ntile : LF_STAGES;
    for (int t = 0; t < pre; ++t) issue(t);
  }
  // r_k of this block: the packs this lane multiplies with (k_gemv_t_sub keeps them in registers for the whole chunk too)
  const int lpc = p.t_lpc, kp = p.t_kp;
  const int cpw = 32 / lpc, sub = lane % lpc, colw = lane / lpc;
  const T* __restrict__ rk = static_cast<const T*>(p.r) + (int64_t)k * mb;
  Pack<T, VEC> rv[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int pkq = sub + q * lpc;
    if (q < kp && pkq < npk)
      rv[q] = *reinterpret_cast<const Pack<T, VEC>*>(rk + pkq * VEC);
    else
#pragma unroll
      for (int e = 0; e < VEC; ++e) rv[q].v[e] = T(0);
  }
  __syncthreads();                                             // mbarrier inits visible
  mbar_wait(&aux_full, 0);

  const bool n_active = tid < npk * 4;
  const int pk = n_active ? tid % npk : 0, cl = n_active ? tid / npk : 0;
  Pack<T, VEC> nacc;
#pragma unroll
  for (int e = 0; e < VEC; ++e) nacc.v[e] = T(0);
  Acc<3, 1> acc, pkacc;                                        // step reductions of the packs this thread owns (tid < TC / VEC)
  acc.clear();
  pkacc.clear();
  T* __restrict__ go = static_cast<T*>(p.grad) + j0;
  T* __restrict__ zo = static_cast<T*>(p.z) + j0;
  T* __restrict__ xo = static_cast<T*>(p.x_next) + j0;

  for (int t = 0; t < ntile; ++t) {
    const int s = t % LF_STAGES;
    mbar_wait(&full[s], (uint32_t)((t / LF_STAGES) & 1));
    const T* tile = ring + (size_t)s * TC * mb;
    const int tc = ncols - t * TC < TC ? ncols - t * TC : TC;
    // ---- A: grad of the tile's columns
    for (int cb = warp * cpw; cb < tc; cb += (LF_BLOCK / 32) * cpw) {
      const int col = cb + colw;
      const bool live = col < tc;
      const T* a = tile + (size_t)(live ? col : 0) * mb;
      T g = T(0);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int pkq = sub + q * lpc;
        if (q < kp && pkq < npk) {
          const Pack<T, VEC> av = *reinterpret_cast<const Pack<T, VEC>*>(a + pkq * VEC);
#pragma unroll
          for (int e = 0; e < VEC; ++e) g = fma(av.v[e], rv[q].v[e], g);
        }
      }
      for (int off = lpc >> 1; off > 0; off >>= 1) g += __shfl_xor_sync(0xffffffffu, g, off);
      if (live && sub == 0) g_sm[col] = g;
    }
    __syncthreads();
    // ---- B: fused step on the tile's columns, one 16-byte pack per thread
    if (tid * VEC < tc) {
      const int jl = t * TC + tid * VEC;                        // offset inside the chunk
      const Pack<T, VEC> xq = *reinterpret_cast<const Pack<T, VEC>*>(xs + jl);
      const Pack<T, VEC> zq = *reinterpret_cast<const Pack<T, VEC>*>(zps + jl);
      const Pack<T, VEC> gq = *reinterpret_cast<const Pack<T, VEC>*>(g_sm + tid * VEC);
      Pack<T, VEC> zn, xn;
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        T yv, rvv;
        StepElem<T, PROX, true>::template run<COMP>(xq.v[e], gq.v[e], zq.v[e], pa, pb, gamma, beta, yv, zn.v[e], rvv, xn.v[e], COMP ? acc : pkacc);
      }
      if constexpr (!COMP) fold_pack<PROX>(acc, pkacc);
      *reinterpret_cast<Pack<T, VEC>*>(xn_sm + tid * VEC) = xn;
      *reinterpret_cast<Pack<T, VEC>*>(go + jl) = gq;
      *reinterpret_cast<Pack<T, VEC>*>(zo + jl) = zn;
      *reinterpret_cast<Pack<T, VEC>*>(xo + jl) = xn;
    }
    __syncthreads();
    // ---- C: the tile's contribution to the chunk partial of A x_next
    if (n_active) {
#pragma unroll 4
      for (int jj = cl; jj < tc; jj += 4) {
        const Pack<T, VEC> a = *reinterpret_cast<const Pack<T, VEC>*>(tile + (size_t)jj * mb + pk * VEC);
        const T xv = xn_sm[jj];
#pragma unroll
        for (int e = 0; e < VEC; ++e) nacc.v[e] = fma(a.v[e], xv, nacc.v[e]);
      }
    }
    __syncthreads();                                            // stage s and g_sm / xn_sm are free again
    if (tid == 0 && t + LF_STAGES < ntile) issue(t + LF_STAGES);
  }
  // ---- chunk partial: lanes added ((l0 + l1) + l2) + l3
  if (n_active) lp[cl * npk + pk] = nacc;
  __syncthreads();
  if (tid < npk) {
    Pack<T, VEC> s_ = lp[tid];
#pragma unroll
    for (int q = 1; q < 4; ++q)
#pragma unroll
      for (int e = 0; e < VEC; ++e) s_.v[e] += lp[q * npk + tid].v[e];
    *reinterpret_cast<Pack<T, VEC>*>(static_cast<T*>(p.partial) + ((int64_t)c * p.nblk + k) * mb + tid * VEC) = s_;
  }
  // ---- the unit's step reductions
  block_reduce<3, 1, LF_BLOCK>(acc);
  if (tid == 0) {
    double* o = p.unit_red + ((size_t)k * p.nchunk + c) * 8;
    o[0] = acc.s[0].hi;
    o[1] = acc.s[0].lo;
    o[2] = acc.s[1].hi;
    o[3] = acc.s[1].lo;
    o[4] = acc.s[2].hi;
    o[5] = acc.s[2].lo;
    o[6] = acc.m[0];
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// The warp-specialised form.  In k_bd_fista above the three phases of a tile run one after the other behind CTA barriers, each
// on a fraction of the threads: ~1 us of dependent latency per tile whatever its size, which caps the sweep at ~0.66 of the HBM
// rate.  Here every phase has its own warps and the tiles flow through them over shared-memory mbarriers, so the phases of
// successive tiles overlap and the only thing a CTA waits for is the TMA ring:
//     warp 0 (one lane)   producer: waits empty[s], issues the bulk copy of tile t into stage s          -> full[s]
//     wa warps            A: grad of the tile's columns (same lanes-per-column order)                     -> gready[s]
//     wb warps            B: the fused step, tile t on warp t % wb, one 16-byte pack per lane          -> xready[s]
//     WC warps            C: acc[row pack][column lane] += a_j * x+_j  (thread = (pack, column lane))     -> empty[s]
// Same arithmetic per column, per pack and per (row pack, column lane) as above, so the results are bit-identical.
// ---------------------------------------------------------------------------------------------------------------------
#define LF_WA_MAX 4
#define LF_WB_MAX 2
#define LF_WC_MAX 4
#define LF_MAX_STAGES 8
#define LF_WS_MAX_THREADS (32 * (1 + LF_WA_MAX + LF_WB_MAX + LF_WC_MAX))

// The gradient warps of k_bd_fista_ws.  LPC (lanes per column) is a template parameter so that the pack loop and the shuffle tree
// unroll, and LF_NP passes (LF_NP x 32/LPC columns) run interleaved: one pass is a dependent chain of 16 FMAs and log2(LPC)
// shuffles; as a rolled loop with run-time LPC it cost ~900 cycles per pass and bounded the whole sweep (profiles/r02_configs.md).
#define LF_NP 4
template <typename T, int LPC>
__device__ __noinline__ void lf_role_grad(const LfParams& p, const T* ring, T* g_sm, uint64_t* full, uint64_t* gready, int aw, int lane, int ncols,
                                          int ntile, int npk, int k) {
  constexpr int VEC = 16 / sizeof(T);
  constexpr int CPW = 32 / LPC;
  const int TC = p.tile_cols, S = p.stages, WA = p.wa;
  const uint32_t hint = (uint32_t)p.wait_ns;
  const int64_t mb = p.mb;
  const int sub = lane % LPC, colw = lane / LPC;
  const T* __restrict__ rk = static_cast<const T*>(p.r) + (int64_t)k * mb;
  Pack<T, VEC> rv[4];
  bool on[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int pkq = sub + q * LPC;
    on[q] = pkq < npk;                      // (q < kp is implied: kp < 4 only for npk <= 2, where pkq >= npk switches the pack off)
    if (on[q])
      rv[q] = *reinterpret_cast<const Pack<T, VEC>*>(rk + pkq * VEC);
    else
#pragma unroll
      for (int e = 0; e < VEC; ++e) rv[q].v[e] = T(0);
  }
  int s = 0, left = ncols;
  uint32_t ph = 0;
  for (int t = 0; t < ntile; ++t) {
    mbar_wait_hint(&full[s], ph, hint);
    const T* tile = ring + (size_t)s * TC * mb;
    T* gs = g_sm + (size_t)s * TC;
    const int tc = left < TC ? left : TC;
    for (int cb = aw * CPW; cb < tc; cb += LF_NP * WA * CPW) {
      int col[LF_NP];
      const T* a[LF_NP];
      T g[LF_NP];
#pragma unroll
      for (int i = 0; i < LF_NP; ++i) {
        col[i] = cb + colw + i * WA * CPW;
        a[i] = tile + (size_t)(col[i] < tc ? col[i] : 0) * mb + sub * VEC;
        g[i] = T(0);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (on[q]) {
          Pack<T, VEC> av[LF_NP];
#pragma unroll
          for (int i = 0; i < LF_NP; ++i) av[i] = *reinterpret_cast<const Pack<T, VEC>*>(a[i] + q * LPC * VEC);
#pragma unroll
          for (int e = 0; e < VEC; ++e)
#pragma unroll
            for (int i = 0; i < LF_NP; ++i) g[i] = fma(av[i].v[e], rv[q].v[e], g[i]);
        }
#pragma unroll
      for (int off = LPC >> 1; off > 0; off >>= 1)
#pragma unroll
        for (int i = 0; i < LF_NP; ++i) g[i] += __shfl_xor_sync(0xffffffffu, g[i], off);
      if (sub == 0) {
#pragma unroll
        for (int i = 0; i < LF_NP; ++i)
          if (col[i] < tc) gs[col[i]] = g[i];
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&gready[s]);
    left -= TC;
    if (++s == S) {
      s = 0;
      ph ^= 1u;
    }
  }
}

template <typename T, int PROX>
__global__ void __launch_bounds__(LF_WS_MAX_THREADS, 2) k_bd_fista_ws(LfParams p) {
  constexpr bool COMP = sizeof(T) == 8;
  constexpr int VEC = 16 / sizeof(T);
  extern __shared__ __align__(128) unsigned char lf_smem[];
  __shared__ uint64_t full[LF_MAX_STAGES], gready[LF_MAX_STAGES], xready[LF_MAX_STAGES], empty[LF_MAX_STAGES];
  __shared__ uint64_t aux_full;
  __shared__ double red[LF_WB_MAX][8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t mb = p.mb, nb = p.nb;
  const int npk = (int)(mb / VEC);
  const int TC = p.tile_cols, S = p.stages, WA = p.wa, WB = p.wb;
  const uint32_t hint = (uint32_t)p.wait_ns;
  const int wc = (int)(blockDim.x >> 5) - (1 + WA + WB);
  const uint32_t col_bytes = (uint32_t)(mb * sizeof(T));
  const int c = blockIdx.x, k = blockIdx.y;
  const int64_t c0 = (int64_t)c * p.chunk_cols;
  int64_t c1 = c0 + p.chunk_cols;
  if (c1 > nb) c1 = nb;
  const int ncols = (int)(c1 - c0);
  const int ntile = (ncols + TC - 1) / TC;
  const int64_t cc4 = (p.chunk_cols + 3) & ~(int64_t)3;

  T* ring = reinterpret_cast<T*>(lf_smem);                     // [S][TC][mb]
  T* xs = ring + (size_t)S * TC * mb;                          // x chunk
  T* zps = xs + cc4;                                           // z_prev chunk
  T* g_sm = zps + cc4;                                         // [S][TC] grad of a tile
  T* xn_sm = g_sm + (size_t)S * TC;                            // [S][TC] x_next of a tile
  Pack<T, VEC>* lp = reinterpret_cast<Pack<T, VEC>*>(xn_sm + (size_t)S * TC);   // [4][npk] lane partials

  const T* __restrict__ A = static_cast<const T*>(p.A);
  const T* __restrict__ src = A + ((int64_t)k * nb + c0) * mb;
  const int64_t j0 = (int64_t)k * nb + c0;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&gready[s], (uint32_t)WA);
      mbar_init(&xready[s], 1);
      mbar_init(&empty[s], (uint32_t)wc);
    }
    mbar_init(&aux_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  Pack<T, VEC> nacc;
#pragma unroll
  for (int e = 0; e < VEC; ++e) nacc.v[e] = T(0);
  const int ctid = tid - 32 * (1 + WA + WB);
  const bool n_active = ctid >= 0 && ctid < npk * 4;
  const int pk = n_active ? ctid % npk : 0, cl = n_active ? ctid / npk : 0;

  if (warp == 0) {
    // ---- producer
    if (lane == 0) {
      const uint32_t vb = (uint32_t)ncols * (uint32_t)sizeof(T);
      mbar_expect_tx(&aux_full, 2 * vb);
      bulk_g2s(xs, static_cast<const T*>(p.x) + j0, vb, &aux_full);
      bulk_g2s(zps, static_cast<const T*>(p.z_prev) + j0, vb, &aux_full);
      int s = 0, left = ncols;
      uint32_t ph = 1;                                           // parity of the PREVIOUS use of the stage
      for (int t = 0; t < ntile; ++t) {
        if (t >= S) mbar_wait_hint(&empty[s], ph, hint);
        const int tc = left < TC ? left : TC;
        const uint32_t bytes = (uint32_t)tc * col_bytes;
        mbar_expect_tx(&full[s], bytes);
        bulk_g2s(ring + (size_t)s * TC * mb, src + (int64_t)t * TC * mb, bytes, &full[s]);
        left -= TC;
        if (++s == S) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
  } else if (warp <= WA) {
    // ---- A: grad of the tile's columns, LPC lanes per column (the order of k_gemv_t_sub)
    switch (p.t_lpc) {
      case 1: lf_role_grad<T, 1>(p, ring, g_sm, full, gready, warp - 1, lane, ncols, ntile, npk, k); break;
      case 2: lf_role_grad<T, 2>(p, ring, g_sm, full, gready, warp - 1, lane, ncols, ntile, npk, k); break;
      case 4: lf_role_grad<T, 4>(p, ring, g_sm, full, gready, warp - 1, lane, ncols, ntile, npk, k); break;
      case 8: lf_role_grad<T, 8>(p, ring, g_sm, full, gready, warp - 1, lane, ncols, ntile, npk, k); break;
      case 16: lf_role_grad<T, 16>(p, ring, g_sm, full, gready, warp - 1, lane, ncols, ntile, npk, k); break;
      default: lf_role_grad<T, 32>(p, ring, g_sm, full, gready, warp - 1, lane, ncols, ntile, npk, k); break;
    }
  } else if (warp <= WA + WB) {
    // ---- B: the fused step on the tile's columns, one 16-byte pack per lane; tiles alternate over the step warps
    const int bw = warp - 1 - WA;
    const T gamma = (T)p.gamma, beta = (T)p.beta, pa = (T)p.pa, pb = (T)p.pb;
    Acc<3, 1> acc, pkacc;
    acc.clear();
    pkacc.clear();
    T* __restrict__ go = static_cast<T*>(p.grad) + j0;
    T* __restrict__ zo = static_cast<T*>(p.z) + j0;
    T* __restrict__ xo = static_cast<T*>(p.x_next) + j0;
    mbar_wait(&aux_full, 0);
    int s = bw % S;
    uint32_t ph = (uint32_t)((bw / S) & 1);
    for (int t = bw; t < ntile; t += WB) {
      const int tc = ncols - t * TC < TC ? ncols - t * TC : TC;
      const bool mine = lane * VEC < tc;
      const int jl = t * TC + lane * VEC;                       // offset inside the chunk
      Pack<T, VEC> xq, zq;
      if (mine) {
        xq = *reinterpret_cast<const Pack<T, VEC>*>(xs + jl);
        zq = *reinterpret_cast<const Pack<T, VEC>*>(zps + jl);
      }
      mbar_wait_hint(&gready[s], ph, hint);
      Pack<T, VEC> gq, zn, xn;
      if (mine) {
        gq = *reinterpret_cast<const Pack<T, VEC>*>(g_sm + (size_t)s * TC + lane * VEC);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          T yv, rvv;
          StepElem<T, PROX, true>::template run<COMP>(xq.v[e], gq.v[e], zq.v[e], pa, pb, gamma, beta, yv, zn.v[e], rvv, xn.v[e], COMP ? acc : pkacc);
        }
        *reinterpret_cast<Pack<T, VEC>*>(xn_sm + (size_t)s * TC + lane * VEC) = xn;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&xready[s]);                   // C can start while this warp folds and stores
      if (mine) {
        if constexpr (!COMP) fold_pack<PROX>(acc, pkacc);
        *reinterpret_cast<Pack<T, VEC>*>(go + jl) = gq;
        *reinterpret_cast<Pack<T, VEC>*>(zo + jl) = zn;
        *reinterpret_cast<Pack<T, VEC>*>(xo + jl) = xn;
      }
      for (int i = 0; i < WB; ++i)
        if (++s == S) {
          s = 0;
          ph ^= 1u;
        }
    }
    warp_reduce<3, 1>(acc);
    if (lane == 0) {
      red[bw][0] = acc.s[0].hi;
      red[bw][1] = acc.s[0].lo;
      red[bw][2] = acc.s[1].hi;
      red[bw][3] = acc.s[1].lo;
      red[bw][4] = acc.s[2].hi;
      red[bw][5] = acc.s[2].lo;
      red[bw][6] = acc.m[0];
    }
  } else {
    // ---- C: the tile's contribution to the chunk partial of A x_next (4 column lanes, sequential FMA chains: k_gemv_n_partial's order)
    int s = 0, left = ncols;
    uint32_t ph = 0;
    const size_t cstep = (size_t)4 * mb;                        // four columns on
    const size_t toff = (size_t)cl * mb + (size_t)pk * VEC;
    for (int t = 0; t < ntile; ++t) {
      const T* ap = ring + (size_t)s * TC * mb + toff;
      const T* xp = xn_sm + (size_t)s * TC + cl;
      const int tc = left < TC ? left : TC;
      mbar_wait_hint(&full[s], ph, hint);
      mbar_wait_hint(&xready[s], ph, hint);
      if (n_active) {
        int jj = cl;
        for (; jj + 12 < tc; jj += 16) {
          const Pack<T, VEC> a0 = *reinterpret_cast<const Pack<T, VEC>*>(ap);
          const Pack<T, VEC> a1 = *reinterpret_cast<const Pack<T, VEC>*>(ap + cstep);
          const Pack<T, VEC> a2 = *reinterpret_cast<const Pack<T, VEC>*>(ap + 2 * cstep);
          const Pack<T, VEC> a3 = *reinterpret_cast<const Pack<T, VEC>*>(ap + 3 * cstep);
          const T x0 = xp[0], x1 = xp[4], x2 = xp[8], x3 = xp[12];
#pragma unroll
          for (int e = 0; e < VEC; ++e) nacc.v[e] = fma(a0.v[e], x0, nacc.v[e]);
#pragma unroll
          for (int e = 0; e < VEC; ++e) nacc.v[e] = fma(a1.v[e], x1, nacc.v[e]);
#pragma unroll
          for (int e = 0; e < VEC; ++e) nacc.v[e] = fma(a2.v[e], x2, nacc.v[e]);
#pragma unroll
          for (int e = 0; e < VEC; ++e) nacc.v[e] = fma(a3.v[e], x3, nacc.v[e]);
          ap += 4 * cstep;
          xp += 16;
        }
        for (; jj < tc; jj += 4) {
          const Pack<T, VEC> a0 = *reinterpret_cast<const Pack<T, VEC>*>(ap);
          const T x0 = xp[0];
#pragma unroll
          for (int e = 0; e < VEC; ++e) nacc.v[e] = fma(a0.v[e], x0, nacc.v[e]);
          ap += cstep;
          xp += 4;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
      left -= TC;
      if (++s == S) {
        s = 0;
        ph ^= 1u;
      }
    }
    if (n_active) lp[cl * npk + pk] = nacc;
  }
  __syncthreads();
  // ---- chunk partial: lanes added ((l0 + l1) + l2) + l3
  if (tid < npk) {
    Pack<T, VEC> s_ = lp[tid];
#pragma unroll
    for (int q = 1; q < 4; ++q)
#pragma unroll
      for (int e = 0; e < VEC; ++e) s_.v[e] += lp[q * npk + tid].v[e];
    *reinterpret_cast<Pack<T, VEC>*>(static_cast<T*>(p.partial) + ((int64_t)c * p.nblk + k) * mb + tid * VEC) = s_;
  }
  // ---- the unit's step reductions
  if (tid == 32) {
    Acc<3, 1> a;
    a.clear();
    for (int w = 0; w < WB; ++w) {
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        dd d;
        d.hi = red[w][2 * q];
        d.lo = red[w][2 * q + 1];
        a.s[q] = dd_sum(a.s[q], d);
      }
      a.m[0] = nanmax(a.m[0], red[w][6]);
    }
    double* o = p.unit_red + ((size_t)k * p.nchunk + c) * 8;
    o[0] = a.s[0].hi;
    o[1] = a.s[0].lo;
    o[2] = a.s[1].hi;
    o[3] = a.s[1].lo;
    o[4] = a.s[2].hi;
    o[5] = a.s[2].lo;
    o[6] = a.m[0];
  }
}

// r_next[i] = (sum over chunks, in chunk order) - b[i], AUX = ||r_next||^2; GSUM / RESSQ / GDR / RESINF = fold of the units' partials
template <typename T>
__global__ void __launch_bounds__(PB_BLOCK) k_bd_fista_combine(const T* __restrict__ partial, int nchunk, int64_t M, const T* __restrict__ b,
                                                               T* __restrict__ r, const double* __restrict__ unit_red, int64_t units,
                                                               PbWorkspace* ws, double* outs, XchgParams xp) {
  constexpr bool COMP = sizeof(T) == 8;
  Acc<4, 1> acc;
  acc.clear();
  for (int64_t i = (int64_t)blockIdx.x * PB_BLOCK + threadIdx.x; i < M; i += (int64_t)gridDim.x * PB_BLOCK) {
    T s = partial[i];
    for (int c = 1; c < nchunk; ++c) s += partial[(int64_t)c * M + i];
    const T rv = b ? sub_rn(s, b[i]) : s;
    r[i] = rv;
    if (COMP)
      dd_add_prod(acc.s[3], (double)rv, (double)rv);
    else
      acc.s[3].hi = __fma_rn((double)rv, (double)rv, acc.s[3].hi);
  }
  for (int64_t u = (int64_t)blockIdx.x * PB_BLOCK + threadIdx.x; u < units; u += (int64_t)gridDim.x * PB_BLOCK) {
    const double* o = unit_red + (size_t)u * 8;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      dd d;
      d.hi = o[2 * q];
      d.lo = o[2 * q + 1];
      acc.s[q] = dd_sum(acc.s[q], d);
    }
    acc.m[0] = nanmax(acc.m[0], o[6]);
  }
  OutMap map;
  map.sum_slot[0] = PB_S_GSUM;
  map.sum_slot[1] = PB_S_RESSQ;
  map.sum_slot[2] = PB_S_GDR;
  map.sum_slot[3] = PB_S_AUX;
  map.max_slot[0] = PB_S_RESINF;
  map.max_slot[1] = -1;
  grid_reduce<4, 1, PB_BLOCK>(acc, ws, outs, map, &xp);
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
struct LfPlan {
  PbLsqOrder ord;
  int tile_cols;
  size_t smem;
  int ws_stages, ws_threads, ws_wa, ws_wb;       // > 0: the warp-specialised kernel with this ring depth
};

template <typename T>
static bool lf_plan(const pb_ctx* ctx, const pb_smooth* f, const pb_prox* g, const void* x, const void* z_prev, const void* grad, const void* z,
                    const void* x_next, LfPlan* plan) {
  constexpr int VEC = 16 / sizeof(T);
  if (ctx->lsq_fista < 0) return false;
  if (f->kind != PB_F_LSQ_BLOCKDIAG || f->nblk < 1 || f->mb < 1 || f->nb < 1) return false;
  if (g->kind != PB_PROX_ZERO && g->kind != PB_PROX_L1 && g->kind != PB_PROX_BOX) return false;
  if (g->kind == PB_PROX_BOX && (g->v0 || g->v1)) return false;
  const int64_t mb = f->mb, nb = f->nb, npk = mb / VEC;
  plan->ord = pb_lsq_order(sizeof(T), f->nblk, mb, nb, mb, mb * nb, f->A, f->r);
  if (plan->ord.n_sub || !plan->ord.t_sub || mb % VEC != 0 || nb % VEC != 0 || plan->ord.chunk_cols % VEC != 0 || npk * 4 > LF_BLOCK) return false;
  if (f->nblk > 65535 || plan->ord.nchunk > 0x7fffffff) return false;
  if (!pb_aligned16(f->A) || !pb_aligned16(f->r) || !pb_aligned16(x) || !pb_aligned16(z_prev) || !pb_aligned16(grad) || !pb_aligned16(z) ||
      !pb_aligned16(x_next))
    return false;
  // worth it only when A does not live in L2 anyway
  if (ctx->lsq_fista == 0 && (double)f->nblk * (double)mb * (double)nb * sizeof(T) < 64.0 * 1024 * 1024) return false;
  int tc = 64;
  size_t tile_cap = 26 * 1024, smem_cap = 110 * 1024;
  if (const char* e = getenv("PROXB200_LF_TC")) {       // tuning hook (tools/tune_lsq_fista.sh): tile columns; large tiles take one CTA per SM
    tc = atoi(e);
    tile_cap = 64 * 1024;
    smem_cap = 200 * 1024;
  }
  while (tc >= 16 && (size_t)tc * mb * sizeof(T) > tile_cap) tc >>= 1;
  if (tc < 16 || tc / VEC > LF_BLOCK) return false;
  const size_t cc4 = ((size_t)plan->ord.chunk_cols + 3) & ~(size_t)3;
  plan->tile_cols = tc;
  plan->ws_stages = 0;
  plan->smem = ((size_t)LF_STAGES * tc * mb + 2 * cc4 + 2 * tc) * sizeof(T) + (size_t)4 * npk * 16 + 128;
  // warp-specialised form: <= 4 C warps (thread = (row pack, column lane)), one pack per lane in the step warp; smaller tiles, deeper ring
  const char* v1 = getenv("PROXB200_LF_V1");
  if (!(v1 && atoi(v1)) && npk * 4 <= 32 * LF_WC_MAX) {
    int wtc = (int)(23040 / (mb * sizeof(T))) & ~3, st = 0;          // ~22 KB tiles: best of the sweep on configs[1] (tools/one_fista_time.py)
    if (wtc > 32 * VEC) wtc = 32 * VEC;
    if (wtc < 16) wtc = 16;
    if (const char* e = getenv("PROXB200_LF_WS_TC"))     // tuning hook
      wtc = atoi(e);
    else
      while (wtc > 16 && (size_t)wtc * mb * sizeof(T) > 26 * 1024) wtc = (wtc >> 1) & ~3;
    if (wtc >= VEC && wtc % VEC == 0 && wtc / VEC <= 32) {
      const size_t fixed = 2 * cc4 * sizeof(T) + (size_t)4 * npk * 16 + 128;
      const size_t per_stage = ((size_t)wtc * mb + 2 * wtc) * sizeof(T);
      st = (int)((110 * 1024 - fixed) / per_stage);
      if (st > LF_MAX_STAGES) st = LF_MAX_STAGES;
      if (const char* e = getenv("PROXB200_LF_WS_STAGES")) st = atoi(e) < st ? atoi(e) : st;
      if (st >= 2) {
        plan->ws_stages = st;
        int wb = 1;                 // one step warp keeps up (the sweep is HBM bound); more warps only add polling
        if (const char* e = getenv("PROXB200_LF_WS_WB")) wb = atoi(e);
        plan->ws_wb = wb < 1 ? 1 : (wb > LF_WB_MAX ? LF_WB_MAX : wb);
        int wa = 4;
        if (const char* e = getenv("PROXB200_LF_WS_WA")) wa = atoi(e);
        plan->ws_wa = wa < 1 ? 1 : (wa > LF_WA_MAX ? LF_WA_MAX : wa);
        plan->ws_threads = 32 * (1 + plan->ws_wa + plan->ws_wb + (int)((npk * 4 + 31) / 32));
        plan->tile_cols = wtc;
        plan->smem = fixed + (size_t)st * per_stage;
        return true;
      }
    }
  }
  return plan->smem <= smem_cap;
}

bool pb_bd_fista_eligible(const pb_ctx* ctx, int dtype, const pb_smooth* f, const pb_prox* g, const void* x, const void* z_prev, const void* grad,
                          const void* z, const void* x_next) {
  LfPlan plan;
  return dtype == PB_F32 ? lf_plan<float>(ctx, f, g, x, z_prev, grad, z, x_next, &plan) : lf_plan<double>(ctx, f, g, x, z_prev, grad, z, x_next, &plan);
}

template <typename T, int PROX>
static int lf_launch_ws(pb_ctx* ctx, const LfParams& p, const LfPlan& plan) {
  auto kern = k_bd_fista_ws<T, PROX>;
  static bool attr_done[PB_MAX_DEVICES] = {};
  const int dev = ctx->device < PB_MAX_DEVICES ? ctx->device : 0;
  if (!attr_done[dev] || ctx->device >= PB_MAX_DEVICES) {
    PB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done[dev] = true;
  }
  dim3 grid((unsigned)p.nchunk, (unsigned)p.nblk);
  kern<<<grid, plan.ws_threads, plan.smem, ctx->stream>>>(p);
  PB_LAUNCH_CHECK(ctx);
  return PB_OK;
}

template <typename T, int PROX>
static int lf_launch(pb_ctx* ctx, const LfParams& p, const LfPlan& plan) {
  if (plan.ws_stages > 0) return lf_launch_ws<T, PROX>(ctx, p, plan);
  auto kern = k_bd_fista<T, PROX>;
  static bool attr_done[PB_MAX_DEVICES] = {};
  const int dev = ctx->device < PB_MAX_DEVICES ? ctx->device : 0;
  if (!attr_done[dev] || ctx->device >= PB_MAX_DEVICES) {
    PB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done[dev] = true;
  }
  dim3 grid((unsigned)p.nchunk, (unsigned)p.nblk);
  kern<<<grid, LF_BLOCK, plan.smem, ctx->stream>>>(p);
  PB_LAUNCH_CHECK(ctx);
  return PB_OK;
}

// One iteration: on entry f->r = A x - b; on return grad = A' r, z, x_next are written, f->r = A x_next - b and the scalar block holds
// GSUM / RESSQ / GDR / RESINF of the step and AUX = ||A x_next - b||^2 (the value of the NEXT point: the caller shifts it).
template <typename T>
static int lf_run(pb_ctx* ctx, const pb_smooth* f, const pb_prox* g, double gamma, double beta, const void* x, const void* z_prev, void* grad, void* z,
                  void* x_next) {
  LfPlan plan;
  if (!lf_plan<T>(ctx, f, g, x, z_prev, grad, z, x_next, &plan)) {
    pb_set_error("pb_bd_fista_step: problem is not eligible");
    return PB_EUNSUPPORTED;
  }
  const int64_t M = f->nblk * f->mb, units = f->nblk * plan.ord.nchunk;
  const size_t part_bytes = ((size_t)plan.ord.nchunk * M * sizeof(T) + 255) & ~(size_t)255;
  int rc = pb_ensure_scratch(ctx, part_bytes + (size_t)units * 8 * sizeof(double));
  if (rc != PB_OK) return rc;
  unsigned char* wsb = static_cast<unsigned char*>(ctx->scratch);
  LfParams p;
  memset(&p, 0, sizeof(p));
  p.A = f->A;
  p.r = f->r;
  p.x = x;
  p.z_prev = z_prev;
  p.grad = grad;
  p.z = z;
  p.x_next = x_next;
  p.partial = wsb;
  p.unit_red = reinterpret_cast<double*>(wsb + part_bytes);
  p.nblk = f->nblk;
  p.mb = f->mb;
  p.nb = f->nb;
  p.chunk_cols = plan.ord.chunk_cols;
  p.nchunk = (int)plan.ord.nchunk;
  p.t_lpc = plan.ord.t_lpc;
  p.t_kp = plan.ord.t_kp;
  p.tile_cols = plan.tile_cols;
  p.stages = plan.ws_stages;
  p.wb = plan.ws_wb;
  p.wa = plan.ws_wa;
  p.wait_ns = 2000;
  if (const char* e = getenv("PROXB200_LF_WS_WAIT_NS")) p.wait_ns = atoi(e);
  p.prox_kind = g->kind;
  p.gamma = (double)(T)gamma;
  p.beta = (double)(T)beta;
  switch (g->kind) {
    case PB_PROX_L1: p.pa = (double)mul_rn_host((T)gamma, (T)g->p0); break;   // gamma*lambda: one rounding in R, like the package
    case PB_PROX_BOX:
      p.pa = g->p0;
      p.pb = g->p1;
      break;
    default: break;
  }
  switch (g->kind) {
    case PB_PROX_L1: rc = lf_launch<T, PB_PROX_L1>(ctx, p, plan); break;
    case PB_PROX_BOX: rc = lf_launch<T, PB_PROX_BOX>(ctx, p, plan); break;
    default: rc = lf_launch<T, PB_PROX_ZERO>(ctx, p, plan); break;
  }
  if (rc != PB_OK) return rc;
  XchgParams xp;
  pb_xchg_next(ctx, &xp, ctx->xchg_fused != 0);
  const int cgrid = pb_stream_grid(ctx, PB_BLOCK, M > units ? M : units, 2);
  k_bd_fista_combine<T><<<cgrid, PB_BLOCK, 0, ctx->stream>>>(static_cast<const T*>(p.partial), p.nchunk, M, static_cast<const T*>(f->b),
                                                             static_cast<T*>(f->r), p.unit_red, units, ctx->ws, ctx->scalars_dev, xp);
  PB_LAUNCH_CHECK(ctx);
  return PB_OK;
}

int pb_bd_fista_step(pb_ctx* ctx, int dtype, const pb_smooth* f, const pb_prox* g, double gamma, double beta, const void* x, const void* z_prev,
                     void* grad, void* z, void* x_next) {
  if (dtype == PB_F32) return lf_run<float>(ctx, f, g, gamma, beta, x, z_prev, grad, z, x_next);
  return lf_run<double>(ctx, f, g, gamma, beta, x, z_prev, grad, z, x_next);
}
