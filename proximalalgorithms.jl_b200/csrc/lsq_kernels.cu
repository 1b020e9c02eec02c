// lsq_kernels.cu -- K4: value and gradient of the least-squares smooth term f(x) = 0.5*||A x - b||^2
// (benchmark/benchmarks.jl:11-17: `res = A*x - b; norm(res)^2/2, A'*res`), dense column-major and block-diagonal.
//
// A single-right-hand-side product has 0.5 (fp32) / 0.25 (fp64) flop per byte of A: it is HBM (or, for the small
// benchmark fixtures, L2/latency) bound and there is no tensor-core formulation that would not change the arithmetic,
// so both products are plain coalesced streaming kernels with fixed-order (deterministic) accumulation:
//   residual : thread per row, column range split into chunks (grid.y) and 4 column lanes per CTA; chunk partials are
//              folded in index order by a second kernel that also subtracts b and reduces ||r||^2 (double-double).
//   gradient : warp per column (columns are contiguous), lanes stride over rows, shuffle tree.
#include "common.cuh"
#include "lsq_order.h"

#define GEMV_ROWS 128  // rows per CTA of the residual kernel
#define GEMV_CL PB_GEMV_CL          // column lanes per CTA
#define GEMV_UNROLL PB_GEMV_UNROLL

// partial[(chunk*nblk + k)*mb + i] = sum_{j in chunk} A_k[i, j] * x_k[j]
template <typename T>
__global__ void __launch_bounds__(GEMV_ROWS* GEMV_CL)
    k_gemv_n_partial(const T* __restrict__ A, int64_t lda, int64_t blk_stride, const T* __restrict__ x,
                     T* __restrict__ partial, int64_t mb, int64_t nb, int64_t nblk, int64_t chunk_cols) {
  __shared__ T sh[GEMV_CL][GEMV_ROWS];
  const int rl = threadIdx.x % GEMV_ROWS, cl = threadIdx.x / GEMV_ROWS;
  const int64_t k = blockIdx.z;
  const int64_t row = (int64_t)blockIdx.x * GEMV_ROWS + rl;
  const int64_t c0 = (int64_t)blockIdx.y * chunk_cols;
  int64_t c1 = c0 + chunk_cols;
  if (c1 > nb) c1 = nb;
  const T* __restrict__ Ak = A + k * blk_stride;
  const T* __restrict__ xk = x + k * nb;
  T acc = T(0);
  if (row < mb) {
    int64_t j = c0 + cl;
    // UNROLL independent loads in flight per thread
    for (; j + (GEMV_UNROLL - 1) * GEMV_CL < c1; j += GEMV_UNROLL * GEMV_CL) {
      T a[GEMV_UNROLL], xv[GEMV_UNROLL];
#pragma unroll
      for (int u = 0; u < GEMV_UNROLL; ++u) {
        a[u] = __ldg(Ak + row + (j + u * GEMV_CL) * lda);
        xv[u] = __ldg(xk + j + u * GEMV_CL);
      }
#pragma unroll
      for (int u = 0; u < GEMV_UNROLL; ++u) acc = fma(a[u], xv[u], acc);
    }
    for (; j < c1; j += GEMV_CL) acc = fma(__ldg(Ak + row + j * lda), __ldg(xk + j), acc);
  }
  sh[cl][rl] = acc;
  __syncthreads();
  if (cl == 0 && row < mb) {
    T s = sh[0][rl];
#pragma unroll
    for (int c = 1; c < GEMV_CL; ++c) s += sh[c][rl];
    partial[((int64_t)blockIdx.y * nblk + k) * mb + row] = s;
  }
}

// Same product, same summation order per row (4 column lanes, each a sequential FMA chain over its columns, lanes added
// ((l0 + l1) + l2) + l3), but every thread owns ONE 16-byte pack of VEC consecutive rows: a warp instruction moves 512 bytes of a
// column instead of 128, and UNROLL packs per thread are in flight.  The thread-per-row kernel above reached 0.59 of the measured HBM
// peak on a tall dense matrix (10000 x 100000 fp32, profiles/r01_ncu_lsq.md: 62 % DRAM, 4-byte loads); results are bit-identical.
// RP = row packs per CTA (block = RP x 4 threads): 128 for tall columns, 32 when a column has at most 32 packs.
template <typename T, int RP>
__global__ void __launch_bounds__(RP* GEMV_CL)
    k_gemv_n_partial_v(const T* __restrict__ A, int64_t lda, int64_t blk_stride, const T* __restrict__ x, T* __restrict__ partial,
                       int64_t mb, int64_t nb, int64_t nblk, int64_t chunk_cols) {
  constexpr int VEC = 16 / sizeof(T);
  __shared__ Pack<T, VEC> sh[GEMV_CL][RP];
  const int rl = threadIdx.x % RP, cl = threadIdx.x / RP;
  const int64_t k = blockIdx.z;
  const int64_t pk = (int64_t)blockIdx.x * RP + rl;          // row pack of this thread
  const int64_t npk = mb / VEC;
  const int64_t c0 = (int64_t)blockIdx.y * chunk_cols;
  int64_t c1 = c0 + chunk_cols;
  if (c1 > nb) c1 = nb;
  const T* __restrict__ Ak = A + k * blk_stride + pk * VEC;
  const T* __restrict__ xk = x + k * nb;
  Pack<T, VEC> acc;
#pragma unroll
  for (int e = 0; e < VEC; ++e) acc.v[e] = T(0);
  if (pk < npk) {
    int64_t j = c0 + cl;
    for (; j + (GEMV_UNROLL - 1) * GEMV_CL < c1; j += GEMV_UNROLL * GEMV_CL) {
      Pack<T, VEC> a[GEMV_UNROLL];
      T xv[GEMV_UNROLL];
#pragma unroll
      for (int u = 0; u < GEMV_UNROLL; ++u) {
        a[u] = ld_pack<T, VEC, false>(Ak + (j + u * GEMV_CL) * lda);
        xv[u] = __ldg(xk + j + u * GEMV_CL);
      }
#pragma unroll
      for (int u = 0; u < GEMV_UNROLL; ++u)
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc.v[e] = fma(a[u].v[e], xv[u], acc.v[e]);
    }
    for (; j < c1; j += GEMV_CL) {
      const Pack<T, VEC> a = ld_pack<T, VEC, false>(Ak + j * lda);
      const T xv = __ldg(xk + j);
#pragma unroll
      for (int e = 0; e < VEC; ++e) acc.v[e] = fma(a.v[e], xv, acc.v[e]);
    }
  }
  sh[cl][rl] = acc;
  __syncthreads();
  if (cl == 0 && pk < npk) {
    Pack<T, VEC> s = sh[0][rl];
#pragma unroll
    for (int c = 1; c < GEMV_CL; ++c)
#pragma unroll
      for (int e = 0; e < VEC; ++e) s.v[e] += sh[c][rl].v[e];
    *reinterpret_cast<Pack<T, VEC>*>(partial + ((int64_t)blockIdx.y * nblk + k) * mb + pk * VEC) = s;
  }
}

// partial product for SHORT columns (mb < 64, e.g. the 4 x 128000 blocks of the group-lasso workload): the thread-per-row
// kernel above would keep mb of its 128 row lanes busy.  Mirror image of k_gemv_t_sub: LPC lanes share one column, each lane
// owns up to KP 16-byte packs of it and accumulates a_ij * x_j into its own row accumulators while the CTA sweeps the
// columns of its chunk (coalesced 16-byte loads of A, 4 sweeps in flight); the column lanes are folded with a fixed
// shuffle tree, the warps through shared memory in warp order.  One CTA = one column chunk of ONE block, so -- like the
// kernel above -- the summation order depends on the block shape only.
template <typename T, int LPC, int KP>
__global__ void __launch_bounds__(PB_BLOCK) k_gemv_n_sub(const T* __restrict__ A, int64_t lda, int64_t blk_stride,
                                                         const T* __restrict__ x, T* __restrict__ partial, int64_t mb,
                                                         int64_t nb, int64_t nblk, int64_t chunk_cols, int64_t chunks_per_blk) {
  constexpr int VEC = 16 / sizeof(T);
  constexpr int CPW = 32 / LPC;
  constexpr int WARPS = PB_BLOCK / 32;
  constexpr int SW = 4;                               // column sweeps in flight
  __shared__ T sh[WARPS][LPC * KP * VEC];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane % LPC, colw = lane / LPC;
  const int64_t k = blockIdx.x / chunks_per_blk, chunk = blockIdx.x % chunks_per_blk;
  const int64_t c0 = chunk * chunk_cols;
  int64_t c1 = c0 + chunk_cols;
  if (c1 > nb) c1 = nb;
  const T* __restrict__ Ak = A + k * blk_stride;
  const T* __restrict__ xk = x + k * nb;
  const int64_t npk = mb / VEC;
  T acc[KP][VEC];
#pragma unroll
  for (int q = 0; q < KP; ++q)
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc[q][e] = T(0);
  const int64_t stride = (int64_t)WARPS * CPW;
  for (int64_t j0 = c0 + (int64_t)warp * CPW + colw; j0 < c1; j0 += SW * stride) {
    Pack<T, VEC> av[SW][KP];
    T xv[SW];
#pragma unroll
    for (int s_ = 0; s_ < SW; ++s_) {
      const int64_t j = j0 + s_ * stride;
      const bool live = j < c1;
      xv[s_] = live ? __ldg(xk + j) : T(0);
      const T* __restrict__ a = Ak + (live ? j : c0) * lda;
#pragma unroll
      for (int q = 0; q < KP; ++q) {
        const int64_t pk = sub + (int64_t)q * LPC;
        if (pk < npk) av[s_][q] = ld_pack<T, VEC, true>(a + pk * VEC);
      }
    }
#pragma unroll
    for (int s_ = 0; s_ < SW; ++s_)
#pragma unroll
      for (int q = 0; q < KP; ++q) {
        const int64_t pk = sub + (int64_t)q * LPC;
        if (pk < npk) {
#pragma unroll
          for (int e = 0; e < VEC; ++e) acc[q][e] = fma(av[s_][q].v[e], xv[s_], acc[q][e]);
        }
      }
  }
  // fold the column lanes of the warp (lanes with equal `sub`), then the warps in order
#pragma unroll
  for (int q = 0; q < KP; ++q)
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      T v = acc[q][e];
#pragma unroll
      for (int off = 16; off >= LPC; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      if (colw == 0) sh[warp][(q * LPC + sub) * VEC + e] = v;
    }
  __syncthreads();
  for (int64_t i = threadIdx.x; i < mb; i += PB_BLOCK) {
    T s_ = sh[0][i];
#pragma unroll
    for (int w = 1; w < WARPS; ++w) s_ += sh[w][i];
    partial[(chunk * nblk + k) * mb + i] = s_;
  }
}

// r[i] = (sum_chunks partial[c][i]) - b[i];  AUX = sum r^2
template <typename T>
__global__ void __launch_bounds__(PB_BLOCK) k_gemv_n_combine(const T* __restrict__ partial, int nchunk, int64_t M,
                                                             const T* __restrict__ b, T* __restrict__ r,
                                                             PbWorkspace* ws, double* outs) {
  constexpr bool COMP = sizeof(T) == 8;
  Acc<1, 1> acc;
  acc.clear();
  for (int64_t i = (int64_t)blockIdx.x * PB_BLOCK + threadIdx.x; i < M; i += (int64_t)gridDim.x * PB_BLOCK) {
    T s = partial[i];
    for (int c = 1; c < nchunk; ++c) s += partial[(int64_t)c * M + i];
    const T rv = b ? sub_rn(s, b[i]) : s;
    r[i] = rv;
    if (COMP)
      dd_add_prod(acc.s[0], (double)rv, (double)rv);
    else
      acc.s[0].hi = __fma_rn((double)rv, (double)rv, acc.s[0].hi);
  }
  OutMap map;
  map.sum_slot[0] = PB_S_AUX;
  map.sum_slot[1] = map.sum_slot[2] = map.sum_slot[3] = -1;
  map.max_slot[0] = map.max_slot[1] = -1;
  grid_reduce<1, 1, PB_BLOCK>(acc, ws, outs, map);
}

// C2 -- column-sharded dense A: combine + all-gather + fold in ONE kernel over NVLink peer memory (SURVEY.md section 8e "Dense-A
// gradient under this partition").  Rank p holds the columns of the GLOBAL chunks [cbase, cbase + nloc) and has just computed their
// partials.  Thread i (row i) 1. pushes its nloc partials to slot [parity][chunk][i] of EVERY peer's vector region (LL words
// {32 data bits, 32-bit sequence}: xchg.cuh), 2. folds ALL nch chunks in GLOBAL chunk order -- its own from the local buffer, the
// others from its landing zone as they arrive -- exactly like k_gemv_n_combine on one GPU, so r, ||r||^2 and everything downstream
// are bit-identical for every shard count (and identical on every rank).  No thread waits before it has pushed: no co-residency
// requirement, no deadlock.  aux_on == 0: this rank contributes 0 to AUX (the native driver sums AUX over the ranks).
template <typename T>
__global__ void __launch_bounds__(PB_BLOCK) k_gemv_n_combine_x(const T* __restrict__ partial, int nloc, int cbase, int nch, int64_t M,
                                                               const T* __restrict__ b, T* __restrict__ r, PbWorkspace* ws, double* outs,
                                                               XchgVecParams xv, int aux_on) {
  constexpr bool COMP = sizeof(T) == 8;
  constexpr int WPE = sizeof(T) / 4;          // 8-byte LL words per element
  constexpr int BATCH = 8;
  const int par = (int)(xv.seq & 1u);
  const unsigned long long tag = (unsigned long long)xv.seq << 32;
  unsigned long long* mine = xv.peer[xv.rank];
  Acc<1, 1> acc;
  acc.clear();
  bool bad = false;
  for (int64_t i = (int64_t)blockIdx.x * PB_BLOCK + threadIdx.x; i < M; i += (int64_t)gridDim.x * PB_BLOCK) {
    // 1. push
    for (int cl = 0; cl < nloc; ++cl) {
      const T v = partial[(int64_t)cl * M + i];
      unsigned int half[WPE];
      if constexpr (COMP) {
        const unsigned long long bits = (unsigned long long)__double_as_longlong((double)v);
        half[0] = (unsigned int)(bits & 0xffffffffull);
        half[WPE - 1] = (unsigned int)(bits >> 32);
      } else {
        half[0] = __float_as_uint((float)v);
      }
      const size_t slot = (((size_t)par * nch + (cbase + cl)) * (size_t)M + (size_t)i) * WPE;
      for (int q = 0; q < xv.world; ++q) {
        if (q == xv.rank) continue;
#pragma unroll
        for (int h = 0; h < WPE; ++h) st_word(xv.peer[q] + slot + h, tag | half[h]);
      }
    }
    // 2. fold in global chunk order
    T s = T(0);
    for (int c0 = 0; c0 < nch; c0 += BATCH) {
      unsigned long long w[BATCH][WPE];
#pragma unroll
      for (int u = 0; u < BATCH; ++u) {
        const int c = c0 + u;
        if (c < nch && (c < cbase || c >= cbase + nloc)) {
          const size_t slot = (((size_t)par * nch + c) * (size_t)M + (size_t)i) * WPE;
#pragma unroll
          for (int h = 0; h < WPE; ++h) w[u][h] = ld_word(mine + slot + h);
        }
      }
#pragma unroll
      for (int u = 0; u < BATCH; ++u) {
        const int c = c0 + u;
        if (c >= nch) break;
        T v;
        if (c >= cbase && c < cbase + nloc) {
          v = partial[(int64_t)(c - cbase) * M + i];
        } else {
          const size_t slot = (((size_t)par * nch + c) * (size_t)M + (size_t)i) * WPE;
#pragma unroll
          for (int h = 0; h < WPE; ++h) {
            if (!bad && (unsigned int)(w[u][h] >> 32) != xv.seq) {      // (after one time-out the launch is lost: do not wait again)
              const unsigned long long t0 = globaltimer_ns();
              unsigned int spins = 0;
              do {
                w[u][h] = ld_word(mine + slot + h);
                if ((++spins & 0x3ff) == 0 && globaltimer_ns() - t0 > PB_XCHG_TIMEOUT_NS) {
                  bad = true;
                  break;
                }
              } while ((unsigned int)(w[u][h] >> 32) != xv.seq);
            }
          }
          if constexpr (COMP) {
            const unsigned long long bits = (w[u][0] & 0xffffffffull) | ((w[u][WPE - 1] & 0xffffffffull) << 32);
            v = (T)__longlong_as_double((long long)bits);
          } else {
            v = (T)__uint_as_float((unsigned int)(w[u][0] & 0xffffffffull));
          }
        }
        s = c == 0 ? v : s + v;
      }
    }
    T rv = b ? sub_rn(s, b[i]) : s;
    if (bad) rv = (T)__longlong_as_double(0x7ff8000000000000ll);     // a peer never published: poison the result, flag the host
    r[i] = rv;
    if (COMP)
      dd_add_prod(acc.s[0], (double)rv, (double)rv);
    else
      acc.s[0].hi = __fma_rn((double)rv, (double)rv, acc.s[0].hi);
  }
  if (bad) st_word(xv.err_word, (unsigned long long)PB_XCHG_ERROR_SEQ << 32);
  if (!aux_on) acc.clear();
  OutMap map;
  map.sum_slot[0] = PB_S_AUX;
  map.sum_slot[1] = map.sum_slot[2] = map.sum_slot[3] = -1;
  map.max_slot[0] = map.max_slot[1] = -1;
  grid_reduce<1, 1, PB_BLOCK>(acc, ws, outs, map);
}

// grad[k*nb + j] = sum_i A_k[i, j] * r[k*mb + i]   (warp per column)
template <typename T>
__global__ void __launch_bounds__(PB_BLOCK) k_gemv_t(const T* __restrict__ A, int64_t lda, int64_t blk_stride,
                                                     const T* __restrict__ r, T* __restrict__ grad, int64_t mb,
                                                     int64_t nb, int64_t nblk) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * PB_BLOCK + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * PB_BLOCK) >> 5;
  const int64_t ncols = nb * nblk;
  for (int64_t col = warp0; col < ncols; col += nwarps) {
    const int64_t k = col / nb, j = col - k * nb;
    const T* __restrict__ a = A + k * blk_stride + j * lda;
    const T* __restrict__ rk = r + k * mb;
    T acc = T(0);
    int64_t i = lane;
    // tall columns: 16 loads per lane in flight (2 KB per warp) -- with 4 the kernel was latency bound at ~0.45 of the HBM peak on a
    // 10000-row column; the FMA order (rows lane, lane + 32, ... ascending) is unchanged
    for (; i + 15 * 32 < mb; i += 16 * 32) {
      T a_[16], r_[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        a_[u] = __ldg(a + i + u * 32);
        r_[u] = __ldg(rk + i + u * 32);
      }
#pragma unroll
      for (int u = 0; u < 16; ++u) acc = fma(a_[u], r_[u], acc);
    }
    for (; i + 3 * 32 < mb; i += 4 * 32) {
      const T a0 = __ldg(a + i), a1 = __ldg(a + i + 32), a2 = __ldg(a + i + 64), a3 = __ldg(a + i + 96);
      const T r0 = __ldg(rk + i), r1 = __ldg(rk + i + 32), r2 = __ldg(rk + i + 64), r3 = __ldg(rk + i + 96);
      acc = fma(a0, r0, acc);
      acc = fma(a1, r1, acc);
      acc = fma(a2, r2, acc);
      acc = fma(a3, r3, acc);
    }
    for (; i < mb; i += 32) acc = fma(__ldg(a + i), __ldg(rk + i), acc);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) grad[col] = acc;
  }
}

// grad = A' r for SHORT columns (mb <= 32 lanes x 4 packs): LPC lanes share one column, each lane owns up to KP 16-byte
// packs of it (rows lane*VEC + k*LPC*VEC), the matching packs of r live in registers for the whole column chunk, a warp
// handles 32/LPC columns per sweep, log2(LPC) shuffle steps finish a column.  One CTA = one chunk of columns of ONE block.
template <typename T, int LPC, int KP, int SW>
__global__ void __launch_bounds__(PB_BLOCK) k_gemv_t_sub(const T* __restrict__ A, int64_t lda, int64_t blk_stride,
                                                         const T* __restrict__ r, T* __restrict__ grad, int64_t mb,
                                                         int64_t nb, int64_t chunk_cols, int64_t chunks_per_blk) {
  constexpr int VEC = 16 / sizeof(T);
  constexpr int CPW = 32 / LPC;                     // columns per warp sweep
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane % LPC, colw = lane / LPC;
  const int64_t k = blockIdx.x / chunks_per_blk;
  const int64_t c0 = (blockIdx.x % chunks_per_blk) * chunk_cols;
  int64_t c1 = c0 + chunk_cols;
  if (c1 > nb) c1 = nb;
  const T* __restrict__ Ak = A + k * blk_stride;
  const T* __restrict__ rk = r + k * mb;
  const int64_t npk = mb / VEC;                     // packs per column (mb % VEC == 0 checked by the launcher)
  Pack<T, VEC> rv[KP];
#pragma unroll
  for (int q = 0; q < KP; ++q) {
    const int64_t pk = sub + (int64_t)q * LPC;
    if (pk < npk)
      rv[q] = *reinterpret_cast<const Pack<T, VEC>*>(rk + pk * VEC);
    else
#pragma unroll
      for (int e = 0; e < VEC; ++e) rv[q].v[e] = T(0);
  }
  constexpr int WARPS = PB_BLOCK / 32;
  // SW column sweeps in flight per thread (independent 16-byte loads): 4 for very short columns (KP <= 2, few registers per
  // column), 1 for the KP = 4 shapes tuned in profiles/r01_tune_lsq.md
  const int64_t stride = (int64_t)WARPS * CPW;
  for (int64_t j0 = c0 + (int64_t)warp * CPW + colw; j0 < c1 + colw; j0 += SW * stride) {
    // (loop bound padded by colw so that all lanes of a warp iterate together; inactive columns contribute nothing)
    Pack<T, VEC> av[SW][KP];
    bool live[SW];
#pragma unroll
    for (int s_ = 0; s_ < SW; ++s_) {
      const int64_t j = j0 + s_ * stride;
      live[s_] = j < c1;
      const T* __restrict__ a = Ak + (live[s_] ? j : c0) * lda;
#pragma unroll
      for (int q = 0; q < KP; ++q) {
        const int64_t pk = sub + (int64_t)q * LPC;
        if (pk < npk) av[s_][q] = ld_pack<T, VEC, true>(a + pk * VEC);
      }
    }
#pragma unroll
    for (int s_ = 0; s_ < SW; ++s_) {
      T acc = T(0);
#pragma unroll
      for (int q = 0; q < KP; ++q) {
        const int64_t pk = sub + (int64_t)q * LPC;
        if (pk < npk) {
#pragma unroll
          for (int e = 0; e < VEC; ++e) acc = fma(av[s_][q].v[e], rv[q].v[e], acc);
        }
      }
#pragma unroll
      for (int off = LPC / 2; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
      if (live[s_] && sub == 0) grad[k * nb + j0 + s_ * stride] = acc;
    }
  }
}

// SquaredDistance: grad = x - b, AUX = ||x - b||^2  -> k_ew OP_SUB lives in step_kernels.cu (pb_sub); thin alias here.
extern "C" int pb_sqdist(pb_ctx* ctx, int dtype, int64_t n, const void* x, const void* b, void* grad) {
  return pb_sub(ctx, dtype, n, x, b, grad);
}

// Column shard of a dense A (C2): the chunking is the one of the GLOBAL matrix and the fold runs over all ranks' chunks.
struct LsqShard {
  int64_t n_global, col_offset;
  int aux_on;
};

template <typename T>
static int residual_t(pb_ctx* ctx, int64_t nblk, int64_t mb, int64_t nb, const T* A, int64_t lda, int64_t blk_stride,
                      const T* x, const T* b, T* r, const LsqShard* shard = nullptr) {
  const int64_t M = nblk * mb;
  if (M == 0) {
    // empty product: AUX = 0
    return pb_nrm2sq(ctx, sizeof(T) == 4 ? PB_F32 : PB_F64, 0, nullptr);
  }
  const int64_t row_tiles = (mb + GEMV_ROWS - 1) / GEMV_ROWS;
  // Column chunking is a function of the BLOCK SHAPE ONLY (never of nblk or the SM count): the summation order of every
  // r_i is then the same on any GPU and for any sharding of the blocks over ranks, so a block-sharded run reproduces the
  // single-GPU bits.  ceil(nb/64) columns per chunk, clamped to [32, 4096]  ->  <= 64 chunks (more only for nb > 262144).
  // (the rule itself lives in lsq_order.h, shared with the persistent driver kernel)
  PbLsqOrder ord = pb_lsq_order(sizeof(T), nblk, mb, nb, lda, blk_stride, A, nullptr);
  int64_t nch_global = 0, cbase = 0;
  if (shard) {
    // chunk columns of the GLOBAL matrix; this rank's columns must start on a chunk boundary and -- unless they are the last -- end on one
    const PbLsqOrder og = pb_lsq_order(sizeof(T), 1, mb, shard->n_global, lda, 0, A, nullptr);
    PB_REQUIRE((shard->col_offset % og.chunk_cols == 0 || (nb == 0 && shard->col_offset == shard->n_global)) &&
                   (shard->col_offset + nb == shard->n_global || nb % og.chunk_cols == 0),
               "column shards of a dense A must be aligned to the column chunks of the global matrix (pb_lsq_dense_chunk_cols)");
    PB_REQUIRE(ord.n_sub == og.n_sub, "shard and global matrix disagree on the residual order");
    ord.chunk_cols = og.chunk_cols;
    ord.nchunk = (nb + og.chunk_cols - 1) / og.chunk_cols;          // 0 for an empty shard
    nch_global = og.nchunk;
    cbase = (shard->col_offset + og.chunk_cols - 1) / og.chunk_cols;      // (= nch_global for an empty shard behind the last column)
    const size_t need = (size_t)2 * (size_t)nch_global * (size_t)M * (sizeof(T) / 4) * 8;
    if (need > PB_XCHG_VEC_BYTES) {
      pb_set_error("pb_lsq_dense_residual_sharded: %lld chunks x %lld rows exceed the vector exchange region", (long long)nch_global, (long long)M);
      return PB_EUNSUPPORTED;
    }
  }
  const int64_t chunk_cols = ord.chunk_cols, nchunk = ord.nchunk;
  PB_REQUIRE(nchunk <= 65535 && nch_global <= 65535, "too many column chunks for one launch");
  (void)row_tiles;
  PB_REQUIRE(nblk <= 65535, "too many blocks for one launch (nblk <= 65535)");
  int rc = pb_ensure_scratch(ctx, (size_t)(nchunk > 0 ? nchunk : 1) * M * sizeof(T));
  if (rc != PB_OK) return rc;
  T* partial = static_cast<T*>(ctx->scratch);
  if (nchunk == 0) {
    // empty shard: nothing to compute, but this rank still takes part in the fold
  } else if (ord.n_sub) {
    const int kp = ord.n_kp, lpc = ord.n_lpc;
#define PB_LAUNCH_NSUB(L, KP_)                                                                                              \
  k_gemv_n_sub<T, L, KP_><<<(unsigned)(nblk * nchunk), PB_BLOCK, 0, ctx->stream>>>(A, lda, blk_stride, x, partial, mb, nb, nblk, \
                                                                                  chunk_cols, nchunk)
    if (kp == 1) {
      PB_LAUNCH_NSUB(1, 1);
    } else if (kp == 2) {
      PB_LAUNCH_NSUB(1, 2);
    } else {
      switch (lpc) {
        case 1: PB_LAUNCH_NSUB(1, 4); break;
        case 2: PB_LAUNCH_NSUB(2, 4); break;
        case 4: PB_LAUNCH_NSUB(4, 4); break;
        default: PB_LAUNCH_NSUB(8, 4); break;
      }
    }
#undef PB_LAUNCH_NSUB
  } else if (ctx->gemv_scalar == 0 && mb % (16 / (int64_t)sizeof(T)) == 0 && lda % (16 / (int64_t)sizeof(T)) == 0 &&
             blk_stride % (16 / (int64_t)sizeof(T)) == 0 && M % (16 / (int64_t)sizeof(T)) == 0 && pb_aligned16(A) && pb_aligned16(partial)) {
    // 16-byte row packs (same order, bit-identical to the thread-per-row kernel below)
    constexpr int VEC = 16 / sizeof(T);
    const int64_t npk = mb / VEC;
    if (npk <= 32) {
      dim3 grid(1, (unsigned)nchunk, (unsigned)nblk);
      k_gemv_n_partial_v<T, 32><<<grid, 32 * GEMV_CL, 0, ctx->stream>>>(A, lda, blk_stride, x, partial, mb, nb, nblk, chunk_cols);
    } else {
      dim3 grid((unsigned)((npk + 127) / 128), (unsigned)nchunk, (unsigned)nblk);
      k_gemv_n_partial_v<T, 128><<<grid, 128 * GEMV_CL, 0, ctx->stream>>>(A, lda, blk_stride, x, partial, mb, nb, nblk, chunk_cols);
    }
  } else {
    dim3 grid((unsigned)row_tiles, (unsigned)nchunk, (unsigned)nblk);
    k_gemv_n_partial<T><<<grid, GEMV_ROWS * GEMV_CL, 0, ctx->stream>>>(A, lda, blk_stride, x, partial, mb, nb, nblk,
                                                                      chunk_cols);
  }
  if (nchunk > 0) PB_LAUNCH_CHECK(ctx);
  int cgrid = pb_stream_grid(ctx, PB_BLOCK, M, 2);
  if (shard && ctx->xchg_shared_device) {
    // peers on the same GPU: their (polling) kernels must fit beside this one
    const int cap = (2 * ctx->sm_count) / (ctx->xchg_world > 0 ? ctx->xchg_world : 1);
    if (cgrid > cap) cgrid = cap < 1 ? 1 : cap;
  }
  if (shard) {
    XchgVecParams xv;
    pb_xchg_vec_next(ctx, &xv);
    k_gemv_n_combine_x<T><<<cgrid, PB_BLOCK, 0, ctx->stream>>>(partial, (int)nchunk, (int)cbase, (int)nch_global, M, b, r, ctx->ws,
                                                               ctx->scalars_dev, xv, shard->aux_on);
  } else {
    k_gemv_n_combine<T><<<cgrid, PB_BLOCK, 0, ctx->stream>>>(partial, (int)nchunk, M, b, r, ctx->ws, ctx->scalars_dev);
  }
  PB_LAUNCH_CHECK(ctx);
  return PB_OK;
}

template <typename T>
static int gradient_t(pb_ctx* ctx, int64_t nblk, int64_t mb, int64_t nb, const T* A, int64_t lda, int64_t blk_stride,
                      const T* r, T* grad) {
  const int64_t ncols = nblk * nb;
  if (ncols == 0) return PB_OK;
  const PbLsqOrder ord = pb_lsq_order(sizeof(T), nblk, mb, nb, lda, blk_stride, A, r);
  if (ord.t_sub) {
    // packs per lane: 1 or 2 for very short columns (whole column in one lane, 4 columns in flight), else 4 with LPC lanes
    const int kp = ord.t_kp, lpc = ord.t_lpc;
    // chunk the columns of a block so that the grid covers the machine ~8x (~32x for the very short columns, whose CTAs
    // are cheap: a finer partition shortens the partial last wave)
    int64_t chunks = ((int64_t)ctx->sm_count * (kp <= 2 ? 32 : 8) + nblk - 1) / nblk;
    const int64_t min_chunk = kp <= 2 ? 2048 : 256;
    if (chunks > (nb + min_chunk - 1) / min_chunk) chunks = (nb + min_chunk - 1) / min_chunk;
    if (chunks < 1) chunks = 1;
    const int64_t chunk_cols = (nb + chunks - 1) / chunks;
    chunks = (nb + chunk_cols - 1) / chunk_cols;
    const int64_t grid = nblk * chunks;
    PB_REQUIRE(grid <= 0x7fffffffLL, "grid too large");
#define PB_LAUNCH_SUB(L, KP_, SW_)                                                                                 \
  k_gemv_t_sub<T, L, KP_, SW_><<<(unsigned)grid, PB_BLOCK, 0, ctx->stream>>>(A, lda, blk_stride, r, grad, mb, nb, chunk_cols, chunks)
    if (kp == 1) {
      PB_LAUNCH_SUB(1, 1, 4);
    } else if (kp == 2) {
      PB_LAUNCH_SUB(1, 2, 4);
    } else {
      switch (lpc) {
        case 1: PB_LAUNCH_SUB(1, 4, 1); break;
        case 2: PB_LAUNCH_SUB(2, 4, 1); break;
        case 4: PB_LAUNCH_SUB(4, 4, 1); break;
        case 8: PB_LAUNCH_SUB(8, 4, 1); break;
        case 16: PB_LAUNCH_SUB(16, 4, 1); break;
        default: PB_LAUNCH_SUB(32, 4, 1); break;
      }
    }
#undef PB_LAUNCH_SUB
    PB_LAUNCH_CHECK(ctx);
    return PB_OK;
  }
  const int grid = pb_stream_grid(ctx, PB_BLOCK / 32, ncols, 8);
  k_gemv_t<T><<<grid, PB_BLOCK, 0, ctx->stream>>>(A, lda, blk_stride, r, grad, mb, nb, nblk);
  PB_LAUNCH_CHECK(ctx);
  return PB_OK;
}

static int check_common(pb_ctx* ctx, int dtype) {
  PB_REQUIRE(ctx != nullptr, "null context");
  PB_REQUIRE(dtype == PB_F32 || dtype == PB_F64, "dtype must be PB_F32 or PB_F64");
  return PB_OK;
}

extern "C" int pb_lsq_dense_residual(pb_ctx* ctx, int dtype, int64_t m, int64_t n, const void* A, int64_t lda,
                                     const void* x, const void* b, void* r) {
  int rc = check_common(ctx, dtype);
  if (rc) return rc;
  PB_REQUIRE(m >= 0 && n >= 0 && lda >= m, "bad shape (need m, n >= 0 and lda >= m)");
  PB_REQUIRE(m == 0 || r != nullptr, "null output");
  PB_REQUIRE(m == 0 || n == 0 || (A && x), "null input");
  if (dtype == PB_F32)
    return residual_t<float>(ctx, 1, m, n, (const float*)A, lda, 0, (const float*)x, (const float*)b, (float*)r);
  return residual_t<double>(ctx, 1, m, n, (const double*)A, lda, 0, (const double*)x, (const double*)b, (double*)r);
}

// Column shard of a dense m x n_global matrix (this rank: columns [col_offset, col_offset + n_local), x its slice): r = A x - b,
// replicated bit-identically on every rank and equal to the single-GPU r; AUX = ||r||^2 (flags & 1: on rank 0 only, 0 elsewhere).
extern "C" int pb_lsq_dense_residual_sharded(pb_ctx* ctx, int dtype, int64_t m, int64_t n_local, const void* A, int64_t lda, const void* x,
                                             const void* b, void* r, int64_t n_global, int64_t col_offset, int flags) {
  int rc = check_common(ctx, dtype);
  if (rc) return rc;
  PB_REQUIRE(m >= 1 && n_local >= 0 && lda >= m, "bad shape (need m >= 1, n_local >= 0 and lda >= m)");
  PB_REQUIRE(n_global >= 1 && col_offset >= 0 && col_offset + n_local <= n_global, "shard outside the matrix");
  PB_REQUIRE(r != nullptr && (n_local == 0 || (A && x)), "null argument");
  PB_REQUIRE(ctx->xchg_world >= 1 && ctx->xchg_connected, "device exchange not initialised / connected (pb_xchg_init, pb_xchg_connect)");
  if (ctx->xchg_world == 1) {
    PB_REQUIRE(n_local == n_global && col_offset == 0, "one rank holds the whole matrix");
    return pb_lsq_dense_residual(ctx, dtype, m, n_local, A, lda, x, b, r);
  }
  LsqShard sh;
  sh.n_global = n_global;
  sh.col_offset = col_offset;
  sh.aux_on = (flags & 1) ? (ctx->xchg_rank == 0) : 1;
  if (dtype == PB_F32)
    return residual_t<float>(ctx, 1, m, n_local, (const float*)A, lda, 0, (const float*)x, (const float*)b, (float*)r, &sh);
  return residual_t<double>(ctx, 1, m, n_local, (const double*)A, lda, 0, (const double*)x, (const double*)b, (double*)r, &sh);
}

// Columns per chunk of the residual order for an m x n matrix: shard boundaries of a column-sharded dense A must be multiples of it.
extern "C" int64_t pb_lsq_dense_chunk_cols(int dtype, int64_t m, int64_t n) {
  const PbLsqOrder o = pb_lsq_order(dtype == PB_F32 ? 4 : 8, 1, m, n, m, 0, nullptr, nullptr);
  return o.chunk_cols;
}

extern "C" int pb_lsq_dense_gradient(pb_ctx* ctx, int dtype, int64_t m, int64_t n, const void* A, int64_t lda,
                                     const void* r, void* grad) {
  int rc = check_common(ctx, dtype);
  if (rc) return rc;
  PB_REQUIRE(m >= 0 && n >= 0 && lda >= m, "bad shape (need m, n >= 0 and lda >= m)");
  PB_REQUIRE(n == 0 || grad != nullptr, "null output");
  PB_REQUIRE(m == 0 || n == 0 || (A && r), "null input");
  if (dtype == PB_F32) return gradient_t<float>(ctx, 1, m, n, (const float*)A, lda, 0, (const float*)r, (float*)grad);
  return gradient_t<double>(ctx, 1, m, n, (const double*)A, lda, 0, (const double*)r, (double*)grad);
}

extern "C" int pb_lsq_blockdiag_residual(pb_ctx* ctx, int dtype, int64_t nblk, int64_t mb, int64_t nb, const void* A,
                                         const void* x, const void* b, void* r) {
  int rc = check_common(ctx, dtype);
  if (rc) return rc;
  PB_REQUIRE(nblk >= 0 && mb >= 0 && nb >= 0, "bad shape");
  PB_REQUIRE(nblk * mb == 0 || r != nullptr, "null output");
  PB_REQUIRE(nblk * mb * nb == 0 || (A && x), "null input");
  if (dtype == PB_F32)
    return residual_t<float>(ctx, nblk, mb, nb, (const float*)A, mb, mb * nb, (const float*)x, (const float*)b, (float*)r);
  return residual_t<double>(ctx, nblk, mb, nb, (const double*)A, mb, mb * nb, (const double*)x, (const double*)b,
                            (double*)r);
}

extern "C" int pb_lsq_blockdiag_gradient(pb_ctx* ctx, int dtype, int64_t nblk, int64_t mb, int64_t nb, const void* A,
                                         const void* r, void* grad) {
  int rc = check_common(ctx, dtype);
  if (rc) return rc;
  PB_REQUIRE(nblk >= 0 && mb >= 0 && nb >= 0, "bad shape");
  PB_REQUIRE(nblk * nb == 0 || grad != nullptr, "null output");
  PB_REQUIRE(nblk * mb * nb == 0 || (A && r), "null input");
  if (dtype == PB_F32)
    return gradient_t<float>(ctx, nblk, mb, nb, (const float*)A, mb, mb * nb, (const float*)r, (float*)grad);
  return gradient_t<double>(ctx, nblk, mb, nb, (const double*)A, mb, mb * nb, (const double*)r, (double*)grad);
}
