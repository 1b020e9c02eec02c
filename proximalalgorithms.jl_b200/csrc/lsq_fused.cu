// lsq_fused.cu -- value AND gradient of the block-diagonal least-squares term with every block of A read from HBM ONCE.
//
//   f(x) = 0.5 * sum_k ||A_k x_k - b_k||^2,   grad_k = A_k' (A_k x_k - b_k)          (benchmark/benchmarks.jl:11-17 per block;
//                                                                                     BASELINE.json configs[1]: 100 blocks of 100 x 1e5 fp32)
// The gradient of block k needs the whole residual r_k, i.e. a second sweep over A_k.  pb_lsq_blockdiag_residual followed by
// pb_lsq_blockdiag_gradient sweeps all of A (4 GB) twice from HBM.  A block is 40 MB and the L2 holds ~96 MB of re-readable data
// (tools/l2_capacity.py), so here the second sweep of a block follows its first one closely enough to be served by L2:
//
//   * ONE persistent cooperative kernel; work units are (block k, column chunk c) of two kinds, handed out dynamically:
//       N(k, c): partial[c][k][:] = A_k[:, chunk c] x_k[chunk c]              (streams the chunk from HBM)
//       T(k, c): grad[chunk c of block k] = A_k[:, chunk c]' r_k                (streams the same bytes again, from L2)
//     The CTA that completes the last N unit of block k assembles r_k = (sum of the chunk partials, in chunk order) - b_k and its
//     share of ||r||^2 and publishes `ready[k]`; T units of ready blocks are taken first; the N unit (k, c) may only start once c + 1
//     T units of block k - LIVE are done, which keeps ~LIVE blocks (80 MB) between the two sweeps.
//   * a chunk of a column-major block is ONE contiguous range of memory: every unit streams it through a shared-memory ring of
//     `cp.async.bulk` tiles (1-D TMA, mbarrier tracked; 32 columns per tile, up to 8 tiles in flight per CTA), so the bytes in flight
//     do not depend on registers and 128 concurrent N units saturate HBM.
//
// Summation orders are those of k_gemv_n_partial(_v) + k_gemv_n_combine and k_gemv_t_sub (lsq_kernels.cu, rule in lsq_order.h):
// r, grad and ||r||^2 are bit-identical to the two-kernel path (tests/test_gpu_lsq.py).
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "lsq_order.h"
#include "tma.cuh"

#define BF_BLOCK 256
#define BF_TILE_COLS 32
#define BF_MAX_STAGES 8

struct BfState {                 // global bookkeeping, zeroed before every launch
  unsigned int n_next;           // next N unit (block-major)
  unsigned int t_front;          // lowest block whose T units are not all done
  unsigned int exit_ticket;
  unsigned int pad[29];
  unsigned long long prof[8];    // ns summed over CTAs: 0 acquiring, 1 N units, 2 T units, 3 assembling r_k; 4 / 5: N / T units done
};
struct BfBlock {                 // per block
  unsigned int n_done, ready, t_next, t_done;
  double sumsq_hi, sumsq_lo;
};

struct BfParams {
  const void* A;
  const void* x;
  const void* b;
  void* r;
  void* grad;
  void* partial;
  int64_t nblk, mb, nb, chunk_cols;
  int nchunk, t_lpc, t_kp, stages, live;
  BfState* st;
  BfBlock* blk;
  PbWorkspace* ws;
  double* out;
};

__device__ __forceinline__ unsigned int bf_ld_acq(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void bf_st_rel(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

enum { BF_UNIT_N = 0, BF_UNIT_T = 1, BF_UNIT_EXIT = 2 };

// thread 0: take the next work unit (T units of ready blocks first, then the next N unit if its throttle allows).  `tf` is this CTA's
// cached lower bound of the first block with unclaimed T units; a claim costs one atomic round trip in the common case.
__device__ int bf_acquire(const BfParams& p, unsigned int& tf, int* k_out, int* c_out) {
  const unsigned int NU = (unsigned int)p.nchunk;
  const unsigned int nblk = (unsigned int)p.nblk;
  const unsigned int total = nblk * NU;
  for (;;) {
    // T units: the oldest block that still has some
    while (tf < nblk && bf_ld_acq(&p.blk[tf].ready)) {
      const unsigned int c = atomicAdd(&p.blk[tf].t_next, 1u);
      if (c < NU) {
        *k_out = (int)tf;
        *c_out = (int)c;
        return BF_UNIT_T;
      }
      tf += 1;                                   // every T unit of this block is taken
    }
    const unsigned int idx = bf_ld_acq(&p.st->n_next);
    if (idx < total) {
      const unsigned int k = idx / NU, c = idx % NU;
      if (k < (unsigned int)p.live || bf_ld_acq(&p.blk[k - p.live].t_done) >= c + 1u) {
        if (atomicCAS(&p.st->n_next, idx, idx + 1u) == idx) {
          *k_out = (int)k;
          *c_out = (int)c;
          return BF_UNIT_N;
        }
        continue;
      }
    } else if (tf >= nblk) {
      return BF_UNIT_EXIT;
    }
    __nanosleep(100);
  }
}

template <typename T>
__global__ void __launch_bounds__(BF_BLOCK, 2) k_bd_fused(BfParams p) {
  constexpr bool COMP = sizeof(T) == 8;
  constexpr int VEC = 16 / sizeof(T);
  extern __shared__ __align__(128) unsigned char bf_smem[];
  __shared__ uint64_t full[BF_MAX_STAGES];
  __shared__ int ctl[4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t mb = p.mb, nb = p.nb;
  const int npk = (int)(mb / VEC);
  const uint32_t col_bytes = (uint32_t)(mb * sizeof(T));
  const uint32_t tile_bytes = BF_TILE_COLS * col_bytes;
  T* ring = reinterpret_cast<T*>(bf_smem);                                        // [stages][BF_TILE_COLS][mb]
  T* xs = ring + (size_t)p.stages * BF_TILE_COLS * mb;                            // [chunk_cols]
  Pack<T, VEC>* lp = reinterpret_cast<Pack<T, VEC>*>(xs + ((p.chunk_cols + 3) & ~(int64_t)3));   // [4][npk] lane partials
  const T* __restrict__ A = static_cast<const T*>(p.A);
  const T* __restrict__ x = static_cast<const T*>(p.x);
  const T* __restrict__ bvec = static_cast<const T*>(p.b);
  T* __restrict__ r = static_cast<T*>(p.r);
  T* __restrict__ grad = static_cast<T*>(p.grad);
  T* __restrict__ partial = static_cast<T*>(p.partial);
  const int64_t M = p.nblk * mb;

  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  unsigned long long it_global = 0;              // tiles consumed so far by this CTA: stage = it % stages, parity = (it / stages) & 1
  unsigned int tf_cache = 0;                     // thread 0: first block that may still have unclaimed T units

  for (;;) {
    unsigned long long t_a = 0;
    if (tid == 0) {
      int k_, c_;
      t_a = globaltimer_ns();
      ctl[0] = bf_acquire(p, tf_cache, &k_, &c_);
      ctl[1] = k_;
      ctl[2] = c_;
      atomicAdd(&p.st->prof[0], globaltimer_ns() - t_a);
      t_a = globaltimer_ns();
    }
    __syncthreads();
    const int type = ctl[0], k = ctl[1], c = ctl[2];
    if (type == BF_UNIT_EXIT) break;
    const int64_t c0 = (int64_t)c * p.chunk_cols;
    int64_t c1 = c0 + p.chunk_cols;
    if (c1 > nb) c1 = nb;
    const int ncols = (int)(c1 - c0);
    const int ntile = (ncols + BF_TILE_COLS - 1) / BF_TILE_COLS;
    const T* __restrict__ src = A + ((int64_t)k * nb + c0) * mb;                  // the chunk: ncols * mb contiguous elements
    const unsigned long long g0 = it_global;      // global index of this unit's tile 0
    auto issue = [&](int t) {                     // thread 0: bulk-load tile t of this unit into its stage
      const int s = (int)((g0 + (unsigned long long)t) % (unsigned long long)p.stages);
      const int tc = ncols - t * BF_TILE_COLS < BF_TILE_COLS ? ncols - t * BF_TILE_COLS : BF_TILE_COLS;
      const uint32_t bytes = (uint32_t)tc * col_bytes;
      mbar_expect_tx(&full[s], bytes);
      bulk_g2s(ring + (size_t)s * BF_TILE_COLS * mb, src + (int64_t)t * BF_TILE_COLS * mb, bytes, &full[s]);
    };
    if (tid == 0) {
      const int pre = ntile < p.stages ? ntile : p.stages;
      for (int t = 0; t < pre; ++t) issue(t);
    }

    if (type == BF_UNIT_N) {
      // ---- partial[c][k][:] = A_k[:, chunk] x_k[chunk]: thread (pk, cl) owns row pack pk and the columns c0 + cl, c0 + cl + 4, ...
      const T* __restrict__ xk = x + (int64_t)k * nb + c0;
      for (int j = tid; j < ncols; j += BF_BLOCK) xs[j] = __ldg(xk + j);
      __syncthreads();
      const bool active = tid < npk * 4;
      const int pk = active ? tid % npk : 0, cl = active ? tid / npk : 0;
      Pack<T, VEC> acc;
#pragma unroll
      for (int e = 0; e < VEC; ++e) acc.v[e] = T(0);
      for (int t = 0; t < ntile; ++t) {
        const int s = (int)(it_global % (unsigned long long)p.stages);
        mbar_wait(&full[s], (uint32_t)((it_global / (unsigned long long)p.stages) & 1ull));
        const T* tile = ring + (size_t)s * BF_TILE_COLS * mb;
        const int tc = ncols - t * BF_TILE_COLS < BF_TILE_COLS ? ncols - t * BF_TILE_COLS : BF_TILE_COLS;
        if (active) {
#pragma unroll 4
          for (int jj = cl; jj < tc; jj += 4) {
            const Pack<T, VEC> a = *reinterpret_cast<const Pack<T, VEC>*>(tile + (size_t)jj * mb + pk * VEC);
            const T xv = xs[t * BF_TILE_COLS + jj];
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc.v[e] = fma(a.v[e], xv, acc.v[e]);
          }
        }
        __syncthreads();                          // every thread is done with stage s
        if (tid == 0 && t + p.stages < ntile) issue(t + p.stages);
        it_global += 1;
      }
      if (active) lp[cl * npk + pk] = acc;
      __syncthreads();
      if (tid < npk) {
        Pack<T, VEC> s_ = lp[tid];
#pragma unroll
        for (int q = 1; q < 4; ++q)
#pragma unroll
          for (int e = 0; e < VEC; ++e) s_.v[e] += lp[q * npk + tid].v[e];
        Pack<T, VEC>* dst = reinterpret_cast<Pack<T, VEC>*>(partial + ((int64_t)c * p.nblk + k) * mb + tid * VEC);
        __stcg(reinterpret_cast<float4*>(dst), *reinterpret_cast<const float4*>(&s_));
      }
      __syncthreads();
      if (tid == 0) {
        __threadfence();
        ctl[3] = (atomicAdd(&p.blk[k].n_done, 1u) + 1u == (unsigned int)p.nchunk) ? 1 : 0;
        atomicAdd(&p.st->prof[1], globaltimer_ns() - t_a);
        atomicAdd(&p.st->prof[4], 1ull);
        t_a = globaltimer_ns();
      }
      __syncthreads();
      if (ctl[3]) {
        // ---- last N unit of block k: r_k = (chunk partials in chunk order) - b_k, this block's share of ||r||^2, then `ready`
        // (on the critical path of the whole pipeline: the nchunk x mb partials are first pulled from L2 by ALL threads with independent
        //  loads into the idle ring, then each row is summed in chunk order from shared memory -- not nchunk dependent L2 round trips)
        __threadfence();
        T* stage = ring;                          // every tile of this unit has been consumed and nothing is in flight
        const int npart = p.nchunk * (int)mb;
        const bool staged = (size_t)npart * sizeof(T) <= (size_t)p.stages * tile_bytes;
        if (staged) {
          for (int q = tid; q < npart; q += BF_BLOCK) {
            const int cc = q / (int)mb, i = q - cc * (int)mb;
            stage[q] = __ldcg(partial + (int64_t)cc * M + (int64_t)k * mb + i);
          }
          __syncthreads();
        }
        Acc<1, 1> a1;
        a1.clear();
        for (int64_t i = tid; i < mb; i += BF_BLOCK) {
          const int64_t gi = (int64_t)k * mb + i;
          T s_;
          if (staged) {
            s_ = stage[i];
            for (int cc = 1; cc < p.nchunk; ++cc) s_ += stage[(size_t)cc * mb + i];
          } else {
            s_ = __ldcg(partial + gi);
            for (int cc = 1; cc < p.nchunk; ++cc) s_ += __ldcg(partial + (int64_t)cc * M + gi);
          }
          const T rv = bvec ? sub_rn(s_, __ldg(bvec + gi)) : s_;
          __stcg(r + gi, rv);
          if (COMP)
            dd_add_prod(a1.s[0], (double)rv, (double)rv);
          else
            a1.s[0].hi = __fma_rn((double)rv, (double)rv, a1.s[0].hi);
        }
        block_reduce<1, 1, BF_BLOCK>(a1);
        if (tid == 0) {
          __stcg(&p.blk[k].sumsq_hi, a1.s[0].hi);
          __stcg(&p.blk[k].sumsq_lo, a1.s[0].lo);
        }
        __syncthreads();
        if (tid == 0) {
          __threadfence();
          bf_st_rel(&p.blk[k].ready, 1u);
          atomicAdd(&p.st->prof[3], globaltimer_ns() - t_a);
        }
      }
    } else {
      // ---- grad[chunk of block k] = A_k[:, chunk]' r_k: LPC lanes share a column, each owning up to KP 16-byte packs (k_gemv_t_sub order)
      const int lpc = p.t_lpc, kp = p.t_kp;
      const int cpw = 32 / lpc, sub = lane % lpc, colw = lane / lpc;
      const T* __restrict__ rk = r + (int64_t)k * mb;
      Pack<T, VEC> rv[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int pkq = sub + q * lpc;
        if (q < kp && pkq < npk) {
          const float4 v = __ldcg(reinterpret_cast<const float4*>(rk + pkq * VEC));
          rv[q] = *reinterpret_cast<const Pack<T, VEC>*>(&v);
        } else {
#pragma unroll
          for (int e = 0; e < VEC; ++e) rv[q].v[e] = T(0);
        }
      }
      T* __restrict__ gk = grad + (int64_t)k * nb + c0;
      for (int t = 0; t < ntile; ++t) {
        const int s = (int)(it_global % (unsigned long long)p.stages);
        mbar_wait(&full[s], (uint32_t)((it_global / (unsigned long long)p.stages) & 1ull));
        const T* tile = ring + (size_t)s * BF_TILE_COLS * mb;
        const int tc = ncols - t * BF_TILE_COLS < BF_TILE_COLS ? ncols - t * BF_TILE_COLS : BF_TILE_COLS;
        for (int cb = warp * cpw; cb < tc; cb += (BF_BLOCK / 32) * cpw) {
          const int col = cb + colw;
          const bool live = col < tc;
          const T* a = tile + (size_t)(live ? col : 0) * mb;
          T acc = T(0);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int pkq = sub + q * lpc;
            if (q < kp && pkq < npk) {
              const Pack<T, VEC> av = *reinterpret_cast<const Pack<T, VEC>*>(a + pkq * VEC);
#pragma unroll
              for (int e = 0; e < VEC; ++e) acc = fma(av.v[e], rv[q].v[e], acc);
            }
          }
          for (int off = lpc >> 1; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
          if (live && sub == 0) gk[t * BF_TILE_COLS + col] = acc;
        }
        __syncthreads();
        if (tid == 0 && t + p.stages < ntile) issue(t + p.stages);
        it_global += 1;
      }
      if (tid == 0) {
        atomicAdd(&p.st->prof[2], globaltimer_ns() - t_a);
        atomicAdd(&p.st->prof[5], 1ull);
        const unsigned int d = atomicAdd(&p.blk[k].t_done, 1u) + 1u;
        if (d == (unsigned int)p.nchunk) {        // block k is finished: move the front past every finished block
          unsigned int tf = bf_ld_acq(&p.st->t_front);
          while (tf < (unsigned int)p.nblk && bf_ld_acq(&p.blk[tf].t_done) == (unsigned int)p.nchunk) {
            atomicCAS(&p.st->t_front, tf, tf + 1u);
            tf = bf_ld_acq(&p.st->t_front);
          }
        }
      }
    }
  }

  // ---- AUX = ||r||^2 (double-double over the blocks), by the last CTA to leave
  __shared__ bool is_last;
  if (tid == 0) {
    __threadfence();
    is_last = atomicAdd(&p.st->exit_ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  Acc<1, 1> a2;
  a2.clear();
  for (int64_t kk = tid; kk < p.nblk; kk += BF_BLOCK) {
    dd o;
    o.hi = __ldcg(&p.blk[kk].sumsq_hi);
    o.lo = __ldcg(&p.blk[kk].sumsq_lo);
    a2.s[0] = dd_sum(a2.s[0], o);
  }
  block_reduce<1, 1, BF_BLOCK>(a2);
  if (tid == 0) {
    p.out[PB_S_AUX] = a2.s[0].hi;
    p.out[PB_S_AUX + 1] = a2.s[0].lo;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
static int bd_fused_t(pb_ctx* ctx, int64_t nblk, int64_t mb, int64_t nb, const T* A, const T* x, const T* b, T* r, T* grad, bool* done) {
  *done = false;
  constexpr int VEC = 16 / sizeof(T);
  if (ctx->lsq_fused < 0 || nblk < 1 || mb < 1 || nb < 1) return PB_OK;
  const PbLsqOrder ord = pb_lsq_order(sizeof(T), nblk, mb, nb, mb, mb * nb, A, r);
  const int64_t npk = mb / VEC;
  const double blk_bytes = (double)mb * (double)nb * sizeof(T), tot_bytes = blk_bytes * (double)nblk;
  // the two-kernel path is as good while everything fits L2 anyway; a block must leave room for LIVE = 2 of them in L2
  if (ord.n_sub || !ord.t_sub || mb % VEC != 0 || npk * 4 > BF_BLOCK || nblk > 0x3fffffLL / ord.nchunk) return PB_OK;
  // MEASURED (profiles/r02_lsq_fused.md): this form is 3-9x SLOWER than the two kernels on configs[1] -- one CTA streams ~20 GB/s through
  // its ring whether the bytes come from HBM or L2, the summation-order rule gives 64 independent units per block, and only ~2 blocks
  // (80 MB) may sit between the sweeps, so ~128 of the 296 CTAs have work at any time.  It is therefore never chosen automatically
  // (PB_OPT_LSQ_FUSED = k > 0 forces it; the tests do, for the bit-parity of its bookkeeping); the single-sweep answer for fixed-stepsize
  // FISTA is lsq_fista.cu, which needs no second sweep at all.
  (void)tot_bytes;
  (void)blk_bytes;
  if (ctx->lsq_fused == 0) return PB_OK;
  if (!pb_aligned16(A) || !pb_aligned16(r) || !pb_aligned16(x)) return PB_OK;
  const size_t col_bytes = (size_t)mb * sizeof(T);
  const size_t tile_bytes = BF_TILE_COLS * col_bytes;
  int stages = (int)((size_t)80 * 1024 / tile_bytes);
  if (stages > BF_MAX_STAGES) stages = BF_MAX_STAGES;
  if (stages < 2) return PB_OK;
  const size_t smem = (size_t)stages * tile_bytes + (((size_t)ord.chunk_cols + 3) & ~(size_t)3) * sizeof(T) + (size_t)4 * npk * 16 + 128;
  if (smem > 110 * 1024) return PB_OK;
  PB_CHECK_CUDA(cudaSetDevice(ctx->device));
  // workspace: chunk partials [nchunk][nblk][mb] | BfState | BfBlock[nblk]
  const size_t part_bytes = ((size_t)ord.nchunk * nblk * mb * sizeof(T) + 255) & ~(size_t)255;
  const size_t st_bytes = sizeof(BfState) + (size_t)nblk * sizeof(BfBlock);
  int rc = pb_ensure_scratch(ctx, part_bytes + st_bytes);
  if (rc != PB_OK) return rc;
  unsigned char* wsb = static_cast<unsigned char*>(ctx->scratch);
  PB_CHECK_CUDA(cudaMemsetAsync(wsb + part_bytes, 0, st_bytes, ctx->stream));
  BfParams p;
  memset(&p, 0, sizeof(p));
  p.A = A;
  p.x = x;
  p.b = b;
  p.r = r;
  p.grad = grad;
  p.partial = wsb;
  p.nblk = nblk;
  p.mb = mb;
  p.nb = nb;
  p.chunk_cols = ord.chunk_cols;
  p.nchunk = (int)ord.nchunk;
  p.t_lpc = ord.t_lpc;
  p.t_kp = ord.t_kp;
  p.stages = stages;
  p.live = ctx->lsq_fused > 0 ? ctx->lsq_fused : 2;
  p.st = reinterpret_cast<BfState*>(wsb + part_bytes);
  p.blk = reinterpret_cast<BfBlock*>(wsb + part_bytes + sizeof(BfState));
  p.ws = ctx->ws;
  p.out = ctx->scalars_dev;
  auto kern = k_bd_fused<T>;
  static bool attr_done[PB_MAX_DEVICES][2] = {};
  const int dev = ctx->device < PB_MAX_DEVICES ? ctx->device : 0;
  if (!attr_done[dev][sizeof(T) == 8] || ctx->device >= PB_MAX_DEVICES) {
    PB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
    attr_done[dev][sizeof(T) == 8] = true;
  }
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, BF_BLOCK, smem) != cudaSuccess || occ < 1) occ = 1;
  if (occ > 2) occ = 2;
  int64_t grid = (int64_t)ctx->sm_count * occ;
  const int64_t units = 2 * nblk * ord.nchunk;
  if (grid > units) grid = units;
  void* args[] = {&p};
  PB_CHECK_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3((unsigned)grid), dim3(BF_BLOCK), args, smem, ctx->stream));
  ctx->launches++;
  *done = true;
  if (getenv("PROXB200_LSQ_FUSED_PROF")) {       // diagnostics: where the CTAs spent their time (ns summed over CTAs)
    unsigned long long prof[8];
    PB_CHECK_CUDA(cudaMemcpyAsync(prof, &p.st->prof[0], sizeof(prof), cudaMemcpyDeviceToHost, ctx->stream));
    PB_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    fprintf(stderr, "lsq_fused: grid %lld stages %d live %d | per CTA us: acquire %.1f  N %.1f  T %.1f  assemble %.1f | N units %llu (%.1f us each)  T units %llu (%.1f us each)\n",
            (long long)grid, stages, p.live, prof[0] / 1e3 / grid, prof[1] / 1e3 / grid, prof[2] / 1e3 / grid, prof[3] / 1e3 / grid, prof[4],
            prof[4] ? prof[1] / 1e3 / prof[4] : 0.0, prof[5], prof[5] ? prof[2] / 1e3 / prof[5] : 0.0);
  }
  return PB_OK;
}

extern "C" int pb_lsq_blockdiag_value_and_gradient(pb_ctx* ctx, int dtype, int64_t nblk, int64_t mb, int64_t nb, const void* A,
                                                   const void* x, const void* b, void* r, void* grad) {
  PB_REQUIRE(ctx != nullptr, "null context");
  PB_REQUIRE(dtype == PB_F32 || dtype == PB_F64, "dtype must be PB_F32 or PB_F64");
  PB_REQUIRE(nblk >= 0 && mb >= 0 && nb >= 0, "bad shape");
  PB_REQUIRE(nblk * mb == 0 || r != nullptr, "null output");
  PB_REQUIRE(nblk * nb == 0 || grad != nullptr, "null output");
  PB_REQUIRE(nblk * mb * nb == 0 || (A && x), "null input");
  bool done = false;
  int rc = dtype == PB_F32 ? bd_fused_t<float>(ctx, nblk, mb, nb, (const float*)A, (const float*)x, (const float*)b, (float*)r, (float*)grad, &done)
                           : bd_fused_t<double>(ctx, nblk, mb, nb, (const double*)A, (const double*)x, (const double*)b, (double*)r, (double*)grad, &done);
  if (rc != PB_OK || done) return rc;
  // not eligible (short columns, everything L2 resident, tiny problem ...): the two sweeps as separate kernels
  rc = pb_lsq_blockdiag_residual(ctx, dtype, nblk, mb, nb, A, x, b, r);
  if (rc != PB_OK) return rc;
  return pb_lsq_blockdiag_gradient(ctx, dtype, nblk, mb, nb, A, r, grad);
}
