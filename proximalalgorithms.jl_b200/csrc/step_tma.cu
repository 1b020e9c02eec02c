// step_tma.cu -- K1/K2 fused step as a TMA bulk-copy shared-memory ring (sm_100a).
//
// Why: the step streams 3 input and 2 output vectors with ~1 flop/byte.  The register (LDG) pipeline keeps bytes in flight
// by holding UNROLL x 3 packs per thread in registers; here the copy engine does it instead: one elected thread issues
// `cp.async.bulk` (1-D TMA, SASS UBLKCP) global -> shared for whole 4 KB tiles with an mbarrier tracking the bytes, all
// threads compute from shared memory, results go to a shared staging tile and leave with `cp.async.bulk` shared -> global.
// Loads of STAGES-1 tiles are in flight per CTA while the current tile is processed; registers hold only one pack.
//
// Per tile i (stage s = i % STAGES, output slot o = i % 3), ONE __syncthreads:
//   all      : wait full[s]  ->  compute in[s] -> out[o]  ->  fence.proxy.async
//   thread 0 : cp.async.bulk.wait_group.read 1      (stores <= i-2 have drained their staging tile => out[(i+1)%3] is free)
//   all      : __syncthreads
//   thread 0 : bulk-store out[o]; commit;  expect_tx(full[s]); bulk-load tile i+STAGES into in[s]
// Arithmetic and reductions are the same code as the register pipeline (step_common.cuh): results are bit-identical.
#include "step_common.cuh"
#include "tma.cuh"

#define TMA_BLOCK 256

template <typename T, int PROX, bool EXTRAP, int STAGES>
__global__ void __launch_bounds__(TMA_BLOCK) k_step_tma(StepParams p) {
  constexpr bool COMP = sizeof(T) == 8;
  constexpr int VEC = 16 / sizeof(T);
  constexpr int TILE = TMA_BLOCK * VEC;              // elements per tile and array: 4 KB
  constexpr int TILE_BYTES = TILE * sizeof(T);
  constexpr int NIN = EXTRAP ? 3 : 2;
  constexpr int NOUT = EXTRAP ? 2 : 1;
  constexpr int NSLOT = 3;                           // output staging slots

  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* in = reinterpret_cast<T*>(smem_raw);                                   // [STAGES][NIN][TILE]
  T* outb = in + (size_t)STAGES * NIN * TILE;                               // [NSLOT][NOUT][TILE]
  uint64_t* full = reinterpret_cast<uint64_t*>(outb + (size_t)NSLOT * NOUT * TILE);   // [STAGES]

  const T* __restrict__ gx = static_cast<const T*>(p.x);
  const T* __restrict__ gg = static_cast<const T*>(p.grad);
  const T* __restrict__ gzp = static_cast<const T*>(p.z_prev);
  T* __restrict__ gz = static_cast<T*>(p.z);
  T* __restrict__ gxn = static_cast<T*>(p.x_next);
  const T gamma = (T)p.gamma, beta = (T)p.beta, pa = (T)p.a, pb = (T)p.b;
  const int64_t n = p.n;
  const int64_t ntiles = n / TILE;
  const int tid = threadIdx.x;

  // tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
  const int64_t my_tiles = ntiles > blockIdx.x ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto issue_load = [&](int64_t it) {   // thread 0 only
    const int s = (int)(it % STAGES);
    const int64_t base = (blockIdx.x + it * gridDim.x) * (int64_t)TILE;
    T* dst = in + (size_t)s * NIN * TILE;
    mbar_expect_tx(&full[s], NIN * TILE_BYTES);
    bulk_g2s(dst, gx + base, TILE_BYTES, &full[s]);
    bulk_g2s(dst + TILE, gg + base, TILE_BYTES, &full[s]);
    if constexpr (EXTRAP) bulk_g2s(dst + 2 * TILE, gzp + base, TILE_BYTES, &full[s]);
  };

  if (tid == 0) {
    const int64_t pre = my_tiles < STAGES ? my_tiles : STAGES;
    for (int64_t it = 0; it < pre; ++it) issue_load(it);
  }

  Acc<3, 1> acc, pk;
  acc.clear();
  pk.clear();

  for (int64_t it = 0; it < my_tiles; ++it) {
    const int s = (int)(it % STAGES);
    const int o = (int)(it % NSLOT);
    mbar_wait(&full[s], (uint32_t)((it / STAGES) & 1));
    const T* sin = in + (size_t)s * NIN * TILE;
    T* sout = outb + (size_t)o * NOUT * TILE;
    const Pack<T, VEC> xv = *reinterpret_cast<const Pack<T, VEC>*>(sin + tid * VEC);
    const Pack<T, VEC> gv = *reinterpret_cast<const Pack<T, VEC>*>(sin + TILE + tid * VEC);
    Pack<T, VEC> zv;
    if constexpr (EXTRAP) zv = *reinterpret_cast<const Pack<T, VEC>*>(sin + 2 * TILE + tid * VEC);
    Pack<T, VEC> zn, xn;
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      T yv, rv;
      StepElem<T, PROX, EXTRAP>::template run<COMP>(xv.v[e], gv.v[e], EXTRAP ? zv.v[e] : T(0), pa, pb, gamma, beta, yv,
                                                    zn.v[e], rv, xn.v[e], COMP ? acc : pk);
    }
    if constexpr (!COMP) fold_pack<PROX>(acc, pk);
    *reinterpret_cast<Pack<T, VEC>*>(sout + tid * VEC) = zn;
    if constexpr (EXTRAP) *reinterpret_cast<Pack<T, VEC>*>(sout + TILE + tid * VEC) = xn;
    fence_proxy_async();
    if (tid == 0) bulk_wait_read<1>();
    __syncthreads();
    if (tid == 0) {
      const int64_t base = (blockIdx.x + it * gridDim.x) * (int64_t)TILE;
      bulk_s2g(gz + base, sout, TILE_BYTES);
      if constexpr (EXTRAP) bulk_s2g(gxn + base, sout + TILE, TILE_BYTES);
      bulk_commit();
      if (it + STAGES < my_tiles) issue_load(it + STAGES);
    }
  }

  // ragged tail (< TILE elements): plain path, spread over the grid, one 16-byte pack per thread and trip so that float sums are
  // grouped exactly as everywhere else (fold_pack); the last < VEC elements are single-element groups like in k_step
  for (int64_t q = ntiles * TMA_BLOCK + (int64_t)blockIdx.x * TMA_BLOCK + tid; q * VEC < n; q += (int64_t)gridDim.x * TMA_BLOCK) {
    const bool whole = q * VEC + VEC <= n;
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const int64_t i = q * VEC + e;
      if (i < n) {
        T yv, zn, rv, xn;
        StepElem<T, PROX, EXTRAP>::template run<COMP>(gx[i], gg[i], EXTRAP ? gzp[i] : T(0), pa, pb, gamma, beta, yv, zn, rv,
                                                      xn, COMP ? acc : pk);
        gz[i] = zn;
        if constexpr (EXTRAP) gxn[i] = xn;
        if constexpr (!COMP) {
          if (!whole) fold_pack<PROX>(acc, pk);
        }
      }
    }
    if constexpr (!COMP) {
      if (whole) fold_pack<PROX>(acc, pk);
    }
  }
  if (tid == 0) bulk_wait_all();   // staging tiles must outlive the last bulk stores

  OutMap map;
  map.sum_slot[0] = PB_S_GSUM;
  map.sum_slot[1] = PB_S_RESSQ;
  map.sum_slot[2] = PB_S_GDR;
  map.sum_slot[3] = -1;
  map.max_slot[0] = PB_S_RESINF;
  map.max_slot[1] = -1;
  grid_reduce<3, 1, TMA_BLOCK>(acc, p.ws, p.out, map, &p.xchg);
}

template <typename T, int PROX, bool EXTRAP>
static int launch_tma(pb_ctx* ctx, const StepParams& p) {
  constexpr int STAGES = 4;
  constexpr int VEC = 16 / sizeof(T);
  constexpr int TILE = TMA_BLOCK * VEC;
  constexpr int NIN = EXTRAP ? 3 : 2, NOUT = EXTRAP ? 2 : 1;
  constexpr size_t smem = (size_t)(STAGES * NIN + 3 * NOUT) * TILE * sizeof(T) + STAGES * sizeof(uint64_t);
  // cudaFuncSetAttribute is per device: one flag per device ordinal (a single process may drive several GPUs)
  static bool attr_done[PB_MAX_DEVICES] = {};
  auto kern = k_step_tma<T, PROX, EXTRAP, STAGES>;
  const int dev = ctx->device < PB_MAX_DEVICES ? ctx->device : 0;
  if (!attr_done[dev] || ctx->device >= PB_MAX_DEVICES) {
    PB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done[dev] = true;
  }
  const int grid = pb_stream_grid(ctx, TILE, p.n, EXTRAP ? 3 : 2);   // co-resident by construction (<= 72 KB smem per CTA)
  kern<<<grid, TMA_BLOCK, smem, ctx->stream>>>(p);
  PB_LAUNCH_CHECK(ctx);
  return PB_OK;
}

template <typename T, bool EXTRAP>
static int launch_tma_prox(pb_ctx* ctx, int prox_kind, const StepParams& p) {
  switch (prox_kind) {
    case PB_PROX_L1:
      return launch_tma<T, PB_PROX_L1, EXTRAP>(ctx, p);
    case PB_PROX_BOX:
      return launch_tma<T, PB_PROX_BOX, EXTRAP>(ctx, p);
    case PB_PROX_SCALE:
      return launch_tma<T, PB_PROX_SCALE, EXTRAP>(ctx, p);
    case PB_PROX_ZERO:
      return launch_tma<T, PB_PROX_ZERO, EXTRAP>(ctx, p);
    default:
      return PB_EUNSUPPORTED;
  }
}

int pb_launch_step_tma(pb_ctx* ctx, int dtype, int prox_kind, bool extrap, const StepParams& p) {
  if (p.y || p.res || p.lo_v || p.hi_v) return PB_EUNSUPPORTED;   // optional streams: register pipeline only
  if (dtype == PB_F32)
    return extrap ? launch_tma_prox<float, true>(ctx, prox_kind, p) : launch_tma_prox<float, false>(ctx, prox_kind, p);
  return extrap ? launch_tma_prox<double, true>(ctx, prox_kind, p) : launch_tma_prox<double, false>(ctx, prox_kind, p);
}
