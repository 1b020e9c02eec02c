// tv_kernels.cu -- K10: one Douglas-Rachford iteration (src/algorithms/douglas_rachford.jl:54-63) of anisotropic
// total-variation denoising   minimize 0.5*||u - b||^2 + lambda*( sum |u[i,j+1]-u[i,j]| + sum |u[i+1,j]-u[i,j]| )
// in product-space (consensus) form -- BASELINE.json configs[4]: 8192 x 8192 image, Float32, 2 x B200.
//
// Total variation is not in the reference (nor is any TV prox in its tests); the splitting is ours and the oracle for it is
// oracle/tv_oracle.py.  The objective is a sum of five terms whose proxes are closed-form and embarrassingly parallel:
//   f_0 = 0.5||u - b||^2                       prox: (x - b)*w + b with w = 1/(1 + gamma) rounded once (no per-pixel divide)
//   f_1 / f_2 = lambda * sum over the EVEN / ODD horizontal pairs (j, j+1) of |u_j+1 - u_j|
//   f_3 / f_4 = lambda * sum over the EVEN / ODD vertical pairs (i, i+1)
//   (pairs of one set are disjoint, so the prox acts on each pair alone: with d = a - c and t = gamma*lambda,
//    |d| <= 2t -> both become (a + c)/2, otherwise a - sign(d) t and c + sign(d) t)
// DR runs on X = (x_0..x_4) with F(X) = sum_k f_k(x_k) and G = indicator{x_0 = ... = x_4} (prox = average of the copies):
//   y_k = prox_{gamma f_k}(x_k);  r_k = 2 y_k - x_k;  z = (r_0 + ... + r_4)*0.2;  res_k = y_k - z;  x_k <- x_k - res_k.
//
// The kernel does the WHOLE iteration in one pass over the image: read the five copies and b, write the five copies
// (11 image passes = 44 B/pixel in Float32; the reference sequence of five broadcasts would move 5 x 13 vectors).  The pair
// partner of a pixel lives in the same 16-byte pack (even horizontal pairs), one element to the left / right (odd horizontal
// pairs) or one row up / down (vertical pairs): those neighbour loads hit L1/L2, not HBM.  x_out must not alias x (the
// partner of a pixel is read by another thread).  HBM-bound stencil; no tensor-core formulation exists.
//
// Multi-GPU (row shards): the only data a rank needs from a neighbour is ONE image row of ONE copy per shard boundary (the
// vertical pair that straddles it).  `halo_prev` / `halo_next` are pointers to that row inside the NEIGHBOUR's x buffer,
// mapped over NVLink with cudaIpc (pb_ipc_*): the kernel reads peer memory directly -- no halo exchange launch, no staging
// copy.  The per-iteration scalar exchange (stop norm) that the driver performs anyway is the inter-GPU barrier.
#include <string.h>

#include "common.cuh"

struct TvParams {
  const void* x;      // [5][H][W]
  void* x_out;        // [5][H][W]
  void* y;            // optional [5][H][W]
  void* z;            // optional [H][W]
  const void* b;      // [H][W]
  const void* halo_prev;   // row (row0 - 1) of the copy that pairs it with row0, or NULL
  const void* halo_next;   // row (row0 + H) of the copy that pairs it with row0 + H - 1, or NULL
  int64_t H, W, row0, Hglob;
  double t, w;        // gamma*lambda and 1/(1 + gamma), in the element type
  PbWorkspace* ws;
  double* outs;
  XchgParams xchg;
};

template <typename T>
__device__ __forceinline__ T tv_pair(T a, T c, T t, T t2) {
  const T d = sub_rn(a, c);
  if (fabs(d) <= t2) return mul_rn(add_rn(a, c), T(0.5));
  return sub_rn(a, copysign(t, d));
}

// MINB: CTAs per SM the register allocation must allow (the pass is latency-bound at 2: ncu shows 16 warps/SM stalled on the
// long scoreboard; 3-4 trade a few spills for more loads in flight)
template <typename T, int VEC, int MINB>
__global__ void __launch_bounds__(PB_BLOCK, MINB) k_dr_tv(TvParams p) {
  const T* __restrict__ x = static_cast<const T*>(p.x);
  const T* __restrict__ b = static_cast<const T*>(p.b);
  const T* __restrict__ hp = static_cast<const T*>(p.halo_prev);
  const T* __restrict__ hn = static_cast<const T*>(p.halo_next);
  T* __restrict__ xo = static_cast<T*>(p.x_out);
  T* __restrict__ yo = static_cast<T*>(p.y);
  T* __restrict__ zo = static_cast<T*>(p.z);
  const int64_t H = p.H, W = p.W, HW = H * W;
  const T t = (T)p.t, t2 = mul_rn(T(2), t), w = (T)p.w;
  T mx = T(0);
  const int64_t npacks = HW / VEC;              // W % VEC == 0 (launcher)
  for (int64_t q = (int64_t)blockIdx.x * PB_BLOCK + threadIdx.x; q < npacks; q += (int64_t)gridDim.x * PB_BLOCK) {
    const int64_t base = q * VEC;
    const int64_t i = base / W, j0 = base - i * W;
    const int64_t gi = p.row0 + i;
    Pack<T, VEC> xv[5], bv, pv3, pv4;
#pragma unroll
    for (int k = 0; k < 5; ++k) xv[k] = ld_pack<T, VEC, true>(x + k * HW + base);
    bv = ld_pack<T, VEC, true>(b + base);
    // vertical partners: copy 3 pairs (even row, even row + 1), copy 4 pairs (odd row, odd row + 1)
    const int64_t g3 = (gi & 1) ? gi - 1 : gi + 1, g4 = (gi & 1) ? gi + 1 : gi - 1;
    bool has3 = g3 >= 0 && g3 < p.Hglob, has4 = g4 >= 0 && g4 < p.Hglob;
    auto load_row = [&](int k, int64_t g, bool& has) -> Pack<T, VEC> {
      Pack<T, VEC> r;
      const int64_t il = g - p.row0;
      if (has && il >= 0 && il < H)
        r = ld_pack<T, VEC, false>(x + k * HW + il * W + j0);
      else if (has && il < 0 && hp)
        r = ld_pack<T, VEC, false>(hp + j0);
      else if (has && il >= H && hn)
        r = ld_pack<T, VEC, false>(hn + j0);
      else
        has = false;
      return r;
    };
    pv3 = load_row(3, g3, has3);
    pv4 = load_row(4, g4, has4);
    // horizontal partners outside the pack (odd pairs; every pair when VEC == 1)
    T left1 = T(0), right1 = T(0), left2 = T(0), right2 = T(0);
    const bool hasl = j0 > 0, hasr = j0 + VEC < W;
    if (VEC == 1) {
      if (hasl) left1 = x[1 * HW + base - 1];
      if (hasr) right1 = x[1 * HW + base + 1];
    }
    if (hasl) left2 = x[2 * HW + base - 1];
    if (hasr) right2 = x[2 * HW + base + VEC];
    Pack<T, VEC> yv[5], xn[5], zv;
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const int par = VEC == 1 ? (int)(j0 & 1) : (e & 1);       // parity of the column (j0 is even when VEC > 1)
      T y[5];
      y[0] = add_rn(mul_rn(sub_rn(xv[0].v[e], bv.v[e]), w), bv.v[e]);
      {  // even horizontal pairs: partner is column j ^ 1
        T c = T(0);
        bool has;
        if (VEC == 1) {
          has = par ? hasl : hasr;
          c = par ? left1 : right1;
        } else {
          has = true;
          c = xv[1].v[e ^ 1];
        }
        y[1] = has ? tv_pair(xv[1].v[e], c, t, t2) : xv[1].v[e];
      }
      {  // odd horizontal pairs: (odd column, odd column + 1)
        T c = T(0);
        bool has;
        if (par) {                       // partner is j + 1
          has = (e + 1 < VEC) ? true : hasr;
          c = (e + 1 < VEC) ? xv[2].v[(e + 1) % VEC] : right2;
        } else {                         // partner is j - 1
          has = (e > 0) ? true : hasl;
          c = (e > 0) ? xv[2].v[(e + VEC - 1) % VEC] : left2;
        }
        y[2] = has ? tv_pair(xv[2].v[e], c, t, t2) : xv[2].v[e];
      }
      y[3] = has3 ? tv_pair(xv[3].v[e], pv3.v[e], t, t2) : xv[3].v[e];
      y[4] = has4 ? tv_pair(xv[4].v[e], pv4.v[e], t, t2) : xv[4].v[e];
      T r[5];
#pragma unroll
      for (int k = 0; k < 5; ++k) r[k] = sub_rn(mul_rn(T(2), y[k]), xv[k].v[e]);
      const T zz = mul_rn(add_rn(add_rn(add_rn(add_rn(r[0], r[1]), r[2]), r[3]), r[4]), T(0.2));
      zv.v[e] = zz;
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const T res = sub_rn(y[k], zz);
        xn[k].v[e] = sub_rn(xv[k].v[e], res);
        yv[k].v[e] = y[k];
        const T ar = fabs(res);
        mx = (ar > mx || ar != ar) ? ar : mx;
      }
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      st_pack<T, VEC, true>(xo + k * HW + base, xn[k]);
      if (yo) st_pack<T, VEC, true>(yo + k * HW + base, yv[k]);
    }
    if (zo) st_pack<T, VEC, true>(zo + base, zv);
  }
  Acc<0, 1> acc;
  acc.clear();
  acc.m[0] = (double)mx;
  OutMap map;
  map.sum_slot[0] = map.sum_slot[1] = map.sum_slot[2] = map.sum_slot[3] = -1;
  map.max_slot[0] = PB_S_RESINF;
  map.max_slot[1] = -1;
  grid_reduce<0, 1, PB_BLOCK>(acc, p.ws, p.outs, map, &p.xchg);
}

extern "C" int pb_dr_tv_step(pb_ctx* ctx, int dtype, int64_t H, int64_t W, const void* x, const void* b, double gamma,
                             double lambda, void* x_out, void* y, void* z, int64_t row0, int64_t Hglob,
                             const void* halo_prev, const void* halo_next) {
  PB_REQUIRE(ctx != nullptr, "null context");
  PB_REQUIRE(dtype == PB_F32 || dtype == PB_F64, "dtype must be PB_F32 or PB_F64");
  PB_REQUIRE(H >= 0 && W >= 0 && row0 >= 0 && Hglob >= row0 + H, "bad shape (need 0 <= row0, row0 + H <= Hglob)");
  PB_REQUIRE(H * W == 0 || (x && b && x_out), "null image");
  PB_REQUIRE(H * W == 0 || x != x_out, "x_out must not alias x (pair partners are read by other threads)");
  PB_REQUIRE(gamma > 0 && lambda >= 0, "need gamma > 0 and lambda >= 0");
  TvParams p;
  p.x = x;
  p.x_out = x_out;
  p.y = y;
  p.z = z;
  p.b = b;
  p.halo_prev = halo_prev;
  p.halo_next = halo_next;
  p.H = H;
  p.W = W;
  p.row0 = row0;
  p.Hglob = Hglob;
  p.ws = ctx->ws;
  p.outs = ctx->scalars_dev;
  if (dtype == PB_F32) {
    p.t = (double)mul_rn_host((float)gamma, (float)lambda);
    volatile float den = 1.0f + (float)gamma;
    volatile float w = 1.0f / den;
    p.w = (double)w;
  } else {
    p.t = mul_rn_host(gamma, lambda);
    volatile double den = 1.0 + gamma;
    volatile double w = 1.0 / den;
    p.w = w;
  }
  pb_xchg_next(ctx, &p.xchg, ctx->xchg_fused != 0 && H * W > 0);
  const int vec = dtype == PB_F32 ? 4 : 2;
  const size_t es = dtype == PB_F32 ? 4 : 8;
  const void* ptrs[] = {x, x_out, y, z, b, halo_prev, halo_next};
  bool vec_ok = W % vec == 0 && ((size_t)H * W * es) % 16 == 0;
  for (const void* q : ptrs) vec_ok = vec_ok && (!q || pb_aligned16(q));
  const int64_t n = H * W;
  const int minb = ctx->ctas_per_sm == 2 || ctx->ctas_per_sm == 4 ? ctx->ctas_per_sm : 3;
  const int grid_v = (int)((int64_t)ctx->sm_count * minb < (n / (PB_BLOCK * vec) + 1) ? (int64_t)ctx->sm_count * minb
                                                                                         : (n / (PB_BLOCK * vec) + 1));
  if (dtype == PB_F32) {
    if (!vec_ok)
      k_dr_tv<float, 1, 2><<<pb_stream_grid(ctx, PB_BLOCK, n, 4), PB_BLOCK, 0, ctx->stream>>>(p);
    else if (minb == 2)
      k_dr_tv<float, 4, 2><<<grid_v, PB_BLOCK, 0, ctx->stream>>>(p);
    else if (minb == 3)
      k_dr_tv<float, 4, 3><<<grid_v, PB_BLOCK, 0, ctx->stream>>>(p);
    else
      k_dr_tv<float, 4, 4><<<grid_v, PB_BLOCK, 0, ctx->stream>>>(p);
  } else {
    if (!vec_ok)
      k_dr_tv<double, 1, 2><<<pb_stream_grid(ctx, PB_BLOCK, n, 4), PB_BLOCK, 0, ctx->stream>>>(p);
    else if (minb == 2)
      k_dr_tv<double, 2, 2><<<grid_v, PB_BLOCK, 0, ctx->stream>>>(p);
    else if (minb == 3)
      k_dr_tv<double, 2, 3><<<grid_v, PB_BLOCK, 0, ctx->stream>>>(p);
    else
      k_dr_tv<double, 2, 4><<<grid_v, PB_BLOCK, 0, ctx->stream>>>(p);
  }
  PB_LAUNCH_CHECK(ctx);
  return PB_OK;
}

// ---- cudaIpc plumbing for peer-readable device buffers (the halo rows above) ------------------------------------------
extern "C" int pb_ipc_export(pb_ctx* ctx, void* dptr, void* handle_out) {
  PB_REQUIRE(ctx != nullptr && dptr != nullptr && handle_out != nullptr, "null argument");
  PB_CHECK_CUDA(cudaSetDevice(ctx->device));
  cudaIpcMemHandle_t h;
  PB_CHECK_CUDA(cudaIpcGetMemHandle(&h, dptr));
  memcpy(handle_out, &h, sizeof(h));
  return PB_OK;
}

extern "C" int pb_ipc_open(pb_ctx* ctx, const void* handle, void** dptr) {
  PB_REQUIRE(ctx != nullptr && handle != nullptr && dptr != nullptr, "null argument");
  PB_CHECK_CUDA(cudaSetDevice(ctx->device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  PB_CHECK_CUDA(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
  return PB_OK;
}

extern "C" int pb_ipc_close(pb_ctx* ctx, void* dptr) {
  PB_REQUIRE(ctx != nullptr, "null context");
  if (!dptr) return PB_OK;
  PB_CHECK_CUDA(cudaSetDevice(ctx->device));
  PB_CHECK_CUDA(cudaIpcCloseMemHandle(dptr));
  return PB_OK;
}
