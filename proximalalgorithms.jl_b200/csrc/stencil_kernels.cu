// stencil_kernels.cu -- K11: forward-difference operator of anisotropic total variation and its adjoint, for the primal-dual
// row of SURVEY.md section 8f (f4: "L = finite-difference operator, h = ||.||_1 conjugate => box projection").
//   L  : R^{H x W} -> R^{2 x H x W},  (L u)[0][i][j] = u[i][j+1] - u[i][j]  (0 in the last column)
//                                      (L u)[1][i][j] = u[i+1][j] - u[i][j]  (0 in the last row)
//   L' : R^{2 x H x W} -> R^{H x W},  (L' (p, q))[i][j] = (p~[i][j-1] - p~[i][j]) + (q~[i-1][j] - q~[i][j])
//        with p~ = p except 0 in the last column and outside the image, q~ = q except 0 in the last row and outside
// so that <L u, (p, q)> = <u, L'(p, q)> exactly in exact arithmetic.  Row-major images; every difference and the final sum are
// rounded separately.  Single coalesced passes (3 images moved per call); neighbour loads hit L1/L2.  HBM-bound.
// STATUS: staged for round 2 -- compiled, specified bit-exactly against oracle/stencil_oracle.py, not yet run on a GPU.
#include "common.cuh"

template <typename T>
__global__ void __launch_bounds__(PB_BLOCK) k_fd2d_forward(const T* __restrict__ u, T* __restrict__ out, int64_t H, int64_t W) {
  const int64_t n = H * W;
  for (int64_t idx = (int64_t)blockIdx.x * PB_BLOCK + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * PB_BLOCK) {
    const int64_t i = idx / W, j = idx - i * W;
    const T c = u[idx];
    out[idx] = (j + 1 < W) ? sub_rn(u[idx + 1], c) : T(0);
    out[n + idx] = (i + 1 < H) ? sub_rn(u[idx + W], c) : T(0);
  }
}

template <typename T>
__global__ void __launch_bounds__(PB_BLOCK) k_fd2d_adjoint(const T* __restrict__ pq, T* __restrict__ out, int64_t H, int64_t W) {
  const int64_t n = H * W;
  const T* __restrict__ p = pq;
  const T* __restrict__ q = pq + n;
  for (int64_t idx = (int64_t)blockIdx.x * PB_BLOCK + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * PB_BLOCK) {
    const int64_t i = idx / W, j = idx - i * W;
    const T pl = (j > 0) ? p[idx - 1] : T(0);
    const T pc = (j + 1 < W) ? p[idx] : T(0);
    const T qu = (i > 0) ? q[idx - W] : T(0);
    const T qc = (i + 1 < H) ? q[idx] : T(0);
    out[idx] = add_rn(sub_rn(pl, pc), sub_rn(qu, qc));
  }
}

extern "C" int pb_fd2d_forward(pb_ctx* ctx, int dtype, int64_t H, int64_t W, const void* u, void* out) {
  PB_REQUIRE(ctx != nullptr, "null context");
  PB_REQUIRE(dtype == PB_F32 || dtype == PB_F64, "dtype must be PB_F32 or PB_F64");
  PB_REQUIRE(H >= 0 && W >= 0, "negative shape");
  PB_REQUIRE(H * W == 0 || (u && out), "null image");
  PB_REQUIRE(u != out, "out must not alias u");
  const int grid = pb_stream_grid(ctx, PB_BLOCK * 4, H * W, 8);
  if (dtype == PB_F32)
    k_fd2d_forward<float><<<grid, PB_BLOCK, 0, ctx->stream>>>((const float*)u, (float*)out, H, W);
  else
    k_fd2d_forward<double><<<grid, PB_BLOCK, 0, ctx->stream>>>((const double*)u, (double*)out, H, W);
  PB_LAUNCH_CHECK(ctx);
  return PB_OK;
}

extern "C" int pb_fd2d_adjoint(pb_ctx* ctx, int dtype, int64_t H, int64_t W, const void* pq, void* out) {
  PB_REQUIRE(ctx != nullptr, "null context");
  PB_REQUIRE(dtype == PB_F32 || dtype == PB_F64, "dtype must be PB_F32 or PB_F64");
  PB_REQUIRE(H >= 0 && W >= 0, "negative shape");
  PB_REQUIRE(H * W == 0 || (pq && out), "null image");
  PB_REQUIRE(pq != out, "out must not alias pq");
  const int grid = pb_stream_grid(ctx, PB_BLOCK * 4, H * W, 8);
  if (dtype == PB_F32)
    k_fd2d_adjoint<float><<<grid, PB_BLOCK, 0, ctx->stream>>>((const float*)pq, (float*)out, H, W);
  else
    k_fd2d_adjoint<double><<<grid, PB_BLOCK, 0, ctx->stream>>>((const double*)pq, (double*)out, H, W);
  PB_LAUNCH_CHECK(ctx);
  return PB_OK;
}
