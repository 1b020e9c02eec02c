// lsq_prox.cu -- K9: proximal operator of f(x) = lambda/2*||A x - b||^2 for a dense A (the `f = LeastSquares(A, b)` of the
// reference's Douglas-Rachford test and benchmark: test/problems/test_lasso_small.jl:39,205-214, benchmark/benchmarks.jl:87-93).
//
// The arithmetic lives in ProximalOperators.jl (LeastSquaresDirect; not under /root/reference, restated from the package's
// published algorithm):   q = lambda*A'b + x/gamma
//   tall A (m >= n):  y = (lambda*A'A + I/gamma)^-1 q
//   wide A (m <  n):  y = gamma*(q - lambda*A'((lambda*AA' + I/gamma)^-1 (A q)))        (matrix-inversion lemma)
// and the returned value is lambda/2*||A y - b||^2.  The package factors S + I/gamma by Cholesky whenever gamma changes and
// does two triangular solves per call.  On a GPU a k x k triangular solve is k dependent steps, so here the factorisation
// step (once per gamma, off the hot path) also forms the explicit inverse W = (S + I/gamma)^-1 in double precision, and the
// per-iteration work is three (tall: one) GEMVs of the K4 family plus two element-wise passes -- all HBM/L2-bound streaming
// kernels.  Results agree with the triangular-solve form to rounding (tolerance stated in tests/test_gpu_dr_lsq.py), not
// bitwise: this term is parity-UNPINNED third-party arithmetic either way.
#include <new>

#include "common.cuh"

struct pb_lsqprox {
  pb_ctx* ctx;
  int dtype;
  int64_t m, n, k;       // k = min(m, n): order of the factored system
  int tall;
  double lambda;
  const void* A;         // caller's device matrix, column-major m x n, lda = m (borrowed)
  const void* b;         // caller's right-hand side (borrowed)
  void* Atb;             // lambda * A'b   (n)
  double* S;             // lambda*A'A or lambda*AA'  (k x k, double)
  double* Lf;            // Cholesky factor scratch (k x k, double)
  void* W;               // (S + I/gamma)^-1 in the element type, column-major k x k
  void *q, *t, *u;       // work vectors: n, m (or k), m (or k)
  double gamma;          // gamma of the current W; < 0: none
};

// S[i,j] = lambda * sum_l a_i[l]*a_j[l] where a_i is column i (tall) or row i (wide) of A; one CTA per entry of the lower
// triangle (mirrored), double accumulation in a fixed order.
template <typename T>
__global__ void __launch_bounds__(128) k_gram(const T* __restrict__ A, int64_t m, int64_t n, int tall, double lambda,
                                              double* __restrict__ S, int64_t k) {
  const int64_t i = blockIdx.y, j = blockIdx.x;
  if (j > i) return;
  const int64_t len = tall ? m : n;
  const int64_t si = tall ? i * m : i, sj = tall ? j * m : j, step = tall ? 1 : m;
  double acc = 0.0;
  for (int64_t l = threadIdx.x; l < len; l += 128) acc = fma((double)A[si + l * step], (double)A[sj + l * step], acc);
  __shared__ double sh[128];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int off = 64; off > 0; off >>= 1) {
    if (threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    S[i + j * k] = lambda * sh[0];
    S[j + i * k] = lambda * sh[0];
  }
}

// In-place lower Cholesky of the k x k double matrix M (column-major) by ONE CTA (right-looking, three barriers per column).
__global__ void __launch_bounds__(1024) k_cholesky(double* __restrict__ M, int64_t k, int* __restrict__ info) {
  for (int64_t j = 0; j < k; ++j) {
    if (threadIdx.x == 0) {
      const double d = M[j + j * k];
      if (!(d > 0.0)) *info = (int)(j + 1);
      M[j + j * k] = sqrt(d);
    }
    __syncthreads();
    const double d = M[j + j * k];
    for (int64_t i = j + 1 + threadIdx.x; i < k; i += blockDim.x) M[i + j * k] /= d;
    __syncthreads();
    // trailing update of the lower triangle: columns c in (j, k), rows i in [c, k)
    const int64_t rem = k - j - 1;
    for (int64_t e = threadIdx.x; e < rem * rem; e += blockDim.x) {
      const int64_t c = j + 1 + e / rem, i = j + 1 + e % rem;
      if (i >= c) M[i + c * k] = fma(-M[i + j * k], M[c + j * k], M[i + c * k]);
    }
    __syncthreads();
  }
}

// Column t of W = (L L')^-1: forward then backward substitution on e_t, one thread per column.
template <typename T>
__global__ void k_chol_inverse(const double* __restrict__ Lm, int64_t k, double* __restrict__ work, T* __restrict__ W) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= k) return;
  double* w = work + t * k;
  for (int64_t i = 0; i < k; ++i) {                   // L z = e_t
    double s = (i == t) ? 1.0 : 0.0;
    for (int64_t c = 0; c < i; ++c) s = fma(-Lm[i + c * k], w[c], s);
    w[i] = s / Lm[i + i * k];
  }
  for (int64_t i = k - 1; i >= 0; --i) {              // L' w = z
    double s = w[i];
    for (int64_t c = i + 1; c < k; ++c) s = fma(-Lm[c + i * k], w[c], s);
    w[i] = s / Lm[i + i * k];
  }
  for (int64_t i = 0; i < k; ++i) W[i + t * k] = (T)w[i];
}

template <typename T>
__global__ void __launch_bounds__(PB_BLOCK) k_add_diag(const double* __restrict__ S, double* __restrict__ M, int64_t k, double dg) {
  for (int64_t e = (int64_t)blockIdx.x * PB_BLOCK + threadIdx.x; e < k * k; e += (int64_t)gridDim.x * PB_BLOCK)
    M[e] = S[e] + ((e / k == e % k) ? (double)(T)dg : 0.0);
}

// q = atb + x / gamma   (`f.q .= f.lambdaAtb .+ x ./ gamma`)
template <typename T>
__global__ void __launch_bounds__(PB_BLOCK) k_q(const T* __restrict__ atb, const T* __restrict__ x, T* __restrict__ q, int64_t n,
                                                double gamma_d) {
  const T gamma = (T)gamma_d;
  for (int64_t i = (int64_t)blockIdx.x * PB_BLOCK + threadIdx.x; i < n; i += (int64_t)gridDim.x * PB_BLOCK)
    q[i] = add_rn(atb[i], x[i] / gamma);
}

// y = gamma * (q - lambda*v)   (`y .*= -lambda; y .+= q; y .*= gamma`)
template <typename T>
__global__ void __launch_bounds__(PB_BLOCK) k_wide_finish(const T* __restrict__ v, const T* __restrict__ q, T* __restrict__ y,
                                                          int64_t n, double lambda_d, double gamma_d) {
  const T ml = (T)(-lambda_d), gamma = (T)gamma_d;
  for (int64_t i = (int64_t)blockIdx.x * PB_BLOCK + threadIdx.x; i < n; i += (int64_t)gridDim.x * PB_BLOCK)
    y[i] = mul_rn(add_rn(mul_rn(v[i], ml), q[i]), gamma);
}

template <typename T>
__global__ void __launch_bounds__(PB_BLOCK) k_scale_inplace(T* __restrict__ v, int64_t n, double s) {
  for (int64_t i = (int64_t)blockIdx.x * PB_BLOCK + threadIdx.x; i < n; i += (int64_t)gridDim.x * PB_BLOCK)
    v[i] = mul_rn(v[i], (T)s);
}

static void lsqprox_free(pb_lsqprox* P) {
  void* ptrs[] = {P->Atb, P->S, P->Lf, P->W, P->q, P->t, P->u};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  delete P;
}

extern "C" int pb_lsq_prox_create(pb_ctx* ctx, int dtype, int64_t m, int64_t n, const void* A, const void* b, double lambda,
                                  pb_lsqprox** out) {
  PB_REQUIRE(ctx != nullptr && out != nullptr, "null argument");
  PB_REQUIRE(dtype == PB_F32 || dtype == PB_F64, "dtype must be PB_F32 or PB_F64");
  PB_REQUIRE(m > 0 && n > 0 && A && b, "need a non-empty matrix and right-hand side");
  PB_REQUIRE(lambda > 0, "lambda must be positive");
  const int64_t k = m < n ? m : n;
  PB_REQUIRE(k <= 4096, "min(m, n) > 4096: the direct (factorised) prox is meant for one small dimension");
  pb_lsqprox* P = new (std::nothrow) pb_lsqprox();
  if (!P) {
    pb_set_error("pb_lsq_prox_create: out of host memory");
    return PB_ENOMEM;
  }
  P->ctx = ctx;
  P->dtype = dtype;
  P->m = m;
  P->n = n;
  P->k = k;
  P->tall = m >= n;
  P->lambda = lambda;
  P->A = A;
  P->b = b;
  P->gamma = -1.0;
  const size_t es = dtype == PB_F32 ? 4 : 8;
  const int64_t mk = m > k ? m : k;
  cudaError_t e = cudaSetDevice(ctx->device);
  if (e == cudaSuccess) e = cudaMalloc(&P->Atb, es * n);
  if (e == cudaSuccess) e = cudaMalloc((void**)&P->S, sizeof(double) * k * k);
  if (e == cudaSuccess) e = cudaMalloc((void**)&P->Lf, sizeof(double) * k * k * 2);     // factor + inverse work space
  if (e == cudaSuccess) e = cudaMalloc(&P->W, es * k * k);
  if (e == cudaSuccess) e = cudaMalloc(&P->q, es * n);
  if (e == cudaSuccess) e = cudaMalloc(&P->t, es * mk);
  if (e == cudaSuccess) e = cudaMalloc(&P->u, es * mk);
  if (e != cudaSuccess) {
    pb_set_error("pb_lsq_prox_create: %s", cudaGetErrorString(e));
    lsqprox_free(P);
    return e == cudaErrorMemoryAllocation ? PB_ENOMEM : PB_ECUDA;
  }
  // Atb = lambda * A'b ; S = lambda * A'A (tall) or lambda * AA' (wide)
  int rc = pb_lsq_dense_gradient(ctx, dtype, m, n, A, m, b, P->Atb);
  if (rc == PB_OK && lambda != 1.0) {
    const int grid = pb_stream_grid(ctx, PB_BLOCK * 4, n, 4);
    if (dtype == PB_F32)
      k_scale_inplace<float><<<grid, PB_BLOCK, 0, ctx->stream>>>((float*)P->Atb, n, lambda);
    else
      k_scale_inplace<double><<<grid, PB_BLOCK, 0, ctx->stream>>>((double*)P->Atb, n, lambda);
    ctx->launches++;
  }
  if (rc == PB_OK) {
    dim3 grid((unsigned)k, (unsigned)k);
    if (dtype == PB_F32)
      k_gram<float><<<grid, 128, 0, ctx->stream>>>((const float*)A, m, n, P->tall, lambda, P->S, k);
    else
      k_gram<double><<<grid, 128, 0, ctx->stream>>>((const double*)A, m, n, P->tall, lambda, P->S, k);
    ctx->launches++;
    if (cudaGetLastError() != cudaSuccess) rc = PB_ECUDA;
  }
  if (rc != PB_OK) {
    lsqprox_free(P);
    return rc;
  }
  *out = P;
  return PB_OK;
}

extern "C" int pb_lsq_prox_destroy(pb_lsqprox* P) {
  if (!P) return PB_OK;
  cudaStreamSynchronize(P->ctx->stream);
  lsqprox_free(P);
  return PB_OK;
}

// factor_step!: W = (S + I/gamma)^-1.  Synchronises (reports a non-positive pivot).
static int lsqprox_factor(pb_lsqprox* P, double gamma) {
  pb_ctx* ctx = P->ctx;
  const int64_t k = P->k;
  int* info = nullptr;
  PB_CHECK_CUDA(cudaMalloc((void**)&info, sizeof(int)));
  PB_CHECK_CUDA(cudaMemsetAsync(info, 0, sizeof(int), ctx->stream));
  const int grid = pb_stream_grid(ctx, PB_BLOCK * 4, k * k, 4);
  // I / gamma is formed in the element type, like `I / gamma` on an R-typed gamma
  if (P->dtype == PB_F32) {
    volatile float dg = 1.0f / (float)gamma;
    k_add_diag<float><<<grid, PB_BLOCK, 0, ctx->stream>>>(P->S, P->Lf, k, (double)dg);
  } else {
    volatile double dg = 1.0 / gamma;
    k_add_diag<double><<<grid, PB_BLOCK, 0, ctx->stream>>>(P->S, P->Lf, k, dg);
  }
  k_cholesky<<<1, 1024, 0, ctx->stream>>>(P->Lf, k, info);
  const int tb = 64;
  if (P->dtype == PB_F32)
    k_chol_inverse<float><<<(unsigned)((k + tb - 1) / tb), tb, 0, ctx->stream>>>(P->Lf, k, P->Lf + k * k, (float*)P->W);
  else
    k_chol_inverse<double><<<(unsigned)((k + tb - 1) / tb), tb, 0, ctx->stream>>>(P->Lf, k, P->Lf + k * k, (double*)P->W);
  ctx->launches += 3;
  int h = 0;
  cudaError_t e = cudaMemcpyAsync(&h, info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(info);
  if (e != cudaSuccess) {
    pb_set_error("pb_lsq_prox: %s", cudaGetErrorString(e));
    return PB_ECUDA;
  }
  if (h != 0) {
    pb_set_error("pb_lsq_prox: S + I/gamma is not positive definite (pivot %d)", h);
    return PB_EINVAL;
  }
  P->gamma = gamma;
  return PB_OK;
}

// prox!(y, f, x, gamma): y as above, AUX = ||A y - b||^2 (value = lambda/2 * AUX).  Asynchronous unless gamma changed.
extern "C" int pb_lsq_prox_apply(pb_ctx* ctx, pb_lsqprox* P, const void* x, double gamma, void* y) {
  PB_REQUIRE(ctx != nullptr && P != nullptr, "null argument");
  PB_REQUIRE(P->ctx == ctx, "operator belongs to another context");
  PB_REQUIRE(x && y, "null vector");
  PB_REQUIRE(gamma > 0, "gamma must be positive");
  if (P->dtype == PB_F32) gamma = (double)(float)gamma;
  if (gamma != P->gamma) {
    const int rc = lsqprox_factor(P, gamma);
    if (rc != PB_OK) return rc;
  }
  const int64_t m = P->m, n = P->n, k = P->k;
  const int grid = pb_stream_grid(ctx, PB_BLOCK * 4, n, 4);
  if (P->dtype == PB_F32)
    k_q<float><<<grid, PB_BLOCK, 0, ctx->stream>>>((const float*)P->Atb, (const float*)x, (float*)P->q, n, gamma);
  else
    k_q<double><<<grid, PB_BLOCK, 0, ctx->stream>>>((const double*)P->Atb, (const double*)x, (double*)P->q, n, gamma);
  PB_LAUNCH_CHECK(ctx);
  int rc;
  if (P->tall) {
    rc = pb_lsq_dense_residual(ctx, P->dtype, k, k, P->W, k, P->q, nullptr, y);           // y = W q
    if (rc != PB_OK) return rc;
  } else {
    rc = pb_lsq_dense_residual(ctx, P->dtype, m, n, P->A, m, P->q, nullptr, P->t);       // t = A q
    if (rc != PB_OK) return rc;
    rc = pb_lsq_dense_residual(ctx, P->dtype, k, k, P->W, k, P->t, nullptr, P->u);       // u = W t
    if (rc != PB_OK) return rc;
    rc = pb_lsq_dense_gradient(ctx, P->dtype, m, n, P->A, m, P->u, y);                    // v = A' u  (into y)
    if (rc != PB_OK) return rc;
    if (P->dtype == PB_F32)
      k_wide_finish<float><<<grid, PB_BLOCK, 0, ctx->stream>>>((const float*)y, (const float*)P->q, (float*)y, n, P->lambda, gamma);
    else
      k_wide_finish<double><<<grid, PB_BLOCK, 0, ctx->stream>>>((const double*)y, (const double*)P->q, (double*)y, n, P->lambda, gamma);
    PB_LAUNCH_CHECK(ctx);
  }
  // value: res = A y - b, AUX = ||res||^2
  return pb_lsq_dense_residual(ctx, P->dtype, m, n, P->A, m, y, P->b, P->t);
}
