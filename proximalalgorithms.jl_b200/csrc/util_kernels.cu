// util_kernels.cu -- counter-based synthetic data: every element is a pure function of its GLOBAL index, so a row shard of an
// n-vector holds exactly the values the unsharded vector has at those indices, whatever the number of GPUs.  bench.py uses
// it so that the per-iteration scalars of the N = 1, 2, 4, 8 runs can be compared bit for bit (its `parity` fingerprint).
// Same generator as the CPU port's first-touch fill (splitmix64 finaliser of index * golden-ratio + seed, mapped to
// [-1, 1) through the top 53 bits), so both arms of the benchmark work on identical data.
#include "common.cuh"

template <typename T>
__global__ void __launch_bounds__(PB_BLOCK) k_fill_counter(T* __restrict__ out, int64_t n, uint64_t offset, uint64_t seed,
                                                           double scale) {
  for (int64_t i = (int64_t)blockIdx.x * PB_BLOCK + threadIdx.x; i < n; i += (int64_t)gridDim.x * PB_BLOCK) {
    uint64_t s = ((uint64_t)i + offset) * 0x9E3779B97F4A7C15ull + seed;
    s ^= s >> 30;
    s *= 0xBF58476D1CE4E5B9ull;
    s ^= s >> 27;
    s *= 0x94D049BB133111EBull;
    s ^= s >> 31;
    const double u = __dsub_rn(__dmul_rn(__dmul_rn((double)(s >> 11), 1.0 / 9007199254740992.0), 2.0), 1.0);
    out[i] = (T)scale * (T)u;
  }
}

extern "C" int pb_fill_counter(pb_ctx* ctx, int dtype, int64_t n, int64_t offset, uint64_t seed, double scale, void* out) {
  PB_REQUIRE(ctx != nullptr, "null context");
  PB_REQUIRE(dtype == PB_F32 || dtype == PB_F64, "dtype must be PB_F32 or PB_F64");
  PB_REQUIRE(n >= 0 && offset >= 0, "negative length / offset");
  PB_REQUIRE(n == 0 || out != nullptr, "null output");
  if (n == 0) return PB_OK;
  const int grid = pb_stream_grid(ctx, PB_BLOCK * 4, n, 8);
  if (dtype == PB_F32)
    k_fill_counter<float><<<grid, PB_BLOCK, 0, ctx->stream>>>((float*)out, n, (uint64_t)offset, seed, scale);
  else
    k_fill_counter<double><<<grid, PB_BLOCK, 0, ctx->stream>>>((double*)out, n, (uint64_t)offset, seed, scale);
  PB_LAUNCH_CHECK(ctx);
  return PB_OK;
}
