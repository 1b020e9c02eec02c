// ctx.cu -- context, memory and scalar-block plumbing of libproxb200 (host side of the C ABI).
#include <stdarg.h>
#include <new>
#include <string.h>

#include "common.cuh"

size_t pb_multi_ws_bytes();   // step_multi.cu

static thread_local char g_err[512] = "";

void pb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* pb_last_error(void) { return g_err; }
extern "C" const char* pb_version(void) { return PB_VERSION_STRING " (sm_100a)"; }

extern "C" int pb_device_count(int* count) {
  PB_REQUIRE(count != nullptr, "null output");
  *count = 0;
  PB_CHECK_CUDA(cudaGetDeviceCount(count));
  return PB_OK;
}

extern "C" int pb_ctx_create(int device, void* stream, int borrow_stream, pb_ctx** out) {
  PB_REQUIRE(out != nullptr, "null output");
  *out = nullptr;
  int ndev = 0;
  PB_CHECK_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) {
    pb_set_error("pb_ctx_create: device %d out of range (%d CUDA devices visible)", device, ndev);
    return PB_EINVAL;
  }
  PB_CHECK_CUDA(cudaSetDevice(device));
  pb_ctx* c = new (std::nothrow) pb_ctx;
  if (!c) {
    pb_set_error("pb_ctx_create: out of host memory");
    return PB_ENOMEM;
  }
  memset(c, 0, sizeof(*c));
  c->device = device;
  c->stream_hints = -1;
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) {
    delete c;
    pb_set_error("cudaGetDeviceProperties -> %s", cudaGetErrorString(e));
    return PB_ECUDA;
  }
  c->sm_count = prop.multiProcessorCount;
  c->l2_bytes = (size_t)prop.l2CacheSize;
  if (borrow_stream) {
    c->stream = static_cast<cudaStream_t>(stream);
    c->owns_stream = false;
  } else {
    e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
      delete c;
      pb_set_error("cudaStreamCreate -> %s", cudaGetErrorString(e));
      return PB_ECUDA;
    }
    c->owns_stream = true;
  }
  bool ok = cudaMalloc(&c->scalars_own, PB_NSCALARS * sizeof(double)) == cudaSuccess &&
            cudaMallocHost(&c->scalars_host, PB_NSCALARS * sizeof(double)) == cudaSuccess &&
            cudaMalloc(&c->ws, sizeof(PbWorkspace)) == cudaSuccess &&
            cudaMalloc(&c->multi_ws, pb_multi_ws_bytes()) == cudaSuccess &&
            cudaMalloc((void**)&c->chain_dev, PB_NSCALARS * sizeof(double)) == cudaSuccess &&
            cudaMemsetAsync(c->scalars_own, 0, PB_NSCALARS * sizeof(double), c->stream) == cudaSuccess &&
            cudaMemsetAsync(c->ws, 0, sizeof(PbWorkspace), c->stream) == cudaSuccess &&
            cudaStreamSynchronize(c->stream) == cudaSuccess;
  if (!ok) {
    pb_set_error("pb_ctx_create: allocation failed -> %s", cudaGetErrorString(cudaGetLastError()));
    pb_ctx_destroy(c);
    return PB_ENOMEM;
  }
  c->scalars_dev = c->scalars_own;
  *out = c;
  return PB_OK;
}

extern "C" int pb_ctx_set_scalars_dev(pb_ctx* c, double* dev) {
  PB_REQUIRE(c != nullptr, "null context");
  c->scalars_dev = dev ? dev : c->scalars_own;
  return PB_OK;
}

extern "C" int pb_ctx_destroy(pb_ctx* c) {
  if (!c) return PB_OK;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  pb_xchg_shutdown(c);
  for (int k = 0; k < 5; ++k)
    if (c->hbuf[k]) cudaFree(c->hbuf[k]);
  if (c->scratch) cudaFree(c->scratch);
  if (c->multi_ws) cudaFree(c->multi_ws);
  if (c->chain_dev) cudaFree(c->chain_dev);
  if (c->ws) cudaFree(c->ws);
  if (c->side_stream) {
    cudaStreamSynchronize(c->side_stream);
    for (int k = 0; k < 2; ++k) {
      if (c->ws_defer[k]) cudaFree(c->ws_defer[k]);
      if (c->ev_main[k]) cudaEventDestroy(c->ev_main[k]);
      if (c->ev_fold[k]) cudaEventDestroy(c->ev_fold[k]);
    }
    cudaStreamDestroy(c->side_stream);
  }
  if (c->scalars_own) cudaFree(c->scalars_own);
  if (c->ahead_host) {
    cudaFreeHost(c->ahead_host);
    cudaEventDestroy(c->ahead_ev[0]);
    cudaEventDestroy(c->ahead_ev[1]);
  }
  if (c->scalars_host) cudaFreeHost(c->scalars_host);
  if (c->owns_stream && c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return PB_OK;
}

extern "C" int pb_ctx_make_current(pb_ctx* c) {
  PB_REQUIRE(c != nullptr, "null context");
  PB_CHECK_CUDA(cudaSetDevice(c->device));
  return PB_OK;
}

extern "C" void* pb_ctx_stream(pb_ctx* c) { return c ? (void*)c->stream : nullptr; }
extern "C" double* pb_ctx_scalars_dev(pb_ctx* c) { return c ? c->scalars_dev : nullptr; }
extern "C" int64_t pb_ctx_launch_count(pb_ctx* c) { return c ? c->launches : 0; }

extern "C" int pb_ctx_sync(pb_ctx* c) {
  PB_REQUIRE(c != nullptr, "null context");
  PB_CHECK_CUDA(cudaStreamSynchronize(c->stream));
  return PB_OK;
}

extern "C" int pb_ctx_set_option(pb_ctx* c, int option, int value) {
  PB_REQUIRE(c != nullptr, "null context");
  switch (option) {
    case PB_OPT_CTAS_PER_SM:
      PB_REQUIRE(value >= 0 && value <= 32, "ctas_per_sm out of range [0, 32]");
      c->ctas_per_sm = value;
      return PB_OK;
    case PB_OPT_STREAM_HINTS:
      PB_REQUIRE(value >= -1 && value <= 1, "stream_hints must be -1, 0 or 1");
      c->stream_hints = value;
      return PB_OK;
    case PB_OPT_UNROLL:
      PB_REQUIRE(value == 0 || value == 1 || value == 2 || value == 4 || value == 8, "unroll must be 0, 1, 2, 4 or 8");
      c->unroll = value;
      return PB_OK;
    case PB_OPT_STEP_IMPL:
      PB_REQUIRE(value >= 0 && value <= 2, "step implementation must be 0, 1 or 2");
      c->step_impl = value;
      return PB_OK;
    case PB_OPT_PERSISTENT:
      PB_REQUIRE(value >= -1 && value <= 32, "persistent mode must be -1, 0 or 1..32");
      c->persist_mode = value;
      return PB_OK;
    case PB_OPT_LSQ_FISTA:
      PB_REQUIRE(value >= -1 && value <= 1, "single-sweep FISTA mode must be -1, 0 or 1");
      c->lsq_fista = value;
      return PB_OK;
    case PB_OPT_LSQ_FUSED:
      PB_REQUIRE(value >= -1 && value <= 8, "fused least-squares mode must be -1, 0 or 1..8");
      c->lsq_fused = value;
      return PB_OK;
    case PB_OPT_GEMV_SCALAR:
      PB_REQUIRE(value == 0 || value == 1, "gemv scalar mode must be 0 or 1");
      c->gemv_scalar = value;
      return PB_OK;
    case PB_OPT_MULTI_ITER:
      PB_REQUIRE(value >= -1 && value <= 1, "multi-iteration mode must be -1, 0 or 1");
      c->multi_mode = value;
      return PB_OK;
    case PB_OPT_FUSED_EXCHANGE:
      PB_REQUIRE(value == 0 || value == 1, "fused exchange must be 0 or 1");
      PB_REQUIRE(value == 0 || (c->xchg_world > 0 && c->xchg_connected), "pb_xchg_init / pb_xchg_connect first");
      c->xchg_fused = value;
      return PB_OK;
    default:
      pb_set_error("pb_ctx_set_option: unknown option %d", option);
      return PB_EINVAL;
  }
}

int pb_ensure_scratch(pb_ctx* c, size_t bytes) {
  if (c->scratch_bytes >= bytes) return PB_OK;
  if (c->scratch) {
    PB_CHECK_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(c->scratch);
    c->scratch = nullptr;
    c->scratch_bytes = 0;
  }
  size_t want = bytes < (1u << 20) ? (1u << 20) : bytes;
  cudaError_t e = cudaMalloc(&c->scratch, want);
  if (e != cudaSuccess) {
    pb_set_error("scratch allocation of %zu bytes failed -> %s", want, cudaGetErrorString(e));
    return PB_ENOMEM;
  }
  c->scratch_bytes = want;
  return PB_OK;
}

extern "C" int pb_malloc(pb_ctx* c, size_t bytes, void** dptr) {
  PB_REQUIRE(c != nullptr && dptr != nullptr, "null argument");
  *dptr = nullptr;
  cudaError_t e = cudaMalloc(dptr, bytes ? bytes : 16);
  if (e != cudaSuccess) {
    pb_set_error("pb_malloc(%zu) -> %s", bytes, cudaGetErrorString(e));
    return PB_ENOMEM;
  }
  return PB_OK;
}

extern "C" int pb_free(pb_ctx* c, void* dptr) {
  PB_REQUIRE(c != nullptr, "null context");
  if (dptr) PB_CHECK_CUDA(cudaFree(dptr));
  return PB_OK;
}

extern "C" int pb_host_alloc(size_t bytes, void** hptr) {
  PB_REQUIRE(hptr != nullptr, "null output");
  *hptr = nullptr;
  cudaError_t e = cudaMallocHost(hptr, bytes ? bytes : 16);
  if (e != cudaSuccess) {
    pb_set_error("pb_host_alloc(%zu) -> %s", bytes, cudaGetErrorString(e));
    return PB_ENOMEM;
  }
  return PB_OK;
}

extern "C" int pb_host_free(void* hptr) {
  if (hptr) PB_CHECK_CUDA(cudaFreeHost(hptr));
  return PB_OK;
}

extern "C" int pb_upload(pb_ctx* c, void* dst, const void* src, size_t bytes) {
  PB_REQUIRE(c != nullptr, "null context");
  PB_REQUIRE(bytes == 0 || (dst && src), "null buffer");
  if (bytes) PB_CHECK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
  return PB_OK;
}

extern "C" int pb_download(pb_ctx* c, void* dst, const void* src, size_t bytes) {
  PB_REQUIRE(c != nullptr, "null context");
  PB_REQUIRE(bytes == 0 || (dst && src), "null buffer");
  if (bytes) PB_CHECK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
  PB_CHECK_CUDA(cudaStreamSynchronize(c->stream));
  return PB_OK;
}

extern "C" int pb_copy(pb_ctx* c, void* dst, const void* src, size_t bytes) {
  PB_REQUIRE(c != nullptr, "null context");
  PB_REQUIRE(bytes == 0 || (dst && src), "null buffer");
  if (bytes) PB_CHECK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, c->stream));
  return PB_OK;
}

extern "C" int pb_memset_zero(pb_ctx* c, void* dst, size_t bytes) {
  PB_REQUIRE(c != nullptr, "null context");
  PB_REQUIRE(bytes == 0 || dst, "null buffer");
  if (bytes) PB_CHECK_CUDA(cudaMemsetAsync(dst, 0, bytes, c->stream));
  return PB_OK;
}

extern "C" int pb_read_scalars(pb_ctx* c, double* out) {
  PB_REQUIRE(c != nullptr && out != nullptr, "null argument");
  PB_CHECK_CUDA(cudaMemcpyAsync(c->scalars_host, c->scalars_dev, PB_NSCALARS * sizeof(double), cudaMemcpyDeviceToHost,
                                c->stream));
  PB_CHECK_CUDA(cudaStreamSynchronize(c->stream));
  memcpy(out, c->scalars_host, PB_NSCALARS * sizeof(double));
  return PB_OK;
}
