// xchg.cu -- host side of C1 (see xchg.cuh): IPC set-up of the per-rank exchange buffers, the stand-alone 1-CTA exchange
// kernel for reads that do not follow a fused step, and the host-side wait on the mapped pinned words.
#include <string.h>
#include <time.h>

#include "common.cuh"

static_assert(sizeof(cudaIpcMemHandle_t) == PB_IPC_HANDLE_BYTES, "PB_IPC_HANDLE_BYTES must match cudaIpcMemHandle_t");
static_assert(PB_MAX_RANKS * PB_XCHG_WORDS_PER_ROW <= PB_BLOCK, "one thread per word");
static_assert(PB_MAX_WORLD == PB_MAX_RANKS, "header and implementation disagree on the maximum world size");

__global__ void __launch_bounds__(PB_BLOCK) k_xchg(XchgParams xp, const double* local_block) {
  xchg_push_wait(xp, local_block);
}

void pb_xchg_next(pb_ctx* ctx, XchgParams* xp, bool want) {
  memset(xp, 0, sizeof(*xp));
  if (!want || ctx->xchg_world <= 0 || !ctx->xchg_connected) return;
  ctx->xchg_seq += 1;
  if (ctx->xchg_seq == 0 || ctx->xchg_seq == PB_XCHG_ERROR_SEQ) ctx->xchg_seq = 1;   // both values are reserved
  ctx->xchg_pending = 1;
  ctx->xchg_pending_launch = ctx->launches + 1;   // the caller launches exactly one kernel next: the one that publishes
  for (int r = 0; r < ctx->xchg_world; ++r) xp->peer[r] = ctx->xchg_peer[r];
  xp->host_words = ctx->xchg_host_words_dev;
  xp->seq = ctx->xchg_seq;
  xp->rank = ctx->xchg_rank;
  xp->world = ctx->xchg_world;
}

// Parameters of the next vector exchange (C2).  The error word is the one behind the scalar landing zone.
void pb_xchg_vec_next(pb_ctx* ctx, XchgVecParams* xv) {
  memset(xv, 0, sizeof(*xv));
  ctx->xchg_vseq += 1;
  if (ctx->xchg_vseq == 0 || ctx->xchg_vseq == PB_XCHG_ERROR_SEQ) ctx->xchg_vseq = 1;
  for (int r = 0; r < ctx->xchg_world; ++r)
    xv->peer[r] = reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(ctx->xchg_peer[r]) + PB_XCHG_BYTES);
  xv->err_word = ctx->xchg_host_words_dev + (size_t)2 * PB_MAX_RANKS * PB_XCHG_WORDS_PER_ROW;
  xv->seq = ctx->xchg_vseq;
  xv->rank = ctx->xchg_rank;
  xv->world = ctx->xchg_world;
}

extern "C" int pb_xchg_init(pb_ctx* ctx, int rank, int world, void* handle_out) {
  PB_REQUIRE(ctx != nullptr, "null context");
  PB_REQUIRE(world >= 1 && world <= PB_MAX_RANKS && rank >= 0 && rank < world, "need 0 <= rank < world <= 8");
  PB_REQUIRE(ctx->xchg_world == 0, "exchange already initialised on this context");
  PB_CHECK_CUDA(cudaSetDevice(ctx->device));
  void* own = nullptr;
  const size_t own_bytes = PB_XCHG_BYTES + (world > 1 ? (size_t)PB_XCHG_VEC_BYTES : 0);   // scalar region + vector region (C2)
  PB_CHECK_CUDA(cudaMalloc(&own, own_bytes));
  PB_CHECK_CUDA(cudaMemset(own, 0, own_bytes));
  ctx->xchg_own = static_cast<unsigned long long*>(own);
  void* host = nullptr;
  const size_t hbytes = (size_t)2 * PB_MAX_RANKS * PB_XCHG_WORDS_PER_ROW * 8 + 8;   // [parity][rank][word] + the vector-exchange error word
  PB_CHECK_CUDA(cudaHostAlloc(&host, hbytes, cudaHostAllocMapped | cudaHostAllocPortable));
  memset(host, 0, hbytes);
  ctx->xchg_host_words = static_cast<unsigned long long*>(host);
  void* dev_alias = nullptr;
  PB_CHECK_CUDA(cudaHostGetDevicePointer(&dev_alias, host, 0));
  ctx->xchg_host_words_dev = static_cast<unsigned long long*>(dev_alias);
  ctx->xchg_rank = rank;
  ctx->xchg_world = world;
  ctx->xchg_seq = 0;
  ctx->xchg_vseq = 0;
  ctx->xchg_pending = 0;
  ctx->xchg_connected = 0;
  for (int r = 0; r < PB_MAX_RANKS; ++r) ctx->xchg_peer[r] = nullptr;
  ctx->xchg_peer[rank] = ctx->xchg_own;
  if (world == 1) ctx->xchg_connected = 1;
  if (handle_out) {
    cudaIpcMemHandle_t h;
    memset(&h, 0, sizeof(h));
    if (world > 1) PB_CHECK_CUDA(cudaIpcGetMemHandle(&h, ctx->xchg_own));
    memcpy(handle_out, &h, sizeof(h));
  }
  PB_CHECK_CUDA(cudaDeviceSynchronize());
  return PB_OK;
}

extern "C" int pb_xchg_connect(pb_ctx* ctx, const void* all_handles) {
  PB_REQUIRE(ctx != nullptr && ctx->xchg_world > 0, "pb_xchg_init first");
  if (ctx->xchg_world == 1) return PB_OK;
  PB_REQUIRE(all_handles != nullptr, "null handles");
  PB_CHECK_CUDA(cudaSetDevice(ctx->device));
  const unsigned char* hs = static_cast<const unsigned char*>(all_handles);
  for (int r = 0; r < ctx->xchg_world; ++r) {
    if (r == ctx->xchg_rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, hs + (size_t)r * PB_IPC_HANDLE_BYTES, sizeof(h));
    void* p = nullptr;
    PB_CHECK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    ctx->xchg_peer[r] = static_cast<unsigned long long*>(p);
  }
  ctx->xchg_connected = 1;
  return PB_OK;
}

// Single-process form of pb_xchg_connect (SURVEY.md section 8b "Threading": one host process drives all GPUs): the n contexts
// live in THIS process (cudaIpc cannot open a handle in the exporting process), so the peers' exchange buffers are reached
// through plain peer access.  ctxs[r] must have been initialised with pb_xchg_init(ctxs[r], r, n, NULL).  Contexts may share
// a device (used by the single-GPU tests of the sharded path: their streams run concurrently).
extern "C" int pb_xchg_connect_local(pb_ctx** ctxs, int n) {
  PB_REQUIRE(ctxs != nullptr && n >= 1 && n <= PB_MAX_RANKS, "need 1 <= n <= 8 contexts");
  for (int r = 0; r < n; ++r) {
    PB_REQUIRE(ctxs[r] != nullptr && ctxs[r]->xchg_world == n && ctxs[r]->xchg_rank == r && ctxs[r]->xchg_own != nullptr,
               "ctxs[r] must be initialised with pb_xchg_init(ctxs[r], r, n, ...)");
  }
  for (int r = 0; r < n; ++r) {
    PB_CHECK_CUDA(cudaSetDevice(ctxs[r]->device));
    for (int q = 0; q < n; ++q) {
      if (ctxs[q]->device != ctxs[r]->device) {
        int can = 0;
        PB_CHECK_CUDA(cudaDeviceCanAccessPeer(&can, ctxs[r]->device, ctxs[q]->device));
        if (!can) {
          pb_set_error("pb_xchg_connect_local: device %d cannot access device %d", ctxs[r]->device, ctxs[q]->device);
          return PB_EUNSUPPORTED;
        }
        const cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[q]->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) PB_CHECK_CUDA(e);
        cudaGetLastError();
      }
      else if (q != r)
        ctxs[r]->xchg_shared_device = 1;
      ctxs[r]->xchg_peer[q] = ctxs[q]->xchg_own;
    }
    ctxs[r]->xchg_connected = 1;
    ctxs[r]->xchg_local = 1;
  }
  return PB_OK;
}

extern "C" int pb_xchg_shutdown(pb_ctx* ctx) {
  if (!ctx || ctx->xchg_world == 0) return PB_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (int r = 0; r < ctx->xchg_world; ++r)
    if (r != ctx->xchg_rank && ctx->xchg_peer[r] && !ctx->xchg_local) cudaIpcCloseMemHandle(ctx->xchg_peer[r]);
  if (ctx->xchg_own) cudaFree(ctx->xchg_own);
  if (ctx->xchg_host_words) cudaFreeHost(ctx->xchg_host_words);
  ctx->xchg_own = nullptr;
  ctx->xchg_host_words = nullptr;
  ctx->xchg_world = 0;
  ctx->xchg_connected = 0;
  ctx->xchg_local = 0;
  ctx->xchg_shared_device = 0;
  ctx->xchg_fused = 0;
  ctx->xchg_pending = 0;
  return PB_OK;
}

// Launch the stand-alone exchange of the current scalar block (used when the last reducing kernel was not a fused step).
extern "C" int pb_exchange(pb_ctx* ctx) {
  PB_REQUIRE(ctx != nullptr && ctx->xchg_world > 0 && ctx->xchg_connected, "exchange not initialised / connected");
  XchgParams xp;
  pb_xchg_next(ctx, &xp, true);
  k_xchg<<<1, PB_BLOCK, 0, ctx->stream>>>(xp, ctx->scalars_dev);
  PB_LAUNCH_CHECK(ctx);
  return PB_OK;
}

static double now_s() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

// Wait (polling mapped pinned memory; no CUDA call on the fast path) for the most recently issued exchange and copy the
// world x PB_NSCALARS rows to `rows_out` (rank order).  If no exchange is pending, one is launched first.
extern "C" int pb_exchange_wait(pb_ctx* ctx, double* rows_out, double timeout_s) {
  PB_REQUIRE(ctx != nullptr && rows_out != nullptr, "null argument");
  PB_REQUIRE(ctx->xchg_world > 0 && ctx->xchg_connected, "exchange not initialised / connected");
  // An exchange is current only if its publishing kernel is the most recent launch of this context: if other kernels
  // (which may have changed the scalar block) were enqueued since, or nothing is pending, publish the block afresh.
  if (!ctx->xchg_pending || ctx->launches != ctx->xchg_pending_launch) {
    int rc = pb_exchange(ctx);
    if (rc != PB_OK) return rc;
  }
  const int rc = pb_xchg_wait_seq(ctx, ctx->xchg_seq, rows_out, timeout_s);
  ctx->xchg_pending = 0;
  return rc;
}

// Poll the landing zone for exchange `want` (a sequence number handed out by pb_xchg_next).  At most two exchanges may be
// outstanding: `want` and its successor (different parity slots).
int pb_xchg_wait_seq(pb_ctx* ctx, unsigned int want, double* rows_out, double timeout_s) {
  const int nwords = ctx->xchg_world * PB_XCHG_WORDS_PER_ROW;
  volatile unsigned long long* words = ctx->xchg_host_words + (size_t)(want & 1u) * PB_MAX_RANKS * PB_XCHG_WORDS_PER_ROW;
  volatile unsigned long long* vec_err = ctx->xchg_host_words + (size_t)2 * PB_MAX_RANKS * PB_XCHG_WORDS_PER_ROW;
  unsigned long long got[PB_MAX_RANKS * PB_XCHG_WORDS_PER_ROW];
  const double t0 = now_s();
  unsigned long long spins = 0;
  int i = nwords - 1;   // scan backwards: the words usually land in ascending order, so the last one arrives last
  while (i >= 0) {
    const unsigned long long v = words[i];
    const unsigned int s = (unsigned int)(v >> 32);
    if (s == want) {
      got[i] = v;
      --i;
      continue;
    }
    if (s == PB_XCHG_ERROR_SEQ) {
      pb_set_error("pb_exchange_wait: a peer did not publish sequence %u within the device time-out", want);
      return PB_ECUDA;
    }
    if ((++spins & 0xfff) == 0 && now_s() - t0 > timeout_s) {
      pb_set_error("pb_exchange_wait: timed out after %.1f s waiting for sequence %u (word %d)", timeout_s, want, i);
      return PB_ECUDA;
    }
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  }
  if ((unsigned int)(*vec_err >> 32) == PB_XCHG_ERROR_SEQ) {
    pb_set_error("pb_exchange_wait: a peer did not publish its chunk partials of A x within the device time-out (vector exchange)");
    return PB_ECUDA;
  }
  for (int k = 0; k < ctx->xchg_world * PB_NSCALARS; ++k) {
    const unsigned long long bits = (got[2 * k] & 0xffffffffull) | ((got[2 * k + 1] & 0xffffffffull) << 32);
    memcpy(&rows_out[k], &bits, 8);
  }
  return PB_OK;
}
