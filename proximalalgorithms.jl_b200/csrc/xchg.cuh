// xchg.cuh -- C1: the per-iteration exchange of the scalar block, done by the GPU itself.
//
// Every rank owns an exchange buffer in device memory (cudaMalloc, exported with cudaIpc and mapped by all peers over
// NVLink).  The last CTA of a reducing kernel (or the 1-CTA k_xchg kernel) pushes this rank's PB_NSCALARS doubles into
// slot [parity][rank] of EVERY rank's buffer with plain peer stores, waits until the P rows of its own buffer have
// arrived and forwards them to mapped pinned host memory, where the host polls for them (pb_exchange_wait).  The host
// needs neither a collective launch nor a cudaMemcpy nor a stream synchronisation, and folds the rows in rank order in
// double-double.  With one rank this degenerates to a zero-copy read-back of the scalar block.
//
// Wire format ("LL" style, as NCCL's low-latency protocol): every double travels as two 8-byte words {32 data bits,
// 32-bit sequence number}.  An aligned 8-byte store is atomic on NVLink and on PCIe, so a consumer that sees the expected
// sequence number in a word also sees its data bits -- no fences, no separate flags, and no reliance on the ORDER in
// which posted writes become visible to a peer or to the host.
//
// Double buffering by the parity of the sequence number is sufficient: a rank can only complete exchange k+1 after every
// peer has pushed k+1, which a peer does in its kernel k+1, i.e. after its kernel k (same stream) finished reading parity k.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "proxb200.h"

#define PB_MAX_RANKS 8                                    // one NVSwitch domain; PB_MAX_RANKS * 32 words <= 256 threads
#define PB_XCHG_WORDS_PER_ROW (2 * PB_NSCALARS)           // 32 eight-byte words per row
#define PB_XCHG_WORDS (2 * PB_MAX_RANKS * PB_XCHG_WORDS_PER_ROW)   // [parity][rank][word]
#define PB_XCHG_BYTES (PB_XCHG_WORDS * 8)
// C2: behind the scalar region (same allocation, hence the same IPC handle) every rank owns a VECTOR region: the landing zone of the
// chunk partials of r = A x for a column-sharded dense A (lsq_kernels.cu: k_gemv_n_combine_x).  [parity][global chunk][row][word]
#define PB_XCHG_VEC_BYTES (128ull << 20)
#define PB_XCHG_TIMEOUT_NS 4000000000ull                  // give up after 4 s (never hang the GPU)
#define PB_XCHG_ERROR_SEQ 0xFFFFFFFFu

struct XchgParams {
  unsigned long long* peer[PB_MAX_RANKS];   // peer[r]: base of rank r's exchange buffer as mapped in THIS process
  unsigned long long* host_words;           // device alias of mapped pinned memory: [parity][PB_MAX_RANKS][32] words
  unsigned int seq;                         // sequence number of this exchange (never 0, never PB_XCHG_ERROR_SEQ)
  int rank, world;                          // world == 0: exchange disabled
  double* gather_out;                       // non-NULL: the reducing CTA folds the ranks' rows itself (rank order, double-double) and
                                            // leaves the GLOBAL sums here (scalar-block layout) instead of forwarding rows to the host:
                                            // the next kernel of a device-side chain reads them (sharded L-BFGS two-loop recursion)
};

struct XchgVecParams {
  unsigned long long* peer[PB_MAX_RANKS];   // peer[r]: base of rank r's VECTOR region as mapped in this process
  unsigned long long* err_word;             // device alias of a mapped pinned word: set to PB_XCHG_ERROR_SEQ << 32 on a time-out
  unsigned int seq;                         // sequence number of this vector exchange (own counter; never 0 / PB_XCHG_ERROR_SEQ)
  int rank, world;
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void st_word(unsigned long long* p, unsigned long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_word(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Executed by ALL threads of one CTA (blockDim.x >= world * 32) after `local_block` is final and visible to the CTA.
__device__ __forceinline__ void xchg_push_wait(const XchgParams& xp, const double* local_block) {
  const int t = threadIdx.x;
  const int par = (int)(xp.seq & 1u);
  const int nwords = xp.world * PB_XCHG_WORDS_PER_ROW;
  if (t >= nwords) return;
  const int r = t / PB_XCHG_WORDS_PER_ROW, w = t % PB_XCHG_WORDS_PER_ROW;
  const unsigned long long bits = (unsigned long long)__double_as_longlong(__ldcg(local_block + (w >> 1)));
  const unsigned int half = (w & 1) ? (unsigned int)(bits >> 32) : (unsigned int)(bits & 0xffffffffull);
  const unsigned long long word = ((unsigned long long)xp.seq << 32) | half;
  unsigned long long v;
  if (r == xp.rank) {
    // my own row needs no trip through memory: forward it straight to the host
    v = word;
  } else {
    // 1. push word w of my row into slot [par][rank] of rank r's buffer (NVLink peer store)
    st_word(xp.peer[r] + ((size_t)(par * PB_MAX_RANKS + xp.rank) * PB_XCHG_WORDS_PER_ROW + w), word);
    // 2. wait for word w of rank r's row in MY buffer (busy poll: 8 bytes from L2, a handful of threads)
    const unsigned long long* mine = xp.peer[xp.rank] + ((size_t)(par * PB_MAX_RANKS + r) * PB_XCHG_WORDS_PER_ROW + w);
    const unsigned long long t0 = globaltimer_ns();
    unsigned int spins = 0;
    v = ld_word(mine);
    while ((unsigned int)(v >> 32) != xp.seq) {
      if ((++spins & 0x3ff) == 0 && globaltimer_ns() - t0 > PB_XCHG_TIMEOUT_NS) {
        v = ((unsigned long long)PB_XCHG_ERROR_SEQ << 32);
        break;
      }
      v = ld_word(mine);
    }
  }
  // 3. forward to the host (landing zone double-buffered by parity too, so that the host may still be reading exchange k
  //    while the kernel of exchange k+1 -- launched ahead by the pipelined driver loop -- publishes)
  st_word(xp.host_words + (size_t)par * PB_MAX_RANKS * PB_XCHG_WORDS_PER_ROW + t, v);
}

// Variant for kernels that consume the rows themselves (the persistent multi-iteration step kernel, step_multi.cu): same push / poll
// protocol, but the P rows are decoded into shared memory (`rows_sh`: world x PB_NSCALARS doubles, rank order) instead of being
// forwarded to the host.  Executed by ALL threads of one CTA (blockDim.x >= world * 32); returns false if a peer timed out.
__device__ __forceinline__ bool xchg_push_gather(const XchgParams& xp, const double* local_block, double* rows_sh) {
  const int t = threadIdx.x;
  const int par = (int)(xp.seq & 1u);
  const int nwords = xp.world * PB_XCHG_WORDS_PER_ROW;
  int bad = 0;
  if (t < nwords) {
    const int r = t / PB_XCHG_WORDS_PER_ROW, w = t % PB_XCHG_WORDS_PER_ROW;
    const unsigned long long bits = (unsigned long long)__double_as_longlong(local_block[w >> 1]);
    const unsigned int half = (w & 1) ? (unsigned int)(bits >> 32) : (unsigned int)(bits & 0xffffffffull);
    const unsigned long long word = ((unsigned long long)xp.seq << 32) | half;
    unsigned long long v = word;
    if (r != xp.rank) {
      st_word(xp.peer[r] + ((size_t)(par * PB_MAX_RANKS + xp.rank) * PB_XCHG_WORDS_PER_ROW + w), word);
      const unsigned long long* mine = xp.peer[xp.rank] + ((size_t)(par * PB_MAX_RANKS + r) * PB_XCHG_WORDS_PER_ROW + w);
      const unsigned long long t0 = globaltimer_ns();
      unsigned int spins = 0;
      v = ld_word(mine);
      while ((unsigned int)(v >> 32) != xp.seq) {
        if ((++spins & 0x3ff) == 0 && globaltimer_ns() - t0 > PB_XCHG_TIMEOUT_NS) {
          bad = 1;
          break;
        }
        v = ld_word(mine);
      }
    }
    reinterpret_cast<unsigned int*>(rows_sh)[r * PB_XCHG_WORDS_PER_ROW + w] = (unsigned int)(v & 0xffffffffull);
  }
  return __syncthreads_or(bad) == 0;
}
