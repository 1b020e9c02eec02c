// xchg.cuh -- C1: the per-iteration exchange of the scalar block, done by the GPU itself.
//
// Every rank owns an exchange buffer in device memory (cudaMalloc, exported with cudaIpc and mapped by all peers over
// NVLink).  The last CTA of a reducing kernel (or the 1-CTA k_xchg kernel) pushes this rank's PB_NSCALARS doubles into
// slot [parity][rank] of EVERY rank's buffer with plain peer stores, publishes a sequence number with a system-scope
// release, waits until all P sequence numbers of its own buffer have arrived, and copies the P rows into mapped pinned
// host memory followed by a host-visible flag.  The host then needs neither a collective launch nor a cudaMemcpy nor a
// stream synchronisation: it polls the pinned flag (pb_exchange_wait) and folds the rows in rank order in double-double.
// With one rank this degenerates to a zero-copy read-back of the scalar block.
//
// Double buffering by the parity of the sequence number is sufficient: a rank can only complete exchange k+1 after every
// peer has pushed k+1, which a peer does after its own exchange-k kernel (same stream) has finished reading parity k.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "proxb200.h"

#define PB_MAX_RANKS 16
#define PB_XCHG_BLOCK_DOUBLES (2 * PB_MAX_RANKS * PB_NSCALARS)                 // blocks[2][MAXR][16]
#define PB_XCHG_BYTES (PB_XCHG_BLOCK_DOUBLES * 8 + 2 * PB_MAX_RANKS * 8)       // + flags[2][MAXR] (u64)
#define PB_XCHG_TIMEOUT_NS 4000000000ull                                       // give up after 4 s (never hang the GPU)
#define PB_XCHG_ERROR_FLAG 0xFFFFFFFFFFFFFFFFull

struct XchgParams {
  double* peer[PB_MAX_RANKS];        // peer[r]: base of rank r's exchange buffer as mapped in THIS process
  double* host_rows;                 // device alias of mapped pinned memory: [world][PB_NSCALARS]
  unsigned long long* host_flag;     // device alias of the mapped pinned sequence flag
  unsigned long long seq;            // sequence number of this exchange (> 0)
  int rank, world;                   // world == 0: exchange disabled
};

__device__ __forceinline__ unsigned long long* xchg_flags(double* base) {
  return reinterpret_cast<unsigned long long*>(base + PB_XCHG_BLOCK_DOUBLES);
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Executed by ALL threads of one CTA (blockDim.x >= world * PB_NSCALARS) after `local_block` is final and visible to the
// CTA.  Contains __syncthreads.
__device__ __forceinline__ void xchg_push_wait(const XchgParams& xp, const double* local_block) {
  const int t = threadIdx.x;
  const int par = (int)(xp.seq & 1ull);
  const int nelem = xp.world * PB_NSCALARS;
  __shared__ int timed_out;
  if (t == 0) timed_out = 0;
  // 1. push my row into slot [par][rank] of every rank's buffer (NVLink peer stores; self included)
  if (t < nelem) {
    const int r = t / PB_NSCALARS, i = t % PB_NSCALARS;
    const double v = __ldcg(local_block + i);
    double* dst = xp.peer[r] + ((size_t)(par * PB_MAX_RANKS + xp.rank) * PB_NSCALARS + i);
    *reinterpret_cast<volatile double*>(dst) = v;
    __threadfence_system();
  }
  __syncthreads();
  // 2. publish the sequence number to every rank, 3. wait for everybody's number in my own buffer
  if (t < xp.world) {
    st_release_sys(xchg_flags(xp.peer[t]) + par * PB_MAX_RANKS + xp.rank, xp.seq);
    const unsigned long long* mine = xchg_flags(xp.peer[xp.rank]) + par * PB_MAX_RANKS + t;
    const unsigned long long t0 = globaltimer_ns();
    while (ld_acquire_sys(mine) != xp.seq) {
      if (globaltimer_ns() - t0 > PB_XCHG_TIMEOUT_NS) {
        timed_out = 1;
        break;
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
  // 4. rows -> mapped pinned host memory, then the host-visible flag
  if (t < nelem) {
    const double v = __ldcg(xp.peer[xp.rank] + ((size_t)par * PB_MAX_RANKS * PB_NSCALARS + t));
    *reinterpret_cast<volatile double*>(xp.host_rows + t) = v;
    __threadfence_system();
  }
  __syncthreads();
  if (t == 0) st_release_sys(xp.host_flag, timed_out ? PB_XCHG_ERROR_FLAG : xp.seq);
}
