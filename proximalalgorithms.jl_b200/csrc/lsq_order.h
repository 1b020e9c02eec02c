// lsq_order.h -- which summation order the least-squares products use for a given block shape.  ONE definition shared by
// the launchers of lsq_kernels.cu and by the on-device driver loop (persist.cu), which restates the same orders inside a
// persistent kernel: both therefore produce bit-identical r = A x - b and grad = A' r, and a solve takes the same
// line-search decisions whichever path runs it.  The order depends on the block shape and on pointer alignment only --
// never on the SM count, the grid, or how blocks are sharded over GPUs.
#pragma once
#include <stdint.h>

#define PB_GEMV_CL 4        // column lanes of the row-per-thread residual order
#define PB_GEMV_UNROLL 8

struct PbLsqOrder {
  int64_t chunk_cols, nchunk;   // column chunking of r = A x - b (partials folded in chunk order)
  int n_sub, n_lpc, n_kp;       // residual: 1 = short-column order (k_gemv_n_sub<LPC, KP>), 0 = row-per-thread order
  int t_sub, t_lpc, t_kp;       // gradient: 1 = sub-warp-per-column order (k_gemv_t_sub<LPC, KP>), 0 = warp-per-column
};

static inline bool pb_order_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// elt = sizeof(T); A: first block; r: the residual vector the gradient kernel reads
static inline PbLsqOrder pb_lsq_order(size_t elt, int64_t nblk, int64_t mb, int64_t nb, int64_t lda, int64_t blk_stride,
                                      const void* A, const void* r) {
  PbLsqOrder o;
  const int64_t VEC = (int64_t)(16 / elt);
  // ceil(nb/64) columns per chunk, clamped to [32, 4096]  ->  <= 64 chunks (more only for nb > 262144)
  int64_t chunk_cols = (nb + 63) / 64;
  if (chunk_cols < PB_GEMV_CL * PB_GEMV_UNROLL) chunk_cols = PB_GEMV_CL * PB_GEMV_UNROLL;
  if (chunk_cols > 4096) chunk_cols = 4096;
  // a multiple of 4 columns (the clamps already are): chunk boundaries then never cut a 16-byte pack of the n-vectors, which lets
  // lsq_fista.cu run the fused step -- whose float reductions are grouped per pack -- on the columns of one chunk
  chunk_cols = (chunk_cols + 3) & ~(int64_t)3;
  o.chunk_cols = chunk_cols;
  o.nchunk = nb > 0 ? (nb + chunk_cols - 1) / chunk_cols : 1;
  const int64_t npk = mb / VEC;
  const bool shape_ok = mb % VEC == 0 && lda % VEC == 0 && blk_stride % VEC == 0 && pb_order_aligned16(A);
  const int kp = npk <= 1 ? 1 : (npk <= 2 ? 2 : 4);
  int lpc = 1;
  while ((int64_t)lpc * kp < npk) lpc <<= 1;
  o.n_sub = (mb < 64 && shape_ok && nblk * o.nchunk <= 0x7fffffffLL) ? 1 : 0;
  o.n_lpc = lpc;
  o.n_kp = kp;
  o.t_sub = (mb > 0 && shape_ok && npk <= 32 * 4 && pb_order_aligned16(r)) ? 1 : 0;
  o.t_lpc = lpc;
  o.t_kp = kp;
  return o;
}
