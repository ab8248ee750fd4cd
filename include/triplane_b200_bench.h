/*
 * triplane_b200_bench.h -- C ABI of libtriplane_b200_bench.so: measurement and bring-up aids for the sm_100a
 * tri-plane renderer.  NOT part of the product library (include/triplane_b200.h); the renderer never loads it.
 * Used by bench.py (the live L2-gather ceiling of the roofline object), profiles/*.py and tests/test_gpu_tc_debug.py.
 * Same conventions as triplane_b200.h: device pointers, caller-owned memory, asynchronous on `stream`.
 */
#ifndef TRIPLANE_B200_BENCH_H_
#define TRIPLANE_B200_BENCH_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- measurement aid: the gather roofline ------------------------- */
/* Fetches random 128-byte lines of buf[n_lines*32 floats] with the render kernels' access shape (8 lanes x
 * 16 bytes per line, 12 lines in flight per thread) from `ctas` CTAs of 512 threads, `iters` rounds each.
 * sink: 65536 floats.  Returns the number of lines fetched (> 0) or a negative error.  The caller times
 * it with CUDA events; lines * 128 B / time is the L2 (small buffer) or DRAM (large buffer) gather bandwidth
 * bench.py reports beside the HBM copy peak. */
int64_t tpr_gather_microbench(const float* buf, int64_t n_lines, int32_t ctas, int32_t iters, float* sink,
                              void* stream);
/* same with the CTA size (multiple of 32, <= 1024) and the lines in flight per thread (4, 6, 12 or 24) chosen
 * by the caller; sink: 65536 floats */
int64_t tpr_gather_microbench_ex(const float* buf, int64_t n_lines, int32_t ctas, int32_t threads,
                                 int32_t in_flight, int32_t iters, float* sink, void* stream);

/* Variants of the gather shape: vec_floats = 4 (LDG.128, eight lanes per line) or 8 (LDG.256, four lanes per line); in_flight loads
 * per burst; pipelined != 0: the next burst is issued before the previous one is consumed.  iters must be even.
 * Supported (vec, in_flight, pipelined): (4, 2|4, 0|1), (4, 6|8, 0), (8, 1|2|4, 0|1).  Returns the lines fetched. */
int64_t tpr_gather_microbench_v2(const float* buf, int64_t n_lines, int32_t ctas, int32_t threads, int32_t vec_floats,
                                 int32_t in_flight, int32_t pipelined, int32_t iters, float* sink, void* stream);

/* The backward pass's scatter shape: eight lanes add 16 bytes each (red.global.add.v4.f32) to random 128-byte lines of
 * buf[n_lines*32 floats]; per_iter lines per eight-lane group and round, `iters` rounds.  Returns the lines added to. */
int64_t tpr_scatter_microbench(float* buf, int64_t n_lines, int32_t ctas, int32_t threads, int32_t per_iter, int32_t iters,
                               void* stream);

/* Issues `count` tcgen05.mma (M = 128, N = n, one K step; tf32 or bf16 operands; A from shared memory or TMEM)
 * from one thread of one CTA; tight != 0 issues them from precomputed descriptors (count % 4 == 0).
 * out_dev[0] = cycles spent issuing, out_dev[1] = cycles until all have completed. */
int tpr_mma_microbench(int32_t n, int32_t bf16, int32_t a_from_tmem, int32_t count, int32_t tight, long long* out_dev,
                       void* stream);

/* Raw tcgen05 decoder plumbing, one 128-row tile per CTA: hidden = x . W1t + b1 (layer 1, SS operands), out = softplus(hidden)
 * . W2t + b2 (layer 2, A from TMEM).  mode 0 = tf32, 1 = 3xTF32, 2 = bf16.  x [n_rows,32], hidden [n_rows,64] (pre-activation),
 * out [n_rows,48] (36 used).  decoder_packed: tpr_pack_decoder's block. */
int tpr_debug_tc_decode(const float* x, int64_t n_rows, const float* decoder_packed, int32_t mode, float* hidden, float* out,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TRIPLANE_B200_BENCH_H_ */
