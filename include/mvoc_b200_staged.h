/*
 * mvoc_b200_staged.h — entry points STAGED for the next round.
 *
 * These kernels compile for sm_100a but HAVE NOT RUN ON HARDWARE YET.  They live in a separate library,
 * mvoc_b200/lib/libmvoc_b200_staged.so, so that the validated product library (libmvoc_b200.so, mvoc_b200.h)
 * is not affected by them; nothing under mvoc_b200/ calls them unless MVOC_STAGED=1 is set, and their tests
 * (tests/test_staged.py) only run on a GPU with MVOC_STAGED=1.  Conventions as in mvoc_b200.h.
 */
#ifndef MVOC_B200_STAGED_H
#define MVOC_B200_STAGED_H

#include <stdint.h>
#include "mvoc_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/*
 * 3x3 convolution, stride 1, padding 1, channels-last, on tcgen05 tensor cores (implicit GEMM: the nine taps
 * are nine shifted TMA boxes, zero-filled at the borders; no im2col buffer).
 * Replaces the cuDNN call behind self.conv1 / self.conv2 of the resnet forward, i2vgen-xl/pnp_utils.py:939, :968
 * (SURVEY §8f-1).
 *
 * x: [N, H, W, Cin] bf16, contiguous.    w_taps: [9, Cout, Cin] bf16 — tap kh*3+kw outermost, i.e.
 * torch weight [Cout, Cin, 3, 3].permute(2, 3, 0, 1).reshape(9, Cout, Cin).
 * bias: [Cout] bf16 or NULL.   residual: [N, H, W, Cout] bf16 added to the result, or NULL.
 * out: [N, H, W, Cout] bf16.   Cin % 64 == 0, Cout % 64 == 0.
 * variant 0: widest output-channel tile dividing Cout (320, 160, 128, 64); 1: never the 320-column tile;
 * 2 (Cout % 320 == 0): 320-column tiles in clusters of two CTAs on neighbouring pixel tiles that share the
 *    weight tile — each CTA loads one half of it and TMA-multicasts it to both (half the L2 traffic of the weights).
 */
int mvoc_conv3x3_nhwc(const void* x, const void* w_taps, const void* bias, const void* residual, void* out,
                      int N, int H, int W, int Cin, int Cout, int dtype, int variant, void* stream);

/*
 * GEGLU feed-forward input projection with the gate applied in the GEMM epilogue:
 *   out[m, j] = (x[m] . w[j] + bias[j]) * gelu(x[m] . w[F + j] + bias[F + j]),   exact (erf) GELU.
 * Replaces ff.net[0] (diffusers GEGLU: Linear(K, 2F) -> chunk -> value * gelu(gate)) reached through
 * i2vgen-xl/pnp_utils.py:335; the [M, 2F] intermediate is never written.
 *
 * x: [M, K] bf16, w: [2F, K] bf16 (torch Linear weight), bias: [2F] bf16 or NULL, out: [M, F] bf16.
 * K % 64 == 0, F % 64 == 0.
 */
int mvoc_linear_geglu(const void* x, const void* w, const void* bias, void* out, int64_t M, int K, int F,
                      int dtype, void* stream);

/*
 * The attention entry point of mvoc_b200.h with every query row split across two softmax threads: eight softmax warps per
 * CTA instead of four, to hide the latency of the score-row dependency chain that the ncu capture of the product
 * kernel shows (no saturated pipe).  Same arguments, layouts and restrictions.
 * variant 0: 3 of every 8 exp2 pairs on the FMA-pipe polynomial (as the product default); 1: all on the MUFU;
 * 2: 4 of 8.
 */
int mvoc_attn_fwd_split(const void* q, const void* k, const void* v, void* o,
                        int B, int H, int Nq, int Nk, int D,
                        int64_t q_sb, int64_t q_sn, int64_t q_sh,
                        int64_t k_sb, int64_t k_sn, int64_t k_sh,
                        int64_t v_sb, int64_t v_sn, int64_t v_sh,
                        int64_t o_sb, int64_t o_sn, int64_t o_sh,
                        float scale, int dtype, int variant, void* stream);

#ifdef __cplusplus
}
#endif
#endif
