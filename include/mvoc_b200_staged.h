/*
 * mvoc_b200_staged.h — entry points STAGED for the next round.
 *
 * Experimental kernel variants kept out of the product library until a hardware run decides their fate.  They live in a separate library,
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
