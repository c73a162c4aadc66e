// mvoc_attn_inject_fwd: the injected self-attention of MVOC's composite processors as one C-ABI call —
// what ModifiedSpaAttnProcessor / ModifiedTmpAttnProcessor do between the q/k/v projections and to_out
// (i2vgen-xl/pnp_utils.py:624-686, :778-864).
//
// Spatial mode with share_p: the blend kernel writes ONE copy of the blended Q', K' (the reference writes the same
// tensor to the uncond and the cond chunk, :664-668), the source branches run the plain attention kernel and the
// two composite branches run the pair kernel (attn_tc.cu, kNV = 2): one softmax, two P.V products.
// q, k, v, o may be column slices of wider row-major buffers (row strides ld_*), e.g. of one fused QKV GEMM output.
//
// Why the blend stays a separate (HBM-bound, ~40 us at the largest level) launch instead of moving into the
// attention kernel's tile loads: the blended K' tile of a key block is consumed by EVERY query tile of the frame
// (32 at 64x64 latents), so blending at load time would redo the select 32 times and read up to 1 + n_obj source
// tiles per key block; blended once, K' costs one write and is then read exactly like any other K.
#include "common.cuh"

extern "C" int mvoc_attn_inject_fwd(void* q, void* k, const void* v, void* o, int64_t ld_q, int64_t ld_k, int64_t ld_v,
                                    int64_t ld_o, int n_obj, int frames, int64_t pixels, int H, int D, const void* mask,
                                    int mask_kind, int base_slot, int mode, int share_p, float scale, int dtype,
                                    int variant, void* stream) {
    using mvoc::set_error;
    if (!q || !k || !v || !o || !mask) {
        set_error("mvoc_attn_inject_fwd: null pointer");
        return MVOC_ERR_INVALID_ARG;
    }
    if (n_obj < 1 || n_obj > MVOC_MAX_OBJECTS || frames < 1 || pixels < 1 || H < 1) {
        set_error("mvoc_attn_inject_fwd: n_obj=%d frames=%d pixels=%lld H=%d out of range", n_obj, frames,
                  (long long)pixels, H);
        return MVOC_ERR_INVALID_ARG;
    }
    if (mode != MVOC_INJECT_SPATIAL && mode != MVOC_INJECT_TEMPORAL) {
        set_error("mvoc_attn_inject_fwd: mode %d (0 = spatial, 1 = temporal)", mode);
        return MVOC_ERR_INVALID_ARG;
    }
    if (share_p != 0 && mode != MVOC_INJECT_SPATIAL) {
        set_error("mvoc_attn_inject_fwd: share_p applies to the spatial mode (the temporal kernel is HBM-bound)");
        return MVOC_ERR_UNSUPPORTED;
    }
    if (mode == MVOC_INJECT_SPATIAL && pixels > 0x7fffffffLL) {
        set_error("mvoc_attn_inject_fwd: %lld tokens per frame", (long long)pixels);
        return MVOC_ERR_UNSUPPORTED;
    }
    const int nb = n_obj + 3;
    const int64_t C = (int64_t)H * D;
    const int64_t lds[4] = {ld_q, ld_k, ld_v, ld_o};
    for (int i = 0; i < 4; ++i)
        if (lds[i] < C || lds[i] % 8 != 0) {
            set_error("mvoc_attn_inject_fwd: row stride #%d = %lld must be a multiple of 8 and >= H*D = %lld", i,
                      (long long)lds[i], (long long)C);
            return MVOC_ERR_INVALID_ARG;
        }
    const int64_t tokens = (int64_t)frames * pixels;
    const int esize = 2;
    int rc;
    // blend Q and K (their row strides may differ: two launches when they do)
    if (ld_q == ld_k) {
        rc = mvoc_qk_blend_strided(q, k, n_obj, tokens, (int)C, ld_q, mask, mask_kind, base_slot, share_p, dtype, stream);
    } else {
        rc = mvoc_qk_blend_strided(q, nullptr, n_obj, tokens, (int)C, ld_q, mask, mask_kind, base_slot, share_p, dtype,
                                   stream);
        if (rc == MVOC_OK)
            rc = mvoc_qk_blend_strided(k, nullptr, n_obj, tokens, (int)C, ld_k, mask, mask_kind, base_slot, share_p,
                                       dtype, stream);
    }
    if (rc != MVOC_OK) return rc;
    if (mode == MVOC_INJECT_SPATIAL) {
        const int64_t qb = pixels * ld_q, kb = pixels * ld_k, vb = pixels * ld_v, ob = pixels * ld_o;
        if (!share_p)
            return mvoc_attn_fwd(q, k, v, o, nb * frames, H, (int)pixels, (int)pixels, D, qb, ld_q, D, kb, ld_k, D, vb,
                                 ld_v, D, ob, ld_o, D, scale, dtype, variant, stream);
        // sources: slots 0..n_obj with their own Q, K, V
        const int src_b = (n_obj + 1) * frames;
        rc = mvoc_attn_fwd(q, k, v, o, src_b, H, (int)pixels, (int)pixels, D, qb, ld_q, D, kb, ld_k, D, vb, ld_v, D, ob,
                           ld_o, D, scale, dtype, variant, stream);
        if (rc != MVOC_OK) return rc;
        // composites: the blended Q', K' of the uncond slot serve both; V and O of slots n_obj+1 and n_obj+2
        const int64_t slot = (int64_t)src_b;   // first batch of the uncond slot
        const char* qu = (const char*)q + slot * qb * esize;
        const char* ku = (const char*)k + slot * kb * esize;
        const char* vu = (const char*)v + slot * vb * esize;
        char* ou = (char*)o + slot * ob * esize;
        return mvoc_attn_pair_fwd(qu, ku, vu, ou, frames, H, (int)pixels, (int)pixels, D, qb, ld_q, D, kb, ld_k, D, vb,
                                  ld_v, D, ob, ld_o, D, frames, scale, dtype, variant, stream);
    }
    // temporal: problem (branch, pixel); its frames are one frame (pixels rows) apart
    int64_t st[16];
    for (int i = 0; i < 4; ++i) {
        st[4 * i + 0] = (int64_t)frames * pixels * lds[i];
        st[4 * i + 1] = lds[i];
        st[4 * i + 2] = pixels * lds[i];
        st[4 * i + 3] = D;
    }
    return mvoc_attn_temporal_strided_fwd(q, k, v, o, nb, pixels, frames, H, D, st, scale, dtype, stream);
}
