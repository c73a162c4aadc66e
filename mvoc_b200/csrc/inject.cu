// mvoc_attn_inject_fwd: the injected self-attention of MVOC's composite processors as one C-ABI call —
// the Q/K mask blend (blend.cu) followed by the attention of all branches (attn_tc.cu / attn_temporal.cu)
// on the same stream.  Two launches today; the entry point is the seam behind which the blend moves into the
// attention kernel's Q/K tile loads.
#include "common.cuh"

extern "C" int mvoc_attn_inject_fwd(void* q, void* k, const void* v, void* o, int n_obj, int frames, int64_t pixels,
                                    int H, int D, const void* mask, int mask_kind, int base_slot, int mode,
                                    int share_p, float scale, int dtype, int variant, void* stream) {
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
    if (share_p != 0) {
        set_error("mvoc_attn_inject_fwd: share_p is reserved (one softmax for the uncond/cond pair) and must be 0");
        return MVOC_ERR_UNSUPPORTED;
    }
    if (mode == MVOC_INJECT_SPATIAL && pixels > 0x7fffffffLL) {
        set_error("mvoc_attn_inject_fwd: %lld tokens per frame", (long long)pixels);
        return MVOC_ERR_UNSUPPORTED;
    }
    const int nb = n_obj + 3;
    const int64_t C = (int64_t)H * D;
    int rc = mvoc_qk_blend(q, k, n_obj, (int64_t)frames * pixels, (int)C, mask, mask_kind, base_slot, dtype, stream);
    if (rc != MVOC_OK) return rc;
    if (mode == MVOC_INJECT_SPATIAL) {
        const int64_t sb = pixels * C, sn = C, sh = D;
        return mvoc_attn_fwd(q, k, v, o, nb * frames, H, (int)pixels, (int)pixels, D, sb, sn, sh, sb, sn, sh, sb, sn, sh,
                             sb, sn, sh, scale, dtype, variant, stream);
    }
    // temporal: problem (branch, pixel); its frames are one frame (pixels * C elements) apart
    int64_t st[16];
    for (int i = 0; i < 4; ++i) {
        st[4 * i + 0] = (int64_t)frames * pixels * C;
        st[4 * i + 1] = C;
        st[4 * i + 2] = pixels * C;
        st[4 * i + 3] = D;
    }
    return mvoc_attn_temporal_strided_fwd(q, k, v, o, nb, pixels, frames, H, D, st, scale, dtype, stream);
}
