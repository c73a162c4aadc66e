// Mask-blend injection kernels (HBM-bound, 128-bit vectorised).
//   mvoc_qk_blend      — Q/K of the composite slots <- blend of source slots
//                        (reference: i2vgen-xl/pnp_utils.py:628-672, :782-850)
//   mvoc_feature_blend — hidden states after resnet/temp-conv/conv_out
//                        (reference: i2vgen-xl/pnp_utils.py:970-1004, :1059-1082, :1114-1146)
//
// Binary masks make `x*(1-m) + obj*m` a select, so the kernels copy the
// selected source bit-exactly and only read the slot that wins; the mask is
// read once per token (qk) / once per 8 pixels (feature), never per channel.
#include "common.cuh"

namespace mvoc {

struct QKBlendParams {
    void* x[2];
    const void* mask;
    int64_t tokens;
    int64_t ld;           // row stride in elements (>= C): rows may be column slices of a wider buffer
    int64_t chunk_elems;  // tokens * ld: distance between branch slots
    int C;
    int n_obj;
    int base_slot;
    int single;           // write the blend to the uncond slot only (the pair kernel reads it for both composites)
};

// One item = one 16-byte piece of one token row.  blockIdx.y picks Q or K.
template <typename T, bool kSoft>
__global__ void __launch_bounds__(256) qk_blend_kernel(QKBlendParams p) {
    T* x = reinterpret_cast<T*>(p.x[blockIdx.y]);
    const int vec_per_tok = p.C >> 3;
    const int64_t items = p.tokens * vec_per_tok;
    const int u_slot = p.n_obj + 1, c_slot = p.n_obj + 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += stride) {
        const int64_t tok = it / vec_per_tok;
        const int64_t off = tok * p.ld + ((it - tok * vec_per_tok) << 3);  // element offset inside a slot
        if (!kSoft) {
            const uint8_t* m = reinterpret_cast<const uint8_t*>(p.mask);
            int src = p.base_slot;
#pragma unroll 4
            for (int j = 0; j < p.n_obj; ++j)
                if (__ldg(m + (int64_t)j * p.tokens + tok)) src = j + 1;
            const Vec16 v = ld_stream16(x + (int64_t)src * p.chunk_elems + off);
            st_stream16(x + (int64_t)u_slot * p.chunk_elems + off, v);
            if (src != c_slot && !p.single) st_stream16(x + (int64_t)c_slot * p.chunk_elems + off, v);
        } else {
            const float* m = reinterpret_cast<const float*>(p.mask);
            float acc[8];
            unpack8<T>(ld_stream16(x + (int64_t)p.base_slot * p.chunk_elems + off), acc);
            bool touched = false;
            for (int j = 0; j < p.n_obj; ++j) {
                const float mj = __ldg(m + (int64_t)j * p.tokens + tok);
                if (mj != 0.0f) {
                    float o[8];
                    unpack8<T>(ld_stream16(x + (int64_t)(j + 1) * p.chunk_elems + off), o);
                    const float w = 1.0f - mj;
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[e] = acc[e] * w + o[e] * mj;
                    touched = true;
                }
            }
            const Vec16 v = pack8<T>(acc);
            st_stream16(x + (int64_t)u_slot * p.chunk_elems + off, v);
            if ((touched || p.base_slot != c_slot) && !p.single)
                st_stream16(x + (int64_t)c_slot * p.chunk_elems + off, v);
        }
    }
}

struct FeatBlendParams {
    void* x;
    const uint8_t* mask;  // [n_obj, T, HW]
    int64_t HW;
    int64_t slot_elems;  // T*C*HW
    int T, C, n_obj;
};

// One item = 8 consecutive pixels of one (frame, channel) plane.
template <typename T>
__global__ void __launch_bounds__(256) feature_blend_kernel(FeatBlendParams p) {
    T* x = reinterpret_cast<T*>(p.x);
    const int64_t vec_per_plane = p.HW >> 3;
    const int64_t items = (int64_t)p.T * p.C * vec_per_plane;
    const int u_slot = p.n_obj + 1, c_slot = p.n_obj + 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += stride) {
        const int64_t plane = it / vec_per_plane;  // f*C + c
        const int64_t pv = it - plane * vec_per_plane;
        const int f = (int)(plane / p.C);
        const int64_t off = it << 3;               // (f*C + c)*HW + pv*8
        const int64_t moff = (int64_t)f * p.HW + (pv << 3);
        Vec16 acc = ld_stream16(x + off);          // slot 0 = background
        uint16_t* a16 = reinterpret_cast<uint16_t*>(&acc);
        for (int j = 0; j < p.n_obj; ++j) {
            const uint2 mm = __ldg(reinterpret_cast<const uint2*>(
                p.mask + (int64_t)j * p.T * p.HW + moff));
            if ((mm.x | mm.y) != 0u) {
                const Vec16 o = ld_stream16(x + (int64_t)(j + 1) * p.slot_elems + off);
                const uint16_t* o16 = reinterpret_cast<const uint16_t*>(&o);
                const uint8_t* mb = reinterpret_cast<const uint8_t*>(&mm);
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    if (mb[e]) a16[e] = o16[e];
            }
        }
        st_stream16(x + (int64_t)u_slot * p.slot_elems + off, acc);
        st_stream16(x + (int64_t)c_slot * p.slot_elems + off, acc);
    }
}

static inline int grid_for(int64_t items, int threads, int ctas_per_sm) {
    int64_t want = (items + threads - 1) / threads;
    int64_t cap = (int64_t)num_sms() * ctas_per_sm;
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace mvoc

using namespace mvoc;

extern "C" int mvoc_qk_blend(void* x0, void* x1, int n_obj, int64_t tokens, int C,
                             const void* mask, int mask_kind, int base_slot, int dtype,
                             void* stream) {
    return mvoc_qk_blend_strided(x0, x1, n_obj, tokens, C, C, mask, mask_kind, base_slot, 0, dtype, stream);
}

extern "C" int mvoc_qk_blend_strided(void* x0, void* x1, int n_obj, int64_t tokens, int C, int64_t ld,
                                     const void* mask, int mask_kind, int base_slot, int single, int dtype,
                                     void* stream) {
    MVOC_REQUIRE(x0 != nullptr && mask != nullptr, MVOC_ERR_INVALID_ARG,
                 "mvoc_qk_blend: null pointer");
    MVOC_REQUIRE(ld >= C && ld % 8 == 0, MVOC_ERR_INVALID_ARG,
                 "mvoc_qk_blend: row stride %lld must be a multiple of 8 and >= C=%d", (long long)ld, C);
    MVOC_REQUIRE(n_obj >= 1 && n_obj <= MVOC_MAX_OBJECTS, MVOC_ERR_INVALID_ARG,
                 "mvoc_qk_blend: n_obj=%d out of range [1,%d]", n_obj, MVOC_MAX_OBJECTS);
    MVOC_REQUIRE(base_slot == 0 || base_slot == n_obj + 2, MVOC_ERR_INVALID_ARG,
                 "mvoc_qk_blend: base_slot=%d must be 0 (background) or n_obj+2 (cond)", base_slot);
    MVOC_REQUIRE(C > 0 && C % 8 == 0, MVOC_ERR_UNSUPPORTED,
                 "mvoc_qk_blend: C=%d must be a positive multiple of 8", C);
    MVOC_REQUIRE(mask_kind == MVOC_MASK_U8 || mask_kind == MVOC_MASK_F32, MVOC_ERR_INVALID_ARG,
                 "mvoc_qk_blend: unknown mask_kind %d", mask_kind);
    MVOC_REQUIRE(dtype == MVOC_BF16 || dtype == MVOC_F16, MVOC_ERR_UNSUPPORTED,
                 "mvoc_qk_blend: dtype %d unsupported (bf16/f16 only)", dtype);
    MVOC_REQUIRE(((uintptr_t)x0 % 16 == 0) && ((uintptr_t)x1 % 16 == 0), MVOC_ERR_INVALID_ARG,
                 "mvoc_qk_blend: pointers must be 16-byte aligned");
    if (tokens == 0) return MVOC_OK;
    QKBlendParams p;
    p.x[0] = x0;
    p.x[1] = x1;
    p.mask = mask;
    p.tokens = tokens;
    p.ld = ld;
    p.chunk_elems = tokens * ld;
    p.C = C;
    p.n_obj = n_obj;
    p.base_slot = base_slot;
    p.single = single ? 1 : 0;
    const int64_t items = tokens * (C / 8);
    dim3 grid(grid_for(items, 256, 16), x1 ? 2 : 1);
    cudaStream_t s = (cudaStream_t)stream;
    const bool soft = mask_kind == MVOC_MASK_F32;
    if (dtype == MVOC_BF16) {
        if (soft) qk_blend_kernel<__nv_bfloat16, true><<<grid, 256, 0, s>>>(p);
        else qk_blend_kernel<__nv_bfloat16, false><<<grid, 256, 0, s>>>(p);
    } else {
        if (soft) qk_blend_kernel<__half, true><<<grid, 256, 0, s>>>(p);
        else qk_blend_kernel<__half, false><<<grid, 256, 0, s>>>(p);
    }
    return check_launch("mvoc_qk_blend");
}

extern "C" int mvoc_feature_blend(void* x, int n_obj, int T, int C, int64_t HW,
                                  const void* mask, int dtype, void* stream) {
    MVOC_REQUIRE(x != nullptr && mask != nullptr, MVOC_ERR_INVALID_ARG,
                 "mvoc_feature_blend: null pointer");
    MVOC_REQUIRE(n_obj >= 1 && n_obj <= MVOC_MAX_OBJECTS, MVOC_ERR_INVALID_ARG,
                 "mvoc_feature_blend: n_obj=%d out of range [1,%d]", n_obj, MVOC_MAX_OBJECTS);
    MVOC_REQUIRE(T > 0 && C > 0 && HW > 0 && HW % 8 == 0, MVOC_ERR_UNSUPPORTED,
                 "mvoc_feature_blend: need T,C>0 and HW%%8==0 (T=%d C=%d HW=%lld)", T, C,
                 (long long)HW);
    MVOC_REQUIRE(dtype == MVOC_BF16 || dtype == MVOC_F16, MVOC_ERR_UNSUPPORTED,
                 "mvoc_feature_blend: dtype %d unsupported (bf16/f16 only)", dtype);
    MVOC_REQUIRE(((uintptr_t)x % 16 == 0) && ((uintptr_t)mask % 8 == 0), MVOC_ERR_INVALID_ARG,
                 "mvoc_feature_blend: x must be 16-byte and mask 8-byte aligned");
    FeatBlendParams p;
    p.x = x;
    p.mask = reinterpret_cast<const uint8_t*>(mask);
    p.HW = HW;
    p.slot_elems = (int64_t)T * C * HW;
    p.T = T;
    p.C = C;
    p.n_obj = n_obj;
    const int64_t items = (int64_t)T * C * (HW / 8);
    const int grid = grid_for(items, 256, 16);
    cudaStream_t s = (cudaStream_t)stream;
    // bf16 and f16 are both moved as raw 16-bit words (pure select).
    feature_blend_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(p);
    return check_launch("mvoc_feature_blend");
}
