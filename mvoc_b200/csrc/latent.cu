// Latent-space elementwise kernels of the step loops (latency/HBM-bound).
//   mvoc_latent_composite  — noise fusion + UNet-input concat
//                            (reference: pipelines/pipeline_i2vgen_xl.py:1644-1663, :1675-1677)
//   mvoc_cfg_ddim_step     — CFG combine + v-prediction DDIM update
//                            (reference: pipelines/pipeline_i2vgen_xl.py:1713-1731 + DDIMScheduler.step)
//   mvoc_ddim_inverse_step — inverse DDIM update
//                            (reference: pipelines/pipeline_i2vgen_xl.py:1967-1984 + DDIMInverseScheduler.step)
// Each replaces ~6-10 ATen launches with one; 8 elements per thread, 128-bit accesses.
#include "common.cuh"

namespace mvoc {

template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&f)[8]);
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&f)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    const float4 b = *reinterpret_cast<const float4*>(p + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
    f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&f)[8]) {
    unpack8<__nv_bfloat16>(ld_global16(p), f);
}
template <>
__device__ __forceinline__ void load8<__half>(const __half* p, float (&f)[8]) {
    unpack8<__half>(ld_global16(p), f);
}

template <typename T>
__device__ __forceinline__ void store8(T* p, const float (&f)[8]);
template <>
__device__ __forceinline__ void store8<float>(float* p, const float (&f)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}
template <>
__device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16* p, const float (&f)[8]) {
    *reinterpret_cast<Vec16*>(p) = pack8<__nv_bfloat16>(f);
}
template <>
__device__ __forceinline__ void store8<__half>(__half* p, const float (&f)[8]) {
    *reinterpret_cast<Vec16*>(p) = pack8<__half>(f);
}

template <typename LT, typename IT>
__global__ void __launch_bounds__(256) latent_composite_kernel(
    LT* __restrict__ z, const LT* __restrict__ bg, const LT* __restrict__ objs,
    const float* __restrict__ mask, IT* __restrict__ unet_in, int n_obj, int64_t E, int64_t THW,
    float r, int do_fusion, int obj_noise_fusion) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) << 3;
    if (i >= E) return;
    float zf[8], bf[8];
    load8<LT>(z + i, zf);
    load8<LT>(bg + i, bf);
    if (unet_in) store8<IT>(unet_in + i, bf);
    if (do_fusion) {
#pragma unroll
        for (int e = 0; e < 8; ++e) zf[e] = r * zf[e] + (1.0f - r) * bf[e];
    }
    const int64_t mi = i % THW;
    for (int j = 0; j < n_obj; ++j) {
        float of[8];
        load8<LT>(objs + (int64_t)j * E + i, of);
        if (unet_in) store8<IT>(unet_in + (int64_t)(j + 1) * E + i, of);
        if (do_fusion) {
            float mf[8];
            load8<float>(mask + (int64_t)j * THW + mi, mf);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float m = mf[e];
                const float back = zf[e] * (1.0f - m);
                // reference order: obj*M first, then the optional noise mix
                const float fg = obj_noise_fusion ? (zf[e] * m) * r + (1.0f - r) * (of[e] * m)
                                                  : of[e] * m;
                zf[e] = back + fg;
            }
        }
    }
    if (do_fusion) store8<LT>(z + i, zf);
    if (unet_in) {
        store8<IT>(unet_in + (int64_t)(n_obj + 1) * E + i, zf);
        store8<IT>(unet_in + (int64_t)(n_obj + 2) * E + i, zf);
    }
}

template <typename PT, typename LT>
__global__ void __launch_bounds__(256) ddim_step_kernel(
    const PT* __restrict__ pu, const PT* __restrict__ pc, LT* __restrict__ x, int64_t E, float g,
    float sa_from, float sb_from, float sa_to, float sb_to) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) << 3;
    if (i >= E) return;
    float u[8], xf[8];
    load8<PT>(pu + i, u);
    load8<LT>(x + i, xf);
    if (pc) {
        float c[8];
        load8<PT>(pc + i, c);
#pragma unroll
        for (int e = 0; e < 8; ++e) u[e] = u[e] + g * (c[e] - u[e]);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const float v = u[e];
        const float x0 = sa_from * xf[e] - sb_from * v;
        const float eps = sa_from * v + sb_from * xf[e];
        xf[e] = sa_to * x0 + sb_to * eps;
    }
    store8<LT>(x + i, xf);
}

template <typename PT, typename LT>
static int launch_ddim(const void* pu, const void* pc, void* x, int64_t E, float g, double a_from,
                       double a_to, cudaStream_t s) {
    const float sa_from = (float)sqrt(a_from), sb_from = (float)sqrt(1.0 - a_from);
    const float sa_to = (float)sqrt(a_to), sb_to = (float)sqrt(1.0 - a_to);
    const int64_t threads = E / 8;
    const int grid = (int)((threads + 255) / 256);
    ddim_step_kernel<PT, LT><<<grid, 256, 0, s>>>(
        reinterpret_cast<const PT*>(pu), reinterpret_cast<const PT*>(pc),
        reinterpret_cast<LT*>(x), E, g, sa_from, sb_from, sa_to, sb_to);
    return check_launch("mvoc_ddim_step");
}

static int ddim_dispatch(const char* name, const void* pu, const void* pc, void* x, int64_t E,
                         float g, double a_from, double a_to, int pred_dtype, int lat_dtype,
                         void* stream) {
    MVOC_REQUIRE(pu != nullptr && x != nullptr, MVOC_ERR_INVALID_ARG, "%s: null pointer", name);
    MVOC_REQUIRE(E > 0 && E % 8 == 0, MVOC_ERR_UNSUPPORTED, "%s: E=%lld must be a positive multiple of 8",
                 name, (long long)E);
    MVOC_REQUIRE(a_from >= 0.0 && a_from <= 1.0 && a_to >= 0.0 && a_to <= 1.0, MVOC_ERR_INVALID_ARG,
                 "%s: alphas must lie in [0,1] (got %g, %g)", name, a_from, a_to);
    MVOC_REQUIRE(((uintptr_t)pu % 16 == 0) && ((uintptr_t)pc % 16 == 0) && ((uintptr_t)x % 16 == 0),
                 MVOC_ERR_INVALID_ARG, "%s: pointers must be 16-byte aligned", name);
    cudaStream_t s = (cudaStream_t)stream;
#define MVOC_DDIM_CASE(PD, PT, LD, LT) \
    if (pred_dtype == PD && lat_dtype == LD) return launch_ddim<PT, LT>(pu, pc, x, E, g, a_from, a_to, s);
    MVOC_DDIM_CASE(MVOC_BF16, __nv_bfloat16, MVOC_F32, float)
    MVOC_DDIM_CASE(MVOC_BF16, __nv_bfloat16, MVOC_BF16, __nv_bfloat16)
    MVOC_DDIM_CASE(MVOC_F16, __half, MVOC_F32, float)
    MVOC_DDIM_CASE(MVOC_F16, __half, MVOC_F16, __half)
    MVOC_DDIM_CASE(MVOC_F32, float, MVOC_F32, float)
#undef MVOC_DDIM_CASE
    set_error("%s: unsupported dtype pair pred=%d lat=%d", name, pred_dtype, lat_dtype);
    return MVOC_ERR_UNSUPPORTED;
}

template <typename LT, typename IT>
static int launch_composite(void* z, const void* bg, const void* objs, const void* mask,
                            void* unet_in, int n_obj, int64_t E, int64_t THW, float r,
                            int do_fusion, int onf, cudaStream_t s) {
    const int64_t threads = E / 8;
    const int grid = (int)((threads + 255) / 256);
    latent_composite_kernel<LT, IT><<<grid, 256, 0, s>>>(
        reinterpret_cast<LT*>(z), reinterpret_cast<const LT*>(bg),
        reinterpret_cast<const LT*>(objs), reinterpret_cast<const float*>(mask),
        reinterpret_cast<IT*>(unet_in), n_obj, E, THW, r, do_fusion, onf);
    return check_launch("mvoc_latent_composite");
}

}  // namespace mvoc

using namespace mvoc;

extern "C" int mvoc_latent_composite(void* z, const void* bg, const void* objs, const void* mask,
                                     void* unet_in, int n_obj, int64_t E, int64_t THW, float ratio,
                                     int do_fusion, int obj_noise_fusion, int lat_dtype,
                                     int in_dtype, void* stream) {
    MVOC_REQUIRE(z != nullptr && bg != nullptr && objs != nullptr, MVOC_ERR_INVALID_ARG,
                 "mvoc_latent_composite: null pointer");
    MVOC_REQUIRE(!do_fusion || mask != nullptr, MVOC_ERR_INVALID_ARG,
                 "mvoc_latent_composite: fusion step needs masks");
    MVOC_REQUIRE(n_obj >= 1 && n_obj <= MVOC_MAX_OBJECTS, MVOC_ERR_INVALID_ARG,
                 "mvoc_latent_composite: n_obj=%d out of range [1,%d]", n_obj, MVOC_MAX_OBJECTS);
    MVOC_REQUIRE(E > 0 && THW > 0 && E % THW == 0 && THW % 8 == 0, MVOC_ERR_UNSUPPORTED,
                 "mvoc_latent_composite: need E%%THW==0 and THW%%8==0 (E=%lld THW=%lld)",
                 (long long)E, (long long)THW);
    MVOC_REQUIRE(((uintptr_t)z % 16 == 0) && ((uintptr_t)bg % 16 == 0) &&
                     ((uintptr_t)objs % 16 == 0) && ((uintptr_t)mask % 16 == 0) &&
                     ((uintptr_t)unet_in % 16 == 0),
                 MVOC_ERR_INVALID_ARG, "mvoc_latent_composite: pointers must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
#define MVOC_LC_CASE(LD, LT, ID, IT)                                                           \
    if (lat_dtype == LD && in_dtype == ID)                                                     \
        return launch_composite<LT, IT>(z, bg, objs, mask, unet_in, n_obj, E, THW, ratio,      \
                                        do_fusion, obj_noise_fusion, s);
    MVOC_LC_CASE(MVOC_F32, float, MVOC_BF16, __nv_bfloat16)
    MVOC_LC_CASE(MVOC_F32, float, MVOC_F16, __half)
    MVOC_LC_CASE(MVOC_F32, float, MVOC_F32, float)
    MVOC_LC_CASE(MVOC_BF16, __nv_bfloat16, MVOC_BF16, __nv_bfloat16)
    MVOC_LC_CASE(MVOC_F16, __half, MVOC_F16, __half)
#undef MVOC_LC_CASE
    set_error("mvoc_latent_composite: unsupported dtype pair lat=%d in=%d", lat_dtype, in_dtype);
    return MVOC_ERR_UNSUPPORTED;
}

extern "C" int mvoc_cfg_ddim_step(const void* pred_uncond, const void* pred_cond, void* x,
                                  int64_t E, float guidance, double alpha_t, double alpha_prev,
                                  int pred_dtype, int lat_dtype, void* stream) {
    return ddim_dispatch("mvoc_cfg_ddim_step", pred_uncond, pred_cond, x, E, guidance, alpha_t,
                         alpha_prev, pred_dtype, lat_dtype, stream);
}

extern "C" int mvoc_ddim_inverse_step(const void* pred_uncond, const void* pred_cond, void* x,
                                      int64_t E, float guidance, double alpha_src,
                                      double alpha_dst, int pred_dtype, int lat_dtype,
                                      void* stream) {
    return ddim_dispatch("mvoc_ddim_inverse_step", pred_uncond, pred_cond, x, E, guidance,
                         alpha_src, alpha_dst, pred_dtype, lat_dtype, stream);
}
