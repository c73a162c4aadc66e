// Row-wise LayerNorm over [M, C] (C = 320 / 640 / 1280 on this path), HBM-bound.
// Reference: norm1 / norm2 / norm3 of BasicTransformerBlock (i2vgen-xl/pnp_utils.py:249-250, :295-296,
// :322; nn.LayerNorm(dim, eps=1e-5, affine)) — 3 per transformer block, 99 launches per UNet forward.
// One warp per row: the row is held in registers (<= 5 x 16 B per lane at C = 1280), mean and centred
// variance by warp shuffles (exact two-pass, no E[x^2]-mean^2 cancellation), one read + one write.
#include "common.cuh"

namespace mvoc {

template <typename T, int VPL>  // VPL = 16-byte vectors per lane (C <= 256 * VPL)
__global__ void __launch_bounds__(256) layernorm_kernel(const T* __restrict__ x, T* __restrict__ y,
                                                        const T* __restrict__ gamma,
                                                        const T* __restrict__ beta, int64_t M, int C,
                                                        float eps) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const int nvec = C >> 3;
    const T* xr = x + row * C;
    float v[VPL][8];
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int vi = lane + 32 * i;
        if (vi < nvec) {
            unpack8<T>(ld_stream16(xr + vi * 8), v[i]);
#pragma unroll
            for (int e = 0; e < 8; ++e) s += v[i][e];
        }
    }
    const float mean = warp_sum(s) / (float)C;
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        if (lane + 32 * i < nvec) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float d = v[i][e] - mean;
                q += d * d;
            }
        }
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
    T* yr = y + row * C;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int vi = lane + 32 * i;
        if (vi < nvec) {
            float g[8], b[8];
            unpack8<T>(ld_global16(gamma + vi * 8), g);
            unpack8<T>(ld_global16(beta + vi * 8), b);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[i][e] = (v[i][e] - mean) * rstd * g[e] + b[e];
            st_stream16(yr + vi * 8, pack8<T>(v[i]));
        }
    }
}

template <typename T>
static int ln_launch(const void* x, void* y, const void* g, const void* b, int64_t M, int C, float eps,
                     cudaStream_t st) {
    const int nvec = C / 8;
    const int vpl = (nvec + 31) / 32;
    const int rows_per_cta = 8;
    const unsigned grid = (unsigned)((M + rows_per_cta - 1) / rows_per_cta);
    const T* xp = (const T*)x;
    T* yp = (T*)y;
    const T* gp = (const T*)g;
    const T* bp = (const T*)b;
    switch (vpl) {
        case 1: layernorm_kernel<T, 1><<<grid, 256, 0, st>>>(xp, yp, gp, bp, M, C, eps); break;
        case 2: layernorm_kernel<T, 2><<<grid, 256, 0, st>>>(xp, yp, gp, bp, M, C, eps); break;
        case 3: layernorm_kernel<T, 3><<<grid, 256, 0, st>>>(xp, yp, gp, bp, M, C, eps); break;
        case 4: layernorm_kernel<T, 4><<<grid, 256, 0, st>>>(xp, yp, gp, bp, M, C, eps); break;
        case 5: layernorm_kernel<T, 5><<<grid, 256, 0, st>>>(xp, yp, gp, bp, M, C, eps); break;
        case 6: case 7: case 8:
            layernorm_kernel<T, 8><<<grid, 256, 0, st>>>(xp, yp, gp, bp, M, C, eps); break;
        default:
            set_error("mvoc_layernorm: C=%d too large (max 2048)", C);
            return MVOC_ERR_UNSUPPORTED;
    }
    return check_launch("mvoc_layernorm");
}

}  // namespace mvoc

using namespace mvoc;

extern "C" int mvoc_layernorm(const void* x, void* y, const void* gamma, const void* beta, int64_t M,
                              int C, float eps, int dtype, void* stream) {
    MVOC_REQUIRE(x && y && gamma && beta, MVOC_ERR_INVALID_ARG, "mvoc_layernorm: null pointer");
    MVOC_REQUIRE(M >= 0 && C > 0 && C % 8 == 0, MVOC_ERR_UNSUPPORTED,
                 "mvoc_layernorm: C=%d must be a positive multiple of 8", C);
    MVOC_REQUIRE(((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0) && ((uintptr_t)gamma % 16 == 0) &&
                     ((uintptr_t)beta % 16 == 0),
                 MVOC_ERR_INVALID_ARG, "mvoc_layernorm: pointers must be 16-byte aligned");
    MVOC_REQUIRE(M < ((int64_t)1 << 31) * 8, MVOC_ERR_UNSUPPORTED, "mvoc_layernorm: too many rows");
    if (M == 0) return MVOC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == MVOC_BF16) return ln_launch<__nv_bfloat16>(x, y, gamma, beta, M, C, eps, st);
    if (dtype == MVOC_F16) return ln_launch<__half>(x, y, gamma, beta, M, C, eps, st);
    set_error("mvoc_layernorm: dtype %d unsupported (bf16/f16 only)", dtype);
    return MVOC_ERR_UNSUPPORTED;
}
