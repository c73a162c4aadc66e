// Standalone checker / timer for the tcgen05 GEMM family of libmvoc_b200.so (mvoc_conv3x3_nhwc, mvoc_temporal_conv3,
// mvoc_linear, mvoc_linear_geglu).  No Python, no torch: starts in milliseconds on a fresh GPU box.  Every case is
// compared with a naive fp32 GPU reference written here.
//
//   tools/gemm_check <section> <variant>      section: linear | conv | tconv | geglu | time
//                                             variant: 0 = one CTA per tile, 1 = CTA pairs (cta_group::2)
// One (section, variant) per process: a trap in one kernel must not take the other sections down with it.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../include/mvoc_b200.h"

#define CK(x)                                                                               \
    do {                                                                                    \
        cudaError_t e_ = (x);                                                               \
        if (e_ != cudaSuccess) {                                                            \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(2);                                                                        \
        }                                                                                   \
    } while (0)

typedef __nv_bfloat16 bf16;

static uint64_t g_seed = 0x9E3779B97F4A7C15ull;
static float frand() {
    g_seed ^= g_seed << 13;
    g_seed ^= g_seed >> 7;
    g_seed ^= g_seed << 17;
    return (float)((g_seed >> 11) & 0xFFFFFF) / 8388608.0f - 1.0f;
}
__global__ void fill_kernel(bf16* p, size_t n, float scale, uint64_t seed) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint64_t z = (i + seed) * 0x9E3779B97F4A7C15ull;
        z ^= z >> 31;
        z *= 0xBF58476D1CE4E5B9ull;
        z ^= z >> 29;
        p[i] = __float2bfloat16(((float)(z & 0xFFFFFF) / 8388608.0f - 1.0f) * scale);
    }
}
static bf16* dev_random(size_t n, float scale) {
    bf16* d;
    CK(cudaMalloc(&d, n * sizeof(bf16)));
    fill_kernel<<<1024, 256>>>(d, n, scale, (uint64_t)(frand() * 1e6f) + 12345);
    CK(cudaDeviceSynchronize());
    return d;
}
template <typename T> static T* dev_alloc(size_t n) {
    T* d;
    CK(cudaMalloc(&d, n * sizeof(T)));
    CK(cudaMemset(d, 0xFF, n * sizeof(T)));   // NaN pattern: unwritten outputs are caught
    return d;
}
__global__ void err_kernel(const bf16* y, const float* ref, size_t n, double* acc) {
    double num = 0.0, den = 0.0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double d = (double)__bfloat162float(y[i]) - (double)ref[i];
        num += d * d;
        den += (double)ref[i] * (double)ref[i];
    }
    atomicAdd(&acc[0], num);
    atomicAdd(&acc[1], den);
}
static double rel_l2(const bf16* y, const float* ref, size_t n) {
    double* acc;
    CK(cudaMalloc(&acc, 2 * sizeof(double)));
    CK(cudaMemset(acc, 0, 2 * sizeof(double)));
    err_kernel<<<256, 256>>>(y, ref, n, acc);
    double h[2];
    CK(cudaMemcpy(h, acc, sizeof(h), cudaMemcpyDeviceToHost));
    CK(cudaFree(acc));
    return sqrt(h[0] / (h[1] > 0 ? h[1] : 1.0));
}

// ------------------------------------------------------------------ naive references
// generic: out[row, c] = bias + sum_taps sum_k x[shifted row, k] * w[tap, c, k] (+ x2 . w2) (+ res)
__global__ void conv_ref(const bf16* x, const bf16* wt, const bf16* bias, const bf16* res, const bf16* x2, const bf16* w2,
                         int ci2, float* out, int N, int H, int W, int ci, int co) {
    const size_t total = (size_t)N * H * W * co;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % co);
        size_t p = i / co;
        const size_t pix = p;
        const int w = (int)(p % W);
        p /= W;
        const int h = (int)(p % H);
        const int n = (int)(p / H);
        float acc = bias ? __bfloat162float(bias[c]) : 0.0f;
        for (int kh = 0; kh < 3; ++kh)
            for (int kw = 0; kw < 3; ++kw) {
                const int hh = h + kh - 1, ww = w + kw - 1;
                if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
                const bf16* xr = x + (((size_t)n * H + hh) * W + ww) * ci;
                const bf16* wr = wt + ((size_t)(kh * 3 + kw) * co + c) * ci;
                for (int k = 0; k < ci; ++k) acc += __bfloat162float(xr[k]) * __bfloat162float(wr[k]);
            }
        if (x2)
            for (int k = 0; k < ci2; ++k) acc += __bfloat162float(x2[pix * ci2 + k]) * __bfloat162float(w2[(size_t)c * ci2 + k]);
        if (res) acc += __bfloat162float(res[i]);
        out[i] = acc;
    }
}
__global__ void tconv_ref(const bf16* x, const bf16* wt, const bf16* bias, const bf16* res, float* out, int B, int T,
                          int64_t S, int ci, int co) {
    const size_t total = (size_t)B * T * S * co;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % co);
        size_t p = i / co;
        const int64_t s = (int64_t)(p % S);
        p /= S;
        const int t = (int)(p % T);
        const int b = (int)(p / T);
        float acc = bias ? __bfloat162float(bias[c]) : 0.0f;
        for (int tap = 0; tap < 3; ++tap) {
            const int tt = t + tap - 1;
            if (tt < 0 || tt >= T) continue;
            const bf16* xr = x + (((size_t)b * T + tt) * S + s) * ci;
            const bf16* wr = wt + ((size_t)tap * co + c) * ci;
            for (int k = 0; k < ci; ++k) acc += __bfloat162float(xr[k]) * __bfloat162float(wr[k]);
        }
        if (res) acc += __bfloat162float(res[i]);
        out[i] = acc;
    }
}
__global__ void linear_ref(const bf16* x, const bf16* w, const bf16* bias, const bf16* res, float* out, int64_t M, int K,
                           int N, int64_t ldx, int64_t ldr) {
    const size_t total = (size_t)M * N;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int n = (int)(i % N);
        const size_t m = i / N;
        float acc = bias ? __bfloat162float(bias[n]) : 0.0f;
        for (int k = 0; k < K; ++k) acc += __bfloat162float(x[m * ldx + k]) * __bfloat162float(w[(size_t)n * K + k]);
        if (res) acc += __bfloat162float(res[m * ldr + n]);
        out[i] = acc;
    }
}
__global__ void geglu_ref(const bf16* x, const bf16* w, const bf16* bias, float* out, int64_t M, int K, int F) {
    const size_t total = (size_t)M * F;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(i % F);
        const size_t m = i / F;
        float v = bias ? __bfloat162float(bias[j]) : 0.0f, g = bias ? __bfloat162float(bias[F + j]) : 0.0f;
        for (int k = 0; k < K; ++k) {
            const float xv = __bfloat162float(x[m * K + k]);
            v += xv * __bfloat162float(w[(size_t)j * K + k]);
            g += xv * __bfloat162float(w[(size_t)(F + j) * K + k]);
        }
        out[i] = v * 0.5f * g * (1.0f + erff(g * 0.70710678118654752f));
    }
}
// strided bf16 output [M, ldo] -> compact copy of the first N columns (for the error kernel)
__global__ void compact_kernel(const bf16* y, bf16* c, int64_t M, int N, int64_t ldo) {
    const size_t total = (size_t)M * N;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
        c[i] = y[(i / N) * ldo + (i % N)];
}

static int report(const char* what, int rc, double err, double bar) {
    if (rc != 0) {
        printf("FAIL %-66s rc=%d: %s\n", what, rc, mvoc_last_error());
        fflush(stdout);
        return 1;
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("FAIL %-66s kernel error: %s\n", what, cudaGetErrorString(e));
        fflush(stdout);
        exit(3);   // the context is gone after a trap
    }
    const bool ok = err <= bar;
    printf("%s %-66s rel L2 %.3e (bar %.0e)\n", ok ? "ok  " : "FAIL", what, err, bar);
    fflush(stdout);
    return ok ? 0 : 1;
}
static double checked_err(int rc, const bf16* y, const float* ref, size_t n) {
    if (rc != 0) return NAN;
    if (cudaDeviceSynchronize() != cudaSuccess) return NAN;
    return rel_l2(y, ref, n);
}

static bf16* g_flush = nullptr;
static void flush_l2() {   // 256 MB write: larger than the 126 MB L2
    if (!g_flush) CK(cudaMalloc(&g_flush, (size_t)256 << 20));
    CK(cudaMemsetAsync(g_flush, 0, (size_t)256 << 20));
}

// ------------------------------------------------------------------ sections
static int section_linear(int variant) {
    // M, K, N, ldx pad, ldo pad, BN override
    const int64_t cases[][6] = {{128, 64, 64, 0, 0, 0},     {256, 64, 64, 0, 0, 0},      {1000, 320, 320, 0, 0, 0},
                                {4096, 320, 960, 0, 0, 0},  {777, 640, 640, 0, 0, 0},    {2048, 1280, 1280, 0, 0, 0},
                                {300, 128, 192, 64, 128, 0}, {5000, 320, 1280, 0, 0, 128}, {11600, 1024, 640, 0, 0, 0},
                                {40960, 320, 320, 0, 0, 0}, {20000, 1280, 320, 0, 0, 64}, {640, 2560, 1280, 0, 0, 0},
                                // a narrower tail tile after the full-width ones: 256+64, 2x256+128, 3x256+192, 192+128, 2x192+64
                                {3000, 320, 320, 0, 0, 256}, {2500, 640, 640, 0, 0, 256}, {1500, 320, 960, 0, 0, 256},
                                {1300, 256, 320, 0, 64, 192}, {900, 128, 448, 64, 0, 192}, {70000, 320, 640, 0, 0, 256}};
    int bad = 0;
    for (auto& c : cases)
        for (int mode = 0; mode < 3; ++mode) {   // 0: plain, 1: + bias, 2: + bias + residual
            const int64_t M = c[0];
            const int K = (int)c[1], N = (int)c[2];
            const int64_t ldx = K + c[3], ldo = N + c[4], ldr = N + (c[4] ? 64 : 0);
            bf16* x = dev_random((size_t)M * ldx, 1.0f);
            bf16* w = dev_random((size_t)N * K, 1.0f / sqrtf((float)K));
            bf16* bias = mode >= 1 ? dev_random(N, 1.0f) : nullptr;
            bf16* res = mode >= 2 ? dev_random((size_t)M * ldr, 1.0f) : nullptr;
            bf16* y = dev_alloc<bf16>((size_t)M * ldo);
            bf16* yc = dev_alloc<bf16>((size_t)M * N);
            float* ref = dev_alloc<float>((size_t)M * N);
            linear_ref<<<1024, 256>>>(x, w, bias, res, ref, M, K, N, ldx, ldr);
            CK(cudaDeviceSynchronize());
            const int rc = mvoc_linear(x, w, bias, res, y, M, K, N, ldx, ldr, ldo, MVOC_BF16, variant | ((int)c[5] << 8),
                                       nullptr);
            if (rc == 0 && cudaDeviceSynchronize() == cudaSuccess) compact_kernel<<<1024, 256>>>(y, yc, M, N, ldo);
            char name[160];
            snprintf(name, sizeof(name), "linear M=%lld K=%d N=%d ldx=%lld ldo=%lld bn=%d mode %d v%d", (long long)M, K, N,
                     (long long)ldx, (long long)ldo, (int)c[5], mode, variant);
            bad += report(name, rc, checked_err(rc, yc, ref, (size_t)M * N), 5e-3);
            cudaFree(x), cudaFree(w), cudaFree(bias), cudaFree(res), cudaFree(y), cudaFree(yc), cudaFree(ref);
        }
    return bad;
}

static int section_conv(int variant) {
    // N, H, W, Cin, Cout, Cin2 (1x1 shortcut source, 0 = none)
    const int cases[][6] = {{2, 64, 64, 64, 64, 0},    {4, 32, 32, 128, 128, 0},  {16, 8, 8, 128, 192, 0},
                            {3, 16, 16, 64, 320, 0},   {2, 11, 20, 64, 64, 0},    {5, 64, 64, 320, 320, 0},
                            {5, 8, 8, 64, 320, 128},   {5, 16, 16, 128, 640, 64}, {3, 22, 40, 64, 128, 0},
                            {5, 32, 32, 640, 640, 0},  {1, 64, 64, 960, 320, 960}};
    int bad = 0;
    for (auto& c : cases)
        for (int with_res = 0; with_res < 2; ++with_res) {
            const int N = c[0], H = c[1], W = c[2], ci = c[3], co = c[4], ci2 = c[5];
            const size_t px = (size_t)N * H * W;
            bf16* x = dev_random(px * ci, 1.0f);
            bf16* wt = dev_random((size_t)9 * co * ci, 1.0f / sqrtf(9.0f * ci));
            bf16* bias = dev_random(co, 1.0f);
            bf16* res = with_res ? dev_random(px * co, 1.0f) : nullptr;
            bf16* x2 = ci2 ? dev_random(px * ci2, 1.0f) : nullptr;
            bf16* w2 = ci2 ? dev_random((size_t)co * ci2, 1.0f / sqrtf((float)ci2)) : nullptr;
            bf16* y = dev_alloc<bf16>(px * co);
            float* ref = dev_alloc<float>(px * co);
            conv_ref<<<2048, 256>>>(x, wt, bias, res, x2, w2, ci2, ref, N, H, W, ci, co);
            CK(cudaDeviceSynchronize());
            const int rc = mvoc_conv3x3_nhwc(x, wt, bias, res, x2, w2, ci2, y, N, H, W, ci, co, MVOC_BF16, variant, nullptr);
            char name[160];
            snprintf(name, sizeof(name), "conv %dx%dx%d %d->%d%s%s v%d", N, H, W, ci, co, ci2 ? " +1x1 shortcut" : "",
                     with_res ? " +residual" : "", variant);
            bad += report(name, rc, checked_err(rc, y, ref, px * co), 5e-3);
            cudaFree(x), cudaFree(wt), cudaFree(bias), cudaFree(res), cudaFree(x2), cudaFree(w2), cudaFree(y), cudaFree(ref);
        }
    return bad;
}

static int section_tconv(int variant) {
    // B, T, S, Cin, Cout
    const int64_t cases[][5] = {{2, 16, 256, 64, 64},  {5, 16, 64, 128, 128}, {5, 16, 4096, 320, 320}, {3, 8, 100, 64, 192},
                                {5, 16, 32, 128, 128}, {5, 32, 8, 64, 64},    {1, 1, 256, 64, 64},     {5, 16, 1024, 640, 640}};
    int bad = 0;
    for (auto& c : cases)
        for (int with_res = 0; with_res < 2; ++with_res) {
            const int B = (int)c[0], T = (int)c[1], ci = (int)c[3], co = (int)c[4];
            const int64_t S = c[2];
            const size_t rows = (size_t)B * T * S;
            bf16* x = dev_random(rows * ci, 1.0f);
            bf16* wt = dev_random((size_t)3 * co * ci, 1.0f / sqrtf(3.0f * ci));
            bf16* bias = dev_random(co, 1.0f);
            bf16* res = with_res ? dev_random(rows * co, 1.0f) : nullptr;
            bf16* y = dev_alloc<bf16>(rows * co);
            float* ref = dev_alloc<float>(rows * co);
            tconv_ref<<<2048, 256>>>(x, wt, bias, res, ref, B, T, S, ci, co);
            CK(cudaDeviceSynchronize());
            const int rc = mvoc_temporal_conv3(x, wt, bias, res, y, B, T, S, ci, co, MVOC_BF16, variant, nullptr);
            char name[160];
            snprintf(name, sizeof(name), "tconv B=%d T=%d S=%lld %d->%d%s v%d", B, T, (long long)S, ci, co,
                     with_res ? " +residual" : "", variant);
            bad += report(name, rc, checked_err(rc, y, ref, rows * co), 5e-3);
            cudaFree(x), cudaFree(wt), cudaFree(bias), cudaFree(res), cudaFree(y), cudaFree(ref);
        }
    return bad;
}

static int section_geglu(int variant) {
    const int64_t cases[][3] = {{256, 64, 64}, {1000, 320, 1280}, {4096, 640, 2560}, {300, 128, 192}, {128, 1280, 5120}};
    int bad = 0;
    for (auto& c : cases) {
        const int64_t M = c[0];
        const int K = (int)c[1], F = (int)c[2];
        bf16* x = dev_random((size_t)M * K, 1.0f);
        bf16* w = dev_random((size_t)2 * F * K, 1.0f / sqrtf((float)K));
        bf16* bias = dev_random((size_t)2 * F, 1.0f);
        bf16* y = dev_alloc<bf16>((size_t)M * F);
        float* ref = dev_alloc<float>((size_t)M * F);
        geglu_ref<<<1024, 256>>>(x, w, bias, ref, M, K, F);
        CK(cudaDeviceSynchronize());
        const int rc = mvoc_linear_geglu(x, w, bias, y, M, K, F, MVOC_BF16, variant, nullptr);
        char name[160];
        snprintf(name, sizeof(name), "linear_geglu M=%lld K=%d F=%d v%d", (long long)M, K, F, variant);
        bad += report(name, rc, checked_err(rc, y, ref, (size_t)M * F), 5e-3);
        cudaFree(x), cudaFree(w), cudaFree(bias), cudaFree(y), cudaFree(ref);
    }
    return bad;
}

template <typename Fn> static float time_ms(Fn fn, int iters) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    for (int i = 0; i < 2; ++i) fn();
    CK(cudaDeviceSynchronize());
    float total = 0.0f;
    for (int i = 0; i < iters; ++i) {
        flush_l2();
        CK(cudaEventRecord(a));
        fn();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        total += ms;
    }
    return total / iters;
}

// shared-memory-port experiment: the same convolutions with the weight and / or activation stages never reloaded
// (variant bits 20 / 21; results are garbage, only the time matters).  If the MMA rate rises when the TMA fill
// traffic disappears, the SM's shared-memory port (TMA writes + MMA operand reads) is what bounds the kernel.
static int section_port(int variant) {
    const int convs[][5] = {{80, 64, 64, 960, 320}, {80, 32, 32, 1280, 640}, {80, 16, 16, 1280, 1280}};
    for (auto& c : convs) {
        const int N = c[0], Hh = c[1], W = c[2], ci = c[3], co = c[4];
        const size_t px = (size_t)N * Hh * W;
        bf16 *x = dev_random(px * ci, 1.0f), *wt = dev_random((size_t)9 * co * ci, 0.02f), *bias = dev_random(co, 1.0f),
             *y = dev_alloc<bf16>(px * co);
        const double fl = 2.0 * px * co * (double)ci * 9;
        for (int dbg = 0; dbg < 4; ++dbg) {
            const int v = variant | (dbg << 20);
            const float t = time_ms([&] { mvoc_conv3x3_nhwc(x, wt, bias, nullptr, nullptr, nullptr, 0, y, N, Hh, W, ci, co, MVOC_BF16, v, nullptr); }, 5);
            printf("conv %dx%dx%d %4d->%4d v%d reload %s%s: %.3f ms  %.0f TF/s\n", N, Hh, W, ci, co, variant,
                   (dbg & 2) ? "-" : "A", (dbg & 1) ? "-" : "W", t, fl / t / 1e9);
            fflush(stdout);
        }
        cudaFree(x), cudaFree(wt), cudaFree(bias), cudaFree(y);
    }
    return 0;
}

// tile-width experiment: ONE problem whose Cout every width divides, each width forced in turn, with and without
// the TMA refills (variant bits 20 / 21) — separates the tensor pipe's rate at N = BN from the fill traffic.
static int section_width(int variant) {
    const int N = 80, Hh = 32, W = 32, ci = 640, co = 3840;
    const size_t px = (size_t)N * Hh * W;
    bf16 *x = dev_random(px * ci, 1.0f), *wt = dev_random((size_t)9 * co * ci, 0.02f), *bias = dev_random(co, 1.0f),
         *y = dev_alloc<bf16>(px * co);
    const double fl = 2.0 * px * co * (double)ci * 9;
    const int widths[] = {256, 192, 160, 128, 64};
    for (int bn : widths)
        for (int dbg = 0; dbg < 4; dbg += 3) {
            const int v = variant | (bn << 8) | (dbg << 20);
            const float t = time_ms([&] { mvoc_conv3x3_nhwc(x, wt, bias, nullptr, nullptr, nullptr, 0, y, N, Hh, W, ci, co, MVOC_BF16, v, nullptr); }, 3);
            printf("conv %dx%dx%d %4d->%4d v%d bn=%3d %s: %.3f ms  %.0f TF/s\n", N, Hh, W, ci, co, variant, bn,
                   dbg ? "no refills" : "normal    ", t, fl / t / 1e9);
            fflush(stdout);
        }
    cudaFree(x), cudaFree(wt), cudaFree(bias), cudaFree(y);
    return 0;
}

// short-K Linears (HBM / epilogue bound): the same launch with parts of the kernel switched off (variant bits 20..23)
static int section_short(int variant) {
    // the last four are the per-rank shapes of an 8-GPU run (latency floor of a launch: one tile per CTA or fewer)
    const int64_t lins[][4] = {{327680, 320, 320, 0}, {327680, 320, 320, 1}, {327680, 320, 960, 0}, {81920, 640, 640, 1},
                               {2560, 1280, 1280, 0}, {2560, 1280, 1280, 1}, {640, 1280, 1280, 0}, {256, 64, 64, 0}};
    const int dbgs[] = {0, 3, 4, 7, 8, 11};
    for (auto& c : lins) {
        const int64_t M = c[0];
        const int K = (int)c[1], N = (int)c[2];
        bf16 *x = dev_random((size_t)M * K, 1.0f), *w = dev_random((size_t)N * K, 0.05f), *bias = dev_random(N, 1.0f),
             *res = c[3] ? dev_random((size_t)M * N, 1.0f) : nullptr, *y = dev_alloc<bf16>((size_t)M * N);
        const double bytes = ((double)M * K + (double)M * N * (c[3] ? 2 : 1) + (double)N * K) * 2;
        for (int dbg : dbgs) {
            if ((dbg & 8) && c[3]) continue;
            const int v = variant | (dbg << 20);
            const float t = time_ms([&] { mvoc_linear(x, w, bias, res, y, M, K, N, K, N, N, MVOC_BF16, v, nullptr); }, 5);
            printf("linear M=%lld K=%d N=%d%s v%d %s%s%s: %.3f ms  %.0f GB/s of algorithmic bytes\n", (long long)M, K, N,
                   c[3] ? " +res" : "", variant, (dbg & 3) == 3 ? "no-refill " : "", (dbg & 4) ? "no-store " : "",
                   (dbg & 8) ? "no-epilogue " : "", t, bytes / t / 1e6);
            fflush(stdout);
        }
        cudaFree(x), cudaFree(w), cudaFree(bias), cudaFree(res), cudaFree(y);
    }
    return 0;
}

static int section_time(int variant) {
    // the UNet's shapes at config 2 (80 frames): l0 64x64 C320, l1 32x32 C640, l2 16x16 C1280, l3 8x8 C1280
    const int convs[][5] = {{80, 64, 64, 320, 320},   {80, 64, 64, 640, 320},   {80, 64, 64, 960, 320},  {80, 32, 32, 640, 640},
                            {80, 32, 32, 1280, 640},  {80, 32, 32, 1920, 640},  {80, 16, 16, 1280, 1280}, {80, 16, 16, 2560, 1280},
                            {80, 8, 8, 1280, 1280},   {80, 8, 8, 2560, 1280}};
    for (auto& c : convs) {
        const int N = c[0], Hh = c[1], W = c[2], ci = c[3], co = c[4];
        const size_t px = (size_t)N * Hh * W;
        bf16 *x = dev_random(px * ci, 1.0f), *wt = dev_random((size_t)9 * co * ci, 0.02f), *bias = dev_random(co, 1.0f),
             *y = dev_alloc<bf16>(px * co);
        const double fl = 2.0 * px * co * (double)ci * 9;
        // auto, auto restricted to widths that divide Cout (variant bit 1), then forced widths (a narrower tail tile
        // covers the remainder where the width does not divide Cout)
        const int bns[] = {0, -1, 256, 192, 128};
        for (int bn : bns) {
            const int v = bn < 0 ? (variant | 2) : (variant | (bn << 8));
            const int rc = mvoc_conv3x3_nhwc(x, wt, bias, nullptr, nullptr, nullptr, 0, y, N, Hh, W, ci, co, MVOC_BF16, v, nullptr);
            if (rc) continue;   // this width cannot tile Cout
            const float t = time_ms([&] { mvoc_conv3x3_nhwc(x, wt, bias, nullptr, nullptr, nullptr, 0, y, N, Hh, W, ci, co, MVOC_BF16, v, nullptr); }, 5);
            printf("conv %dx%dx%d %4d->%4d bn=%3d%s v%d : %.3f ms  %.0f TF/s\n", N, Hh, W, ci, co, bn < 0 ? 0 : bn,
                   bn < 0 ? " exact widths" : "", variant, t, fl / t / 1e9);
            fflush(stdout);
        }
        cudaFree(x), cudaFree(wt), cudaFree(bias), cudaFree(y);
    }
    // Linears: M, K, N, residual
    const int64_t lins[][4] = {{327680, 320, 960, 0},   {327680, 320, 320, 1},  {327680, 1280, 320, 1}, {81920, 640, 1920, 0},
                               {81920, 640, 640, 1},    {81920, 2560, 640, 1},  {20480, 1280, 3840, 0}, {20480, 1280, 1280, 1},
                               {20480, 5120, 1280, 1},  {11600, 1024, 640, 0},  {5120, 1280, 1280, 1}};
    for (auto& c : lins) {
        const int64_t M = c[0];
        const int K = (int)c[1], N = (int)c[2];
        bf16 *x = dev_random((size_t)M * K, 1.0f), *w = dev_random((size_t)N * K, 0.05f), *bias = dev_random(N, 1.0f),
             *res = c[3] ? dev_random((size_t)M * N, 1.0f) : nullptr, *y = dev_alloc<bf16>((size_t)M * N);
        const double fl = 2.0 * M * (double)K * N;
        const double bytes = ((double)M * K + (double)M * N * (c[3] ? 2 : 1) + (double)N * K) * 2;
        const int bns[] = {0, -1, 256, 192};
        for (int bn : bns) {
            const int v = bn < 0 ? (variant | 2) : (variant | (bn << 8));
            const int rc = mvoc_linear(x, w, bias, res, y, M, K, N, K, N, N, MVOC_BF16, v, nullptr);
            if (rc) continue;
            const float t = time_ms([&] { mvoc_linear(x, w, bias, res, y, M, K, N, K, N, N, MVOC_BF16, v, nullptr); }, 5);
            printf("linear M=%lld K=%d N=%d%s bn=%3d%s v%d : %.3f ms  %.0f TF/s  %.0f GB/s\n", (long long)M, K, N,
                   c[3] ? " +res" : "", bn < 0 ? 0 : bn, bn < 0 ? " exact widths" : "", variant, t, fl / t / 1e9, bytes / t / 1e6);
            fflush(stdout);
        }
        cudaFree(x), cudaFree(w), cudaFree(bias), cudaFree(res), cudaFree(y);
    }
    const int64_t ffs[][2] = {{327680, 320}, {81920, 640}, {20480, 1280}};
    for (auto& c : ffs) {
        const int64_t M = c[0];
        const int K = (int)c[1], F = 4 * K;
        bf16 *x = dev_random((size_t)M * K, 1.0f), *w = dev_random((size_t)2 * F * K, 0.05f), *bias = dev_random(2 * F, 1.0f),
             *y = dev_alloc<bf16>((size_t)M * F);
        const float t = time_ms([&] { mvoc_linear_geglu(x, w, bias, y, M, K, F, MVOC_BF16, variant, nullptr); }, 5);
        printf("linear_geglu M=%lld K=%d F=%d v%d : %.3f ms  %.0f TF/s, %.0f GB/s of x+out\n", (long long)M, K, F, variant, t,
               2.0 * M * K * 2.0 * F / t / 1e9, ((double)M * K + (double)M * F) * 2 / t / 1e6);
        fflush(stdout);
        cudaFree(x), cudaFree(w), cudaFree(bias), cudaFree(y);
    }
    const int64_t tcs[][4] = {{5, 16, 4096, 320}, {5, 16, 1024, 640}, {5, 16, 256, 1280}, {5, 16, 64, 1280}};
    for (auto& c : tcs) {
        const int B = (int)c[0], T = (int)c[1], C = (int)c[3];
        const int64_t S = c[2];
        const size_t rows = (size_t)B * T * S;
        bf16 *x = dev_random(rows * C, 1.0f), *wt = dev_random((size_t)3 * C * C, 0.02f), *bias = dev_random(C, 1.0f),
             *y = dev_alloc<bf16>(rows * C);
        for (int notail = 0; notail < 2; ++notail) {
            const int v = variant | (notail << 1);
            const float t = time_ms([&] { mvoc_temporal_conv3(x, wt, bias, nullptr, y, B, T, S, C, C, MVOC_BF16, v, nullptr); }, 5);
            printf("tconv B=%d T=%d S=%lld C=%d v%d %s: %.3f ms  %.0f TF/s\n", B, T, (long long)S, C, variant,
                   notail ? "exact widths" : "auto        ", t, 2.0 * rows * C * (double)C * 3 / t / 1e9);
            fflush(stdout);
        }
        cudaFree(x), cudaFree(wt), cudaFree(bias), cudaFree(y);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("time section: %s\n", e == cudaSuccess ? "done" : cudaGetErrorString(e));
    return e == cudaSuccess ? 0 : 1;
}

// one <linear|conv> <variant> ...: a single shape launched a few times (the target of an ncu capture)
//   one linear <variant> M K N res        |   one conv <variant> N H W Cin Cout
static int section_one(int argc, char** argv) {
    if (argc < 4) return 1;
    const int variant = atoi(argv[3]);
    if (!strcmp(argv[2], "linear") && argc >= 8) {
        const int64_t M = atoll(argv[4]);
        const int K = atoi(argv[5]), N = atoi(argv[6]), res = atoi(argv[7]);
        bf16 *x = dev_random((size_t)M * K, 1.0f), *w = dev_random((size_t)N * K, 0.05f), *bias = dev_random(N, 1.0f),
             *r = res ? dev_random((size_t)M * N, 1.0f) : nullptr, *y = dev_alloc<bf16>((size_t)M * N);
        for (int i = 0; i < 4; ++i) {
            flush_l2();
            const int rc = mvoc_linear(x, w, bias, r, y, M, K, N, K, N, N, MVOC_BF16, variant, nullptr);
            if (rc) printf("rc=%d %s\n", rc, mvoc_last_error());
        }
    } else if (!strcmp(argv[2], "conv") && argc >= 9) {
        const int N = atoi(argv[4]), H = atoi(argv[5]), W = atoi(argv[6]), ci = atoi(argv[7]), co = atoi(argv[8]);
        const size_t px = (size_t)N * H * W;
        bf16 *x = dev_random(px * ci, 1.0f), *wt = dev_random((size_t)9 * co * ci, 0.02f), *bias = dev_random(co, 1.0f),
             *y = dev_alloc<bf16>(px * co);
        for (int i = 0; i < 4; ++i) {
            flush_l2();
            const int rc = mvoc_conv3x3_nhwc(x, wt, bias, nullptr, nullptr, nullptr, 0, y, N, H, W, ci, co, MVOC_BF16, variant, nullptr);
            if (rc) printf("rc=%d %s\n", rc, mvoc_last_error());
        }
    } else {
        return 1;
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("one: %s\n", cudaGetErrorString(e));
    return e == cudaSuccess ? 0 : 1;
}

int main(int argc, char** argv) {
    const char* what = argc > 1 ? argv[1] : "linear";
    if (!strcmp(what, "one")) return section_one(argc, argv);
    const int variant = argc > 2 ? atoi(argv[2]) : 0;
    int rc = mvoc_device_check(0);
    if (rc != 0) {
        printf("device check failed: %s\n", mvoc_last_error());
        return 2;
    }
    int bad;
    if (!strcmp(what, "linear")) bad = section_linear(variant);
    else if (!strcmp(what, "conv")) bad = section_conv(variant);
    else if (!strcmp(what, "tconv")) bad = section_tconv(variant);
    else if (!strcmp(what, "geglu")) bad = section_geglu(variant);
    else if (!strcmp(what, "time")) bad = section_time(variant);
    else if (!strcmp(what, "port")) bad = section_port(variant);
    else if (!strcmp(what, "width")) bad = section_width(variant);
    else if (!strcmp(what, "short")) bad = section_short(variant);
    else {
        printf("usage: gemm_check linear|conv|tconv|geglu|time [variant]\n");
        return 2;
    }
    printf("%s v%d: %d failing case(s)\n", what, variant, bad);
    return bad ? 1 : 0;
}
