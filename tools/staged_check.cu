// Standalone checker for the kernels staged in libmvoc_b200_staged.so (no Python, no torch: starts in
// milliseconds on a fresh GPU box).  Every case compares against a naive fp32 GPU reference written here.
//
//   built by __graft_entry__.build() next to the libraries it links (tools/staged_check, rpath = ../mvoc_b200/lib)
//   tools/staged_check conv | geglu | attn | time        one section per process: a trap in one kernel must not
//                                                         take the other sections down with it
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../include/mvoc_b200.h"
#include "../include/mvoc_b200_staged.h"

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);     \
            exit(2);                                                                            \
        }                                                                                       \
    } while (0)

typedef __nv_bfloat16 bf16;

static uint64_t g_seed = 0x9E3779B97F4A7C15ull;
static float frand() {  // uniform in [-1, 1)
    g_seed ^= g_seed << 13;
    g_seed ^= g_seed >> 7;
    g_seed ^= g_seed << 17;
    return (float)((g_seed >> 11) & 0xFFFFFF) / 8388608.0f - 1.0f;
}
static bf16* dev_random(size_t n, float scale) {
    std::vector<bf16> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = __float2bfloat16(frand() * scale);
    bf16* d;
    CK(cudaMalloc(&d, n * sizeof(bf16)));
    CK(cudaMemcpy(d, h.data(), n * sizeof(bf16), cudaMemcpyHostToDevice));
    return d;
}
template <typename T> static T* dev_alloc(size_t n) {
    T* d;
    CK(cudaMalloc(&d, n * sizeof(T)));
    CK(cudaMemset(d, 0xFF, n * sizeof(T)));   // NaN pattern: unwritten outputs are caught
    return d;
}

// relative L2 error of a bf16 result against an fp32 reference, both on the device
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
    double* acc = dev_alloc<double>(2);
    CK(cudaMemset(acc, 0, 2 * sizeof(double)));
    err_kernel<<<256, 256>>>(y, ref, n, acc);
    double h[2];
    CK(cudaMemcpy(h, acc, sizeof(h), cudaMemcpyDeviceToHost));
    CK(cudaFree(acc));
    return sqrt(h[0] / (h[1] > 0 ? h[1] : 1.0));   // NaN in y propagates
}

// ------------------------------------------------------------------ naive references
__global__ void conv_ref(const bf16* x, const bf16* wt, const bf16* bias, const bf16* res, float* out, int N, int H,
                         int W, int ci, int co) {
    const size_t total = (size_t)N * H * W * co;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % co);
        size_t p = i / co;
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
        if (res) acc += __bfloat162float(res[i]);
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

// one thread per (b, h, query): two passes over the keys
__global__ void attn_ref(const bf16* q, const bf16* k, const bf16* v, float* out, int B, int H, int Nq, int Nk,
                         float scale) {
    const int C = H * 64;
    const size_t total = (size_t)B * H * Nq;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int n = (int)(i % Nq);
        const int h = (int)((i / Nq) % H);
        const int b = (int)(i / ((size_t)Nq * H));
        const bf16* qr = q + ((size_t)b * Nq + n) * C + h * 64;
        float qf[64];
        for (int d = 0; d < 64; ++d) qf[d] = __bfloat162float(qr[d]);
        float mx = -INFINITY;
        for (int j = 0; j < Nk; ++j) {
            const bf16* kr = k + ((size_t)b * Nk + j) * C + h * 64;
            float s = 0.0f;
            for (int d = 0; d < 64; ++d) s += qf[d] * __bfloat162float(kr[d]);
            mx = fmaxf(mx, s * scale);
        }
        float o[64], l = 0.0f;
        for (int d = 0; d < 64; ++d) o[d] = 0.0f;
        for (int j = 0; j < Nk; ++j) {
            const bf16* kr = k + ((size_t)b * Nk + j) * C + h * 64;
            const bf16* vr = v + ((size_t)b * Nk + j) * C + h * 64;
            float s = 0.0f;
            for (int d = 0; d < 64; ++d) s += qf[d] * __bfloat162float(kr[d]);
            const float p = expf(s * scale - mx);
            l += p;
            for (int d = 0; d < 64; ++d) o[d] += p * __bfloat162float(vr[d]);
        }
        float* orow = out + ((size_t)b * Nq + n) * C + h * 64;
        for (int d = 0; d < 64; ++d) orow[d] = o[d] / l;
    }
}

static int report(const char* what, int rc, double err, double bar) {
    if (rc != 0) {
        printf("FAIL %-58s rc=%d: %s\n", what, rc, mvoc_last_error());
        return 1;
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("FAIL %-58s kernel error: %s\n", what, cudaGetErrorString(e));
        exit(3);   // context is gone after a trap
    }
    const bool ok = err <= bar;   // false for NaN
    printf("%s %-58s rel L2 %.3e (bar %.0e)\n", ok ? "ok  " : "FAIL", what, err, bar);
    return ok ? 0 : 1;
}

// ------------------------------------------------------------------ sections
static int section_conv() {
    const int cases[][5] = {{2, 64, 64, 64, 64},   {4, 32, 32, 128, 128}, {16, 8, 8, 128, 160},
                            {3, 16, 16, 64, 320},  {2, 11, 20, 64, 64},   {5, 64, 64, 320, 320},
                            {5, 8, 8, 64, 320},    {5, 16, 16, 128, 640}};
    int bad = 0;
    for (auto& c : cases)
        for (int variant = 0; variant < 3; ++variant)
            for (int with_res = 0; with_res < 2; ++with_res) {
                const int N = c[0], H = c[1], W = c[2], ci = c[3], co = c[4];
                if (variant == 2 && co % 320 != 0) continue;   // CTA-pair variant: 320-column tiles only
                const size_t px = (size_t)N * H * W;
                bf16* x = dev_random(px * ci, 1.0f);
                bf16* wt = dev_random((size_t)9 * co * ci, 1.0f / sqrtf(9.0f * ci));
                bf16* bias = dev_random(co, 1.0f);
                bf16* res = with_res ? dev_random(px * co, 1.0f) : nullptr;
                bf16* y = dev_alloc<bf16>(px * co);
                float* ref = dev_alloc<float>(px * co);
                conv_ref<<<1024, 256>>>(x, wt, bias, res, ref, N, H, W, ci, co);
                CK(cudaDeviceSynchronize());
                const int rc = mvoc_conv3x3_nhwc(x, wt, bias, res, y, N, H, W, ci, co, MVOC_BF16, variant, nullptr);
                char name[128];
                snprintf(name, sizeof(name), "conv %dx%dx%d %d->%d variant %d%s", N, H, W, ci, co, variant,
                         with_res ? " +residual" : "");
                const double err = rc == 0 && cudaDeviceSynchronize() == cudaSuccess ? rel_l2(y, ref, px * co) : NAN;
                bad += report(name, rc, err, 5e-3);
                cudaFree(x), cudaFree(wt), cudaFree(bias), cudaFree(res), cudaFree(y), cudaFree(ref);
            }
    return bad;
}

static int section_geglu() {
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
        const int rc = mvoc_linear_geglu(x, w, bias, y, M, K, F, MVOC_BF16, nullptr);
        char name[128];
        snprintf(name, sizeof(name), "linear_geglu M=%lld K=%d F=%d", (long long)M, K, F);
        const double err = rc == 0 && cudaDeviceSynchronize() == cudaSuccess ? rel_l2(y, ref, (size_t)M * F) : NAN;
        bad += report(name, rc, err, 5e-3);
        cudaFree(x), cudaFree(w), cudaFree(bias), cudaFree(y), cudaFree(ref);
    }
    return bad;
}

static int section_attn() {
    const int cases[][4] = {{2, 5, 256, 256}, {1, 1, 128, 128}, {2, 2, 64, 64},   {2, 5, 256, 145},
                            {1, 3, 200, 77},  {1, 2, 880, 880}, {1, 2, 300, 200}, {1, 5, 4096, 4096}};
    int bad = 0;
    for (auto& c : cases)
        for (int variant = 0; variant < 3; ++variant) {
            const int B = c[0], H = c[1], Nq = c[2], Nk = c[3], C = H * 64;
            bf16* q = dev_random((size_t)B * Nq * C, 1.0f);
            bf16* k = dev_random((size_t)B * Nk * C, 1.0f);
            bf16* v = dev_random((size_t)B * Nk * C, 1.0f);
            bf16* y = dev_alloc<bf16>((size_t)B * Nq * C);
            float* ref = dev_alloc<float>((size_t)B * Nq * C);
            attn_ref<<<512, 128>>>(q, k, v, ref, B, H, Nq, Nk, 0.125f);
            CK(cudaDeviceSynchronize());
            const int rc = mvoc_attn_fwd_split(q, k, v, y, B, H, Nq, Nk, 64, (int64_t)Nq * C, C, 64, (int64_t)Nk * C, C, 64,
                                               (int64_t)Nk * C, C, 64, (int64_t)Nq * C, C, 64, 0.125f, MVOC_BF16, variant,
                                               nullptr);
            char name[128];
            snprintf(name, sizeof(name), "attn_split B=%d H=%d Nq=%d Nk=%d variant %d", B, H, Nq, Nk, variant);
            const double err = rc == 0 && cudaDeviceSynchronize() == cudaSuccess ? rel_l2(y, ref, (size_t)B * Nq * C) : NAN;
            bad += report(name, rc, err, 1e-2);
            cudaFree(q), cudaFree(k), cudaFree(v), cudaFree(y), cudaFree(ref);
        }
    return bad;
}

template <typename Fn> static float time_ms(Fn fn, int iters) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    for (int i = 0; i < 2; ++i) fn();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int i = 0; i < iters; ++i) fn();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    return ms / iters;
}

static int section_time() {
    {  // l0 spatial self-attention of config 2: product kernel vs row-split
        const int B = 80, H = 5, N = 4096, C = H * 64;
        bf16 *q = dev_random((size_t)B * N * C, 1.0f), *k = dev_random((size_t)B * N * C, 1.0f),
             *v = dev_random((size_t)B * N * C, 1.0f), *y = dev_alloc<bf16>((size_t)B * N * C);
        const double fl = 4.0 * B * H * (double)N * N * 64;
        const int64_t sb = (int64_t)N * C;
        float t = time_ms([&] { mvoc_attn_fwd(q, k, v, y, B, H, N, N, 64, sb, C, 64, sb, C, 64, sb, C, 64, sb, C, 64, 0.125f,
                                               MVOC_BF16, 0, nullptr); }, 5);
        printf("attention l0 product      : %.3f ms  %.0f TF/s\n", t, fl / t / 1e9);
        for (int variant = 0; variant < 3; ++variant) {
            t = time_ms([&] { mvoc_attn_fwd_split(q, k, v, y, B, H, N, N, 64, sb, C, 64, sb, C, 64, sb, C, 64, sb, C, 64,
                                                  0.125f, MVOC_BF16, variant, nullptr); }, 5);
            printf("attention l0 split v%d     : %.3f ms  %.0f TF/s\n", variant, t, fl / t / 1e9);
        }
        cudaFree(q), cudaFree(k), cudaFree(v), cudaFree(y);
    }
    const int convs[][5] = {{80, 64, 64, 320, 320}, {80, 64, 64, 960, 320}, {80, 32, 32, 640, 640}, {80, 32, 32, 1920, 640},
                            {80, 16, 16, 1280, 1280}, {80, 16, 16, 2560, 1280}, {80, 8, 8, 1280, 1280}};
    for (auto& c : convs) {
        const int N = c[0], Hh = c[1], W = c[2], ci = c[3], co = c[4];
        const size_t px = (size_t)N * Hh * W;
        bf16 *x = dev_random(px * ci, 1.0f), *wt = dev_random((size_t)9 * co * ci, 0.02f), *bias = dev_random(co, 1.0f),
             *y = dev_alloc<bf16>(px * co);
        const double fl = 2.0 * px * co * (double)ci * 9;
        for (int variant = 0; variant < 3; ++variant) {
            if (variant == 2 && co % 320 != 0) continue;
            const float t = time_ms([&] { mvoc_conv3x3_nhwc(x, wt, bias, nullptr, y, N, Hh, W, ci, co, MVOC_BF16, variant, nullptr); }, 5);
            printf("conv %dx%dx%d %4d->%4d v%d : %.3f ms  %.0f TF/s\n", N, Hh, W, ci, co, variant, t, fl / t / 1e9);
        }
        cudaFree(x), cudaFree(wt), cudaFree(bias), cudaFree(y);
    }
    const int64_t ffs[][2] = {{327680, 320}, {81920, 640}, {20480, 1280}};
    for (auto& c : ffs) {
        const int64_t M = c[0];
        const int K = (int)c[1], F = 4 * K;
        bf16 *x = dev_random((size_t)M * K, 1.0f), *w = dev_random((size_t)2 * F * K, 0.05f), *bias = dev_random(2 * F, 1.0f),
             *y = dev_alloc<bf16>((size_t)M * F);
        const float t = time_ms([&] { mvoc_linear_geglu(x, w, bias, y, M, K, F, MVOC_BF16, nullptr); }, 5);
        printf("linear_geglu M=%lld K=%d F=%d : %.3f ms  %.0f TF/s, %.0f GB/s of x+out\n", (long long)M, K, F, t,
               2.0 * M * K * 2.0 * F / t / 1e9, ((double)M * K + (double)M * F) * 2 / t / 1e6);
        cudaFree(x), cudaFree(w), cudaFree(bias), cudaFree(y);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("time section: %s\n", e == cudaSuccess ? "done" : cudaGetErrorString(e));
    return e == cudaSuccess ? 0 : 1;
}

int main(int argc, char** argv) {
    const char* what = argc > 1 ? argv[1] : "conv";
    int rc = mvoc_device_check(0);
    if (rc != 0) {
        printf("device check failed: %s\n", mvoc_last_error());
        return 2;
    }
    int bad;
    if (!strcmp(what, "conv")) bad = section_conv();
    else if (!strcmp(what, "geglu")) bad = section_geglu();
    else if (!strcmp(what, "attn")) bad = section_attn();
    else if (!strcmp(what, "time")) bad = section_time();
    else {
        printf("usage: staged_check conv|geglu|attn|time\n");
        return 2;
    }
    printf("%s: %d failing case(s)\n", what, bad);
    return bad ? 1 : 0;
}
