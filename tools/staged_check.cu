// Standalone checker for the kernels staged in libmvoc_b200_staged.so (no Python, no torch: starts in
// milliseconds on a fresh GPU box).  Every case compares against a naive fp32 GPU reference written here.
//
//   built by __graft_entry__.build() next to the libraries it links (tools/staged_check, rpath = ../mvoc_b200/lib)
//   tools/staged_check attn | time        one section per process: a trap in one kernel must not
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
    cudaError_t e = cudaDeviceSynchronize();
    printf("time section: %s\n", e == cudaSuccess ? "done" : cudaGetErrorString(e));
    return e == cudaSuccess ? 0 : 1;
}

int main(int argc, char** argv) {
    const char* what = argc > 1 ? argv[1] : "attn";
    int rc = mvoc_device_check(0);
    if (rc != 0) {
        printf("device check failed: %s\n", mvoc_last_error());
        return 2;
    }
    int bad;
    if (!strcmp(what, "attn")) bad = section_attn();
    else if (!strcmp(what, "time")) bad = section_time();
    else {
        printf("usage: staged_check attn|time\n");
        return 2;
    }
    printf("%s: %d failing case(s)\n", what, bad);
    return bad ? 1 : 0;
}
