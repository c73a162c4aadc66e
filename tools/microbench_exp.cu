// Micro-benchmark (development aid): MUFU throughput of ex2.approx.f32 vs packed f16x2 / bf16x2 forms,
// and the FMA-pipe polynomial exp2.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/mb_exp tools/microbench_exp.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>

template <int MODE>
__global__ void k(float* out, int iters) {
    float a[8];
    for (int i = 0; i < 8; ++i) a[i] = -0.001f * (threadIdx.x + i);
    unsigned h[8];
    for (int i = 0; i < 8; ++i) h[i] = 0xb800b800u + i;  // two small negative halves
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
            if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
            if (MODE == 3) {  // Cody-Waite + degree-3 polynomial on the FMA pipe
                float x = a[i];
                float xi = floorf(x);
                float f = x - xi;
                float p = fmaf(fmaf(fmaf(0.0555041f, f, 0.2402265f), f, 0.6931472f), f, 1.0f);
                a[i] = __int_as_float(__float_as_int(p) + ((int)xi << 23)) - 1.5f;
            }
        }
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(h[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int per_iter) {
    float* d;
    const int blocks = 148 * 8, threads = 256, iters = 4096;
    cudaMalloc(&d, blocks * threads * 4);
    k<MODE><<<blocks, threads>>>(d, 16);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(d, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)blocks * threads * iters * 8 * per_iter;
    printf("%-28s %8.3f ms  %8.1f Gexp/s  (%.2f exp/clk/SM at 1.9 GHz)\n", name, ms, ops / ms / 1e6,
           ops / (ms * 1e-3) / 148 / 1.9e9);
    cudaFree(d);
}

int main() {
    run<0>("ex2.approx.ftz.f32", 1);
    run<1>("ex2.approx.f16x2", 2);
    run<2>("ex2.approx.ftz.bf16x2", 2);
    run<3>("poly3 exp2 on FMA pipe", 1);
    return 0;
}
