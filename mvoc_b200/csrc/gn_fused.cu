// GroupNorm (+SiLU, + fused per-(frame, channel) add) on channels-last activations [N, S, C] in ONE launch and
// TWO DRAM passes (read X once, write Y once) — the algorithmic traffic.  Same reference call sites as gn_nhwc.cu
// (i2vgen-xl/pnp_utils.py:909-910, :941-965, :1048-1051, :430, :185-188; pipelines/pipeline_i2vgen_xl.py:351-352).
//
// The three-launch form (gn_nhwc.cu: stats, finalize, apply) reads X twice from HBM: by the time `apply` starts,
// the 126 MB L2 has long lost the head of a 210-630 MB tensor.  Here a persistent, co-resident grid walks the tensor
// in SLABS of whole statistics groups sized for L2 (<= ~40 MB):
//   phase 1   every CTA reduces its (frame, chunk) items of the slab to per-group (mean, M2) partials  [HBM -> L2]
//   barrier   grid-wide (one atomic counter + generation word, self-resetting)
//   phase 2   every CTA Chan-merges the partials of its item's group (all chunks, all frames of the group), then
//             normalises the SAME items it read in phase 1                                            [L2 -> HBM]
// so the second read of X is served by L2.  Statistics over T frames (the temporal GroupNorms) are just groups that
// span `frames` consecutive n.  The launch is cooperative (co-residency of the grid is what makes the barrier safe).
// Multi-GPU pixel shards, whose partials must cross GPUs between the phases, keep the three-launch form.
#include <cooperative_groups.h>
#include "common.cuh"

namespace mvoc {

struct GNFParams {
    const void* x;
    void* y;
    const void* gamma;
    const void* beta;
    const void* add;        // [N, C] or null
    float2* partial;        // workspace [N][chunks][G] (mean, M2)
    unsigned* bar;          // {arrival count, generation}; zero-initialised once, self-resetting
    int64_t S;
    int64_t tpc;            // tokens per chunk
    int N, C, G, chunks;
    int R;                  // token lanes per CTA; blockDim = (C/8) * R
    int frames;             // statistics shared by `frames` consecutive n
    int slab_frames;        // frames per slab (a multiple of `frames`)
    float eps;
    int silu;
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Grid barrier number `k` (1-based) of this launch; gen0 = generation word read before the first arrival.
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned gen0, unsigned k) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned arrived = atomicAdd(&bar[0], 1u);
        if (arrived == gridDim.x - 1) {
            bar[0] = 0u;                 // everybody has arrived: nobody touches the counter until the next barrier,
            __threadfence();             // which starts only after the generation below is published
            atomicAdd(&bar[1], 1u);
        } else {
            unsigned spins = 0;
            while ((unsigned)(ld_acquire_u32(&bar[1]) - gen0) < k) {
                __nanosleep(64);
                if (++spins > (1u << 26)) {
                    printf("mvoc gn_fused_kernel: grid barrier %u timed out (block %d)\n", k, blockIdx.x);
                    __trap();
                }
            }
        }
        __threadfence();
    }
    __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(320, 2) gn_fused_kernel(GNFParams p) {
    // smem: part1[R][C], part2[R][C], s1[C], s2[C], sh[C], stat[G] (float2), gen0
    extern __shared__ float gnf_smem[];
    const int C = p.C, VC = C >> 3, Cg = C / p.G, R = p.R;
    float* part1 = gnf_smem;
    float* part2 = part1 + (size_t)R * C;
    float* s1 = part2 + (size_t)R * C;
    float* s2 = s1 + C;
    float* sh = s2 + C;
    float2* stat_s = reinterpret_cast<float2*>(sh + C);
    __shared__ unsigned gen0_s;
    if (threadIdx.x == 0) gen0_s = ld_acquire_u32(&p.bar[1]);
    __syncthreads();
    const unsigned gen0 = gen0_s;
    const int vc = threadIdx.x % VC, r = threadIdx.x / VC;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    unsigned n_bar = 0;

    for (int slab0 = 0; slab0 < p.N; slab0 += p.slab_frames) {
        const int nf = min(p.slab_frames, p.N - slab0);
        const int items = nf * p.chunks;
        // ---------------- phase 1: per-(frame, chunk) partial statistics ----------------
        for (int it = blockIdx.x; it < items; it += gridDim.x) {
            const int n = slab0 + it / p.chunks, chunk = it % p.chunks;
            const int64_t t0 = (int64_t)chunk * p.tpc;
            const int64_t t1 = min(p.S, t0 + p.tpc);
            const T* xb = reinterpret_cast<const T*>(p.x) + (int64_t)n * p.S * C + vc * 8;
            float addv[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) addv[e] = 0.0f;
            if (p.add) unpack8<T>(ld_global16(reinterpret_cast<const T*>(p.add) + (int64_t)n * C + vc * 8), addv);
            float shift[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) shift[e] = 0.0f;
            if (t0 < t1) {  // common per-channel shift (first token of the chunk) against cancellation
                unpack8<T>(ld_global16(xb + t0 * C), shift);
#pragma unroll
                for (int e = 0; e < 8; ++e) shift[e] += addv[e];
            }
            float a1[8], a2[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) a1[e] = a2[e] = 0.0f;
            int64_t t = t0 + r;
            for (; t + 7 * R < t1; t += 8 * R) {   // eight independent 16-byte loads in flight per thread
                Vec16 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = ld_global16(xb + (t + u * R) * C);   // allocating: phase 2 re-reads it from L2
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    float f[8];
                    unpack8<T>(v[u], f);
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float d = f[e] + addv[e] - shift[e];
                        a1[e] += d;
                        a2[e] += d * d;
                    }
                }
            }
            for (; t < t1; t += R) {
                float f[8];
                unpack8<T>(ld_global16(xb + t * C), f);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float d = f[e] + addv[e] - shift[e];
                    a1[e] += d;
                    a2[e] += d * d;
                }
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                part1[(size_t)r * C + vc * 8 + e] = a1[e];
                part2[(size_t)r * C + vc * 8 + e] = a2[e];
            }
            if (r == 0) {
#pragma unroll
                for (int e = 0; e < 8; ++e) sh[vc * 8 + e] = shift[e];
            }
            __syncthreads();
            for (int c = threadIdx.x; c < C; c += blockDim.x) {   // fixed order: deterministic
                float x1 = 0.0f, x2 = 0.0f;
                for (int rr = 0; rr < R; ++rr) {
                    x1 += part1[(size_t)rr * C + c];
                    x2 += part2[(size_t)rr * C + c];
                }
                s1[c] = x1;
                s2[c] = x2;
            }
            __syncthreads();
            if (threadIdx.x < p.G) {   // one thread per group: Chan-merge the per-channel statistics of its channels
                const int g = threadIdx.x;
                const float cnt = (float)(t1 > t0 ? (t1 - t0) : 0);
                float na = 0.f, ma = 0.f, qa = 0.f;
                if (cnt > 0.f) {
                    for (int c = g * Cg; c < (g + 1) * Cg; ++c) {
                        const float m = s1[c] / cnt;
                        const float mb = sh[c] + m;
                        const float qb = fmaxf(s2[c] - s1[c] * m, 0.0f);
                        const float nt = na + cnt;
                        const float d = mb - ma;
                        ma += d * (cnt / nt);
                        qa += qb + d * d * (na * cnt / nt);
                        na = nt;
                    }
                }
                p.partial[((int64_t)n * p.chunks + chunk) * p.G + g] = make_float2(ma, qa);
            }
            __syncthreads();   // smem is reused by the next item
        }
        grid_barrier(p.bar, gen0, ++n_bar);
        // ---------------- phase 2: merge the group's partials, normalise the same items ----------------
        for (int it = blockIdx.x; it < items; it += gridDim.x) {
            const int n = slab0 + it / p.chunks, chunk = it % p.chunks;
            const int n_first = (n / p.frames) * p.frames;
            const int total = p.frames * p.chunks;
            for (int g = warp; g < p.G; g += n_warps) {   // one warp per group: frames x chunks partials
                float na = 0.f, ma = 0.f, qa = 0.f;
                for (int i = lane; i < total; i += 32) {
                    const int fr = i / p.chunks, ch = i - fr * p.chunks;
                    const int64_t c0 = (int64_t)ch * p.tpc;
                    const float nb = (float)(max((int64_t)0, min(p.S, c0 + p.tpc) - c0) * Cg);
                    const float2 pm = p.partial[((int64_t)(n_first + fr) * p.chunks + ch) * p.G + g];
                    if (nb > 0.f) {
                        const float nt = na + nb;
                        const float d = pm.x - ma;
                        ma += d * (nb / nt);
                        qa += pm.y + d * d * (na * nb / nt);
                        na = nt;
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float nb = __shfl_xor_sync(0xffffffffu, na, o);
                    const float mb = __shfl_xor_sync(0xffffffffu, ma, o);
                    const float qb = __shfl_xor_sync(0xffffffffu, qa, o);
                    if (nb > 0.f) {
                        const float nt = na + nb;
                        const float d = mb - ma;
                        ma += d * (nb / nt);
                        qa += qb + d * d * (na * nb / nt);
                        na = nt;
                    }
                }
                if (lane == 0) stat_s[g] = make_float2(ma, rsqrtf(qa / fmaxf(na, 1.0f) + p.eps));
            }
            __syncthreads();
            const int64_t t0 = (int64_t)chunk * p.tpc;
            const int64_t t1 = min(p.S, t0 + p.tpc);
            const T* xb = reinterpret_cast<const T*>(p.x) + (int64_t)n * p.S * C + vc * 8;
            T* yb = reinterpret_cast<T*>(p.y) + (int64_t)n * p.S * C + vc * 8;
            float ga[8], be[8], sc[8], sf[8], addv[8];
            unpack8<T>(ld_global16(reinterpret_cast<const T*>(p.gamma) + vc * 8), ga);
            unpack8<T>(ld_global16(reinterpret_cast<const T*>(p.beta) + vc * 8), be);
#pragma unroll
            for (int e = 0; e < 8; ++e) addv[e] = 0.0f;
            if (p.add) unpack8<T>(ld_global16(reinterpret_cast<const T*>(p.add) + (int64_t)n * C + vc * 8), addv);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float2 ms = stat_s[(vc * 8 + e) / Cg];
                sc[e] = ga[e] * ms.y;
                sf[e] = be[e] + (addv[e] - ms.x) * sc[e];
            }
            int64_t t = t0 + r;
            for (; t + 7 * R < t1; t += 8 * R) {
                Vec16 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = ld_stream16(xb + (t + u * R) * C);   // last use of X
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    float f[8];
                    unpack8<T>(v[u], f);
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float w = f[e] * sc[e] + sf[e];
                        f[e] = p.silu ? silu_f(w) : w;
                    }
                    st_stream16(yb + (t + u * R) * C, pack8<T>(f));
                }
            }
            for (; t < t1; t += R) {
                float f[8];
                unpack8<T>(ld_stream16(xb + t * C), f);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float w = f[e] * sc[e] + sf[e];
                    f[e] = p.silu ? silu_f(w) : w;
                }
                st_stream16(yb + t * C, pack8<T>(f));
            }
            __syncthreads();   // stat_s is rewritten by the next item
        }
        // no barrier needed here: the next slab writes partials of OTHER frames
    }
}

// L2-sized slabs: the slab's X must survive in the 126 MB L2 from phase 1 to phase 2 while phase 2 also writes as
// many bytes of Y through it.
constexpr int64_t GNF_SLAB_BYTES = 40ll << 20;
constexpr int GNF_MAX_CHUNKS = 256;

struct GNFGeometry {
    int R, threads, chunks, slab_frames, grid;
    int64_t tpc;
    size_t smem;
};

template <typename T>
static int gnf_plan(int64_t N, int64_t S, int C, int G, int frames, GNFGeometry* geo, const char* what) {
    const int VC = C / 8;
    int r = 320 / VC;
    if (r < 1) r = 1;
    if (r > 32) r = 32;
    while (VC * r > 320 && r > 1) --r;
    geo->R = r;
    geo->threads = VC * r;
    MVOC_REQUIRE(geo->threads <= 320, MVOC_ERR_UNSUPPORTED, "%s: C=%d exceeds 2560 channels", what, C);
    MVOC_REQUIRE(geo->threads >= G && geo->threads >= 32, MVOC_ERR_UNSUPPORTED, "%s: C=%d too small for G=%d", what, C, G);
    geo->smem = ((2 * (size_t)r + 3) * (size_t)C) * sizeof(float) + (size_t)G * sizeof(float2);
    MVOC_REQUIRE(geo->smem <= 96 * 1024, MVOC_ERR_UNSUPPORTED, "%s: C=%d needs %zu B of smem", what, C, geo->smem);
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        cudaFuncSetAttribute(gn_fused_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        cudaFuncSetAttribute(gn_fused_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        attr_set[dev] = true;
    }
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gn_fused_kernel<T>, geo->threads, geo->smem);
    MVOC_REQUIRE(e == cudaSuccess && per_sm >= 1, MVOC_ERR_CUDA, "%s: occupancy query failed (%s)", what,
                 cudaGetErrorString(e));
    if (per_sm > 4) per_sm = 4;
    const int64_t frame_bytes = S * (int64_t)C * 2;
    const int64_t group_bytes = frame_bytes * frames;
    // whole statistics groups per slab, about GNF_SLAB_BYTES (a single group may exceed it: nothing to split)
    int64_t groups_per_slab = GNF_SLAB_BYTES / group_bytes;
    if (groups_per_slab < 1) groups_per_slab = 1;
    int64_t slab_frames = groups_per_slab * frames;
    if (slab_frames > N) slab_frames = N;
    geo->slab_frames = (int)slab_frames;
    // items of a slab = slab_frames x chunks: a small whole number per CTA, 24-96 KB each
    int grid = per_sm * num_sms();
    int64_t chunks = 1;
    for (int k = 1; k <= 8; ++k) {
        chunks = ((int64_t)grid * k) / slab_frames;
        if (chunks < 1) chunks = 1;
        if (frame_bytes / chunks <= 96 * 1024) break;
    }
    if (chunks > GNF_MAX_CHUNKS) chunks = GNF_MAX_CHUNKS;
    if (chunks > S) chunks = S;
    const int64_t tpc = (S + chunks - 1) / chunks;
    geo->chunks = (int)((S + tpc - 1) / tpc);
    geo->tpc = tpc;
    const int64_t items = slab_frames * geo->chunks;
    geo->grid = (int)(items < grid ? items : grid);
    return MVOC_OK;
}

}  // namespace mvoc

using namespace mvoc;

extern "C" int64_t mvoc_groupnorm_nhwc_fused_workspace_bytes(int64_t N, int G) {
    return N * (int64_t)G * GNF_MAX_CHUNKS * (int64_t)sizeof(float2) + 256;
}

// workspace: mvoc_groupnorm_nhwc_fused_workspace_bytes(N, G) bytes whose FIRST 256 bytes were zeroed once when the
// buffer was created (the grid barrier's counter / generation words; the kernel leaves the counter at zero).
extern "C" int mvoc_groupnorm_nhwc_fused(const void* x, void* y, const void* gamma, const void* beta, const void* add,
                                         int64_t N, int64_t S, int C, int G, int frames_per_stat, float eps, int silu,
                                         int dtype, void* workspace, void* stream) {
    const char* what = "mvoc_groupnorm_nhwc_fused";
    MVOC_REQUIRE(x && y && gamma && beta && workspace, MVOC_ERR_INVALID_ARG, "%s: null pointer", what);
    MVOC_REQUIRE(N > 0 && S > 0 && C > 0 && G > 0 && C % G == 0, MVOC_ERR_INVALID_ARG,
                 "%s: bad shape N=%lld S=%lld C=%d G=%d", what, (long long)N, (long long)S, C, G);
    MVOC_REQUIRE(C % 8 == 0 && C / 8 <= 1024, MVOC_ERR_UNSUPPORTED, "%s: C=%d must be a multiple of 8 (<= 8192)", what, C);
    MVOC_REQUIRE(G <= 128, MVOC_ERR_UNSUPPORTED, "%s: G=%d too large", what, G);
    MVOC_REQUIRE(frames_per_stat >= 1 && N % frames_per_stat == 0, MVOC_ERR_INVALID_ARG,
                 "%s: N=%lld not a multiple of frames_per_stat=%d", what, (long long)N, frames_per_stat);
    MVOC_REQUIRE(N <= 0x7fffffffLL / 256, MVOC_ERR_UNSUPPORTED, "%s: N=%lld too large", what, (long long)N);
    MVOC_REQUIRE(dtype == MVOC_BF16 || dtype == MVOC_F16, MVOC_ERR_UNSUPPORTED, "%s: dtype %d unsupported (bf16/f16 only)",
                 what, dtype);
    MVOC_REQUIRE((uintptr_t)x % 16 == 0 && (uintptr_t)y % 16 == 0 && (uintptr_t)workspace % 256 == 0 &&
                     (uintptr_t)add % 16 == 0,
                 MVOC_ERR_INVALID_ARG, "%s: x / y / add must be 16-byte and the workspace 256-byte aligned", what);
    GNFGeometry geo;
    int rc = dtype == MVOC_BF16 ? gnf_plan<__nv_bfloat16>(N, S, C, G, frames_per_stat, &geo, what)
                                : gnf_plan<__half>(N, S, C, G, frames_per_stat, &geo, what);
    if (rc != MVOC_OK) return rc;
    GNFParams p{};
    p.x = x, p.y = y, p.gamma = gamma, p.beta = beta, p.add = add;
    p.bar = reinterpret_cast<unsigned*>(workspace);
    p.partial = reinterpret_cast<float2*>(reinterpret_cast<char*>(workspace) + 256);
    p.S = S, p.tpc = geo.tpc;
    p.N = (int)N, p.C = C, p.G = G, p.chunks = geo.chunks, p.R = geo.R;
    p.frames = frames_per_stat, p.slab_frames = geo.slab_frames;
    p.eps = eps, p.silu = silu;
    void* args[] = {&p};
    const void* fn = dtype == MVOC_BF16 ? (const void*)gn_fused_kernel<__nv_bfloat16> : (const void*)gn_fused_kernel<__half>;
    cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3(geo.grid), dim3(geo.threads), args, geo.smem, (cudaStream_t)stream);
    MVOC_REQUIRE(e == cudaSuccess, MVOC_ERR_CUDA, "%s: cooperative launch failed: %s", what, cudaGetErrorString(e));
    return MVOC_OK;
}
