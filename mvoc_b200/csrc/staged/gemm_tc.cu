// STAGED FOR ROUND 2 — compiled (sm_100a) but NOT YET RUN ON HARDWARE, not linked into libmvoc_b200.so and not
// on any product path.  Built into mvoc_b200/lib/libmvoc_b200_staged.so; see include/mvoc_b200_staged.h.
//
// Implicit-GEMM 3x3 convolution (stride 1, pad 1) and GEGLU-fused Linear on channels-last bf16 activations with
// tcgen05 tensor cores: SURVEY §8f-1 (the dense work the reference leaves to cuDNN / cuBLAS through
// i2vgen-xl/pnp_utils.py:939, :968 (resnet conv1 / conv2), :1048-1051 (temporal convs) and :335 (GEGLU feed-forward)).
//
// One CTA = 128 output pixels (TMEM lanes) x BN output channels (fp32 accumulators in TMEM columns).
//   conv : D[p, co] = sum_{tap, ci} X[p + (kh-1, kw-1), ci] * Wt[tap, co, ci]
//          The A tile of a tap is a 4-D TMA box {64 ch, bw, bh, bn} of X[N, H, W, Cin] whose start coordinate is
//          shifted by the tap offset; out-of-range rows / columns are ZERO-FILLED by TMA, which is the padding.
//          No im2col buffer exists anywhere.  128 pixel rows of 128 B = the K-major 128B-swizzled operand layout
//          the attention kernel's Q tile uses (attn_tc.cu) — descriptors and instruction descriptor are the same.
//   geglu: taps = 1, the two accumulator halves are the value and the gate columns of one Linear; the epilogue
//          writes value * gelu(gate), so the [M, 2F] intermediate never reaches HBM.
// Warps 0-3 epilogue (thread i = TMEM lane i = pixel i), warp 4 TMA producer, warp 5 MMA issuer.
#include <cuda.h>
#include "../common.cuh"
#include "../ptx.cuh"
#include "../../../include/mvoc_b200_staged.h"

namespace mvoc {
namespace gemm {

constexpr int BM = 128;                   // pixels per CTA
constexpr int KC = 64;                    // channels per K chunk = one 128-byte row
constexpr int A_BYTES = BM * KC * 2;      // 16 KB
constexpr int THREADS = 192;
constexpr int SMEM_LIMIT = 232448;        // 227 KB per CTA

enum Epilogue { EPI_BIAS = 0, EPI_GEGLU = 1 };

// BN = accumulator columns of the CTA; issued as kPieces MMAs of N = BN / kPieces (N <= 256 per instruction and
// per TMA box).  kCtas = CTAs per SM the shared-memory budget is sized for.
template <int BN_, int kPieces_, int kStages_, int kCtas_>
struct Cfg {
    static constexpr int BN = BN_, kPieces = kPieces_, kStages = kStages_, kCtas = kCtas_;
    static constexpr int NP = BN / kPieces;                 // N of one MMA / rows of one B box
    static constexpr int B_BYTES = BN * KC * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int bar_off = kStages * STAGE_BYTES;   // full[], empty[], acc_full
    static constexpr int n_bars = 2 * kStages + 1;
    static constexpr int tmem_ptr_off = bar_off + n_bars * 8;
    static constexpr int alloc = tmem_ptr_off + 16;
    static constexpr uint32_t tmem_cols = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : BN <= 256 ? 256 : 512;
    static_assert(NP % 16 == 0 && NP >= 16 && NP <= 256, "tcgen05.mma with M = 128 needs N % 16 == 0, N <= 256");
    static_assert((NP * KC * 2) % 1024 == 0, "B pieces must keep the 1024-byte swizzle-atom alignment");
    static_assert(kCtas * (alloc + 1024) <= 233472, "shared memory budget of the SM exceeded");
    static_assert(alloc <= SMEM_LIMIT, "shared memory budget of the CTA exceeded");
    static_assert(kCtas * tmem_cols <= 512, "TMEM columns of the SM exceeded");
};

struct Params {
    __nv_bfloat16* out;            // [pixels, out_ld]
    const __nv_bfloat16* bias;     // [Cout] (conv) / [2F] (geglu) or nullptr
    const __nv_bfloat16* residual; // [pixels, out_ld] added in the epilogue, or nullptr
    int N, H, W;                   // activation frames / rows / columns (geglu: 1, 1, M)
    int bn, bh, bw;                // pixel box of a CTA, bn*bh*bw == 128
    int tiles_w, tiles_h;          // ceil(W / bw), ceil(H / bh)
    int n_tiles;                   // column tiles of the output
    int k_chunks;                  // Cin / 64
    int taps;                      // 9 (3x3, pad 1) or 1
    int out_ld;                    // row stride of out / residual in elements (Cout, or F for geglu)
    int gate_row_offset;           // geglu: row of Wt where the gate half starts (F); conv: unused
};

// ---- thread-block-cluster pieces of the kCluster variant (pairs of CTAs that share the weight tile) ----------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA tile load written to the same shared-memory offset of every CTA in cta_mask; each destination CTA's mbarrier
// (same offset) receives the complete_tx of the bytes it got.
__device__ __forceinline__ void tma_load_4d_multicast(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1,
                                                      int c2, int c3, uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
        "[%0], [%1, {%3, %4, %5, %6}], [%2], %7;" ::"r"(dst),
        "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "h"(cta_mask)
        : "memory");
}
// tcgen05.commit whose mbarrier arrive is delivered to the barrier at this offset in every CTA of cta_mask
__device__ __forceinline__ void tc_commit_multicast(uint32_t bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::
                     "r"(bar), "h"(cta_mask)
                 : "memory");
}

// kCluster: the grid is launched in clusters of two CTAs that work on two neighbouring pixel tiles and the SAME
// column tile.  Each CTA loads its own activation tile, but only ONE of the two weight pieces, multicast into both
// CTAs: the L2 -> SM traffic of the weights, the larger operand, halves.
template <typename C, int kEpi, bool kCluster>
__device__ __forceinline__ void gemm_tc_body(const CUtensorMap& tm_x, const CUtensorMap& tm_w, const Params& prm) {
    static_assert(!kCluster || C::kPieces == 2, "the cluster variant splits the two weight pieces between two CTAs");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t sbase = smem_u32(smem_raw);
    if ((sbase & 1023u) != 0u) {
        if (threadIdx.x == 0) printf("mvoc gemm_tc_kernel: dynamic smem base 0x%x is not 1024-byte aligned\n", sbase);
        __trap();
    }
    const uint32_t bars = sbase + C::bar_off;
    const uint32_t b_full = bars, b_empty = bars + 8 * C::kStages, b_acc = bars + 16 * C::kStages;
    const uint32_t s_tmem_ptr = sbase + C::tmem_ptr_off;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // column tile fastest: CTAs that share an activation tile are launched next to each other (L2 reuse)
    const uint32_t crank = kCluster ? cluster_ctarank() : 0u;
    int n_tile, m_tile;
    if (kCluster) {   // CTA pair p: column tile p % n_tiles, pixel tiles 2 * (p / n_tiles) + {0, 1}
        const int pair = blockIdx.x >> 1;
        n_tile = pair % prm.n_tiles;
        m_tile = (pair / prm.n_tiles) * 2 + (int)crank;   // may be one past the last tile: all loads OOB, no stores
    } else {
        n_tile = blockIdx.x % prm.n_tiles;
        m_tile = blockIdx.x / prm.n_tiles;
    }
    const int tw = m_tile % prm.tiles_w;
    m_tile /= prm.tiles_w;
    const int th = m_tile % prm.tiles_h;
    const int tn = m_tile / prm.tiles_h;
    const int w0 = tw * prm.bw, h0 = th * prm.bh, n0 = tn * prm.bn;
    const int pad = prm.taps == 9 ? 1 : 0;
    const int n_chunks = prm.k_chunks * prm.taps;
    // geglu: the CTA's BN accumulator columns are NP value columns followed by NP gate columns
    const int out_col0 = kEpi == EPI_GEGLU ? n_tile * C::NP : n_tile * C::BN;

    if (warp == 4 && lane == 0) {
        ptx::prefetch_tensormap(&tm_x);
        ptx::prefetch_tensormap(&tm_w);
        for (int s = 0; s < C::kStages; ++s) {
            ptx::mbar_init(b_full + 8 * s, 1);
            ptx::mbar_init(b_empty + 8 * s, kCluster ? 2 : 1);   // cluster: both CTAs' MMAs read what I multicast
        }
        ptx::mbar_init(b_acc, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 5) {
        ptx::tmem_alloc(s_tmem_ptr, C::tmem_cols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    if (kCluster) cluster_sync_all();   // the peer's barriers exist before anything is multicast to them
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem_raw + C::tmem_ptr_off);

    if (warp == 4) {
        // ===================== TMA producer =====================
        for (int c = 0; c < n_chunks; ++c) {
            const int s = c % C::kStages;
            const uint32_t ph = (uint32_t)(c / C::kStages) & 1u;
            const int kc = c / prm.taps, tap = c - kc * prm.taps;   // taps innermost: the 9 shifted boxes of one
            const int kh = tap / 3, kw = tap - kh * 3;              // channel chunk hit the same L2 lines
            ptx::mbar_wait(b_empty + 8 * s, ph ^ 1u, 1);
            if (lane == 0) {
                const uint32_t sA = sbase + s * C::STAGE_BYTES, sB = sA + A_BYTES;
                ptx::mbar_expect_tx(b_full + 8 * s, C::STAGE_BYTES);
                ptx::tma_load_4d(sA, &tm_x, b_full + 8 * s, kc * KC, w0 + kw - pad, h0 + kh - pad, n0);
#pragma unroll
                for (int p = 0; p < C::kPieces; ++p) {
                    int wrow;
                    if (kEpi == EPI_GEGLU) wrow = out_col0 + p * prm.gate_row_offset;   // piece 0 value, 1 gate
                    else wrow = out_col0 + p * C::NP;
                    if (!kCluster)
                        ptx::tma_load_4d(sB + p * (C::NP * KC * 2), &tm_w, b_full + 8 * s, kc * KC, wrow, tap, 0);
                    else if (p == (int)crank)   // my piece goes to both CTAs; the peer sends me the other one
                        tma_load_4d_multicast(sB + p * (C::NP * KC * 2), &tm_w, b_full + 8 * s, kc * KC, wrow, tap, 0,
                                              (uint16_t)0x3);
                }
            }
            __syncwarp();
        }
    } else if (warp == 5) {
        // ===================== MMA issuer =====================
        constexpr uint32_t IDESC = ptx::idesc_bf16(BM, C::NP, 0, 0);   // A and B K-major
        for (int c = 0; c < n_chunks; ++c) {
            const int s = c % C::kStages;
            const uint32_t ph = (uint32_t)(c / C::kStages) & 1u;
            ptx::mbar_wait(b_full + 8 * s, ph, 2);
            ptx::tc_fence_after();
            if (lane == 0) {
                const uint32_t sA = sbase + s * C::STAGE_BYTES, sB = sA + A_BYTES;
                const uint64_t a0 = ptx::smem_desc_sw128(sA, 16, 1024);
#pragma unroll
                for (int p = 0; p < C::kPieces; ++p) {
                    const uint64_t b0 = ptx::smem_desc_sw128(sB + p * (C::NP * KC * 2), 16, 1024);
#pragma unroll
                    for (int ks = 0; ks < KC / 16; ++ks)   // +32 B per 16-element K step (encoded >> 4)
                        ptx::mma_ss(tmem + p * C::NP, a0 + (uint64_t)(ks * 2), b0 + (uint64_t)(ks * 2), IDESC,
                                    (c > 0 || ks > 0) ? 1u : 0u);
                }
                if (kCluster) tc_commit_multicast(b_empty + 8 * s, (uint16_t)0x3);   // frees the stage in BOTH producers
                else ptx::tc_commit(b_empty + 8 * s);            // stage reusable once these MMAs have read it
                if (c == n_chunks - 1) ptx::tc_commit(b_acc);    // accumulators complete
            }
            __syncwarp();
        }
    } else {
        // ===================== epilogue (warps 0-3) =====================
        const int row = threadIdx.x;                       // pixel inside the tile == TMEM lane
        const int iw = row % prm.bw, ih = (row / prm.bw) % prm.bh, in = row / (prm.bw * prm.bh);
        const int pw = w0 + iw, phh = h0 + ih, pn = n0 + in;
        const bool valid = pw < prm.W && phh < prm.H && pn < prm.N;
        const int64_t pixel = ((int64_t)pn * prm.H + phh) * prm.W + pw;
        const uint32_t tacc = tmem + ((uint32_t)(warp * 32) << 16);
        __nv_bfloat16* orow = prm.out + pixel * prm.out_ld + out_col0;
        const __nv_bfloat16* rrow = prm.residual ? prm.residual + pixel * prm.out_ld + out_col0 : nullptr;
        ptx::mbar_wait(b_acc, 0, 3);
        ptx::tc_fence_after();
        constexpr int OUT_COLS = kEpi == EPI_GEGLU ? C::NP : C::BN;
        static_assert(OUT_COLS % 32 == 0, "epilogue works on 32-column chunks");
#pragma unroll 1
        for (int c0 = 0; c0 < OUT_COLS; c0 += 32) {
            uint32_t r[32];
            ptx::tmem_ld32(tacc + c0, r);
            uint32_t g[32];
            if (kEpi == EPI_GEGLU) ptx::tmem_ld32(tacc + C::NP + c0, g);
            ptx::tmem_wait_ld();
            if (valid) {
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    float f[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(r[i + e]);
                    if (prm.bias) {
                        float bv[8];
                        unpack8<__nv_bfloat16>(ld_global16(prm.bias + out_col0 + c0 + i), bv);
#pragma unroll
                        for (int e = 0; e < 8; ++e) f[e] += bv[e];
                    }
                    if (kEpi == EPI_GEGLU) {
                        float gv[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) gv[e] = __uint_as_float(g[i + e]);
                        if (prm.bias) {
                            float bg[8];
                            unpack8<__nv_bfloat16>(ld_global16(prm.bias + prm.gate_row_offset + out_col0 + c0 + i), bg);
#pragma unroll
                            for (int e = 0; e < 8; ++e) gv[e] += bg[e];
                        }
#pragma unroll
                        for (int e = 0; e < 8; ++e)   // exact (erf) GELU, as torch.nn.functional.gelu
                            f[e] *= 0.5f * gv[e] * (1.0f + erff(gv[e] * 0.70710678118654752f));
                    }
                    if (rrow) {
                        float rv[8];
                        unpack8<__nv_bfloat16>(ld_stream16(rrow + c0 + i), rv);
#pragma unroll
                        for (int e = 0; e < 8; ++e) f[e] += rv[e];
                    }
                    *reinterpret_cast<Vec16*>(orow + c0 + i) = pack8<__nv_bfloat16>(f);
                }
            }
        }
        ptx::tc_fence_before();
    }

    __syncthreads();
    if (warp == 5) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem, C::tmem_cols);
    }
    if (kCluster) cluster_sync_all();   // no CTA leaves while its peer may still multicast into it / arrive on it
}

template <typename C, int kEpi>
__global__ void __launch_bounds__(THREADS, C::kCtas)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w, const Params prm) {
    gemm_tc_body<C, kEpi, false>(tm_x, tm_w, prm);
}

template <typename C, int kEpi>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, C::kCtas)
gemm_tc_cluster_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                       const Params prm) {
    gemm_tc_body<C, kEpi, true>(tm_x, tm_w, prm);
}

// ------------------------------------------------------------------ host ---
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
        if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
    }
    return fn;
}

// 4-D bf16 tensor, dims / box innermost first, strides (bytes) of dims 1..3; 128-byte swizzle, zero OOB fill.
static int make_map4(CUtensorMap* m, const void* base, const cuuint64_t (&dims)[4], const cuuint64_t (&strides)[3],
                     const cuuint32_t (&box)[4], const char* what) {
    EncodeTiledFn fn = get_encode_fn();
    MVOC_REQUIRE(fn != nullptr, MVOC_ERR_DRIVER, "%s: cuTensorMapEncodeTiled unavailable", what);
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MVOC_REQUIRE(r == CUDA_SUCCESS, MVOC_ERR_DRIVER,
                 "%s: cuTensorMapEncodeTiled failed with CUresult %d (dims %llu,%llu,%llu,%llu box %u,%u,%u,%u)", what,
                 (int)r, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
                 (unsigned long long)dims[3], box[0], box[1], box[2], box[3]);
    return MVOC_OK;
}

// Pixel box (bn, bh, bw), powers of two with product 128, that wastes the fewest zero-filled pixels on the
// (N, H, W) grid; ties go to the widest box (longest contiguous runs).
static void choose_box(int N, int H, int W, int* bn, int* bh, int* bw) {
    int64_t best = -1;
    for (int w = 128; w >= 1; w >>= 1)
        for (int h = 128 / w; h >= 1; h >>= 1) {
            const int n = 128 / (w * h);
            const int64_t padded = (int64_t)((W + w - 1) / w) * w * ((H + h - 1) / h) * h * ((N + n - 1) / n) * n;
            if (best < 0 || padded < best) {
                best = padded;
                *bn = n, *bh = h, *bw = w;
            }
        }
}

template <typename C, int kEpi>
static int launch(const CUtensorMap& mx, const CUtensorMap& mw, const Params& prm, int64_t ctas, cudaStream_t s,
                  const char* what) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<C, kEpi>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             C::alloc);
        MVOC_REQUIRE(e == cudaSuccess, MVOC_ERR_CUDA, "%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
        attr_set = true;
    }
    MVOC_REQUIRE(ctas > 0 && ctas <= 0x7fffffffLL, MVOC_ERR_UNSUPPORTED, "%s: %lld CTAs", what, (long long)ctas);
    gemm_tc_kernel<C, kEpi><<<(unsigned)ctas, THREADS, C::alloc, s>>>(mx, mw, prm);
    return check_launch(what);
}

// clusters of two CTAs along the pixel axis: m_tiles is rounded up to an even number
template <typename C, int kEpi>
static int launch_cluster(const CUtensorMap& mx, const CUtensorMap& mw, const Params& prm, int64_t m_tiles,
                          cudaStream_t s, const char* what) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc_cluster_kernel<C, kEpi>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, C::alloc);
        MVOC_REQUIRE(e == cudaSuccess, MVOC_ERR_CUDA, "%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
        attr_set = true;
    }
    const int64_t ctas = ((m_tiles + 1) / 2) * 2 * prm.n_tiles;
    MVOC_REQUIRE(ctas > 0 && ctas <= 0x7fffffffLL, MVOC_ERR_UNSUPPORTED, "%s: %lld CTAs", what, (long long)ctas);
    gemm_tc_cluster_kernel<C, kEpi><<<(unsigned)ctas, THREADS, C::alloc, s>>>(mx, mw, prm);
    return check_launch(what);
}

// tile configurations: accumulator columns, MMA pieces, pipeline stages, CTAs per SM
using Cfg320 = Cfg<320, 2, 4, 1>;   // 128 x 320: one activation tile feeds 320 channels (least L2 traffic per FLOP)
using Cfg160 = Cfg<160, 1, 3, 2>;   // 128 x 160, two CTAs per SM (the other CTA's epilogue hides behind MMAs)
using Cfg128 = Cfg<128, 1, 3, 2>;
using Cfg64 = Cfg<64, 1, 4, 2>;
using CfgG128 = Cfg<128, 2, 3, 2>;  // GEGLU: 64 value + 64 gate columns

}  // namespace gemm
}  // namespace mvoc

using namespace mvoc;

extern "C" int mvoc_conv3x3_nhwc(const void* x, const void* w_taps, const void* bias, const void* residual, void* out,
                                 int N, int H, int W, int Cin, int Cout, int dtype, int variant, void* stream) {
    const char* what = "mvoc_conv3x3_nhwc";
    MVOC_REQUIRE(x && w_taps && out, MVOC_ERR_INVALID_ARG, "%s: null pointer", what);
    MVOC_REQUIRE(dtype == MVOC_BF16, MVOC_ERR_UNSUPPORTED, "%s: dtype %d unsupported (bf16 only)", what, dtype);
    MVOC_REQUIRE(N > 0 && H > 0 && W > 0, MVOC_ERR_INVALID_ARG, "%s: empty activation N=%d H=%d W=%d", what, N, H, W);
    MVOC_REQUIRE(Cin > 0 && Cin % 64 == 0 && Cout > 0 && Cout % 64 == 0, MVOC_ERR_UNSUPPORTED,
                 "%s: Cin=%d / Cout=%d must be multiples of 64", what, Cin, Cout);
    MVOC_REQUIRE(variant >= 0 && variant <= 2, MVOC_ERR_INVALID_ARG, "%s: unknown variant %d", what, variant);
    MVOC_REQUIRE(variant != 2 || Cout % 320 == 0, MVOC_ERR_UNSUPPORTED,
                 "%s: variant 2 (CTA pairs sharing the weight tile) needs Cout %% 320 == 0, got %d", what, Cout);
    MVOC_REQUIRE(((uintptr_t)x % 16 == 0) && ((uintptr_t)w_taps % 16 == 0) && ((uintptr_t)out % 16 == 0) &&
                     ((uintptr_t)bias % 16 == 0) && ((uintptr_t)residual % 16 == 0),
                 MVOC_ERR_INVALID_ARG, "%s: pointers must be 16-byte aligned", what);
    gemm::Params prm{};
    prm.out = (__nv_bfloat16*)out;
    prm.bias = (const __nv_bfloat16*)bias;
    prm.residual = (const __nv_bfloat16*)residual;
    prm.N = N, prm.H = H, prm.W = W;
    gemm::choose_box(N, H, W, &prm.bn, &prm.bh, &prm.bw);
    prm.tiles_w = (W + prm.bw - 1) / prm.bw;
    prm.tiles_h = (H + prm.bh - 1) / prm.bh;
    const int64_t m_tiles = (int64_t)prm.tiles_w * prm.tiles_h * ((N + prm.bn - 1) / prm.bn);
    prm.k_chunks = Cin / 64;
    prm.taps = 9;
    prm.out_ld = Cout;
    prm.gate_row_offset = 0;
    // variant 0: widest tile that divides Cout; variant 1: never the single-CTA 320-column tile
    int BN = Cout % 320 == 0 && variant != 1 ? 320 : Cout % 160 == 0 ? 160 : Cout % 128 == 0 ? 128 : 64;
    prm.n_tiles = Cout / BN;
    const int NP = BN == 320 ? 160 : BN;

    CUtensorMap mx, mw;
    int rc;
    {
        const cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
        const cuuint64_t str[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
        const cuuint32_t box[4] = {64, (cuuint32_t)prm.bw, (cuuint32_t)prm.bh, (cuuint32_t)prm.bn};
        if ((rc = gemm::make_map4(&mx, x, dims, str, box, what)) != MVOC_OK) return rc;
    }
    {
        const cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)Cout, 9, 1};
        const cuuint64_t str[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)Cout * Cin * 2, (cuuint64_t)9 * Cout * Cin * 2};
        const cuuint32_t box[4] = {64, (cuuint32_t)NP, 1, 1};
        if ((rc = gemm::make_map4(&mw, w_taps, dims, str, box, what)) != MVOC_OK) return rc;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t ctas = m_tiles * prm.n_tiles;
    if (variant == 2) return gemm::launch_cluster<gemm::Cfg320, gemm::EPI_BIAS>(mx, mw, prm, m_tiles, s, what);
    switch (BN) {
        case 320: return gemm::launch<gemm::Cfg320, gemm::EPI_BIAS>(mx, mw, prm, ctas, s, what);
        case 160: return gemm::launch<gemm::Cfg160, gemm::EPI_BIAS>(mx, mw, prm, ctas, s, what);
        case 128: return gemm::launch<gemm::Cfg128, gemm::EPI_BIAS>(mx, mw, prm, ctas, s, what);
        default: return gemm::launch<gemm::Cfg64, gemm::EPI_BIAS>(mx, mw, prm, ctas, s, what);
    }
}

extern "C" int mvoc_linear_geglu(const void* x, const void* w, const void* bias, void* out, int64_t M, int K, int F,
                                 int dtype, void* stream) {
    const char* what = "mvoc_linear_geglu";
    MVOC_REQUIRE(x && w && out, MVOC_ERR_INVALID_ARG, "%s: null pointer", what);
    MVOC_REQUIRE(dtype == MVOC_BF16, MVOC_ERR_UNSUPPORTED, "%s: dtype %d unsupported (bf16 only)", what, dtype);
    MVOC_REQUIRE(M > 0 && M <= 0x7fffffffLL, MVOC_ERR_INVALID_ARG, "%s: M=%lld", what, (long long)M);
    MVOC_REQUIRE(K > 0 && K % 64 == 0, MVOC_ERR_UNSUPPORTED, "%s: K=%d must be a multiple of 64", what, K);
    MVOC_REQUIRE(F > 0 && F % 64 == 0, MVOC_ERR_UNSUPPORTED, "%s: F=%d must be a multiple of 64", what, F);
    MVOC_REQUIRE(((uintptr_t)x % 16 == 0) && ((uintptr_t)w % 16 == 0) && ((uintptr_t)out % 16 == 0) &&
                     ((uintptr_t)bias % 16 == 0),
                 MVOC_ERR_INVALID_ARG, "%s: pointers must be 16-byte aligned", what);
    gemm::Params prm{};
    prm.out = (__nv_bfloat16*)out;
    prm.bias = (const __nv_bfloat16*)bias;
    prm.residual = nullptr;
    prm.N = 1, prm.H = 1, prm.W = (int)M;
    prm.bn = 1, prm.bh = 1, prm.bw = 128;
    prm.tiles_w = (int)((M + 127) / 128);
    prm.tiles_h = 1;
    prm.k_chunks = K / 64;
    prm.taps = 1;
    prm.out_ld = F;
    prm.gate_row_offset = F;
    // value + gate columns per CTA: 160 + 160 (one CTA per SM), else 64 + 64 (two CTAs per SM)
    const int NP = F % 160 == 0 ? 160 : 64;
    prm.n_tiles = F / NP;

    CUtensorMap mx, mw;
    int rc;
    {
        const cuuint64_t dims[4] = {(cuuint64_t)K, (cuuint64_t)M, 1, 1};
        const cuuint64_t str[3] = {(cuuint64_t)K * 2, (cuuint64_t)M * K * 2, (cuuint64_t)M * K * 2};
        const cuuint32_t box[4] = {64, 128, 1, 1};
        if ((rc = gemm::make_map4(&mx, x, dims, str, box, what)) != MVOC_OK) return rc;
    }
    {
        const cuuint64_t dims[4] = {(cuuint64_t)K, (cuuint64_t)(2 * (int64_t)F), 1, 1};
        const cuuint64_t str[3] = {(cuuint64_t)K * 2, (cuuint64_t)2 * F * K * 2, (cuuint64_t)2 * F * K * 2};
        const cuuint32_t box[4] = {64, (cuuint32_t)NP, 1, 1};
        if ((rc = gemm::make_map4(&mw, w, dims, str, box, what)) != MVOC_OK) return rc;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t ctas = (int64_t)prm.tiles_w * prm.n_tiles;
    if (NP == 160) return gemm::launch<gemm::Cfg320, gemm::EPI_GEGLU>(mx, mw, prm, ctas, s, what);
    return gemm::launch<gemm::CfgG128, gemm::EPI_GEGLU>(mx, mw, prm, ctas, s, what);
}
