// STAGED FOR ROUND 2 — compiled (sm_100a) but NOT YET RUN ON HARDWARE; built into libmvoc_b200_staged.so only.
//
// attn_tc.cu with every query row split across TWO softmax threads (eight softmax warps per CTA instead of four).
// Why: the final ncu capture of attn_fwd_kernel shows no saturated pipe (tensor 38 %, MUFU 48 %, issue 55 %): with
// one softmax warp per scheduler and CTA, the ~630-instruction dependent chain of a 128-column score row is
// latency-bound (≈ 2970 clocks per pair of KV blocks against ≈ 1280 for the busiest pipe).  Half a row per thread
// halves that chain and doubles the warps the schedulers can interleave.  The two halves of a row agree on the
// running maximum through a 16-bit shared-memory exchange per block: the maximum is rounded UP to bf16 — any
// common stabiliser >= the true row maximum gives the same softmax.  Row sums stay per half until the epilogue.
//
// Reference: F.scaled_dot_product_attention at i2vgen-xl/pnp_utils.py:684-686 and inside diffusers'
// AttnProcessor2_0 (i2vgen-xl/pnp_utils.py:348-385) — non-causal, no mask, no dropout, D = 64.
//
// One CTA = one 128-row query tile of one (batch, head); two CTAs per SM.
//   warps 0-3   softmax of key columns [0, 64) of every block: thread i owns query row i (= TMEM lane i)
//   warps 4-7   softmax of key columns [64, 128): thread 128 + i owns the same row i
//   warp  8     TMA producer: Q once, then a ring of K / V tiles
//   warp  9     MMA issuer (one lane): S = Q K^T, O += P V; owns the TMEM allocation
//   warps 10-11 idle (they complete the third warpgroup for setmaxnreg)
// TMEM columns (fp32): S [0,128)  P [128,192) (bf16 pairs)  O [192,256).
#include <cuda.h>
#include "../common.cuh"
#include "../ptx.cuh"
#include "../../../include/mvoc_b200_staged.h"

namespace mvoc {
namespace attn_split {

constexpr int BM = 128;  // query rows per CTA
constexpr int BN = 128;  // keys per iteration
constexpr int HALF = 64; // keys per softmax thread and iteration
constexpr int HD = 64;   // head dim
constexpr int TILE_BYTES = BM * HD * 2;  // 16 KB: 128 rows x 128 B
constexpr int THREADS = 384;
constexpr uint32_t TMEM_COLS = 256;
constexpr uint32_t COL_S = 0, COL_P = 128, COL_O = 192;
constexpr float RESCALE_LOG2_THRESHOLD = 8.0f;
constexpr int kStages = 3;

struct Smem {
    static constexpr int q_off = 0;
    static constexpr int k_off = TILE_BYTES;
    static constexpr int v_off = k_off + kStages * TILE_BYTES;
    static constexpr int xchg_off = v_off + kStages * TILE_BYTES;   // uint16 [2 halves][128 rows]
    static constexpr int bar_off = xchg_off + 2 * BM * 2;
    // barriers: q_full, s_full, s_free, p_full, pv_done, k_full[], k_empty[], v_full[], v_empty[]
    static constexpr int n_bars = 5 + 4 * kStages;
    static constexpr int tmem_ptr_off = bar_off + n_bars * 8;
    static constexpr int alloc = tmem_ptr_off + 16;
    static_assert(2 * (alloc + 1024) <= 233472, "two CTAs per SM no longer fit in shared memory");
};

struct Params {
    __nv_bfloat16* o;
    int64_t o_sb, o_sn, o_sh;
    int Nq, Nk;
    float scale_log2;
};

// named barrier shared by the two warps (w, w + 4) that own the same 32 rows
__device__ __forceinline__ void pair_sync(int id) {
    asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory");
}

// smallest bf16-representable value >= x, as its upper 16 bits (x finite or -inf)
__device__ __forceinline__ uint32_t ceil_bf16_bits(float x) {
    const uint32_t b = __float_as_uint(x);
    const uint32_t up = (b & 0x80000000u) ? b : b + 0xFFFFu;   // negative: truncation already rounds up
    return up >> 16;
}

template <uint32_t kEmuMask>
__global__ void __launch_bounds__(THREADS, 2)
attn_split_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                  const __grid_constant__ CUtensorMap tm_v, const Params prm) {
    using L = Smem;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t sbase = smem_u32(smem_raw);
    uint8_t* sgen = smem_raw;
    if ((sbase & 1023u) != 0u) {
        if (threadIdx.x == 0) printf("mvoc attn_split_kernel: dynamic smem base 0x%x is not 1024-byte aligned\n", sbase);
        __trap();
    }

    const uint32_t sQ = sbase + L::q_off;
    const uint32_t sK = sbase + L::k_off;
    const uint32_t sV = sbase + L::v_off;
    const uint32_t bars = sbase + L::bar_off;
    const uint32_t b_q_full = bars, b_s_full = bars + 8, b_s_free = bars + 16, b_p_full = bars + 24,
                   b_pv_done = bars + 32;
    const uint32_t b_k_full = bars + 40, b_k_empty = b_k_full + 8 * kStages,
                   b_v_full = b_k_empty + 8 * kStages, b_v_empty = b_v_full + 8 * kStages;
    const uint32_t s_tmem_ptr = sbase + L::tmem_ptr_off;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_blk = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int n_blocks = (prm.Nk + BN - 1) / BN;

    if (warp == 8 && lane == 0) {
        ptx::prefetch_tensormap(&tm_q);
        ptx::prefetch_tensormap(&tm_k);
        ptx::prefetch_tensormap(&tm_v);
        ptx::mbar_init(b_q_full, 1);
        ptx::mbar_init(b_s_full, 1);
        ptx::mbar_init(b_s_free, 256);
        ptx::mbar_init(b_p_full, 256);
        ptx::mbar_init(b_pv_done, 1);
        for (int s = 0; s < kStages; ++s) {
            ptx::mbar_init(b_k_full + 8 * s, 1);
            ptx::mbar_init(b_k_empty + 8 * s, 1);
            ptx::mbar_init(b_v_full + 8 * s, 1);
            ptx::mbar_init(b_v_empty + 8 * s, 1);
        }
        ptx::fence_mbar_init();
    }
    if (warp == 9) {
        ptx::tmem_alloc(s_tmem_ptr, TMEM_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sgen + L::tmem_ptr_off);

    // 2 CTAs/SM leave 80 registers per thread at launch (384 threads); the data-movement warpgroup keeps 32 and
    // the two softmax warpgroups take 104 each: 256*104 + 128*32 = 30720 = 384*80.
    if (warp >= 8) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 32;" ::: "memory");
    if (warp == 8) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            ptx::mbar_expect_tx(b_q_full, TILE_BYTES);
            ptx::tma_load_4d(sQ, &tm_q, b_q_full, 0, h, m_blk * BM, b);
        }
        for (int j = 0; j < n_blocks; ++j) {
            const int s = j % kStages;
            const uint32_t ph = (uint32_t)(j / kStages) & 1u;
            ptx::mbar_wait(b_k_empty + 8 * s, ph ^ 1u, 1);
            if (lane == 0) {
                ptx::mbar_expect_tx(b_k_full + 8 * s, TILE_BYTES);
                ptx::tma_load_4d(sK + s * TILE_BYTES, &tm_k, b_k_full + 8 * s, 0, h, j * BN, b);
            }
            ptx::mbar_wait(b_v_empty + 8 * s, ph ^ 1u, 2);
            if (lane == 0) {
                ptx::mbar_expect_tx(b_v_full + 8 * s, TILE_BYTES);
                ptx::tma_load_4d(sV + s * TILE_BYTES, &tm_v, b_v_full + 8 * s, 0, h, j * BN, b);
            }
            __syncwarp();
        }
    } else if (warp == 9) {
        // ===================== MMA issuer =====================
        constexpr uint32_t IDESC_QK = ptx::idesc_bf16(BM, BN, 0, 0);  // A,B K-major
        constexpr uint32_t IDESC_PV = ptx::idesc_bf16(BM, HD, 0, 1);  // B (=V) MN-major
        const uint32_t tS = tmem + COL_S, tP = tmem + COL_P, tO = tmem + COL_O;
        auto issue_qk = [&](int j) {
            const int s = j % kStages;
            const uint64_t a0 = ptx::smem_desc_sw128(sQ, 16, 1024);
            const uint64_t b0 = ptx::smem_desc_sw128(sK + s * TILE_BYTES, 16, 1024);
#pragma unroll
            for (int ks = 0; ks < HD / 16; ++ks)
                ptx::mma_ss(tS, a0 + (uint64_t)(ks * 2), b0 + (uint64_t)(ks * 2), IDESC_QK, ks > 0);
            ptx::tc_commit(b_k_empty + 8 * s);
            ptx::tc_commit(b_s_full);
        };
        ptx::mbar_wait(b_q_full, 0, 3);
        ptx::mbar_wait(b_k_full, 0, 4);
        ptx::tc_fence_after();
        if (lane == 0) issue_qk(0);
        __syncwarp();
        for (int j = 0; j < n_blocks; ++j) {
            const int s = j % kStages;
            const uint32_t ph = (uint32_t)(j / kStages) & 1u;
            if (j + 1 < n_blocks) {
                const int s1 = (j + 1) % kStages;
                ptx::mbar_wait(b_k_full + 8 * s1, (uint32_t)((j + 1) / kStages) & 1u, 5);
                ptx::mbar_wait(b_s_free, (uint32_t)j & 1u, 6);
                ptx::tc_fence_after();
                if (lane == 0) issue_qk(j + 1);
                __syncwarp();
            }
            ptx::mbar_wait(b_v_full + 8 * s, ph, 7);
            ptx::mbar_wait(b_p_full, (uint32_t)j & 1u, 8);
            ptx::tc_fence_after();
            if (lane == 0) {
                const uint64_t bv = ptx::smem_desc_sw128(sV + s * TILE_BYTES, 16384, 1024);
#pragma unroll
                for (int ks = 0; ks < BN / 16; ++ks) {
                    const uint64_t bd = bv + (uint64_t)((ks * 2048) >> 4);
                    const uint32_t acc = (j > 0 || ks > 0) ? 1u : 0u;
                    ptx::mma_ts(tO, tP + ks * 8, bd, IDESC_PV, acc);
                }
                ptx::tc_commit(b_v_empty + 8 * s);
                ptx::tc_commit(b_pv_done);
            }
            __syncwarp();
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 104;" ::: "memory");
        // ===================== softmax + epilogue (warps 0-7) =====================
        const int half = warp >> 2;                 // which 64 key columns of every block
        const int wq = warp & 3;                    // TMEM lane quarter
        const int row = wq * 32 + lane;             // query row inside the tile == TMEM lane
        const int bar_a = 1 + wq, bar_b = 5 + wq;   // named barriers of the warp pair (wq, wq + 4)
        const uint32_t lane_base = tmem + ((uint32_t)(wq * 32) << 16);
        const uint32_t tS = lane_base + COL_S + half * HALF, tP = lane_base + COL_P + half * (HALF / 2),
                       tO = lane_base + COL_O + half * (HD / 2);
        const uint32_t x_mine = sbase + L::xchg_off + (half * BM + row) * 2;
        const uint32_t x_other = sbase + L::xchg_off + ((half ^ 1) * BM + row) * 2;
        const float sl2 = prm.scale_log2;
        float m_used = -INFINITY, l_sum = 0.0f;
        for (int j = 0; j < n_blocks; ++j) {
            const int valid = min(HALF, max(0, prm.Nk - j * BN - half * HALF));   // real keys among my 64 columns
            ptx::mbar_wait(b_s_full, (uint32_t)j & 1u, 9);
            ptx::tc_fence_after();
            uint32_t r[HALF];
#pragma unroll
            for (int c = 0; c < HALF / 32; ++c)
                ptx::tmem_ld32(tS + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&r[c * 32]));
            ptx::tmem_wait_ld();
            ptx::tc_fence_before();
            ptx::mbar_arrive(b_s_free);
            if (valid < HALF) {  // keys past Nk (zero-filled by TMA) must not take part: exp2(-inf) = 0
#pragma unroll
                for (int i = 0; i < HALF; ++i)
                    if (i >= valid) r[i] = 0xff800000u;
            }
            float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
            for (int i = 0; i < HALF; i += 8) {
                mx0 = ptx::max3(mx0, __uint_as_float(r[i + 0]), __uint_as_float(r[i + 1]));
                mx1 = ptx::max3(mx1, __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
                mx2 = ptx::max3(mx2, __uint_as_float(r[i + 4]), __uint_as_float(r[i + 5]));
                mx3 = ptx::max3(mx3, __uint_as_float(r[i + 6]), __uint_as_float(r[i + 7]));
            }
            // both halves of the row must use ONE stabiliser: exchange the (bf16-rounded-up) half maxima
            const uint32_t mine = ceil_bf16_bits(fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)));
            pair_sync(bar_b);   // the partner has read the previous block's value
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(x_mine), "h"((uint16_t)mine) : "memory");
            pair_sync(bar_a);
            uint16_t other16;
            asm volatile("ld.shared.u16 %0, [%1];" : "=h"(other16) : "r"(x_other) : "memory");
            const float mx = fmaxf(__uint_as_float(mine << 16), __uint_as_float((uint32_t)other16 << 16));
            bool pv_waited = false;
            if (j == 0) {
                m_used = mx;
            } else {
                const float m_new = fmaxf(m_used, mx);
                const bool need = (m_new - m_used) * sl2 > RESCALE_LOG2_THRESHOLD;
                if (__any_sync(0xffffffffu, need)) {   // same rows, same values => same decision in both warps
                    ptx::mbar_wait(b_pv_done, (uint32_t)(j - 1) & 1u, 10);
                    ptx::tc_fence_after();
                    pv_waited = true;
                    const float f = need ? ptx::ex2_approx((m_used - m_new) * sl2) : 1.0f;
                    if (need) m_used = m_new;
                    l_sum *= f;
                    uint32_t o[32];                    // my half of the O row
                    ptx::tmem_ld32(tO, o);
                    ptx::tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
                    ptx::tmem_st32(tO, o);
                    ptx::tmem_wait_st();
                }
            }
            const uint64_t sl2_2 = ptx::pack2(sl2, sl2);
            const uint64_t negm_2 = ptx::pack2(-m_used * sl2, -m_used * sl2);
            uint32_t pk[HALF / 2];
            uint64_t la = ptx::pack2(0.0f, 0.0f), lb = la, lc = la, ld = la;
#pragma unroll
            for (int k = 0; k < HALF / 2; ++k) {
                const uint64_t x2 = ptx::fma2(ptx::pack2(__uint_as_float(r[2 * k]), __uint_as_float(r[2 * k + 1])),
                                              sl2_2, negm_2);
                float p0, p1;
                if ((kEmuMask >> (k & 7)) & 1u) {
                    ptx::ex2_poly2(x2, p0, p1);
                } else {
                    float x0, x1;
                    ptx::unpack2(x2, x0, x1);
                    p0 = ptx::ex2_approx(x0);
                    p1 = ptx::ex2_approx(x1);
                }
                const uint64_t p2 = ptx::pack2(p0, p1);
                if ((k & 3) == 0) la = ptx::add2(la, p2);
                else if ((k & 3) == 1) lb = ptx::add2(lb, p2);
                else if ((k & 3) == 2) lc = ptx::add2(lc, p2);
                else ld = ptx::add2(ld, p2);
                __nv_bfloat162 pb = __floats2bfloat162_rn(p0, p1);
                pk[k] = *reinterpret_cast<uint32_t*>(&pb);
            }
            {
                float s0, s1;
                ptx::unpack2(ptx::add2(ptx::add2(la, lb), ptx::add2(lc, ld)), s0, s1);
                l_sum += s0 + s1;
            }
            if (j > 0 && !pv_waited) {   // the P buffer is free once the previous P V has completed
                ptx::mbar_wait(b_pv_done, (uint32_t)(j - 1) & 1u, 11);
                ptx::tc_fence_after();
            }
            ptx::tmem_st32(tP, pk);
            ptx::tmem_wait_st();
            ptx::tc_fence_before();
            ptx::mbar_arrive(b_p_full);
        }
        // ---- epilogue: row sum of both halves, O / l -> bf16 -> global --------------------------------
        ptx::mbar_wait(b_pv_done, (uint32_t)(n_blocks - 1) & 1u, 12);
        ptx::tc_fence_after();
        // every MMA has completed, the K ring is free: fp32 exchange of the partial row sums through it
        const uint32_t l_mine = sK + (half * BM + row) * 4, l_other = sK + ((half ^ 1) * BM + row) * 4;
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(l_mine), "f"(l_sum) : "memory");
        pair_sync(bar_a);
        float l_partner;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(l_partner) : "r"(l_other) : "memory");
        const float inv_l = 1.0f / (l_sum + l_partner);
        const int q_row = m_blk * BM + row;
        __nv_bfloat16* orow = prm.o + (int64_t)b * prm.o_sb + (int64_t)q_row * prm.o_sn +
                              (int64_t)h * prm.o_sh + half * (HD / 2);
        uint32_t o[32];
        ptx::tmem_ld32(tO, o);
        ptx::tmem_wait_ld();
        if (q_row < prm.Nq) {
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
                float f[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(o[i + e]) * inv_l;
                *reinterpret_cast<Vec16*>(orow + i) = pack8<__nv_bfloat16>(f);
            }
        }
        ptx::tc_fence_before();
    }

    __syncthreads();
    if (warp == 9) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem, TMEM_COLS);
    }
}

// ------------------------------------------------------------------ host ---
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

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

// [B, N, H, 64] bf16 with element strides (sb, sn, sh); box = 128 tokens x 64 of one head.
static int make_map(CUtensorMap* m, const void* base, int B, int H, int N, int64_t sb, int64_t sn,
                    int64_t sh, const char* what) {
    EncodeTiledFn fn = get_encode_fn();
    MVOC_REQUIRE(fn != nullptr, MVOC_ERR_DRIVER, "mvoc_attn_fwd_split: cuTensorMapEncodeTiled unavailable");
    cuuint64_t dims[4] = {(cuuint64_t)HD, (cuuint64_t)H, (cuuint64_t)N, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)sh * 2, (cuuint64_t)sn * 2, (cuuint64_t)sb * 2};
    if (H == 1) strides[0] = 128;
    if (B == 1) strides[2] = (cuuint64_t)sn * 2 * (cuuint64_t)N;
    cuuint32_t box[4] = {(cuuint32_t)HD, 1, (cuuint32_t)BM, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides,
                    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MVOC_REQUIRE(r == CUDA_SUCCESS, MVOC_ERR_DRIVER,
                 "mvoc_attn_fwd_split: cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r);
    return MVOC_OK;
}

template <uint32_t kEmuMask>
static int launch(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv,
                  const Params& prm, int B, int H, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(attn_split_kernel<kEmuMask>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, Smem::alloc);
        MVOC_REQUIRE(e == cudaSuccess, MVOC_ERR_CUDA, "mvoc_attn_fwd_split: cudaFuncSetAttribute: %s",
                     cudaGetErrorString(e));
        attr_set = true;
    }
    dim3 grid((prm.Nq + BM - 1) / BM, H, B);
    attn_split_kernel<kEmuMask><<<grid, THREADS, Smem::alloc, s>>>(mq, mk, mv, prm);
    return check_launch("mvoc_attn_fwd_split");
}

}  // namespace attn_split
}  // namespace mvoc

using namespace mvoc;

extern "C" int mvoc_attn_fwd_split(const void* q, const void* k, const void* v, void* o, int B, int H,
                                   int Nq, int Nk, int D, int64_t q_sb, int64_t q_sn, int64_t q_sh,
                                   int64_t k_sb, int64_t k_sn, int64_t k_sh, int64_t v_sb, int64_t v_sn,
                                   int64_t v_sh, int64_t o_sb, int64_t o_sn, int64_t o_sh, float scale,
                                   int dtype, int variant, void* stream) {
    MVOC_REQUIRE(q && k && v && o, MVOC_ERR_INVALID_ARG, "mvoc_attn_fwd_split: null pointer");
    MVOC_REQUIRE(dtype == MVOC_BF16, MVOC_ERR_UNSUPPORTED, "mvoc_attn_fwd_split: dtype %d unsupported (bf16 only)", dtype);
    MVOC_REQUIRE(D == attn_split::HD, MVOC_ERR_UNSUPPORTED, "mvoc_attn_fwd_split: head_dim %d unsupported (64 only)", D);
    MVOC_REQUIRE(B > 0 && H > 0 && Nq > 0 && Nk > 0, MVOC_ERR_INVALID_ARG,
                 "mvoc_attn_fwd_split: empty problem B=%d H=%d Nq=%d Nk=%d", B, H, Nq, Nk);
    MVOC_REQUIRE(B <= 65535 && H <= 65535, MVOC_ERR_UNSUPPORTED, "mvoc_attn_fwd_split: B=%d / H=%d exceed the grid limits", B, H);
    MVOC_REQUIRE(variant >= 0 && variant <= 2, MVOC_ERR_INVALID_ARG, "mvoc_attn_fwd_split: unknown variant %d", variant);
    const int64_t strides[12] = {q_sb, q_sn, q_sh, k_sb, k_sn, k_sh, v_sb, v_sn, v_sh, o_sb, o_sn, o_sh};
    for (int i = 0; i < 12; ++i)
        MVOC_REQUIRE(strides[i] % 8 == 0 && strides[i] >= 0, MVOC_ERR_UNSUPPORTED,
                     "mvoc_attn_fwd_split: stride #%d = %lld is not a non-negative multiple of 8 elements", i,
                     (long long)strides[i]);
    MVOC_REQUIRE(((uintptr_t)q % 16 == 0) && ((uintptr_t)k % 16 == 0) && ((uintptr_t)v % 16 == 0) &&
                     ((uintptr_t)o % 16 == 0),
                 MVOC_ERR_INVALID_ARG, "mvoc_attn_fwd_split: pointers must be 16-byte aligned");
    CUtensorMap mq, mk, mv;
    int rc;
    if ((rc = attn_split::make_map(&mq, q, B, H, Nq, q_sb, q_sn, q_sh, "q")) != MVOC_OK) return rc;
    if ((rc = attn_split::make_map(&mk, k, B, H, Nk, k_sb, k_sn, k_sh, "k")) != MVOC_OK) return rc;
    if ((rc = attn_split::make_map(&mv, v, B, H, Nk, v_sb, v_sn, v_sh, "v")) != MVOC_OK) return rc;
    attn_split::Params prm;
    prm.o = (__nv_bfloat16*)o;
    prm.o_sb = o_sb;
    prm.o_sn = o_sn;
    prm.o_sh = o_sh;
    prm.Nq = Nq;
    prm.Nk = Nk;
    prm.scale_log2 = scale * 1.4426950408889634f;
    cudaStream_t s = (cudaStream_t)stream;
    // variant 0: 3 of 8 exp2 pairs on the FMA pipe (as the product kernel); 1: all on the MUFU; 2: 4 of 8
    switch (variant) {
        case 1: return attn_split::launch<0x00u>(mq, mk, mv, prm, B, H, s);
        case 2: return attn_split::launch<0xAAu>(mq, mk, mv, prm, B, H, s);
        default: return attn_split::launch<0xA8u>(mq, mk, mv, prm, B, H, s);
    }
}
