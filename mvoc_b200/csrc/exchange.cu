// Frame-shard <-> pixel-shard exchange of the frame-parallel partition as one-sided puts over NVLink peer memory
// (SURVEY §8b `mvoc_exchange_*`, §8e).  The temporal operators of the UNet (TemporalConvLayer, the temporal
// transformers; i2vgen-xl/pnp_utils.py:1042-1057, :170-220) need every frame of a pixel, everything else every
// pixel of a frame; the reference runs on one GPU and expresses the same re-layout as the
// [(b t) c h w] <-> [(b h w) t c] permutes at :189, :207-213, :1044-1046, :1055-1057.
//
// Each rank owns an ARENA: one cudaMalloc'ed region whose CUDA-IPC handle every peer opens, so that a kernel on
// rank r can store straight into rank d's memory through NVLink / NVSwitch.  All ranks allocate from their arenas in
// lockstep (same sizes, same order), so a destination buffer has the SAME offset on every rank (a symmetric heap).
//
//   put kernel   reads the local shard once and writes every 16-byte vector directly at its final position in the
//                destination rank's layout (the pack copy, the transfer and the unpack copy of an NCCL all-to-all
//                collapse into one pass); the last CTA to finish publishes this rank's epoch in every peer's flag
//                row with a system-scope release store
//   wait kernel  one warp spins (system-scope acquire loads) until every source rank has published the epoch
//                (or, with MVOC_EXCHANGE_WAIT_FUSED, the put kernel's last CTA does the same after publishing:
//                one launch per exchange instead of two)
// Epochs live in device memory (one counter per exchange site), so a captured CUDA graph replays correctly.
#include "common.cuh"

namespace mvoc {
namespace exch {

constexpr int MAX_RANKS = 16;
constexpr int MAX_SITES = 1024;
// arena header: [0, 64 KB): flags[site][src rank] (uint32) written by peers; [64 KB, 68 KB): this rank's own epoch
// counters per site; [68 KB, 72 KB): CTA arrival counters per site; payload from 128 KB on.
constexpr int64_t HDR_FLAGS = 0, HDR_EPOCH = 65536, HDR_ARRIVE = 65536 + 4096, HDR_BYTES = 131072;

constexpr int SITE_MASK = MVOC_EXCHANGE_WAIT_FUSED - 1;   // the site argument may carry MVOC_EXCHANGE_WAIT_FUSED

struct Peers {
    void* base[MAX_RANKS];
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// After this CTA's peer stores: fence, count the CTA in; the last one bumps this rank's epoch for the site and
// publishes it in flags[site][rank] of every peer.
// With `wait_here` the same thread then waits until every peer has published the site in THIS rank's flag row: the
// put and the wait of an exchange are one launch (all of this rank's puts are issued by then, so nobody can
// dead-lock on a CTA that is still moving data).
__device__ __forceinline__ void publish(const Peers& peers, int rank, int world, int site, bool wait_here) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        char* mine = reinterpret_cast<char*>(peers.base[rank]);
        unsigned* arrive = reinterpret_cast<unsigned*>(mine + HDR_ARRIVE) + site;
        const unsigned n = atomicAdd(arrive, 1u);
        if (n == gridDim.x - 1) {
            *arrive = 0u;
            unsigned* ep = reinterpret_cast<unsigned*>(mine + HDR_EPOCH) + site;
            const unsigned e = *ep + 1u;
            *ep = e;
            __threadfence_system();
            for (int d = 0; d < world; ++d) {
                unsigned* flag = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(peers.base[d]) + HDR_FLAGS) +
                                 (size_t)site * MAX_RANKS + rank;
                st_release_sys(flag, e);
            }
            if (wait_here) {
                const unsigned* row = reinterpret_cast<const unsigned*>(mine + HDR_FLAGS) + (size_t)site * MAX_RANKS;
                for (int d = 0; d < world; ++d) {
                    unsigned spins = 0;
                    while ((int)(ld_acquire_sys(row + d) - e) < 0) {
                        __nanosleep(64);
                        if (++spins > (1u << 25)) {
                            printf("mvoc exchange put+wait: site %d rank-slot %d stuck at %u (want %u)\n", site, d,
                                   ld_acquire_sys(row + d), e);
                            __trap();
                        }
                    }
                }
            }
        }
    }
}

// frame shards -> pixel shards.  Local x [b, tl, S, C] (this rank's tl = T / P frames, all S pixels); destination
// rank d receives pixels [d * sp, (d + 1) * sp) into its buffer [b, T, sp, C] at frames rank * tl ...
__global__ void __launch_bounds__(256) put_pixel_shards_kernel(const Vec16* __restrict__ x, Peers peers, int64_t dst_off,
                                                                int rank, int world, int b, int tl, int64_t S, int vc,
                                                                int site) {
    const int64_t sp = S / world;
    const int64_t row_vecs = (int64_t)vc;                    // 16-byte vectors per token row
    const int64_t total = (int64_t)b * tl * S * row_vecs;
    const int T = tl * world;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = i % row_vecs;
        int64_t r = i / row_vecs;
        const int64_t pix = r % S;
        r /= S;
        const int t = (int)(r % tl);
        const int bb = (int)(r / tl);
        const int d = (int)(pix / sp);
        const int64_t drow = ((int64_t)bb * T + (rank * tl + t)) * sp + (pix - (int64_t)d * sp);
        Vec16* dst = reinterpret_cast<Vec16*>(reinterpret_cast<char*>(peers.base[d]) + dst_off) + drow * row_vecs + v;
        st_stream16(dst, ld_stream16(x + i));
    }
    publish(peers, rank, world, site & SITE_MASK, (site & MVOC_EXCHANGE_WAIT_FUSED) != 0);
}

// pixel shards -> frame shards.  Local y [b, T, sp, C]; destination rank d owns frames [d * tl, (d + 1) * tl) and
// receives this rank's pixels into its buffer [b, tl, S, C] at pixels rank * sp ...
__global__ void __launch_bounds__(256) put_frame_shards_kernel(const Vec16* __restrict__ y, Peers peers, int64_t dst_off,
                                                                int rank, int world, int b, int T, int64_t sp, int vc,
                                                                int site) {
    const int tl = T / world;
    const int64_t S = sp * world;
    const int64_t row_vecs = (int64_t)vc;
    const int64_t total = (int64_t)b * T * sp * row_vecs;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = i % row_vecs;
        int64_t r = i / row_vecs;
        const int64_t pix = r % sp;
        r /= sp;
        const int t = (int)(r % T);
        const int bb = (int)(r / T);
        const int d = t / tl;
        const int64_t drow = ((int64_t)bb * tl + (t - d * tl)) * S + ((int64_t)rank * sp + pix);
        Vec16* dst = reinterpret_cast<Vec16*>(reinterpret_cast<char*>(peers.base[d]) + dst_off) + drow * row_vecs + v;
        st_stream16(dst, ld_stream16(y + i));
    }
    publish(peers, rank, world, site & SITE_MASK, (site & MVOC_EXCHANGE_WAIT_FUSED) != 0);
}

// all-gather of a small buffer (GroupNorm partial statistics of a pixel shard): slot `rank` of every peer's
// [world, n_vec] buffer receives this rank's n_vec vectors.
__global__ void __launch_bounds__(256) put_allgather_kernel(const Vec16* __restrict__ src, Peers peers, int64_t dst_off,
                                                             int rank, int world, int64_t n_vec, int site) {
    const int64_t total = n_vec * world;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int d = (int)(i / n_vec);
        const int64_t j = i - (int64_t)d * n_vec;
        Vec16* dst = reinterpret_cast<Vec16*>(reinterpret_cast<char*>(peers.base[d]) + dst_off) + (int64_t)rank * n_vec + j;
        st_stream16(dst, ld_global16(src + j));
    }
    publish(peers, rank, world, site & SITE_MASK, (site & MVOC_EXCHANGE_WAIT_FUSED) != 0);
}

__global__ void wait_kernel(const char* mine, int world, int site) {
    const int lane = threadIdx.x;
    if (lane < world) {
        const unsigned want = *(reinterpret_cast<const unsigned*>(mine + HDR_EPOCH) + site);   // bumped by my own put
        const unsigned* flag = reinterpret_cast<const unsigned*>(mine + HDR_FLAGS) + (size_t)site * MAX_RANKS + lane;
        unsigned spins = 0;
        while ((int)(ld_acquire_sys(flag) - want) < 0) {
            __nanosleep(128);
            if (++spins > (1u << 25)) {
                printf("mvoc exchange wait: site %d rank-slot %d stuck at %u (want %u)\n", site, lane,
                       ld_acquire_sys(flag), want);
                __trap();
            }
        }
    }
}

static int check_common(const char* what, void* const* peer_bases, int rank, int world, int site_arg) {
    const int site = site_arg & SITE_MASK;
    MVOC_REQUIRE(peer_bases != nullptr, MVOC_ERR_INVALID_ARG, "%s: null peer table", what);
    MVOC_REQUIRE(world >= 1 && world <= MAX_RANKS && rank >= 0 && rank < world, MVOC_ERR_INVALID_ARG,
                 "%s: rank %d of %d (at most %d ranks)", what, rank, world, MAX_RANKS);
    MVOC_REQUIRE(site >= 0 && site < MAX_SITES, MVOC_ERR_INVALID_ARG, "%s: site %d out of range [0, %d)", what, site,
                 MAX_SITES);
    for (int d = 0; d < world; ++d)
        MVOC_REQUIRE(peer_bases[d] != nullptr, MVOC_ERR_INVALID_ARG, "%s: peer %d has no arena", what, d);
    return MVOC_OK;
}

static int grid_for_vecs(int64_t vecs) {
    int64_t want = (vecs + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 8;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    return (int)want;
}

}  // namespace exch
}  // namespace mvoc

using namespace mvoc;

extern "C" int64_t mvoc_exchange_header_bytes(void) { return exch::HDR_BYTES; }
extern "C" int mvoc_exchange_max_sites(void) { return exch::MAX_SITES; }

extern "C" int mvoc_exchange_arena_create(int64_t bytes, void** base, void* ipc_handle64) {
    const char* what = "mvoc_exchange_arena_create";
    MVOC_REQUIRE(base && ipc_handle64 && bytes > exch::HDR_BYTES, MVOC_ERR_INVALID_ARG, "%s: bad arguments", what);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, (size_t)bytes);
    MVOC_REQUIRE(e == cudaSuccess, MVOC_ERR_CUDA, "%s: cudaMalloc(%lld): %s", what, (long long)bytes, cudaGetErrorString(e));
    e = cudaMemset(p, 0, (size_t)exch::HDR_BYTES);
    MVOC_REQUIRE(e == cudaSuccess, MVOC_ERR_CUDA, "%s: cudaMemset: %s", what, cudaGetErrorString(e));
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        set_error("%s: cudaIpcGetMemHandle: %s", what, cudaGetErrorString(e));
        return MVOC_ERR_CUDA;
    }
    memcpy(ipc_handle64, &h, 64);
    *base = p;
    cudaDeviceSynchronize();
    return MVOC_OK;
}

extern "C" int mvoc_exchange_arena_open(const void* ipc_handle64, void** peer_base) {
    const char* what = "mvoc_exchange_arena_open";
    MVOC_REQUIRE(ipc_handle64 && peer_base, MVOC_ERR_INVALID_ARG, "%s: null pointer", what);
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle64, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    MVOC_REQUIRE(e == cudaSuccess, MVOC_ERR_CUDA, "%s: cudaIpcOpenMemHandle: %s", what, cudaGetErrorString(e));
    *peer_base = p;
    return MVOC_OK;
}

extern "C" int mvoc_exchange_arena_close(void* peer_base) {
    if (peer_base) cudaIpcCloseMemHandle(peer_base);
    return MVOC_OK;
}

extern "C" int mvoc_exchange_arena_destroy(void* base) {
    if (base) cudaFree(base);
    return MVOC_OK;
}

extern "C" int mvoc_exchange_to_pixel_shards(const void* x, void* const* peer_bases, int64_t dst_offset, int rank, int world,
                                             int b, int frames_local, int64_t S, int C, int dtype, int site, void* stream) {
    const char* what = "mvoc_exchange_to_pixel_shards";
    int rc = exch::check_common(what, peer_bases, rank, world, site);
    if (rc != MVOC_OK) return rc;
    MVOC_REQUIRE(x && b > 0 && frames_local > 0 && S > 0 && C > 0, MVOC_ERR_INVALID_ARG, "%s: empty problem", what);
    MVOC_REQUIRE(dtype == MVOC_BF16 || dtype == MVOC_F16, MVOC_ERR_UNSUPPORTED, "%s: dtype %d (16-bit only)", what, dtype);
    MVOC_REQUIRE(C % 8 == 0 && S % world == 0, MVOC_ERR_UNSUPPORTED,
                 "%s: C=%d must be a multiple of 8 and S=%lld divisible by the %d ranks", what, C, (long long)S, world);
    MVOC_REQUIRE((uintptr_t)x % 16 == 0 && dst_offset >= exch::HDR_BYTES && dst_offset % 16 == 0, MVOC_ERR_INVALID_ARG,
                 "%s: x must be 16-byte aligned and the destination offset inside the arena payload", what);
    exch::Peers peers{};
    for (int d = 0; d < world; ++d) peers.base[d] = peer_bases[d];
    const int64_t vecs = (int64_t)b * frames_local * S * (C / 8);
    exch::put_pixel_shards_kernel<<<exch::grid_for_vecs(vecs), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const Vec16*>(x), peers, dst_offset, rank, world, b, frames_local, S, C / 8, site);
    return check_launch(what);
}

extern "C" int mvoc_exchange_to_frame_shards(const void* y, void* const* peer_bases, int64_t dst_offset, int rank, int world,
                                             int b, int frames_total, int64_t S_local, int C, int dtype, int site,
                                             void* stream) {
    const char* what = "mvoc_exchange_to_frame_shards";
    int rc = exch::check_common(what, peer_bases, rank, world, site);
    if (rc != MVOC_OK) return rc;
    MVOC_REQUIRE(y && b > 0 && frames_total > 0 && S_local > 0 && C > 0, MVOC_ERR_INVALID_ARG, "%s: empty problem", what);
    MVOC_REQUIRE(dtype == MVOC_BF16 || dtype == MVOC_F16, MVOC_ERR_UNSUPPORTED, "%s: dtype %d (16-bit only)", what, dtype);
    MVOC_REQUIRE(C % 8 == 0 && frames_total % world == 0, MVOC_ERR_UNSUPPORTED,
                 "%s: C=%d must be a multiple of 8 and T=%d divisible by the %d ranks", what, C, frames_total, world);
    MVOC_REQUIRE((uintptr_t)y % 16 == 0 && dst_offset >= exch::HDR_BYTES && dst_offset % 16 == 0, MVOC_ERR_INVALID_ARG,
                 "%s: y must be 16-byte aligned and the destination offset inside the arena payload", what);
    exch::Peers peers{};
    for (int d = 0; d < world; ++d) peers.base[d] = peer_bases[d];
    const int64_t vecs = (int64_t)b * frames_total * S_local * (C / 8);
    exch::put_frame_shards_kernel<<<exch::grid_for_vecs(vecs), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const Vec16*>(y), peers, dst_offset, rank, world, b, frames_total, S_local, C / 8, site);
    return check_launch(what);
}

extern "C" int mvoc_exchange_allgather(const void* src, int64_t bytes, void* const* peer_bases, int64_t dst_offset, int rank,
                                       int world, int site, void* stream) {
    const char* what = "mvoc_exchange_allgather";
    int rc = exch::check_common(what, peer_bases, rank, world, site);
    if (rc != MVOC_OK) return rc;
    MVOC_REQUIRE(src && bytes > 0 && bytes % 16 == 0 && (uintptr_t)src % 16 == 0, MVOC_ERR_INVALID_ARG,
                 "%s: need a 16-byte aligned source of a multiple of 16 bytes", what);
    MVOC_REQUIRE(dst_offset >= exch::HDR_BYTES && dst_offset % 16 == 0, MVOC_ERR_INVALID_ARG,
                 "%s: destination offset outside the arena payload", what);
    exch::Peers peers{};
    for (int d = 0; d < world; ++d) peers.base[d] = peer_bases[d];
    const int64_t n_vec = bytes / 16;
    exch::put_allgather_kernel<<<exch::grid_for_vecs(n_vec * world), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const Vec16*>(src), peers, dst_offset, rank, world, n_vec, site);
    return check_launch(what);
}

extern "C" int mvoc_exchange_wait(const void* my_base, int world, int site, void* stream) {
    const char* what = "mvoc_exchange_wait";
    MVOC_REQUIRE(my_base != nullptr && world >= 1 && world <= exch::MAX_RANKS && site >= 0 && site < exch::MAX_SITES,
                 MVOC_ERR_INVALID_ARG, "%s: bad arguments", what);
    exch::wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(reinterpret_cast<const char*>(my_base), world, site);
    return check_launch(what);
}
