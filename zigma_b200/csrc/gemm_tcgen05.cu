// bf16 GEMM on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), hand written for sm_100a.
//
//     C[M, N] = A[M, K] * B[N, K]^T (+ bias[N])        A, B, C row-major bf16, fp32 accumulation in TMEM
//
// This is the shape of every dense projection on the ZigMa hot path in the token-major layout (DESIGN.md
// section 3): in_proj (mamba_simple.py:290-294), x_proj / dt_proj (selective_scan_interface.py:322-323)
// and out_proj (:365) -- activations (B*L, K) times an nn.Linear weight (N, K).  The reference leaves them
// to cuBLAS behind F.linear / `@`.
//
// Structure (one CTA per SM, persistent over output tiles; 192 threads):
//   warp 0   TMA producer: one elected lane issues cp.async.bulk.tensor loads of the A (128 x 64) and
//            B (BN x 64) k-blocks into a 4-stage shared-memory ring (128-byte swizzle), mbarrier tx-count
//   warp 1   MMA issuer: one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16),
//            4 per k-block, accumulating into one of two TMEM accumulator buffers; tcgen05.commit releases
//            the smem stage / signals the epilogue.  Also owns tcgen05.alloc / dealloc.
//   warps 2-5 epilogue: tcgen05.ld (32 lanes x 32 columns per instruction) -> registers -> (+bias) ->
//            bf16 -> 16-byte global stores, optionally through `out_rowmap` (row scatter of the out_proj
//            result back to raster order, mamba_simple.py:388-394).  Runs one tile behind the MMA warp.
// K, M, N remainders are handled by TMA out-of-bounds zero fill (loads) and predicated stores.
//
// Thread-block clusters (CL = 2 or 4 CTAs along M): at the projection shapes of this model the single-CTA
// kernel is bound by L2 -> SM bandwidth, not by the tensor pipe (round-1 measurement: in_proj 214 us = 2.46 GB
// of operand tiles at ~11.5 TB/s): every 128 x 256 output tile re-reads its 256 x K
// weight tile from L2.  The CTAs of a cluster work on CL vertically adjacent output tiles that share the
// weight tile; each CTA fetches 1/CL of it and TMA-multicasts the slice into the shared memory of all CL CTAs,
// cutting the per-CTA L2 traffic per k-block from 48 KB to 16 + 32/CL KB.  A stage may only be overwritten
// once EVERY CTA of the cluster has consumed it: each CTA's MMA warp signals "stage consumed" with a MULTICAST
// tcgen05.commit to the `free` barrier (arrival count CL) of all CTAs.  (A first version relayed the local commit
// with remote mbarrier arrives from a helper warp: correct but 2.2x / 5.6x slower at CL = 2 / 4, round-1 run 10.)
#include "zg_common.cuh"
#include <cuda.h>
#include <stdlib.h>

namespace zg {

#ifndef ZG_GEMM_2CTA_DEFAULT
#define ZG_GEMM_2CTA_DEFAULT 1
#endif
constexpr int G_BM = 128, G_BK = 64, G_UMMA_K = 16, G_THREADS = 192;   // 6 warps: TMA, MMA, 4 epilogue
// smem ring depth: as many (128 + BN) x 64 bf16 stages as fit next to the 32 KB epilogue staging
__host__ __device__ constexpr int gemm_stages(int BN, bool pair = false) {      // pair: each CTA stages half of the B tile
    const int stage = (G_BM + (pair ? BN / 2 : BN)) * G_BK * 2, avail = 227 * 1024 - 4 * 2 * 32 * 128 - 2048;
    return avail / stage > 8 ? 8 : avail / stage;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, uint16_t mask) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
                     smem_u32(dst)),
                 "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t *bar, uint32_t cta) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(bar)), "r"(cta));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
// ---- cta_group::2 (CTA pair) variants: one MMA of M = 256 over the two CTAs of a cluster; PTX forms as in the vendored
// CUTLASS headers (cute/arch/mma_sm100_umma.hpp SM100_MMA_F16BF16_2x1SM_SS, cutlass/arch/barrier.h umma_arrive_multicast_2x1SM,
// cute/arch/copy_sm100_tma.hpp SM100_TMA_2SM_LOAD_2D, cute/arch/tmem_allocator_sm100.hpp Allocator2Sm) ----
__device__ __forceinline__ uint32_t leader_addr(const void *p) {      // the same shared-memory offset in CTA 0 of the cluster
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(p)), "r"(0u));
    return ra;
}
__device__ __forceinline__ void tma_load_2d_pair(void *dst, const CUtensorMap *map, uint32_t leader_bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    const uint32_t z = 0u;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(z) : "memory");
}
__device__ __forceinline__ void umma_commit_pair_mc(uint64_t *bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster_addr(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
// shared-memory matrix descriptor: K-major operand, 128-byte swizzle, rows of 64 bf16 (128 B), 8-row groups
// 1024 B apart (cute/arch/mma_sm100_desc.hpp: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1
// [46,48), layout_type SWIZZLE_128B=2 [61,64)).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffff) >> 4);
    d |= (uint64_t)1 << 16;                 // LBO (unused for swizzled K-major), canonical value 1
    d |= (uint64_t)(1024 >> 4) << 32;       // SBO
    d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
    return d;
}
// instruction descriptor, kind::f16: D = fp32, A = B = bf16, both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// the same completion signalled to the barrier at this offset in EVERY CTA of `mask` (hardware multicast:
// no software relay on the stage-release path)
__device__ __forceinline__ void umma_commit_mc(uint64_t *bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// the same load without the wait: the caller issues several and waits once (tmem_ld_wait) before it touches the registers
__device__ __forceinline__ void tmem_ld_32x32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
#ifndef ZG_GEMM_EPI_LD2
#define ZG_GEMM_EPI_LD2 0      // 1: both 32-column halves of a 64-column epilogue chunk are loaded from TMEM before ONE wait (timing experiment)
#endif

struct GemmArgs {
    __nv_bfloat16 *C;
    const __nv_bfloat16 *bias;
    const int32_t *out_rowmap;
    int64_t ldc;
    int M, N, K, rows_per_batch;
    int tma_store;      // 1: epilogue goes through smem + TMA store (needs 16-byte aligned C rows)
};

// TWO (needs CL == 2): the two CTAs of the cluster form a CTA PAIR -- one tcgen05.mma.cta_group::2 of M = 256 per k-step, issued by
// CTA 0, reading each CTA's own 128 x 64 A tile and its HALF (BN / 2 rows) of the B tile: half the B bytes per SM and twice the
// math per pipeline stage compared with two independent M = 128 MMAs on a multicast B tile.
template <int BN, int CL, bool TWO = false>
__global__ void __launch_bounds__(G_THREADS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const GemmArgs g) {
    static_assert(!TWO || CL == 2, "a CTA pair is a cluster of two");
    constexpr int BM = G_BM, BK = G_BK, STAGES = gemm_stages(BN, TWO);
    constexpr uint32_t A_BYTES = BM * BK * 2, B_BYTES = (TWO ? BN / 2 : BN) * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr uint32_t TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
    extern __shared__ __align__(1024) unsigned char gsm[];
    unsigned char *tiles = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(gsm) + 1023) & ~(uintptr_t)1023);
    // epilogue staging: per epilogue warp two (32 rows x 128 B) swizzled buffers for the TMA store
    constexpr uint32_t EPI_BYTES = 4 * 2 * 32 * 128;
    unsigned char *epi = tiles + STAGES * STAGE_BYTES;
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(epi + EPI_BYTES);
    uint64_t *empty_bar = full_bar + STAGES;
    uint64_t *tfull_bar = empty_bar + STAGES;     // [2] accumulator ready
    uint64_t *tempty_bar = tfull_bar + 2;         // [2] accumulator drained
    uint64_t *free_bar = tempty_bar + 2;          // [STAGES] stage consumed by every CTA of the cluster (CL > 1)
    uint32_t *tmem_base_slot = reinterpret_cast<uint32_t *>(free_bar + STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = (CL > 1) ? cluster_ctarank() : 0u;
    const int cluster_id = blockIdx.x / CL, num_clusters = gridDim.x / CL;
    const int m_tiles = (g.M + BM - 1) / BM, n_tiles = (g.N + BN - 1) / BN;
    const int sm_tiles = (m_tiles + CL - 1) / CL;             // super tiles of CL vertically adjacent output tiles
    const int num_tiles = sm_tiles * n_tiles;                 // per cluster work list (identical for all its CTAs)
    const int k_blocks = (g.K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); mbar_init(&free_bar[i], CL); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], TWO ? 256 : 128); }   // pair: both CTAs' epilogue threads
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if constexpr (TWO) {      // (the same warp of both CTAs, the same destination offset: cute Allocator2Sm)
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "r"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "r"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CL > 1) cluster_sync_all();          // every CTA's barriers are initialised before any remote arrive / multicast
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_base_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one()) {
            uint32_t it = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                const int m_blk = (tile / n_tiles) * CL + (int)rank, n_blk = tile % n_tiles;
                for (int kb = 0; kb < k_blocks; ++kb, ++it) {
                    const int st = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait((CL > 1 && !TWO) ? &free_bar[st] : &empty_bar[st], ph ^ 1);
                    unsigned char *sa = tiles + st * STAGE_BYTES;
                    if constexpr (TWO) {
                        // both CTAs load their own A rows and their half of the B rows; every byte is counted on CTA 0's barrier,
                        // which CTA 0 arms for the pair
                        if (rank == 0) mbar_expect_tx(&full_bar[st], 2 * STAGE_BYTES);
                        const uint32_t lbar = leader_addr(&full_bar[st]);
                        tma_load_2d_pair(sa, &tmA, lbar, kb * BK, m_blk * BM);
                        tma_load_2d_pair(sa + A_BYTES, &tmB, lbar, kb * BK, n_blk * BN + (int)rank * (BN / 2));
                        continue;
                    }
                    mbar_expect_tx(&full_bar[st], STAGE_BYTES);
                    tma_load_2d(sa, &tmA, &full_bar[st], kb * BK, m_blk * BM);
                    if (CL > 1) {   // my 1/CL slice of the weight tile, multicast to the whole cluster
                        constexpr int SL = BN / CL;
                        tma_load_2d_mc(sa + A_BYTES + rank * (SL * BK * 2), &tmB, &full_bar[st], kb * BK, n_blk * BN + (int)rank * SL,
                                       (uint16_t)((1u << CL) - 1));
                    } else {
                        tma_load_2d(sa + A_BYTES, &tmB, &full_bar[st], kb * BK, n_blk * BN);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (pair: CTA 0 only) =====
        if ((!TWO || rank == 0) && elect_one()) {
            constexpr uint32_t idesc = make_idesc(TWO ? 2 * BM : BM, BN);
            uint32_t it = 0, tcount = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++tcount) {
                const int acc = tcount & 1;
                const uint32_t aph = (tcount >> 1) & 1;
                mbar_wait(&tempty_bar[acc], aph ^ 1);            // epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_c = tmem_base + acc * BN;
                for (int kb = 0; kb < k_blocks; ++kb, ++it) {
                    const int st = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(&full_bar[st], ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sa = smem_u32(tiles + st * STAGE_BYTES);
                    const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sa + A_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / G_UMMA_K; ++k) {    // +32 B per K=16 step inside the 128-byte swizzle atom
                        if constexpr (TWO) umma_f16_pair(tmem_c, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
                        else umma_f16(tmem_c, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
                    }
                    // smem stage free once these MMAs retire; with clusters every CTA of the cluster is told
                    if constexpr (TWO) umma_commit_pair_mc(&empty_bar[st], (uint16_t)3);
                    else if (CL > 1) umma_commit_mc(&free_bar[st], (uint16_t)((1u << CL) - 1));
                    else umma_commit(&empty_bar[st]);
                }
                if constexpr (TWO) umma_commit_pair_mc(&tfull_bar[acc], (uint16_t)3);      // accumulator complete, in both CTAs
                else umma_commit(&tfull_bar[acc]);
            }
        }
    } else if (warp < 6) {
        // ===== epilogue warps 2..5: TMEM lane quadrant = warp % 4 =====
        const int quad = warp & 3;
        uint32_t tcount = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++tcount) {
            const int m_blk = (tile / n_tiles) * CL + (int)rank, n_blk = tile % n_tiles;
            const int acc = tcount & 1;
            const uint32_t aph = (tcount >> 1) & 1;
            mbar_wait(&tfull_bar[acc], aph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int row = m_blk * BM + quad * 32 + lane;
            int64_t drow = row;
            if (g.out_rowmap && row < g.M) {
                const int bidx = row / g.rows_per_batch;
                drow = (int64_t)bidx * g.rows_per_batch + g.out_rowmap[row - bidx * g.rows_per_batch];
            }
            __nv_bfloat16 *crow = g.C + drow * g.ldc;
            constexpr int NFULL = BN / 64;                    // 64-column chunks that go through the TMA store
            const int col_lim = min(g.N, (n_blk + 1) * BN);   // first column that is NOT this tile's
            // direct store of accumulator columns [c0, c0 + 32) of this thread's row (bounds: the matrix and the tile)
            auto direct_store32 = [&](int c0) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + c0), v);
                const int col0 = n_blk * BN + c0;
                if (row >= g.M || col0 >= col_lim) return;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int cj = col0 + j * 8;
                    if (cj + 8 <= col_lim && ((reinterpret_cast<uintptr_t>(crow + cj) & 15) == 0)) {
                        uint32_t o[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            float a = __uint_as_float(v[j * 8 + 2 * i]), b2 = __uint_as_float(v[j * 8 + 2 * i + 1]);
                            if (g.bias) {
                                a += __bfloat162float(g.bias[cj + 2 * i]);
                                b2 += __bfloat162float(g.bias[cj + 2 * i + 1]);
                            }
                            __nv_bfloat162 h = __floats2bfloat162_rn(a, b2);
                            o[i] = *reinterpret_cast<uint32_t *>(&h);
                        }
                        *reinterpret_cast<uint4 *>(crow + cj) = make_uint4(o[0], o[1], o[2], o[3]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            if (cj + i < col_lim) {
                                float a = __uint_as_float(v[j * 8 + i]);
                                if (g.bias) a += __bfloat162float(g.bias[cj + i]);
                                crow[cj + i] = __float2bfloat16_rn(a);
                            }
                    }
                }
            };
            if (g.tma_store) {
                // ---- coalesced path: TMEM -> registers -> bf16 -> swizzled smem -> TMA store (clips M/N edges) ----
                unsigned char *ebuf = epi + (warp - 2) * (2 * 32 * 128);
#pragma unroll 1
                for (int c0 = 0; c0 < NFULL * 64; c0 += 64) {
                    if (n_blk * BN + c0 >= g.N) break;                     // whole chunk outside the matrix
                    unsigned char *buf = ebuf + ((c0 >> 6) & 1) * (32 * 128);
                    // the TMA store that last read this buffer (two chunks ago) must have finished reading it
                    if (lane == 0) {
                        if (NFULL >= 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                        else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // one chunk per tile: same buffer every time
                    }
                    __syncwarp();
#if ZG_GEMM_EPI_LD2
                    uint32_t vv[2][32];
                    tmem_ld_32x32_nowait(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + c0), vv[0]);
                    tmem_ld_32x32_nowait(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + c0 + 32), vv[1]);
                    tmem_ld_wait();
#endif
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
#if ZG_GEMM_EPI_LD2
                        uint32_t (&v)[32] = vv[hh];
#else
                        uint32_t v[32];
                        tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + c0 + hh * 32), v);
#endif
                        const int col0 = n_blk * BN + c0 + hh * 32;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint32_t o[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                float a = __uint_as_float(v[j * 8 + 2 * i]), b2 = __uint_as_float(v[j * 8 + 2 * i + 1]);
                                if (g.bias) {
                                    const int cc = col0 + j * 8 + 2 * i;
                                    if (cc < g.N) a += __bfloat162float(g.bias[cc]);
                                    if (cc + 1 < g.N) b2 += __bfloat162float(g.bias[cc + 1]);
                                }
                                __nv_bfloat162 h = __floats2bfloat162_rn(a, b2);
                                o[i] = *reinterpret_cast<uint32_t *>(&h);
                            }
                            // row = lane, 16-byte chunk index (hh * 4 + j) XOR (row % 8): the 128-byte swizzle the TMA expects
                            const int chunk = (hh * 4 + j) ^ (lane & 7);
                            *reinterpret_cast<uint4 *>(buf + lane * 128 + chunk * 16) = make_uint4(o[0], o[1], o[2], o[3]);
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> async proxy
                    __syncwarp();
                    if (lane == 0) {
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tmC), "r"(smem_u32(buf)),
                                     "r"(n_blk * BN + c0), "r"(m_blk * BM + quad * 32)
                                     : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
                // tile widths that are not a multiple of 64 (80, 160): the last 16 / 32 columns leave by direct stores
                if constexpr (BN % 64 != 0) direct_store32(NFULL * 64);
            } else {
                // ---- direct path (row scatter through out_rowmap, or unaligned C): per-thread 16-byte stores ----
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32) direct_store32(c0);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            if constexpr (TWO) mbar_arrive_cluster_addr(leader_addr(&tempty_bar[acc]));     // the pair's MMA issuer lives in CTA 0
            else mbar_arrive(&tempty_bar[acc]);
        }
    }
    if (warp >= 2 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // my TMA stores have landed
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CL > 1) cluster_sync_all();          // no CTA leaves while a peer may still multicast into it / arrive on its barriers
    if (warp == 1) {
        if constexpr (TWO) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D row-major bf16 matrix (rows x cols, leading dimension ld elements) -> tensor map with a (box_rows x 64) box
static int make_map(CUtensorMap *m, const void *base, int64_t rows, int64_t cols, int64_t ld, int box_rows, bool l2_promote = true) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return zg_set_error("gemm_bf16_tn: cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)G_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, l2_promote ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return zg_set_error("gemm_bf16_tn: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 0;
}

template <int BN, int CL, bool TWO = false> static int launch_gemm(const zg_gemm_params &p, cudaStream_t s) {
    CUtensorMap tmA, tmB, tmC;
    if (int rc = make_map(&tmA, p.A, p.M, p.K, p.lda, G_BM)) return rc;
    if (int rc = make_map(&tmB, p.B, p.N, p.K, p.ldb, BN / CL)) return rc;
    const bool tma_store_ok = !p.out_rowmap && p.ldc % 8 == 0 && (reinterpret_cast<uintptr_t>(p.C) & 15) == 0;
    if (tma_store_ok) {
        if (int rc = make_map(&tmC, p.C, p.M, p.N, p.ldc, 32, false)) return rc;
    } else {
        tmC = tmA;     // unused by the direct-store epilogue
    }
    GemmArgs g{reinterpret_cast<__nv_bfloat16 *>(p.C), reinterpret_cast<const __nv_bfloat16 *>(p.bias), p.out_rowmap, p.ldc, p.M, p.N, p.K,
               p.rows_per_batch > 0 ? p.rows_per_batch : p.M, tma_store_ok ? 1 : 0};
    const int smem = gemm_stages(BN, TWO) * (G_BM * G_BK * 2 + (TWO ? BN / 2 : BN) * G_BK * 2) + 4 * 2 * 32 * 128 + 1024 + 512;
    auto kern = gemm_bf16_tn_kernel<BN, CL, TWO>;
    // per DEVICE (function attributes and SM counts are device state; one process may drive several GPUs)
    static bool attr_dev[64] = {};
    static int sms_dev[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    const int di = dev & 63;
    if (!attr_dev[di]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return zg_set_error("gemm_bf16_tn: cudaFuncSetAttribute(%d): %s", smem, cudaGetErrorString(e));
        attr_dev[di] = true;
    }
    if (!sms_dev[di]) cudaDeviceGetAttribute(&sms_dev[di], cudaDevAttrMultiProcessorCount, dev);
    const int sms = sms_dev[di];
    const int m_tiles = (p.M + G_BM - 1) / G_BM, n_tiles = (p.N + BN - 1) / BN;
    const int work = ((m_tiles + CL - 1) / CL) * n_tiles;          // cluster work items
    int clusters = sms / CL;
    if (work < clusters) clusters = work;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(clusters * CL);
    cfg.blockDim = dim3(G_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, g);
    zg_count_launch();
    if (e != cudaSuccess) return zg_set_error("gemm_bf16_tn: launch failed: %s", cudaGetErrorString(e));
    return zg_check_launch("gemm_bf16_tn");
}

// cluster size: ZG_GEMM_CLUSTER = 1 | 2 | 4 (default 2 when the problem has at least that many row tiles)
static int gemm_cluster_setting() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("ZG_GEMM_CLUSTER");
        v = e ? atoi(e) : 2;
        if (v != 1 && v != 2 && v != 4) v = 2;
    }
    return v;
}

template <int BN> static int launch_gemm_cl(const zg_gemm_params &p, cudaStream_t s) {
    int cl = gemm_cluster_setting();
    const int m_tiles = (p.M + G_BM - 1) / G_BM;
    while (cl > 1 && (m_tiles < cl || (BN / cl) % 8 != 0)) cl >>= 1;     // (a CTA's multicast slice must be whole 8-row swizzle groups)
    static int pair = -1;        // ZG_GEMM_2CTA = 1: the cluster of two runs as a CTA pair (one M = 256 MMA, cta_group::2)
    if (pair < 0) { const char *e = getenv("ZG_GEMM_2CTA"); pair = e ? atoi(e) : ZG_GEMM_2CTA_DEFAULT; }
    if (cl == 4) return launch_gemm<BN, 4>(p, s);
    if (cl == 2) {
        // (each CTA's half of the B tile must be whole 8-row swizzle groups, N % 16 == 0.  Short-K products are memory-bound and
        // lose from the coupling of the pair: dt_proj, K = 40, 60.2 us as a pair vs 41.5 us -- pairs from 4 k-blocks on.)
        if constexpr (BN % 32 == 0) { if (pair && p.K >= 4 * G_BK) return launch_gemm<BN, 2, true>(p, s); }
        return launch_gemm<BN, 2>(p, s);
    }
    return launch_gemm<BN, 1>(p, s);
}

}  // namespace zg

extern "C" int zg_gemm_bf16_tn(const zg_gemm_params *pp, void *stream) {
    ZG_REQUIRE(pp != nullptr, "gemm_bf16_tn: null params");
    const zg_gemm_params &p = *pp;
    ZG_REQUIRE(p.A && p.B && p.C, "gemm_bf16_tn: null tensor pointer");
    ZG_REQUIRE(p.M > 0 && p.N > 0 && p.K > 0, "gemm_bf16_tn: bad shape (%d, %d, %d)", p.M, p.N, p.K);
    ZG_REQUIRE(p.lda % 8 == 0 && p.ldb % 8 == 0 && p.lda >= p.K && p.ldb >= p.K && p.ldc >= p.N, "gemm_bf16_tn: leading dimensions must be multiples of 8 elements");
    ZG_REQUIRE((reinterpret_cast<uintptr_t>(p.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.B) & 15) == 0, "gemm_bf16_tn: A and B must be 16-byte aligned");
    ZG_REQUIRE(!p.out_rowmap || (p.rows_per_batch > 0 && p.M % p.rows_per_batch == 0), "gemm_bf16_tn: out_rowmap needs rows_per_batch dividing M");
    cudaStream_t s = (cudaStream_t)stream;
    // tile width: among {256, 160, 128, 80, 64} the WIDEST whose ragged last column tile wastes <= 5 % of the MMA work, else the
    // one that wastes least (wider tiles move fewer operand bytes per MAC through L2 and re-read A fewer times).  160 exists for
    // N = 640 (out_proj of the D = 640 models: 4 x 160 exactly; 256-wide tiles pad it to 768 = 17 % idle MMAs, round 1), 80 for the
    // x_proj widths 72 / 80 (one pass over A instead of two 64-wide ones).
    static int bn_env = -1;
    if (bn_env < 0) { const char *e = getenv("ZG_GEMM_BN"); bn_env = e ? atoi(e) : 0; }
    int bn = bn_env;
    if (bn != 64 && bn != 80 && bn != 128 && bn != 160 && bn != 256) {
        const int cand[5] = {256, 160, 128, 80, 64};
        auto waste = [&](int b) { return (double)(((p.N + b - 1) / b) * b - p.N) / (double)(((p.N + b - 1) / b) * b); };
        bn = 0;
        for (int i = 0; i < 5 && !bn; ++i)
            if (waste(cand[i]) <= 0.05) bn = cand[i];
        if (!bn) {
            bn = 64;
            for (int i = 0; i < 5; ++i)
                if (waste(cand[i]) < waste(bn) - 1e-9) bn = cand[i];
        }
    }
    if (bn == 256) return zg::launch_gemm_cl<256>(p, s);
    if (bn == 160) return zg::launch_gemm_cl<160>(p, s);
    if (bn == 128) return zg::launch_gemm_cl<128>(p, s);
    if (bn == 80) return zg::launch_gemm_cl<80>(p, s);
    return zg::launch_gemm_cl<64>(p, s);
}
