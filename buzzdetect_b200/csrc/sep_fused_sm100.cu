// Fused separable block (sm_100a): depthwise 3x3 + BN + ReLU -> pointwise 1x1 + BN + ReLU in ONE kernel, with the
// depthwise INPUT staged through shared memory by TMA and the pointwise accumulators for up to 512 output channels
// resident in TMEM, so the stencil runs ONCE per output tile whatever the layer width.
//     C[M,N] = relu( relu(DW3x3(X) + b_dw)[M,K] * W[N,K]^T * out_scale + b_pw )
// Reference op: _separable_conv, embedders/yamnet/yamnet.py:52-74 (BN folded on the host).
//
// Data flow per CTA (512 threads, one CTA per SM, persistent over output tiles):
//   warp 3      input TMA: 4-D boxes [32 channels, BW columns, BH rows, PB patches] of the float32 NHWC input into a ring
//               of shared-memory tiles, several boxes ahead of the stencil.  A box starts one pixel outside the image
//               for stride 1 (pad 1/1) and ends one pixel outside for stride 2 (TensorFlow SAME on even sizes pads
//               0/1); TMA zero-fills out-of-bounds elements, so the stencil has no boundary tests.
//   warps 8-15  stencil: two groups of four warps take alternate boxes; a thread owns 4 channels of a 4x2 (stride 1)
//               or 2x2 (stride 2) block of output pixels (3 / 6.25 LDS.128 per output vector), and writes hi/lo fp16
//               straight into the SWIZZLE_128B K-major A tile (2 stages of 128 rows x 64 channels).
//   warp 0      weight TMA: 128-row x 64-channel boxes of the hi/lo weight planes into a ring of 32 KB slots.
//   warp 1      tcgen05.mma issuer: per k-block, NACC/128 weight slots x 4 k-steps x (1 | 3) MMAs (M=128, N=128, K=16)
//               into TMEM columns [slot*128, +128) -- the A tile is read NACC/128 times, the stencil computed once.
//               (An MMA costs about (4 KB of A + 32 B x N of B) / 64 B per clock of shared-memory operand fetch, more
//               than its math below N = 256: N = 64 slots ran the tensor pipe at a third of its rate.)
//   warp 2      TMEM allocator: NACC <= 256: two accumulator stages; NACC = 512: one stage = all 512 columns.
//   warps 4-7   epilogue: tcgen05.ld -> scale + bias (constant bank) + ReLU -> swizzled smem block -> TMA store.
//
// An output tile is either TRt image rows x Wo columns of ONE patch, or PT WHOLE patches (6x4 layers: 5 patches =
// 120 of the 128 accumulator rows), so tiles never cut a patch and one box covers a tile's input.
//
// Why these choices (measured, profiles/r1_summary.md): the kernels of this family are bound by the SM's L1/shared
// data pipe (one 128-byte wavefront per clock), not by HBM or FMA issue -- hence block-shaped stencils, bias from the
// constant bank, and TMA stores instead of LDS + STG in the epilogue.
#include <cstdlib>

#include "bd_common.cuh"
#include "bd_kernels.cuh"

namespace bd {

namespace {

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kCB = 32;                                          // channels per input box (128-byte pixel rows)
constexpr int kBNs = 128;                                        // weight rows per ring slot = MMA N
constexpr int kF3Threads = 512;
constexpr int kGroupThreads = 128;                               // stencil group = 4 warps
constexpr int kEpiBufBytes = 32 * 128;                           // one [32 rows x 32 float] swizzled block per epilogue warp
constexpr int kEpiBytes = 4 * kEpiBufBytes;
constexpr int kAStages = 2;
constexpr int kMaxInStages = 6;
constexpr int kMaxBSlots = 8;
constexpr int kSmemMax = 227 * 1024;
constexpr int kBarBytes = 320;
constexpr int kMaxN = 512;
constexpr int kATile = kBM * kBK * 2;                            // one fp16 plane of the A tile, 16 KB
constexpr int kBSlotPlane = kBNs * kBK * 2;                      // 16 KB

struct BiasParam { float v[kMaxN]; };                            // kernel parameter = constant bank (no L1 traffic)

struct F3Params {
    const float* dw_w;
    const float* dw_b;
    int P, K, N, Ho, Wo;
    int PT;                  // patches per tile (1, or whole patches when > 1)
    int TRt;                 // output rows per tile (PT == 1) or Ho
    int PB, TR;              // patches / output rows per input box
    int BW, BH;              // box extent in input pixels (per patch)
    int parts;               // boxes per (tile, 32-channel group)
    int rows_per_part;       // A-tile rows produced from one box
    int valid_rows;          // live accumulator rows per tile
    int tiles_per_patch;     // PT == 1
    int m_tiles;
    int in_stages, in_stride;   // input ring depth, bytes between slots
    int b_slots;
    int prefetch_boxes;      // how far the L2 prefetch cursor runs ahead of the loads (0 = off)
    int nohalo;              // 1: boxes hold whole patches without the padding ring; the stencil masks its border taps
    int H, W;                // input extent per patch
    int items;               // stencil blocks per box (x 8 channel quads)
    int blocks_w, blocks_h;  // stencil blocks per patch-part in W and H
    float out_scale;
};

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const void* map, uint64_t* bar, int32_t c0, int32_t c1,
                                            int32_t c2, int32_t c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

__device__ __forceinline__ void tma_prefetch_4d(const void* map, int32_t c0, int32_t c1, int32_t c2, int32_t c3) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(
                     reinterpret_cast<uint64_t>(map)),
                 "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}

// hi/lo fp16 of 4 channels of one pixel -> the swizzled A tile (row = pixel, 128-byte rows, 16-byte chunks XOR row&7).
// NSPLIT = 2 ("fp16f8"): the second plane holds e5m2 bytes instead, [lo * 2^11 (64 channels) | hi (64 channels)] per row.
template <int NSPLIT>
__device__ __forceinline__ void store_a(uint32_t a_hi, uint32_t a_lo, uint32_t row, uint32_t chunk, uint32_t half8,
                                        uint32_t c8chunk, uint32_t c8off, float4 a) {
    a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f);
    const uint32_t rowoff = row << 7;                  // = (row >> 3) * 1024 + (row & 7) * 128: 8-row groups of 1 KB, 128-byte rows
    const uint32_t off = rowoff + ((chunk ^ (row & 7u)) << 4) + half8;
    __half2 hp[2] = {half2_sat(a.x, a.y), half2_sat(a.z, a.w)};
    const __half h0 = __low2half(hp[0]), h1 = __high2half(hp[0]);
    const __half h2 = __low2half(hp[1]), h3 = __high2half(hp[1]);
    sts64(a_hi + off, reinterpret_cast<uint32_t*>(hp)[0], reinterpret_cast<uint32_t*>(hp)[1]);
    if (NSPLIT == 3) {
        __half2 lp[2] = {half2_sat(a.x - __half2float(h0), a.y - __half2float(h1)),
                         half2_sat(a.z - __half2float(h2), a.w - __half2float(h3))};
        sts64(a_lo + off, reinterpret_cast<uint32_t*>(lp)[0], reinterpret_cast<uint32_t*>(lp)[1]);
    }
    if (NSPLIT == 2) {
        const float f0 = __half2float(h0), f1 = __half2float(h1), f2 = __half2float(h2), f3 = __half2float(h3);
        const uint32_t lo8 = pack_e5m2x4((a.x - f0) * kF8LoScale, (a.y - f1) * kF8LoScale, (a.z - f2) * kF8LoScale,
                                         (a.w - f3) * kF8LoScale);
        const uint32_t hi8 = pack_e5m2x4(f0, f1, f2, f3);
        sts32u(a_lo + rowoff + ((c8chunk ^ (row & 7u)) << 4) + c8off, lo8);
        sts32u(a_lo + rowoff + (((c8chunk + 4u) ^ (row & 7u)) << 4) + c8off, hi8);
    }
}

__device__ __forceinline__ void fma4(float4& acc, const float4& x, const float4& w) {
    acc.x = fmaf(x.x, w.x, acc.x);
    acc.y = fmaf(x.y, w.y, acc.y);
    acc.z = fmaf(x.z, w.z, acc.z);
    acc.w = fmaf(x.w, w.w, acc.w);
}

template <int NSPLIT, int STRIDE, int NACC, bool NOHALO, bool CTA2>
__global__ void __launch_bounds__(kF3Threads, 1)
sep_fused3_kernel(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_b_hi,
                  const __grid_constant__ CUtensorMap map_b_lo, const __grid_constant__ CUtensorMap map_c,
                  const __grid_constant__ CUtensorMap map_c8, const __grid_constant__ BiasParam biasp,
                  const F3Params prm) {
    constexpr int kPlanes = NSPLIT == 1 ? 1 : 2;
    constexpr int kAStageBytes = kPlanes * kATile;
    constexpr int kBSlotBytes = kPlanes * kBSlotPlane;
    constexpr int kAccStages = NACC == 512 ? 1 : 2;
    constexpr int kTmemCols = NACC == 128 ? 256 : 512;
    constexpr int kMmaN = CTA2 ? 2 * kBNs : kBNs;      // CTA pair: each CTA holds 128 of the MMA's 256 weight rows
    constexpr int kSlotsPerKb = NACC / kMmaN;
    constexpr int BWc = STRIDE == 1 ? 4 : 2;            // stencil block: BWc columns x 2 rows of output pixels
    constexpr int NCW = (BWc - 1) * STRIDE + 3;         // input columns per block row
    constexpr int NRW = STRIDE + 3;                     // input rows per block (2 output rows)
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* a_base = smem;                                            // [2][planes][16 KB]
    unsigned char* epi_base = a_base + kAStages * kAStageBytes;              // [4 warps][4 KB], 1024-byte aligned
    unsigned char* b_base = epi_base + kEpiBytes;                            // [b_slots][planes][8 KB]
    unsigned char* in_ring = b_base + prm.b_slots * kBSlotBytes;             // [in_stages][in_stride]
    uint64_t* bars = reinterpret_cast<uint64_t*>(in_ring + prm.in_stages * prm.in_stride);
    uint64_t* a_full = bars;                            // [2]
    uint64_t* a_empty = bars + 2;                       // [2]
    uint64_t* b_full = bars + 4;                        // [8]
    uint64_t* b_empty = bars + 12;                      // [8]
    uint64_t* in_full = bars + 20;                      // [6]
    uint64_t* in_empty = bars + 26;                     // [6]
    uint64_t* tmem_full = bars + 32;                    // [2]
    uint64_t* tmem_empty = bars + 34;                   // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 36);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int K = prm.K, Ho = prm.Ho, Wo = prm.Wo;
    const int n_groups = prm.N / NACC;
    // a pass = (output tile, group of NACC output channels); a CTA pair walks pairs of adjacent tiles together (the odd
    // CTA of the last pair may get a tile past the end: its loads are zero-filled and its stores clipped by TMA)
    const int cta_rank = CTA2 ? static_cast<int>(blockIdx.x & 1) : 0;
    const bool leader = cta_rank == 0;
    const int num_pass = (CTA2 ? (prm.m_tiles + 1) / 2 : prm.m_tiles) * n_groups;
    const int pass0 = CTA2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
    const int pass_step = CTA2 ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
    auto tile_of = [&](int ps) { return CTA2 ? 2 * (ps / n_groups) + cta_rank : ps / n_groups; };
    const int num_kb = K / kBK;
    const int parts = prm.parts;
    const int in_stages = prm.in_stages;
    const int b_slots = prm.b_slots;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_b_hi);
        if (NSPLIT > 1) tma_prefetch_desc(&map_b_lo);
    }
    if (warp == 3 && lane == 0) {
        tma_prefetch_desc(&map_in);
        tma_prefetch_desc(&map_c);
        tma_prefetch_desc(&map_c8);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kAStages; ++i) {
            mbar_init(&a_full[i], CTA2 ? 16 : 8);       // every stencil warp (of both CTAs) arrives once per k-block
            mbar_init(&a_empty[i], 1);
        }
        for (int i = 0; i < kMaxBSlots; ++i) {
            mbar_init(&b_full[i], 1);
            mbar_init(&b_empty[i], 1);
        }
        for (int i = 0; i < kMaxInStages; ++i) {
            mbar_init(&in_full[i], 1);
            mbar_init(&in_empty[i], 4);                 // the four warps of the consuming stencil group
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], CTA2 ? 256 : 128);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        if (CTA2) tmem_alloc_pair<kTmemCols>(tmem_slot);
        else tmem_alloc<kTmemCols>(tmem_slot);
    }
    if (prm.valid_rows < kBM) {
        // rows the stencil never writes: zero both A stages once (their accumulator rows are discarded, but keep
        // uninitialised bit patterns away from the tensor pipe)
        uint4* a = reinterpret_cast<uint4*>(a_base);
        for (int i = threadIdx.x; i < kAStages * kAStageBytes / 16; i += kF3Threads) a[i] = make_uint4(0, 0, 0, 0);
        fence_proxy_async_smem();
    }
    tc_fence_before();
    if (CTA2) cluster_sync_all();                       // the peer's barriers are initialised before anyone signals them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================================================= weight TMA
        if (lane == 0) {
            int bs = 0;
            uint32_t bphase = 0;
            for (int ps = pass0; ps < num_pass; ps += pass_step) {
                const int n0 = (ps % n_groups) * NACC;
                for (int kb = 0; kb < num_kb; ++kb) {
#pragma unroll 1
                    for (int nb = 0; nb < kSlotsPerKb; ++nb) {
                        mbar_wait(&b_empty[bs], bphase ^ 1);
                        unsigned char* dst = b_base + bs * kBSlotBytes;
                        if (CTA2) {
                            // both CTAs load their half of the 256-row tile; completion counts on the leader's barrier
                            const int row = n0 + nb * kMmaN + cta_rank * kBNs;
                            if (leader) mbar_arrive_expect_tx(&b_full[bs], 2 * kBSlotBytes);
                            tma_load_2d_pair(dst, &map_b_hi, &b_full[bs], kb * kBK, row);
                            if (NSPLIT > 1) tma_load_2d_pair(dst + kBSlotPlane, &map_b_lo, &b_full[bs], kb * kBK * (NSPLIT == 2 ? 2 : 1), row);
                            if (++bs == b_slots) { bs = 0; bphase ^= 1; }
                            continue;
                        }
                        mbar_arrive_expect_tx(&b_full[bs], kBSlotBytes);
                        tma_load_2d(dst, &map_b_hi, &b_full[bs], kb * kBK, n0 + nb * kBNs);
                        // second plane: the fp16 lo weights, or (NSPLIT = 2) the e5m2 plane [w_hi 2^-11 | w_lo], 128 bytes per k-block
                        if (NSPLIT > 1) tma_load_2d(dst + kBSlotPlane, &map_b_lo, &b_full[bs], kb * kBK * (NSPLIT == 2 ? 2 : 1), n0 + nb * kBNs);
                        if (++bs == b_slots) { bs = 0; bphase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 3) {
        // ================================================================= input TMA (float32 NHWC boxes)
        if (lane == 0) {
            const uint32_t box_bytes = static_cast<uint32_t>(kCB * prm.BW * prm.BH * prm.PB * 4);
            // Two cursors over the same box sequence: `pf` runs kPrefetchBoxes ahead and only asks L2 for the box
            // (cp.async.bulk.prefetch.tensor), `cur` issues the real loads.  The ring holds 2-4 boxes (shared memory is
            // spent on the GEMM operands), which covers an L2 hit but not a DRAM round trip; with the prefetch the ring
            // only ever waits on L2.
            struct Cursor { int ps, kb, sb, part; };
            auto advance = [&](Cursor& c) {
                if (++c.part == parts) {
                    c.part = 0;
                    if (++c.sb == kBK / kCB) {
                        c.sb = 0;
                        if (++c.kb == num_kb) { c.kb = 0; c.ps += pass_step; }
                    }
                }
            };
            auto coords = [&](const Cursor& c, int& c0, int& cw, int& ch, int& cp) {
                const int m_tile = tile_of(c.ps);
                int p0, oh;
                if (prm.PT == 1) {
                    p0 = m_tile / prm.tiles_per_patch;
                    oh = (m_tile - p0 * prm.tiles_per_patch) * prm.TRt + c.part * prm.TR;
                } else {
                    p0 = m_tile * prm.PT + c.part * prm.PB;
                    oh = 0;
                }
                c0 = c.kb * kBK + c.sb * kCB;
                cw = (STRIDE == 1 && !NOHALO) ? -1 : 0;
                ch = NOHALO ? 0 : (STRIDE == 1 ? oh - 1 : 2 * oh);
                cp = p0;
            };
            Cursor cur{pass0, 0, 0, 0}, pf = cur;
            int c0, cw, ch, cp;
            for (int i = 0; i < prm.prefetch_boxes && pf.ps < num_pass; ++i) {
                coords(pf, c0, cw, ch, cp);
                tma_prefetch_4d(&map_in, c0, cw, ch, cp);
                advance(pf);
            }
            int is = 0;
            uint32_t iphase = 0;
            while (cur.ps < num_pass) {
                if (pf.ps < num_pass && prm.prefetch_boxes > 0) {
                    coords(pf, c0, cw, ch, cp);
                    tma_prefetch_4d(&map_in, c0, cw, ch, cp);
                    advance(pf);
                }
                coords(cur, c0, cw, ch, cp);
                mbar_wait(&in_empty[is], iphase ^ 1);
                mbar_arrive_expect_tx(&in_full[is], box_bytes);
                tma_load_4d(in_ring + is * prm.in_stride, &map_in, &in_full[is], c0, cw, ch, cp);
                if (++is == in_stages) { is = 0; iphase ^= 1; }
                advance(cur);
            }
        }
    } else if (warp == 1) {
        // ================================================================= MMA issuer
        // The WHOLE warp walks the loops and waits on the barriers; only the tcgen05 instructions sit under elect_one().
        // With the loop inside `if (lane == 0)` every address and descriptor lives in per-thread registers and each
        // MMA drags a chain of R2UR moves behind it (~120 clk per issue, measured): with 32-clock N=64 MMAs the
        // tensor pipe then idles three quarters of the time.  Warp-uniform control flow keeps them in uniform registers.
        if (leader) {
            constexpr uint32_t idesc = umma_idesc_f16(CTA2 ? 2 * kBM : kBM, kMmaN);
            int stage = 0, bs = 0, acc = 0;
            uint32_t phase = 0, bphase = 0, acc_phase = 0;
            for (int ps = pass0; ps < num_pass; ps += pass_step) {
                if (CTA2) mbar_wait_cluster(&tmem_empty[acc], acc_phase ^ 1);
                else mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_acc = tmem_base + static_cast<uint32_t>(acc * NACC);
                for (int kb = 0; kb < num_kb; ++kb) {
                    if (CTA2) mbar_wait_cluster(&a_full[stage], phase);
                    else mbar_wait(&a_full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(a_base + stage * kAStageBytes);
                    const uint64_t da_hi = umma_desc_k128(a_hi), da_lo = umma_desc_k128(a_hi + kATile);
#pragma unroll 1
                    for (int nb = 0; nb < kSlotsPerKb; ++nb) {
                        if (CTA2) mbar_wait_cluster(&b_full[bs], bphase);
                        else mbar_wait(&b_full[bs], bphase);
                        tc_fence_after();
                        const uint32_t b_hi = smem_u32(b_base + bs * kBSlotBytes);
                        const uint64_t db_hi = umma_desc_k128(b_hi), db_lo = umma_desc_k128(b_hi + kBSlotPlane);
                        const uint32_t d_tmem = d_acc + static_cast<uint32_t>(nb * kMmaN);
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < kBK / 16; ++k) {
                                const uint64_t koff = static_cast<uint64_t>(k) * 2u;     // 16 fp16 = 32 bytes, >> 4
                                const uint32_t accum = (kb | k) != 0 ? 1u : 0u;
                                if (CTA2) {
                                    umma_f16_ss_pair(d_tmem, da_hi + koff, db_hi + koff, idesc, accum);
                                    if (NSPLIT == 3) {
                                        umma_f16_ss_pair(d_tmem, da_lo + koff, db_hi + koff, idesc, 1u);
                                        umma_f16_ss_pair(d_tmem, da_hi + koff, db_lo + koff, idesc, 1u);
                                    }
                                } else {
                                    umma_f16_ss(d_tmem, da_hi + koff, db_hi + koff, idesc, accum);
                                    if (NSPLIT == 3) {
                                        umma_f16_ss(d_tmem, da_lo + koff, db_hi + koff, idesc, 1u);
                                        umma_f16_ss(d_tmem, da_hi + koff, db_lo + koff, idesc, 1u);
                                    }
                                }
                            }
                            if (NSPLIT == 2 && !CTA2) {
                                // both correction products as one e5m2 contraction over the second planes (K = 128 bytes)
                                constexpr uint32_t idesc8 = umma_idesc_e5m2(kBM, kMmaN);
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    umma_f8_ss(d_tmem, da_lo + static_cast<uint64_t>(k) * 2u, db_lo + static_cast<uint64_t>(k) * 2u, idesc8, 1u);
                            }
                            if (CTA2) umma_commit_pair(&b_empty[bs]); else umma_commit(&b_empty[bs]);
                            if (nb == kSlotsPerKb - 1) {
                                if (CTA2) umma_commit_pair(&a_empty[stage]); else umma_commit(&a_empty[stage]);
                                if (kb == num_kb - 1) {
                                    if (CTA2) umma_commit_pair(&tmem_full[acc]); else umma_commit(&tmem_full[acc]);
                                }
                            }
                        }
                        __syncwarp();
                        if (++bs == b_slots) { bs = 0; bphase ^= 1; }
                    }
                    if (++stage == kAStages) { stage = 0; phase ^= 1; }
                }
                if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 8) {
        // ================================================================= stencil (smem boxes -> smem A tile)
        const int g = (warp - 8) >> 2;                  // group 0 / 1: alternate boxes
        const int tg = threadIdx.x - 256 - g * kGroupThreads;
        const int quad = tg & 7;                        // 4 channels of the 32-channel box
        // this thread's (up to two) blocks of a box: input byte offset and first A-tile row, fixed for the kernel
        // (whole-patch boxes carry no padding ring: taps outside the patch are masked instead -- bit i of rmask / bit c
        // of cmask says whether window row i / column c exists)
        uint32_t blk_in[2], blk_row[2], rmask[2], cmask[2];
        bool blk_ok[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int item = tg + k * kGroupThreads;
            blk_ok[k] = item < prm.items;
            const int b = item >> 3;
            const int bcol = b % prm.blocks_w;
            const int b2 = b / prm.blocks_w;
            const int brow = b2 % prm.blocks_h;
            const int pl = b2 / prm.blocks_h;
            const int pad = (NOHALO && STRIDE == 1) ? 1 : 0;
            const int wr0 = 2 * brow * STRIDE - pad, wc0 = bcol * BWc * STRIDE - pad;     // window origin in box pixels
            blk_in[k] = static_cast<uint32_t>((((pl * prm.BH + wr0) * prm.BW + wc0) * kCB + quad * 4) * 4);
            blk_row[k] = static_cast<uint32_t>(pl * Ho * Wo + 2 * brow * Wo + bcol * BWc);
            rmask[k] = cmask[k] = 0xFFFFFFFFu;
            if (NOHALO) {
                rmask[k] = cmask[k] = 0;
                for (int i = 0; i < NRW; ++i) if (wr0 + i >= 0 && wr0 + i < prm.H) rmask[k] |= 1u << i;
                for (int c = 0; c < NCW; ++c) if (wc0 + c >= 0 && wc0 + c < prm.W) cmask[k] |= 1u << c;
            }
        }
        const uint32_t rowpitch = static_cast<uint32_t>(prm.BW * kCB * 4);     // bytes between box rows
        const uint32_t in_ring_u32 = smem_u32(in_ring);
        const uint32_t a_u32 = smem_u32(a_base);
        const uint32_t half8 = static_cast<uint32_t>(quad & 1) << 3;
        int stage = 0, is = g;
        uint32_t phase = 0, iphase = 0;
        if (is >= in_stages) { is -= in_stages; iphase ^= 1; }
        for (int ps = pass0; ps < num_pass; ps += pass_step) {
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait_sleepy(&a_empty[stage], phase ^ 1);
                const uint32_t a_hi = a_u32 + static_cast<uint32_t>(stage * kAStageBytes);
                const uint32_t a_lo = a_hi + kATile;
                int sb_loaded = -1;
                float4 kk[9], bdw;
#pragma unroll 1
                for (int j = g; j < 2 * parts; j += 2) {
                    const int sb = j / parts, part = j - sb * parts;
                    if (sb != sb_loaded) {
                        const int c = kb * kBK + sb * kCB + quad * 4;
#pragma unroll
                        for (int i = 0; i < 9; ++i) kk[i] = __ldg(reinterpret_cast<const float4*>(prm.dw_w + i * K + c));
                        bdw = __ldg(reinterpret_cast<const float4*>(prm.dw_b + c));
                        sb_loaded = sb;
                    }
                    const uint32_t chunk = static_cast<uint32_t>(sb * 4 + (quad >> 1));
                    const uint32_t c8chunk = static_cast<uint32_t>(sb * 2 + (quad >> 2)), c8off = static_cast<uint32_t>(quad & 3) << 2;
                    mbar_wait_sleepy(&in_full[is], iphase);
                    const uint32_t tile = in_ring_u32 + static_cast<uint32_t>(is * prm.in_stride);
                    const uint32_t part_row = static_cast<uint32_t>(part * prm.rows_per_part);
#pragma unroll 1
                    for (int k = 0; k < 2; ++k) {
                        if (!blk_ok[k]) break;
                        const uint32_t base = tile + blk_in[k];
                        float4 acc[2][BWc];
#pragma unroll
                        for (int o = 0; o < 2; ++o)
#pragma unroll
                            for (int r = 0; r < BWc; ++r) acc[o][r] = bdw;
#pragma unroll
                        for (int i = 0; i < NRW; ++i) {
                            float4 v[NCW];
                            const bool r_ok = (rmask[k] >> i) & 1u;
#pragma unroll
                            for (int c = 0; c < NCW; ++c) {
                                const uint32_t addr = base + i * rowpitch + c * (kCB * 4);
                                v[c] = NOHALO ? lds128_pred(addr, r_ok && ((cmask[k] >> c) & 1u)) : lds128(addr);
                            }
#pragma unroll
                            for (int o = 0; o < 2; ++o) {
                                const int kh = i - o * STRIDE;
                                if (kh < 0 || kh > 2) continue;
#pragma unroll
                                for (int r = 0; r < BWc; ++r)
#pragma unroll
                                    for (int kw = 0; kw < 3; ++kw) fma4(acc[o][r], v[r * STRIDE + kw], kk[kh * 3 + kw]);
                            }
                        }
                        const uint32_t row0 = part_row + blk_row[k];
                        // Whole-patch tiles (Wo = 4): the four blocks of a warp start 8 A-tile rows apart, so their
                        // stores would meet in the same banks (row & 7 and the swizzle equal: 4 wavefronts instead of
                        // 2 per STS.64).  Odd 8-lane groups therefore store their second output row first.
                        const bool swap_rows = NOHALO && ((lane >> 3) & 1);
#pragma unroll
                        for (int o = 0; o < 2; ++o)
#pragma unroll
                            for (int r = 0; r < BWc; ++r) {
                                float4 val = acc[o][r];
                                if (NOHALO) {
                                    const float4 alt = acc[o ^ 1][r];
                                    val.x = swap_rows ? alt.x : val.x; val.y = swap_rows ? alt.y : val.y;
                                    val.z = swap_rows ? alt.z : val.z; val.w = swap_rows ? alt.w : val.w;
                                }
                                const uint32_t oo = NOHALO ? (static_cast<uint32_t>(o) ^ (swap_rows ? 1u : 0u)) : static_cast<uint32_t>(o);
                                store_a<NSPLIT>(a_hi, a_lo, row0 + oo * static_cast<uint32_t>(Wo) + static_cast<uint32_t>(r), chunk, half8, c8chunk, c8off, val);
                            }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&in_empty[is]);               // box consumed (release orders the reads)
                    is += 2;
                    if (is >= in_stages) { is -= in_stages; iphase ^= 1; }
                }
                fence_proxy_async_smem();            // generic-proxy smem writes -> visible to the tensor-core proxy
                __syncwarp();
                if (lane == 0) {
                    if (CTA2) mbar_arrive_leader(&a_full[stage]);
                    else mbar_arrive(&a_full[stage]);
                }
                if (++stage == kAStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ================================================================= epilogue
        const int q = warp & 3;
        const uint32_t stg = smem_u32(epi_base) + static_cast<uint32_t>(q * kEpiBufBytes);
        const int live = prm.valid_rows - q * 32;       // rows of this warp's lane quad that exist (<= 0: none)
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int ps = pass0; ps < num_pass; ps += pass_step) {
            const int m_tile = tile_of(ps), n0 = (ps % n_groups) * NACC;
            int row0;                                    // < 2^31: checked by the launcher
            if (prm.PT == 1) {
                const int p = m_tile / prm.tiles_per_patch;
                row0 = (p * Ho + (m_tile - p * prm.tiles_per_patch) * prm.TRt) * Wo + q * 32;
            } else {
                row0 = m_tile * prm.PT * Ho * Wo + q * 32;
            }
            mbar_wait_sleepy(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * NACC);
            if (live > 0) {
#pragma unroll 1
                for (int c0 = 0; c0 < NACC; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(taddr + static_cast<uint32_t>(c0), r);
                    const float* bp = biasp.v + n0 + c0;
                    if (lane == 0) tma_store_wait_read<0>();          // previous block's store has read the buffer
                    __syncwarp();
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 o;
                        o.x = fmaxf(fmaf(__uint_as_float(r[4 * j + 0]), prm.out_scale, bp[4 * j + 0]), 0.f);
                        o.y = fmaxf(fmaf(__uint_as_float(r[4 * j + 1]), prm.out_scale, bp[4 * j + 1]), 0.f);
                        o.z = fmaxf(fmaf(__uint_as_float(r[4 * j + 2]), prm.out_scale, bp[4 * j + 2]), 0.f);
                        o.w = fmaxf(fmaf(__uint_as_float(r[4 * j + 3]), prm.out_scale, bp[4 * j + 3]), 0.f);
                        sts128(epi_swz_addr(stg, lane, j), o);
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        if (live >= 32) {
                            tma_store_2d(&map_c, stg, n0 + c0, row0);
                        } else {                                     // partial quad: whole 8-row groups only
                            for (int r8 = 0; r8 + 8 <= live; r8 += 8)
                                tma_store_2d(&map_c8, stg + static_cast<uint32_t>(r8 * 128), n0 + c0, row0 + r8);
                        }
                        tma_store_commit();
                    }
                }
            }
            tc_fence_before();
            if (CTA2) mbar_arrive_leader(&tmem_empty[acc]);
            else mbar_arrive(&tmem_empty[acc]);
            if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
        }
        if (lane == 0) tma_store_wait_all();
    }

    tc_fence_before();
    if (CTA2) cluster_sync_all();                       // nobody leaves while the pair's MMAs may still read its smem
    else __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        if (CTA2) tmem_dealloc_pair<kTmemCols>(tmem_base);
        else tmem_dealloc<kTmemCols>(tmem_base);
    }
}

bool encode_4d_f32(CUtensorMap* map, const float* ptr, int C, int W, int H, int P, int box_c, int box_w, int box_h,
                   int box_p) {
    TensorMapEncodeFn fn = tensor_map_encode_fn();
    if (fn == nullptr) return false;
    cuuint64_t gdim[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                          static_cast<cuuint64_t>(P)};
    cuuint64_t gstride[3] = {static_cast<cuuint64_t>(C) * 4, static_cast<cuuint64_t>(W) * C * 4,
                             static_cast<cuuint64_t>(H) * W * C * 4};
    cuuint32_t box[4] = {static_cast<cuuint32_t>(box_c), static_cast<cuuint32_t>(box_w), static_cast<cuuint32_t>(box_h),
                         static_cast<cuuint32_t>(box_p)};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(ptr), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

template <int NSPLIT, int STRIDE, int NACC, bool NOHALO>
cudaError_t launch_f3_t(const CUtensorMap& map_in, const CUtensorMap& map_c, const CUtensorMap& map_c8, const BiasParam& bp,
                        const PwGemmPlan& p, const F3Params& prm, int smem_bytes, int grid, cudaStream_t stream) {
    sep_fused3_kernel<NSPLIT, STRIDE, NACC, NOHALO, false><<<grid, kF3Threads, smem_bytes, stream>>>(
        map_in, p.b_hi, NSPLIT == 2 ? p.b_c8 : p.b_lo, map_c, map_c8, bp, prm);
    return cudaGetLastError();
}

// CTA-pair launch: clusters of two CTAs (one TPC), grid = an even number of CTAs
template <int NSPLIT, int STRIDE, int NACC, bool NOHALO>
cudaError_t launch_f3_pair_t(const CUtensorMap& map_in, const CUtensorMap& map_c, const CUtensorMap& map_c8,
                             const BiasParam& bp, const PwGemmPlan& p, const F3Params& prm, int smem_bytes, int grid,
                             cudaStream_t stream) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(grid));
    cfg.blockDim = dim3(kF3Threads);
    cfg.dynamicSmemBytes = static_cast<size_t>(smem_bytes);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, sep_fused3_kernel<NSPLIT, STRIDE, NACC, NOHALO, true>, map_in, p.b_hi, p.b_lo, map_c,
                              map_c8, bp, prm);
}

template <int NSPLIT, int STRIDE>
cudaError_t set_attr_f3() {
    cudaError_t e;
#define BD_F3_ATTR(NACC, NH)                                                                               \
    if ((e = cudaFuncSetAttribute(sep_fused3_kernel<NSPLIT, STRIDE, NACC, NH, false>,                      \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax)) != cudaSuccess)  \
        return e;
    BD_F3_ATTR(128, false) BD_F3_ATTR(256, false) BD_F3_ATTR(512, false) BD_F3_ATTR(512, true) BD_F3_ATTR(256, true)
    if ((e = cudaFuncSetAttribute(sep_fused3_kernel<NSPLIT, STRIDE, 512, true, true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax)) != cudaSuccess)
        return e;
    if ((e = cudaFuncSetAttribute(sep_fused3_kernel<NSPLIT, STRIDE, 256, false, true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax)) != cudaSuccess)
        return e;
#undef BD_F3_ATTR
    return cudaSuccess;
}

// tiling of one layer: fills the geometry fields of F3Params; false if the layer does not fit this kernel
bool plan_geometry(int K, int N, int H, int W, int stride, int box_limit, F3Params* out) {
    if (K % kBK != 0 || N % 128 != 0 || N > kMaxN) return false;
    if (stride != 1 && stride != 2) return false;
    if ((H % stride) || (W % stride)) return false;
    const int Ho = H / stride, Wo = W / stride;
    const int bwc = stride == 1 ? 4 : 2;
    if (Wo < bwc || (Wo & (Wo - 1)) || Wo > kBM || (Ho & 1)) return false;
    F3Params g{};
    g.K = K; g.N = N; g.Ho = Ho; g.Wo = Wo;
    g.BW = (Wo - 1) * stride + 3;
    const int px = Ho * Wo;
    g.H = H; g.W = W;
    if (px >= 96) {
        // one patch per tile, TRt rows at a time; input boxes of TR rows with a one-pixel padding ring: the largest
        // even divisor of TRt whose box stays under `box_limit` (small enough for a 4-deep ring: two stencil groups,
        // each with a box in use and one in flight)
        g.PT = 1; g.PB = 1; g.nohalo = 0;
        g.TRt = Ho < kBM / Wo ? Ho : kBM / Wo;
        if (g.TRt < 2 || Ho % g.TRt || (g.TRt & 1)) return false;
        g.tiles_per_patch = Ho / g.TRt;
        int tr = g.TRt;
        while (tr > 2 && (kCB * g.BW * ((tr - 1) * stride + 3) * 4 > box_limit || g.TRt % tr || (tr & 1))) --tr;
        if (g.TRt % tr || (tr & 1)) return false;
        g.TR = tr;
        g.parts = g.TRt / tr;
        g.rows_per_part = tr * Wo;
        g.valid_rows = g.TRt * Wo;
        g.BH = (g.TR - 1) * stride + 3;
    } else {
        // small patches: PT whole patches per tile, boxes of PB whole patches WITHOUT the padding ring (a 6x4 patch
        // with its ring would be twice the bytes); the stencil masks the taps that fall outside the patch
        g.PT = kBM / px; g.nohalo = 1;
        g.TRt = Ho; g.TR = Ho;
        g.tiles_per_patch = 1;
        g.BW = W;
        g.BH = H;
        const int per_patch = kCB * W * H * 4;
        int pb = g.PT;
        while (pb > 1 && (pb * per_patch > box_limit || g.PT % pb)) --pb;
        g.PB = pb;
        g.parts = g.PT / pb;
        g.rows_per_part = pb * px;
        g.valid_rows = g.PT * px;
    }
    g.blocks_w = Wo / bwc;
    g.blocks_h = g.TR / 2;
    g.items = g.PB * g.blocks_h * g.blocks_w * 8;
    if (g.items > 2 * kGroupThreads || g.valid_rows % 8) return false;
    if (kCB * g.BW * g.BH * g.PB * 4 > 64 * 1024) return false;
    *out = g;
    return true;
}

}  // namespace

cudaError_t sep_fused3_init_device() {
    cudaError_t e;
    if ((e = set_attr_f3<1, 1>()) != cudaSuccess) return e;
    if ((e = set_attr_f3<1, 2>()) != cudaSuccess) return e;
    if ((e = set_attr_f3<3, 1>()) != cudaSuccess) return e;
    if ((e = set_attr_f3<2, 1>()) != cudaSuccess) return e;
    if ((e = set_attr_f3<2, 2>()) != cudaSuccess) return e;
    return set_attr_f3<3, 2>();
}

bool sep_fused3_supported(int K, int N, int H, int W, int stride) {
    F3Params g;
    return plan_geometry(K, N, H, W, stride, 40 * 1024, &g);
}

cudaError_t launch_sep_fused3(const PwGemmPlan& p, const float* X, const float* dw_w, const float* dw_b,
                              const float* bias_host, float* C, int P, int H, int W, int stride, int num_sms,
                              cudaStream_t stream, bool cta_pairs) {
    if (P <= 0) return cudaSuccess;
    if (bias_host == nullptr || p.N % 128 != 0) return cudaErrorInvalidValue;
    static const int nacc_cap = [] { const char* e = getenv("BD_F3_NACC"); return e ? atoi(e) : 512; }();
    int nacc = p.N >= 512 ? 512 : p.N;                   // 128, 256 or 512 accumulator columns per pass
    if (nacc > nacc_cap && (nacc_cap == 128 || nacc_cap == 256) && p.N % nacc_cap == 0) nacc = nacc_cap;
    if (nacc != 128 && nacc != 256 && nacc != 512) return cudaErrorInvalidValue;
    const int planes = p.nsplit == 1 ? 1 : 2;
    const int fixed = 1024 + kAStages * planes * kATile + kEpiBytes + kBarBytes;
    const int b_slot_bytes = planes * kBSlotPlane;
    // weight ring: enough 16 KB slots in flight to cover the L2 round trip at the MMA's consumption rate (the wider
    // the accumulator, the more slots one k-block eats); the rest of shared memory is the input ring
    // Measured (profiles/r1_summary.md): fewer, larger boxes beat a deeper ring of small ones (per-box barrier and
    // tap-weight reload cost), and shared memory left to L1 matters because the tap weights are re-read per box.
    int b_slots = nacc == 512 ? 3 : 2;
    const int ring_bytes = kSmemMax - fixed - b_slots * b_slot_bytes;
    F3Params prm;
    if (!plan_geometry(p.K, p.N, H, W, stride, 40 * 1024, &prm)) return cudaErrorInvalidValue;
    prm.dw_w = dw_w; prm.dw_b = dw_b;
    prm.P = P;
    prm.out_scale = p.out_scale;
    prm.m_tiles = prm.PT == 1 ? P * prm.tiles_per_patch : (P + prm.PT - 1) / prm.PT;
    const int box_bytes = kCB * prm.BW * prm.BH * prm.PB * 4;
    prm.in_stride = (box_bytes + 127) & ~127;
    int in_stages = ring_bytes / prm.in_stride;
    if (in_stages > kMaxInStages) in_stages = kMaxInStages;
    if (in_stages < 2) return cudaErrorInvalidValue;
    prm.b_slots = b_slots;
    prm.in_stages = in_stages;
    {
        // Measured on B200: asking L2 for the boxes ahead of time does not help (layers 3-6 within +-2 %, layer 3
        // slower) -- the ring is not what these kernels wait on.  Kept as an experiment knob, off by default.
        static const int pf_env = [] { const char* e = getenv("BD_F3_PREFETCH"); return e ? atoi(e) : 0; }();
        prm.prefetch_boxes = pf_env;
    }
    const int smem_bytes = fixed + b_slots * b_slot_bytes + in_stages * prm.in_stride;
    const long long passes = static_cast<long long>(prm.m_tiles) * (p.N / nacc);
    const long long m_rows = static_cast<long long>(P) * prm.Ho * prm.Wo;
    if (passes >= (1LL << 31) || m_rows + kBM >= (1LL << 31)) return cudaErrorInvalidValue;
    int grid = static_cast<int>(passes < num_sms ? passes : num_sms);
    // CTA pairs (cta_group::2: one MMA of M = 256, N = 256 per pair, each CTA holding half of the weight tile).
    // Parity-tested, but measured SLOWER than the one-CTA kernel on B200 (layers 8-12: 0.218 vs 0.187 ms per
    // audio-hour, layer 6: 0.305 vs 0.213): the pair couples two stencil / epilogue pipelines behind one issuer and the
    // peer's half of B does not arrive faster than a second local copy would.  Opt-in (BD_FUSE_PAIR / BD_F3_PAIR=1).
    static const int pair_env = [] { const char* e = getenv("BD_F3_PAIR"); return e ? atoi(e) : 0; }();
    const bool pair = (cta_pairs || pair_env != 0) && num_sms >= 2 && nacc >= 256 && passes >= 2 && p.nsplit != 2;
    if (pair) {
        const long long pair_passes = (static_cast<long long>(prm.m_tiles) + 1) / 2 * (p.N / nacc);
        grid = static_cast<int>(2 * pair_passes < num_sms ? 2 * pair_passes : num_sms);
    }
    CUtensorMap map_in, map_c, map_c8;
    if (!encode_4d_f32(&map_in, X, p.K, W, H, P, kCB, prm.BW, prm.BH, prm.PB)) return cudaErrorUnknown;   // box <= tensor in nohalo mode
    if (!encode_store_map_f32(&map_c, C, m_rows, p.N, 32) || !encode_store_map_f32(&map_c8, C, m_rows, p.N, 8))
        return cudaErrorUnknown;
    BiasParam bp;
    for (int i = 0; i < kMaxN; ++i) bp.v[i] = i < p.N ? bias_host[i] : 0.f;
#define BD_F3(NS, S)                                                                                             \
    do {                                                                                                         \
        if (pair && prm.nohalo && nacc == 512)                                                                   \
            return launch_f3_pair_t<NS, S, 512, true>(map_in, map_c, map_c8, bp, p, prm, smem_bytes, grid & ~1, stream); \
        if (pair && !prm.nohalo && nacc == 256)                                                                  \
            return launch_f3_pair_t<NS, S, 256, false>(map_in, map_c, map_c8, bp, p, prm, smem_bytes, grid & ~1, stream); \
        if (prm.nohalo) {                                                                                        \
            if (nacc == 256) return launch_f3_t<NS, S, 256, true>(map_in, map_c, map_c8, bp, p, prm, smem_bytes, grid, stream); \
            if (nacc != 512) return cudaErrorInvalidValue;                                                       \
            return launch_f3_t<NS, S, 512, true>(map_in, map_c, map_c8, bp, p, prm, smem_bytes, grid, stream);    \
        }                                                                                                        \
        if (nacc == 128) return launch_f3_t<NS, S, 128, false>(map_in, map_c, map_c8, bp, p, prm, smem_bytes, grid, stream); \
        if (nacc == 256) return launch_f3_t<NS, S, 256, false>(map_in, map_c, map_c8, bp, p, prm, smem_bytes, grid, stream); \
        return launch_f3_t<NS, S, 512, false>(map_in, map_c, map_c8, bp, p, prm, smem_bytes, grid, stream);      \
    } while (0)
    if (p.nsplit == 1) {
        if (stride == 1) BD_F3(1, 1);
        BD_F3(1, 2);
    }
    if (p.nsplit == 2) {
        if (pair) return cudaErrorInvalidValue;           // the fp16 + fp8 plan is a one-CTA kernel
        if (stride == 1) BD_F3(2, 1);
        BD_F3(2, 2);
    }
    if (stride == 1) BD_F3(3, 1);
    BD_F3(3, 2);
#undef BD_F3
}

}  // namespace bd
