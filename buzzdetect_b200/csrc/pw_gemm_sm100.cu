// Pointwise (1x1) convolution as a tcgen05 GEMM for sm_100a:
//     C[M,N] = relu(A[M,K] * W[N,K]^T + bias[N])         M = patches*H*W, K = Cin, N = Cout
// Reference op: Conv2D 1x1 + FusedBatchNormV3 + Relu of _separable_conv (embedders/yamnet/yamnet.py:62-73); BN is
// folded into W/bias on the host.
//
// Precision: the reference computes in float32.  Tensor cores take fp16 operands, so each float32 operand x is
// carried as two fp16 planes hi = fp16(x), lo = fp16(x - hi).  NSPLIT = 3 issues A_hi*W_hi + A_lo*W_hi + A_hi*W_lo
// into one float32 TMEM accumulator (error ~2^-22 relative, same as float32 storage); NSPLIT = 1 issues only
// A_hi*W_hi (error ~2^-11).  See DESIGN.md "precision".
//
// Structure (one CTA per SM, persistent over 128 x BN output tiles, warp-specialised):
//   warp 0   : TMA producer  -- cp.async.bulk.tensor 2-D loads of 128x64 / BNx64 fp16 boxes, SWIZZLE_128B,
//              STAGES-deep ring guarded by full/empty mbarriers
//   warp 1   : MMA issuer    -- one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16),
//              tcgen05.commit releases smem slots and publishes the accumulator
//   warp 2   : TMEM allocator (2 accumulator stages x BN columns)
//   warps 4-7: epilogue      -- tcgen05.ld 32x32b -> scale, +bias, ReLU -> smem transpose -> coalesced float4 stores
//
// Used by the stand-alone GEMMs of layers 7, 13, 14 (default plan) and by every layer when fusion is switched off.
// This file also holds the first-generation fused kernels (sep_fused_kernel: register-fed stencil producers;
// l12_fused_kernel: layers 1+2 behind CTA-wide barriers), kept selectable through fuse_mask and parity-tested; the
// default plan uses sep_fused_sm100.cu and l12_fused_sm100.cu instead.
#include <cmath>
#include <cuda_fp8.h>

#include "bd_common.cuh"
#include "bd_kernels.cuh"

namespace bd {

namespace {

constexpr int kBM = 128;
constexpr int kBK = 64;                      // 64 fp16 = 128 B = one swizzle row
constexpr int kGemmThreads = 256;
constexpr int kEpiStride = 36;                                   // floats per staged row (32 + 4 pad: conflict-free v4)
constexpr int kEpiBytes = 4 * 32 * kEpiStride * 4;               // 4 epilogue warps x 32 rows
constexpr int kSmemBudget = 227 * 1024 - 2048 - kEpiBytes;

template <int BN, int NSPLIT>
struct GemmCfg {
    static constexpr int kPlanes = NSPLIT == 1 ? 1 : 2;
    static constexpr int kATile = kBM * kBK * 2;                 // 16 KB
    static constexpr int kBTile = BN * kBK * 2;
    static constexpr int kStageBytes = kPlanes * (kATile + kBTile);
    static constexpr int kStagesRaw = kSmemBudget / kStageBytes;
    static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
    static constexpr int kTmemCols = 2 * BN;                     // power of two for BN in {64,128,256}
    static constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + 1024 /*align slack*/ + 256 /*barriers*/;
    static_assert(kStages >= 2, "need at least a double buffer");
    static_assert(kTmemCols >= 32 && kTmemCols <= 512 && (kTmemCols & (kTmemCols - 1)) == 0, "TMEM columns");
};

template <int BN, int NSPLIT>
__global__ void __launch_bounds__(kGemmThreads, 1)
pw_gemm_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
               const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
               const float* __restrict__ bias, float* __restrict__ C, int M, int N, int K, float out_scale) {
    using Cfg = GemmCfg<BN, NSPLIT>;
    constexpr int STAGES = Cfg::kStages;
    extern __shared__ unsigned char smem_raw[];
    // SWIZZLE_128B operands need 1024-byte aligned tiles
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float* epi_stage = reinterpret_cast<float*>(smem + STAGES * Cfg::kStageBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStageBytes + kEpiBytes);
    uint64_t* full_bar = bars;                    // [STAGES]
    uint64_t* empty_bar = bars + STAGES;          // [STAGES]
    uint64_t* tmem_full = bars + 2 * STAGES;      // [2]
    uint64_t* tmem_empty = bars + 2 * STAGES + 2; // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tiles = (M + kBM - 1) / kBM, n_tiles = N / BN;
    const int num_tiles = m_tiles * n_tiles;
    const int num_kb = (K + kBK - 1) / kBK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a_hi);
        tma_prefetch_desc(&map_b_hi);
        if (NSPLIT > 1) {
            tma_prefetch_desc(&map_a_lo);
            tma_prefetch_desc(&map_b_lo);
        }
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 128);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================================================= TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                const int m_blk = t / n_tiles, n_blk = t % n_tiles;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    unsigned char* st = smem + stage * Cfg::kStageBytes;
                    mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
                    // second planes: fp16 lo, or (NSPLIT = 2) the e5m2 correction planes, 128 bytes per k-block
                    constexpr int kc2 = NSPLIT == 2 ? 2 : 1;
                    tma_load_2d(st, &map_a_hi, &full_bar[stage], kb * kBK, m_blk * kBM);
                    if (NSPLIT > 1) tma_load_2d(st + Cfg::kATile, &map_a_lo, &full_bar[stage], kb * kBK * kc2, m_blk * kBM);
                    unsigned char* sb = st + Cfg::kPlanes * Cfg::kATile;
                    tma_load_2d(sb, &map_b_hi, &full_bar[stage], kb * kBK, n_blk * BN);
                    if (NSPLIT > 1) tma_load_2d(sb + Cfg::kBTile, &map_b_lo, &full_bar[stage], kb * kBK * kc2, n_blk * BN);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================================================= MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(kBM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(smem + stage * Cfg::kStageBytes);
                    const uint32_t a_lo = a_hi + Cfg::kATile;
                    const uint32_t b_hi = a_hi + Cfg::kPlanes * Cfg::kATile;
                    const uint32_t b_lo = b_hi + Cfg::kBTile;
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k) {
                        const uint32_t koff = static_cast<uint32_t>(k) * 32u;     // 16 fp16 = 32 bytes along K
                        const uint64_t da_hi = umma_desc_k128(a_hi + koff);
                        const uint64_t db_hi = umma_desc_k128(b_hi + koff);
                        umma_f16_ss(d_tmem, da_hi, db_hi, idesc, (kb | k) != 0 ? 1u : 0u);
                        if (NSPLIT == 3) {
                            const uint64_t da_lo = umma_desc_k128(a_lo + koff);
                            const uint64_t db_lo = umma_desc_k128(b_lo + koff);
                            umma_f16_ss(d_tmem, da_lo, db_hi, idesc, 1u);
                            umma_f16_ss(d_tmem, da_hi, db_lo, idesc, 1u);
                        }
                    }
                    if (NSPLIT == 2) {
                        constexpr uint32_t idesc8 = umma_idesc_e5m2(kBM, BN);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint32_t koff = static_cast<uint32_t>(k) * 32u;  // 32 e5m2 = 32 bytes along K
                            umma_f8_ss(d_tmem, umma_desc_k128(a_lo + koff), umma_desc_k128(b_lo + koff), idesc8, 1u);
                        }
                    }
                    umma_commit(&empty_bar[stage]);          // smem slot reusable once these MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tmem_full[acc]);                // accumulator complete -> epilogue
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ================================================================= epilogue (warps 4..7 <-> TMEM lane quads)
        const int q = warp & 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
            const int m_blk = t / n_tiles, n_blk = t % n_tiles;
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const int row0 = m_blk * kBM + q * 32;                 // first row of this warp's TMEM lane quad
            const int n0 = n_blk * BN;
            float* stg = epi_stage + q * 32 * kEpiStride;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN);
            const int srow = lane >> 3, scol = (lane & 7) * 4;     // store mapping: 8 lanes cover one 128-byte row run
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(taddr + static_cast<uint32_t>(c0), r);
                tmem_ld_wait();
                // thread = row: scale, bias, ReLU, stage the 32 columns of this row in shared memory
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + n0 + c0 + j));
                    float4 o;
                    o.x = fmaxf(fmaf(__uint_as_float(r[j + 0]), out_scale, bv.x), 0.f);
                    o.y = fmaxf(fmaf(__uint_as_float(r[j + 1]), out_scale, bv.y), 0.f);
                    o.z = fmaxf(fmaf(__uint_as_float(r[j + 2]), out_scale, bv.z), 0.f);
                    o.w = fmaxf(fmaf(__uint_as_float(r[j + 3]), out_scale, bv.w), 0.f);
                    *reinterpret_cast<float4*>(stg + lane * kEpiStride + j) = o;
                }
                __syncwarp();
                // transposed read-back: each store instruction writes 4 rows x 128 contiguous bytes
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int rl = i * 4 + srow;
                    const float4 o = *reinterpret_cast<const float4*>(stg + rl * kEpiStride + scol);
                    const int grow = row0 + rl;
                    if (grow < M) *reinterpret_cast<float4*>(C + static_cast<long long>(grow) * N + n0 + c0 + scol) = o;
                }
                __syncwarp();
            }
            tc_fence_before();
            mbar_arrive(&tmem_empty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------ fused depthwise + pointwise
// One separable block in one kernel:
//     C[M,N] = relu( relu(DW3x3(X) + b_dw)[M,K] * W[N,K]^T * out_scale + b_pw )
// The depthwise output never goes to HBM: eight producer warps compute it on the CUDA cores straight into the
// SWIZZLE_128B K-major shared-memory tile that tcgen05.mma reads as its A operand (hi/lo fp16 planes), while warp 0
// streams the weight tile with TMA.  Everything else (TMEM double buffering, epilogue) is the plain GEMM above.
//
// Producer mapping per 128-pixel x 64-channel k-block: thread = (channel quad, strip of 4 consecutive output pixels
// along W); 256 threads x 2 strips.  The 3 x ((4-1)*STRIDE+3) input window and the 9 tap vectors sit in registers.
// Element (row r, channel c) of the tile lives at  (r/8)*1024 + (r%8)*128 + (((c/8) ^ (r%8)) * 16) + (c%8)*2  bytes.
constexpr int kFusedThreads = 512;
constexpr int kProducerWarps = 8;

template <int BN, int NSPLIT, int STRIDE>
__global__ void __launch_bounds__(kFusedThreads, 1)
sep_fused_kernel(const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                 const float* __restrict__ X, const float* __restrict__ dw_w, const float* __restrict__ dw_b,
                 const float* __restrict__ bias, float* __restrict__ C, int P, int H, int W, int K, int N,
                 float out_scale) {
    using Cfg = GemmCfg<BN, NSPLIT>;
    constexpr int STAGES = Cfg::kStages;
    constexpr int PB = STRIDE == 1 ? 1 : 0;
    constexpr int R = 4;
    constexpr int NC = (R - 1) * STRIDE + 3;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float* epi_stage = reinterpret_cast<float*>(smem + STAGES * Cfg::kStageBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStageBytes + kEpiBytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + STAGES;
    uint64_t* tmem_full = bars + 2 * STAGES;
    uint64_t* tmem_empty = bars + 2 * STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Ho = H / STRIDE, Wo = W / STRIDE;
    const int M = P * Ho * Wo;
    const int m_tiles = (M + kBM - 1) / kBM, n_tiles = N / BN;
    const int num_tiles = m_tiles * n_tiles;
    const int num_kb = (K + kBK - 1) / kBK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_b_hi);
        if (NSPLIT > 1) tma_prefetch_desc(&map_b_lo);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 1 + kProducerWarps);     // TMA expect_tx arrive + one arrive per producer warp
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 128);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================================================= TMA producer (weights only)
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                const int n_blk = t % n_tiles;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    unsigned char* sb = smem + stage * Cfg::kStageBytes + Cfg::kPlanes * Cfg::kATile;
                    mbar_arrive_expect_tx(&full_bar[stage], Cfg::kPlanes * Cfg::kBTile);
                    tma_load_2d(sb, &map_b_hi, &full_bar[stage], kb * kBK, n_blk * BN);
                    if (NSPLIT > 1) tma_load_2d(sb + Cfg::kBTile, &map_b_lo, &full_bar[stage], kb * kBK, n_blk * BN);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================================================= MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(kBM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(smem + stage * Cfg::kStageBytes);
                    const uint32_t a_lo = a_hi + Cfg::kATile;
                    const uint32_t b_hi = a_hi + Cfg::kPlanes * Cfg::kATile;
                    const uint32_t b_lo = b_hi + Cfg::kBTile;
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k) {
                        const uint32_t koff = static_cast<uint32_t>(k) * 32u;
                        const uint64_t da_hi = umma_desc_k128(a_hi + koff);
                        const uint64_t db_hi = umma_desc_k128(b_hi + koff);
                        umma_f16_ss(d_tmem, da_hi, db_hi, idesc, (kb | k) != 0 ? 1u : 0u);
                        if (NSPLIT > 1) {
                            const uint64_t da_lo = umma_desc_k128(a_lo + koff);
                            const uint64_t db_lo = umma_desc_k128(b_lo + koff);
                            umma_f16_ss(d_tmem, da_lo, db_hi, idesc, 1u);
                            umma_f16_ss(d_tmem, da_hi, db_lo, idesc, 1u);
                        }
                    }
                    umma_commit(&empty_bar[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tmem_full[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 8) {
        // ================================================================= depthwise producers (CUDA cores -> smem A)
        const int pt = threadIdx.x - 256;
        const int quad = pt & 15;                       // 4 channels of the 64-channel k-block
        const int wo_bits = 31 - __clz(Wo);
        const float inv_ho = 1.0f / static_cast<float>(Ho);
        const int strip0 = pt >> 4;                     // strips strip0 and strip0 + 16 (4 output pixels each)
        int stage = 0;
        uint32_t phase = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
            const int m_blk = t / n_tiles;
            for (int kb = 0; kb < num_kb; ++kb) {
                const int c = kb * kBK + quad * 4;
                const bool c_ok = c < K;
                float4 kk[9];
                float4 bdw = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c_ok) {
#pragma unroll
                    for (int i = 0; i < 9; ++i) kk[i] = __ldg(reinterpret_cast<const float4*>(dw_w + i * K + c));
                    bdw = __ldg(reinterpret_cast<const float4*>(dw_b + c));
                } else {
#pragma unroll
                    for (int i = 0; i < 9; ++i) kk[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                mbar_wait(&empty_bar[stage], phase ^ 1);
                unsigned char* a_hi = smem + stage * Cfg::kStageBytes;
                unsigned char* a_lo = a_hi + Cfg::kATile;
#pragma unroll 1
                for (int s2 = 0; s2 < 2; ++s2) {
                    const int strip = strip0 + 16 * s2;
                    const int m = m_blk * kBM + strip * R;
                    float4 acc[R];
#pragma unroll
                    for (int r = 0; r < R; ++r) acc[r] = bdw;
                    const bool live = c_ok && m < M;            // M % 4 == 0: a strip is all valid or all padding
                    if (live) {
                        // Wo is a power of two; tq < 2^23 so the float reciprocal division by Ho is exact
                        const int ow0 = m & (Wo - 1);
                        const int tq = m >> wo_bits;
                        const int pq = static_cast<int>((static_cast<float>(tq) + 0.5f) * inv_ho);
                        const int oh = tq - pq * Ho;
                        const long long p = pq;
                        const float* inp = X + p * H * W * K + c;
#pragma unroll
                        for (int kh = 0; kh < 3; ++kh) {
                            const int ih = oh * STRIDE + kh - PB;
                            if (ih < 0 || ih >= H) continue;
                            const float* rowp = inp + static_cast<long long>(ih) * W * K;
                            float4 v[NC];
#pragma unroll
                            for (int j = 0; j < NC; ++j) {
                                const int iw = ow0 * STRIDE - PB + j;
                                v[j] = (iw >= 0 && iw < W)
                                           ? __ldg(reinterpret_cast<const float4*>(rowp + static_cast<long long>(iw) * K))
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
                            }
#pragma unroll
                            for (int r = 0; r < R; ++r) {
#pragma unroll
                                for (int kw = 0; kw < 3; ++kw) {
                                    const float4 x = v[r * STRIDE + kw];
                                    const float4 w4 = kk[kh * 3 + kw];
                                    acc[r].x = fmaf(x.x, w4.x, acc[r].x);
                                    acc[r].y = fmaf(x.y, w4.y, acc[r].y);
                                    acc[r].z = fmaf(x.z, w4.z, acc[r].z);
                                    acc[r].w = fmaf(x.w, w4.w, acc[r].w);
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        float4 a = acc[r];
                        if (live) {
                            a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f);
                        } else {
                            a = make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                        const int row = strip * R + r;
                        const uint32_t off = static_cast<uint32_t>((row >> 3) * 1024 + (row & 7) * 128 +
                                                                   ((((quad >> 1) ^ (row & 7))) << 4) + ((quad & 1) << 3));
                        __half2 hp[2] = {half2_sat(a.x, a.y), half2_sat(a.z, a.w)};
                        const __half h0 = __low2half(hp[0]), h1 = __high2half(hp[0]);
                        const __half h2 = __low2half(hp[1]), h3 = __high2half(hp[1]);
                        *reinterpret_cast<uint2*>(a_hi + off) = *reinterpret_cast<uint2*>(hp);
                        if (NSPLIT > 1) {
                            __half2 lp[2] = {half2_sat(a.x - __half2float(h0), a.y - __half2float(h1)),
                                             half2_sat(a.z - __half2float(h2), a.w - __half2float(h3))};
                            *reinterpret_cast<uint2*>(a_lo + off) = *reinterpret_cast<uint2*>(lp);
                        }
                    }
                }
                fence_proxy_async_smem();            // generic-proxy smem writes -> visible to the tensor-core proxy
                __syncwarp();
                if (lane == 0) mbar_arrive(&full_bar[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ================================================================= epilogue (identical to pw_gemm_kernel)
        const int q = warp & 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
            const int m_blk = t / n_tiles, n_blk = t % n_tiles;
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const int row0 = m_blk * kBM + q * 32;
            const int n0 = n_blk * BN;
            float* stg = epi_stage + q * 32 * kEpiStride;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN);
            const int srow = lane >> 3, scol = (lane & 7) * 4;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(taddr + static_cast<uint32_t>(c0), r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + n0 + c0 + j));
                    float4 o;
                    o.x = fmaxf(fmaf(__uint_as_float(r[j + 0]), out_scale, bv.x), 0.f);
                    o.y = fmaxf(fmaf(__uint_as_float(r[j + 1]), out_scale, bv.y), 0.f);
                    o.z = fmaxf(fmaf(__uint_as_float(r[j + 2]), out_scale, bv.z), 0.f);
                    o.w = fmaxf(fmaf(__uint_as_float(r[j + 3]), out_scale, bv.w), 0.f);
                    *reinterpret_cast<float4*>(stg + lane * kEpiStride + j) = o;
                }
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int rl = i * 4 + srow;
                    const float4 o = *reinterpret_cast<const float4*>(stg + rl * kEpiStride + scol);
                    const int grow = row0 + rl;
                    if (grow < M) *reinterpret_cast<float4*>(C + static_cast<long long>(grow) * N + n0 + c0 + scol) = o;
                }
                __syncwarp();
            }
            tc_fence_before();
            mbar_arrive(&tmem_empty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------ layers 1 + 2 in one kernel
// conv 3x3/2 (1->32) -> depthwise 3x3 (32) -> pointwise 32->64, all with folded BN + ReLU, per 128-pixel tile
// (4 image rows x 32 columns of one patch).  Reads 3.4 KB of log-mel, writes 32 KB of layer-2 output; the layer-1
// activation (196 KB/patch) and the depthwise output (2 x 98 KB/patch of fp16 planes) never leave the SM:
//   1. stage 13 log-mel rows in smem          2. layer-1 tile 6 x 34 x 32 (1-pixel halo, zero outside the image)
//   3. depthwise from smem -> hi/lo fp16 straight into the SWIZZLE_128B A tile        4. one thread issues the
//   tcgen05 MMAs (K = 32: two k-steps x 3 products) against the resident weight tile (TMA, loaded once per CTA)
//   5. all 16 warps drain TMEM (32 rows x 16 columns each) through a smem transpose into coalesced stores.
// No warp specialisation: two such CTAs share an SM and overlap each other's phases.
constexpr int kL12Threads = 512;
constexpr int kL12Rows = 4;                                   // image rows per tile (x 32 columns = 128 pixels)
constexpr int kL12TileH = kL12Rows + 2, kL12TileW = 34;
constexpr int kL12LmRows = 2 * kL12TileH + 1, kL12LmStride = 66;
constexpr int kL12EpiStride = 20;                             // floats per staged row (16 + 4 pad)
constexpr int kL12ABytes = 2 * kBM * kBK * 2;                 // A hi + lo
constexpr int kL12BBytes = 2 * 64 * kBK * 2;                  // W hi + lo (64 x 64 box, columns 32..63 are OOB zeros)
constexpr int kL12ScratchBytes = 16 * 32 * kL12EpiStride * 4; // epilogue staging; aliases the layer-1 tile + log-mel rows
static_assert(kL12ScratchBytes >= (kL12TileH * kL12TileW * 32 + kL12LmRows * kL12LmStride) * 4, "scratch too small");
constexpr int kL12SmemBytes = kL12ABytes + kL12BBytes + kL12ScratchBytes + 2 * (9 * 32 + 32) * 4 + 64 + 1024;

template <int NSPLIT>
__global__ void __launch_bounds__(kL12Threads, 2)
l12_fused_kernel(const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                 const float* __restrict__ logmel, int hop_frames, int P, const float* __restrict__ w1,
                 const float* __restrict__ b1, const float* __restrict__ dw_w, const float* __restrict__ dw_b,
                 const float* __restrict__ bias, float* __restrict__ C, float out_scale) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* a_hi = smem;
    unsigned char* a_lo = smem + kBM * kBK * 2;
    unsigned char* b_hi = smem + kL12ABytes;
    unsigned char* b_lo = b_hi + 64 * kBK * 2;
    float* scratch = reinterpret_cast<float*>(smem + kL12ABytes + kL12BBytes);
    float* c1 = scratch;                                       // [6][34][32]
    float* lm = c1 + kL12TileH * kL12TileW * 32;               // [13][66]
    float* sw = reinterpret_cast<float*>(smem + kL12ABytes + kL12BBytes + kL12ScratchBytes);   // [9][32]
    float* sb = sw + 9 * 32;                                   // [32]
    float* skw = sb + 32;                                      // depthwise taps [9][32]
    float* skb = skw + 9 * 32;                                 // depthwise bias [32]
    uint64_t* b_bar = reinterpret_cast<uint64_t*>(skb + 32);
    uint64_t* mma_bar = b_bar + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cg = tid & 7;
    for (int i = tid; i < 9 * 32; i += kL12Threads) { sw[i] = w1[i]; skw[i] = dw_w[i]; }
    if (tid < 32) { sb[tid] = b1[tid]; skb[tid] = dw_b[tid]; }
    if (tid == 0) {
        tma_prefetch_desc(&map_b_hi);
        mbar_init(b_bar, 1);
        mbar_init(mma_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<64>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (tid == 0) {                                            // weights: once per CTA
        mbar_arrive_expect_tx(b_bar, NSPLIT > 1 ? kL12BBytes : kL12BBytes / 2);
        tma_load_2d(b_hi, &map_b_hi, b_bar, 0, 0);
        if (NSPLIT > 1) tma_load_2d(b_lo, &map_b_lo, b_bar, 0, 0);
    }
    constexpr uint32_t idesc = umma_idesc_f16(kBM, 64);
    uint32_t mma_phase = 0;
    bool b_ready = false;

    const unsigned tiles_per_patch = 48 / kL12Rows;           // 12
    const unsigned n_tiles = static_cast<unsigned>(P) * tiles_per_patch;
    for (unsigned t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const long long p = t / tiles_per_patch;
        const int r0 = static_cast<int>(t - static_cast<unsigned>(p) * tiles_per_patch) * kL12Rows;
        const float* in = logmel + p * hop_frames * kMel;
        // ---- 1. log-mel rows 2*(r0-1) .. +12, columns 0..64
        const int lm_row0 = 2 * (r0 - 1);
        for (int i = tid; i < kL12LmRows * 65; i += kL12Threads) {
            const int rr = i / 65, cc = i - rr * 65;
            const int gr = lm_row0 + rr;
            lm[rr * kL12LmStride + cc] = (gr >= 0 && gr < kPatchFrames && cc < kMel) ? __ldg(in + gr * kMel + cc) : 0.f;
        }
        __syncthreads();
        // ---- 2. layer-1 tile with halo
        {
            float4 w1r[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) w1r[i] = *reinterpret_cast<const float4*>(sw + i * 32 + cg * 4);
            const float4 b1r = *reinterpret_cast<const float4*>(sb + cg * 4);
            for (int i = tid; i < kL12TileH * kL12TileW * 8; i += kL12Threads) {
                const int px = i >> 3;
                const int tr = px / kL12TileW, tc = px - tr * kL12TileW;
                const int ir = r0 - 1 + tr, ic = tc - 1;
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ir >= 0 && ir < 48 && ic >= 0 && ic < 32) {
                    a = b1r;
                    const float* l0 = lm + (2 * tr) * kL12LmStride + 2 * ic;
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
                        for (int kw = 0; kw < 3; ++kw) {
                            const float v = l0[kh * kL12LmStride + kw];
                            const float4 wk = w1r[kh * 3 + kw];
                            a.x = fmaf(v, wk.x, a.x);
                            a.y = fmaf(v, wk.y, a.y);
                            a.z = fmaf(v, wk.z, a.z);
                            a.w = fmaf(v, wk.w, a.w);
                        }
                    }
                    a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f);
                }
                *reinterpret_cast<float4*>(c1 + px * 32 + cg * 4) = a;
            }
        }
        __syncthreads();
        // ---- 3. depthwise: thread = (row 0..3, strip of 2 columns 0..15, channel quad) -> A tile rows
        {
            const int ws = (tid >> 3) & 15, orow = tid >> 7;
            const float4 bdw = *reinterpret_cast<const float4*>(skb + cg * 4);
            float4 acc[2] = {bdw, bdw};
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const float* rowp = c1 + ((orow + kh) * kL12TileW + ws * 2) * 32 + cg * 4;
                float4 v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = *reinterpret_cast<const float4*>(rowp + j * 32);
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const float4 w4 = *reinterpret_cast<const float4*>(skw + (kh * 3 + kw) * 32 + cg * 4);
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const float4 x = v[r + kw];
                        acc[r].x = fmaf(x.x, w4.x, acc[r].x);
                        acc[r].y = fmaf(x.y, w4.y, acc[r].y);
                        acc[r].z = fmaf(x.z, w4.z, acc[r].z);
                        acc[r].w = fmaf(x.w, w4.w, acc[r].w);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                float4 a = acc[r];
                a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f);
                const int row = orow * 32 + ws * 2 + r;
                const uint32_t off = static_cast<uint32_t>((row >> 3) * 1024 + (row & 7) * 128 +
                                                           ((((cg >> 1) ^ (row & 7))) << 4) + ((cg & 1) << 3));
                __half2 hp[2] = {half2_sat(a.x, a.y), half2_sat(a.z, a.w)};
                const __half h0 = __low2half(hp[0]), h1 = __high2half(hp[0]);
                const __half h2 = __low2half(hp[1]), h3 = __high2half(hp[1]);
                *reinterpret_cast<uint2*>(a_hi + off) = *reinterpret_cast<uint2*>(hp);
                if (NSPLIT > 1) {
                    __half2 lp[2] = {half2_sat(a.x - __half2float(h0), a.y - __half2float(h1)),
                                     half2_sat(a.z - __half2float(h2), a.w - __half2float(h3))};
                    *reinterpret_cast<uint2*>(a_lo + off) = *reinterpret_cast<uint2*>(lp);
                }
            }
        }
        fence_proxy_async_smem();
        __syncthreads();
        // ---- 4. MMA: K = 32 -> k-steps 0 and 1 only
        if (tid == 0) {
            if (!b_ready) { mbar_wait(b_bar, 0); }
            tc_fence_after();
            const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const uint32_t koff = static_cast<uint32_t>(k) * 32u;
                umma_f16_ss(tmem_base, umma_desc_k128(ah + koff), umma_desc_k128(bh + koff), idesc, k != 0 ? 1u : 0u);
                if (NSPLIT > 1) {
                    umma_f16_ss(tmem_base, umma_desc_k128(al + koff), umma_desc_k128(bh + koff), idesc, 1u);
                    umma_f16_ss(tmem_base, umma_desc_k128(ah + koff), umma_desc_k128(bl + koff), idesc, 1u);
                }
            }
            umma_commit(mma_bar);
        }
        b_ready = true;
        mbar_wait(mma_bar, mma_phase);
        mma_phase ^= 1;
        tc_fence_after();
        // ---- 5. epilogue: warp = (TMEM lane quad q, 16-column group)
        {
            const int q = warp & 3, colgrp = warp >> 2;
            float* stg = scratch + warp * 32 * kL12EpiStride;     // aliases c1/lm: every thread is past phase 3
            uint32_t r[16];
            tmem_ld_32x32b_x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(colgrp * 16), r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
                const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + colgrp * 16 + j));
                float4 o;
                o.x = fmaxf(fmaf(__uint_as_float(r[j + 0]), out_scale, bv.x), 0.f);
                o.y = fmaxf(fmaf(__uint_as_float(r[j + 1]), out_scale, bv.y), 0.f);
                o.z = fmaxf(fmaf(__uint_as_float(r[j + 2]), out_scale, bv.z), 0.f);
                o.w = fmaxf(fmaf(__uint_as_float(r[j + 3]), out_scale, bv.w), 0.f);
                *reinterpret_cast<float4*>(stg + lane * kL12EpiStride + j) = o;
            }
            __syncwarp();
            const long long m0 = static_cast<long long>(t) * kBM + q * 32;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int rl = i * 8 + (lane >> 2);
                const float4 o = *reinterpret_cast<const float4*>(stg + rl * kL12EpiStride + (lane & 3) * 4);
                *reinterpret_cast<float4*>(C + (m0 + rl) * 64 + colgrp * 16 + (lane & 3) * 4) = o;
            }
        }
        tc_fence_before();
        __syncthreads();                                       // TMEM drained, scratch free for the next tile
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<64>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------ host side

typedef TensorMapEncodeFn EncodeTiledFn;
EncodeTiledFn get_encode_fn() { return tensor_map_encode_fn(); }

bool encode_2d_f16(CUtensorMap* map, const void* ptr, int rows, int cols, int box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (fn == nullptr) return false;
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t gstride[1] = {static_cast<cuuint64_t>(cols) * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(kBK), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

bool encode_2d_u8(CUtensorMap* map, const void* ptr, int rows, int row_bytes, int box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (fn == nullptr || row_bytes % 128 != 0) return false;
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(row_bytes), static_cast<cuuint64_t>(rows)};
    cuuint64_t gstride[1] = {static_cast<cuuint64_t>(row_bytes)};
    cuuint32_t box[2] = {128u, static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

template <int BN, int NSPLIT>
cudaError_t set_attr() {
    return cudaFuncSetAttribute(pw_gemm_kernel<BN, NSPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                GemmCfg<BN, NSPLIT>::kSmemBytes);
}

template <int BN, int NSPLIT>
cudaError_t launch_t(const PwGemmPlan& p, const float* bias, float* C, int M, int num_sms, cudaStream_t stream) {
    const int tiles = ((M + kBM - 1) / kBM) * (p.N / BN);
    const int grid = tiles < num_sms ? tiles : num_sms;
    pw_gemm_kernel<BN, NSPLIT><<<grid, kGemmThreads, GemmCfg<BN, NSPLIT>::kSmemBytes, stream>>>(
        p.a_hi, NSPLIT == 2 ? p.a_c8 : p.a_lo, p.b_hi, NSPLIT == 2 ? p.b_c8 : p.b_lo, bias, C, M, p.N, p.K, p.out_scale);
    return cudaGetLastError();
}

template <int BN, int NSPLIT, int STRIDE>
cudaError_t set_attr_fused() {
    return cudaFuncSetAttribute(sep_fused_kernel<BN, NSPLIT, STRIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                GemmCfg<BN, NSPLIT>::kSmemBytes);
}

template <int BN, int NSPLIT, int STRIDE>
cudaError_t launch_fused_t(const PwGemmPlan& p, const float* X, const float* dw_w, const float* dw_b, const float* bias,
                           float* C, int P, int H, int W, int num_sms, cudaStream_t stream) {
    const int M = P * (H / STRIDE) * (W / STRIDE);
    const int tiles = ((M + kBM - 1) / kBM) * (p.N / BN);
    const int grid = tiles < num_sms ? tiles : num_sms;
    sep_fused_kernel<BN, NSPLIT, STRIDE><<<grid, kFusedThreads, GemmCfg<BN, NSPLIT>::kSmemBytes, stream>>>(
        p.b_hi, p.b_lo, X, dw_w, dw_b, bias, C, P, H, W, p.K, p.N, p.out_scale);
    return cudaGetLastError();
}

}  // namespace

TensorMapEncodeFn tensor_map_encode_fn() {
    static TensorMapEncodeFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess) {
            fn = reinterpret_cast<TensorMapEncodeFn>(p);
        }
    }
    return fn;
}

bool encode_kmajor_f16_map(CUtensorMap* map, const void* ptr, int rows, int cols, int box_rows) {
    return encode_2d_f16(map, ptr, rows, cols, box_rows);
}

bool encode_kmajor_u8_map(CUtensorMap* map, const void* ptr, int rows, int row_bytes, int box_rows) {
    return encode_2d_u8(map, ptr, rows, row_bytes, box_rows);
}

float split_weights_f16f8(const float* w, size_t rows, size_t K, __half* hi, unsigned char* c8) {
    const size_t n = rows * K;
    float mx = 0.f;
    for (size_t i = 0; i < n; ++i) mx = fmaxf(mx, fabsf(w[i]));
    int ex = 0;
    float scale = 1.f;
    if (mx > 0.f && std::isfinite(mx)) {
        std::frexp(mx, &ex);
        scale = std::ldexp(1.0f, 10 - ex);    // mx*scale in [512,1024), as in split_weights_f16
    }
    for (size_t r = 0; r < rows; ++r) {
        for (size_t k = 0; k < K; ++k) {
            const float v = w[r * K + k] * scale;
            const __half h = __float2half_rn(v);
            hi[r * K + k] = h;
            const float hf = __half2float(h);
            unsigned char* row = c8 + r * 2 * K + (k / 64) * 128 + (k % 64);
            row[0] = static_cast<unsigned char>(__nv_cvt_float_to_fp8(hf / 2048.f, __NV_SATFINITE, __NV_E5M2));
            row[64] = static_cast<unsigned char>(__nv_cvt_float_to_fp8(v - hf, __NV_SATFINITE, __NV_E5M2));
        }
    }
    return 1.0f / scale;
}

bool encode_store_map_f32(CUtensorMap* map, float* ptr, long long rows, int cols, int box_rows) {
    TensorMapEncodeFn fn = tensor_map_encode_fn();
    if (fn == nullptr || cols % 32 != 0 || rows < 1 || box_rows < 8 || box_rows % 8) return false;
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t gstride[1] = {static_cast<cuuint64_t>(cols) * 4};
    cuuint32_t box[2] = {32, static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, ptr, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

float split_weights_f16(const float* w, size_t n, __half* hi, __half* lo) {
    float mx = 0.f;
    for (size_t i = 0; i < n; ++i) mx = fmaxf(mx, fabsf(w[i]));
    int ex = 0;
    float scale = 1.f;
    if (mx > 0.f && std::isfinite(mx)) {
        std::frexp(mx, &ex);                  // mx = f * 2^ex, f in [0.5,1)
        scale = std::ldexp(1.0f, 10 - ex);    // mx*scale in [512,1024)
    }
    for (size_t i = 0; i < n; ++i) {
        const float v = w[i] * scale;         // exact (power of two)
        const __half h = __float2half_rn(v);
        hi[i] = h;
        lo[i] = __float2half_rn(v - __half2float(h));
    }
    return 1.0f / scale;
}

cudaError_t pw_gemm_init_device() {
    cudaError_t e;
    if ((e = set_attr<64, 1>()) != cudaSuccess) return e;
    if ((e = set_attr<128, 1>()) != cudaSuccess) return e;
    if ((e = set_attr<256, 1>()) != cudaSuccess) return e;
    if ((e = set_attr<64, 2>()) != cudaSuccess) return e;
    if ((e = set_attr<128, 2>()) != cudaSuccess) return e;
    if ((e = set_attr<256, 2>()) != cudaSuccess) return e;
    if ((e = set_attr<64, 3>()) != cudaSuccess) return e;
    if ((e = set_attr<128, 3>()) != cudaSuccess) return e;
    if ((e = set_attr<256, 3>()) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(l12_fused_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kL12SmemBytes)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(l12_fused_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kL12SmemBytes)) != cudaSuccess) return e;
#define BD_FUSED_ATTR(BN)                                                           \
    if ((e = set_attr_fused<BN, 1, 1>()) != cudaSuccess) return e;                  \
    if ((e = set_attr_fused<BN, 1, 2>()) != cudaSuccess) return e;                  \
    if ((e = set_attr_fused<BN, 3, 1>()) != cudaSuccess) return e;                  \
    if ((e = set_attr_fused<BN, 3, 2>()) != cudaSuccess) return e;
    BD_FUSED_ATTR(64)
    BD_FUSED_ATTR(128)
    BD_FUSED_ATTR(256)
#undef BD_FUSED_ATTR
    return cudaSuccess;
}

cudaError_t launch_l12_fused(const PwGemmPlan& p, const float* logmel, int hop_frames, int P, const float* w1,
                             const float* b1, const float* dw_w, const float* dw_b, const float* bias, float* C,
                             int num_sms, cudaStream_t stream) {
    if (P <= 0) return cudaSuccess;
    if (p.N != 64 || p.K != 32 || p.block_n != 64) return cudaErrorInvalidValue;
    const long long tiles = static_cast<long long>(P) * (48 / kL12Rows);
    if (tiles >= (1LL << 31)) return cudaErrorInvalidValue;
    const int grid = static_cast<int>(tiles < 2LL * num_sms ? tiles : 2LL * num_sms);
    if (p.nsplit == 1)
        l12_fused_kernel<1><<<grid, kL12Threads, kL12SmemBytes, stream>>>(p.b_hi, p.b_lo, logmel, hop_frames, P, w1, b1,
                                                                          dw_w, dw_b, bias, C, p.out_scale);
    else
        l12_fused_kernel<3><<<grid, kL12Threads, kL12SmemBytes, stream>>>(p.b_hi, p.b_lo, logmel, hop_frames, P, w1, b1,
                                                                          dw_w, dw_b, bias, C, p.out_scale);
    return cudaGetLastError();
}

cudaError_t launch_sep_fused(const PwGemmPlan& p, const float* X, const float* dw_w, const float* dw_b,
                             const float* bias, float* C, int P, int H, int W, int stride, int num_sms,
                             cudaStream_t stream) {
    if (P <= 0) return cudaSuccess;
    if ((stride != 1 && stride != 2) || (W / stride) % 4 != 0 || p.K % 4 != 0) return cudaErrorInvalidValue;
    if (((W / stride) & (W / stride - 1)) != 0) return cudaErrorInvalidValue;      // producers index with shifts
#define BD_FUSED(BN, NS)                                                                                         \
    return stride == 1 ? launch_fused_t<BN, NS, 1>(p, X, dw_w, dw_b, bias, C, P, H, W, num_sms, stream)          \
                       : launch_fused_t<BN, NS, 2>(p, X, dw_w, dw_b, bias, C, P, H, W, num_sms, stream)
    if (p.nsplit == 1) {
        if (p.block_n == 64) BD_FUSED(64, 1);
        if (p.block_n == 128) BD_FUSED(128, 1);
        BD_FUSED(256, 1);
    }
    if (p.block_n == 64) BD_FUSED(64, 3);
    if (p.block_n == 128) BD_FUSED(128, 3);
    BD_FUSED(256, 3);
#undef BD_FUSED
}

cudaError_t pw_gemm_make_plan(PwGemmPlan* plan, const __half* a_hi, const __half* a_lo, int M_max, int K,
                              const __half* b_hi, const __half* b_lo, int N, int nsplit, int block_n,
                              float out_scale, const char** err) {
    *err = nullptr;
    if (nsplit != 1 && nsplit != 3 && nsplit != 2) { *err = "nsplit must be 1, 2 or 3"; return cudaErrorInvalidValue; }
    if (nsplit == 2 && (K % 64 != 0 || a_lo == nullptr || b_lo == nullptr)) {
        *err = "the fp16 + fp8 plan needs K % 64 == 0 and both e5m2 planes";
        return cudaErrorInvalidValue;
    }
    if (K % 8 != 0) { *err = "K must be a multiple of 8 (16-byte TMA row stride)"; return cudaErrorInvalidValue; }
    int bn = block_n;
    if (bn <= 0) bn = N % 128 == 0 ? 128 : 64;
    if ((bn != 64 && bn != 128 && bn != 256) || N % bn != 0) { *err = "N must be a multiple of block_n in {64,128,256}"; return cudaErrorInvalidValue; }
    plan->M_max = M_max; plan->N = N; plan->K = K; plan->block_n = bn; plan->nsplit = nsplit; plan->out_scale = out_scale;
    if (a_lo == nullptr) a_lo = a_hi;
    if (b_lo == nullptr) b_lo = b_hi;
    if (!encode_2d_f16(&plan->a_hi, a_hi, M_max, K, kBM) || !encode_2d_f16(&plan->a_lo, a_lo, M_max, K, kBM) ||
        !encode_2d_f16(&plan->b_hi, b_hi, N, K, bn) || !encode_2d_f16(&plan->b_lo, b_lo, N, K, bn) ||
        !encode_2d_f16(&plan->b64_hi, b_hi, N, K, 64) || !encode_2d_f16(&plan->b64_lo, b_lo, N, K, 64)) {
        *err = "cuTensorMapEncodeTiled failed";
        return cudaErrorUnknown;
    }
    if (nsplit == 2 && (!encode_2d_u8(&plan->a_c8, a_lo, M_max, 2 * K, kBM) || !encode_2d_u8(&plan->b_c8, b_lo, N, 2 * K, bn))) {
        *err = "cuTensorMapEncodeTiled failed (e5m2 planes)";
        return cudaErrorUnknown;
    }
    return cudaSuccess;
}

cudaError_t launch_pw_gemm(const PwGemmPlan& p, const float* bias, float* C, int M, int num_sms, cudaStream_t stream) {
    if (M <= 0) return cudaSuccess;
    if (M > p.M_max) return cudaErrorInvalidValue;
    if (p.nsplit == 1) {
        if (p.block_n == 64) return launch_t<64, 1>(p, bias, C, M, num_sms, stream);
        if (p.block_n == 128) return launch_t<128, 1>(p, bias, C, M, num_sms, stream);
        return launch_t<256, 1>(p, bias, C, M, num_sms, stream);
    }
    if (p.nsplit == 2) {
        if (p.block_n == 64) return launch_t<64, 2>(p, bias, C, M, num_sms, stream);
        if (p.block_n == 128) return launch_t<128, 2>(p, bias, C, M, num_sms, stream);
        return launch_t<256, 2>(p, bias, C, M, num_sms, stream);
    }
    if (p.block_n == 64) return launch_t<64, 3>(p, bias, C, M, num_sms, stream);
    if (p.block_n == 128) return launch_t<128, 3>(p, bias, C, M, num_sms, stream);
    return launch_t<256, 3>(p, bias, C, M, num_sms, stream);
}

}  // namespace bd
