// Fused separable block, version 3 (sm_100a): depthwise 3x3 + BN + ReLU -> pointwise 1x1 + BN + ReLU in ONE kernel,
// with the depthwise INPUT staged through shared memory by TMA.
//     C[M,N] = relu( relu(DW3x3(X) + b_dw)[M,K] * W[N,K]^T * out_scale + b_pw )
// Reference op: _separable_conv, embedders/yamnet/yamnet.py:52-74 (BN folded on the host).
//
// Why a third version: sep_fused_kernel (pw_gemm_sm100.cu) feeds its stencil from registers, so every producer thread
// exposes a full DRAM round trip per strip and the kernel is latency bound (profiles/fusion_r1.md).  Here a dedicated
// thread streams 4-D TMA boxes  [32 channels, BW columns, BH rows, 1 patch]  of the float32 NHWC input into a ring of
// shared-memory tiles, several boxes ahead of the stencil warps.  The box starts one pixel outside the image for
// stride 1 (pad 1/1) and ends one pixel outside for stride 2 (TensorFlow SAME on even sizes pads 0/1); TMA zero-fills
// out-of-bounds elements, so the stencil has no boundary tests at all.
//
// Warp roles (512 threads, one CTA per SM, persistent over output tiles):
//   warp 0      weight TMA (SWIZZLE_128B K-major boxes, as in pw_gemm_kernel)
//   warp 1      tcgen05.mma issuer (M=128, N=BN, K=16; 3 MMAs per k-step in the fp16x3 split mode)
//   warp 2      TMEM allocator
//   warp 3      input TMA (float32 boxes -> in_ring)
//   warps 4-7   epilogue: tcgen05.ld -> scale + bias (constant bank) + ReLU -> swizzled smem block -> TMA store
//   warps 8-15  stencil producers: depthwise from in_ring, hi/lo fp16 straight into the swizzled A tile
//
// An output tile is TRt image rows x Wo columns of ONE patch (<= 128 pixels, so tiles never straddle patches and one
// box covers a tile's input); layers whose patch has 96 output pixels run with 96 of the 128 accumulator rows live.
#include "bd_common.cuh"
#include "bd_kernels.cuh"

namespace bd {

namespace {

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kCB = 32;                                          // channels per input box (128-byte pixel rows)
constexpr int kF3Threads = 512;
constexpr int kF3ProdWarps = 8;
constexpr int kF3ProdThreads = kF3ProdWarps * 32;
constexpr int kEpiBufBytes = 32 * 128;                           // one [32 rows x 32 float] swizzled block per epilogue warp
constexpr int kEpiBytes = 4 * kEpiBufBytes;
constexpr int kMaxN = 512;

struct BiasParam { float v[kMaxN]; };                            // kernel parameter = constant bank (no L1 traffic)
constexpr int kABStages = 2;
constexpr int kMaxInStages = 6;
constexpr int kSmemMax = 227 * 1024;
constexpr int kBarBytes = 256;

template <int BN, int NSPLIT>
struct F3Cfg {
    static constexpr int kPlanes = NSPLIT == 1 ? 1 : 2;
    static constexpr int kATile = kBM * kBK * 2;
    static constexpr int kBTile = BN * kBK * 2;
    static constexpr int kStageBytes = kPlanes * (kATile + kBTile);
    static constexpr int kTmemCols = 2 * BN;
};

struct F3Params {
    const float* dw_w;
    const float* dw_b;
    int P, K, N, Ho, Wo;
    int TRt;                 // output rows per tile
    int TR;                  // output rows per input box
    int BW, BH;              // box extent in input pixels
    int tiles_per_patch;
    int in_stages, in_stride;   // ring depth, bytes between ring slots
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

template <int BN, int NSPLIT, int STRIDE, int R>
__global__ void __launch_bounds__(kF3Threads, 1)
sep_fused3_kernel(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_b_hi,
                  const __grid_constant__ CUtensorMap map_b_lo, const __grid_constant__ CUtensorMap map_c,
                  const __grid_constant__ BiasParam biasp, const F3Params prm) {
    using Cfg = F3Cfg<BN, NSPLIT>;
    constexpr int NC = (R - 1) * STRIDE + 3;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* epi_base = smem + kABStages * Cfg::kStageBytes;           // 1024-byte aligned (swizzled store blocks)
    unsigned char* in_ring = epi_base + kEpiBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(in_ring + prm.in_stages * prm.in_stride);
    uint64_t* full_bar = bars;                          // [2]  A (stencil) + B (TMA) ready
    uint64_t* empty_bar = bars + 2;                     // [2]  MMAs of the stage retired
    uint64_t* in_full = bars + 4;                       // [6]
    uint64_t* in_empty = bars + 10;                     // [6]
    uint64_t* tmem_full = bars + 16;                    // [2]
    uint64_t* tmem_empty = bars + 18;                   // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int K = prm.K, N = prm.N, Ho = prm.Ho, Wo = prm.Wo;
    const int n_tiles = N / BN;
    const int num_tiles = prm.P * prm.tiles_per_patch * n_tiles;
    const int num_kb = K / kBK;
    const int parts = prm.TRt / prm.TR;
    const int valid_rows = prm.TRt * Wo;
    const int in_stages = prm.in_stages;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_b_hi);
        if (NSPLIT > 1) tma_prefetch_desc(&map_b_lo);
    }
    if (warp == 3 && lane == 0) {
        tma_prefetch_desc(&map_in);
        tma_prefetch_desc(&map_c);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kABStages; ++i) {
            mbar_init(&full_bar[i], 1 + kF3ProdWarps);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < kMaxInStages; ++i) {
            mbar_init(&in_full[i], 1);
            mbar_init(&in_empty[i], kF3ProdWarps);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 128);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    if (valid_rows < kBM) {
        // rows the stencil never writes must not hold NaN/Inf bit patterns (their accumulator rows are discarded,
        // but keep the tensor pipe away from garbage): zero both A stages once
        for (int s = 0; s < kABStages; ++s) {
            uint4* a = reinterpret_cast<uint4*>(smem + s * Cfg::kStageBytes);
            for (int i = threadIdx.x; i < Cfg::kPlanes * Cfg::kATile / 16; i += kF3Threads) a[i] = make_uint4(0, 0, 0, 0);
        }
        fence_proxy_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================================================= weight TMA
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
                    if (++stage == kABStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 3) {
        // ================================================================= input TMA (float32 NHWC boxes)
        if (lane == 0) {
            const uint32_t box_bytes = static_cast<uint32_t>(kCB * prm.BW * prm.BH * 4);
            int is = 0;
            uint32_t iphase = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                const int m_tile = t / n_tiles;
                const int p = m_tile / prm.tiles_per_patch;
                const int oh0 = (m_tile - p * prm.tiles_per_patch) * prm.TRt;
                for (int kb = 0; kb < num_kb; ++kb) {
                    for (int sb = 0; sb < kBK / kCB; ++sb) {
                        for (int part = 0; part < parts; ++part) {
                            const int oh = oh0 + part * prm.TR;
                            mbar_wait(&in_empty[is], iphase ^ 1);
                            mbar_arrive_expect_tx(&in_full[is], box_bytes);
                            tma_load_4d(in_ring + is * prm.in_stride, &map_in, &in_full[is], kb * kBK + sb * kCB,
                                        STRIDE == 1 ? -1 : 0, STRIDE == 1 ? oh - 1 : 2 * oh, p);
                            if (++is == in_stages) { is = 0; iphase ^= 1; }
                        }
                    }
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
                    if (++stage == kABStages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tmem_full[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 8) {
        // ================================================================= stencil producers (smem -> smem A tile)
        const int pt = threadIdx.x - 256;
        const int quad = pt & 7;                        // 4 channels of the 32-channel box
        const int strip0 = pt >> 3;                     // strips strip0, strip0 + 32, ...
        const int wo_bits = 31 - __clz(Wo);
        const int strips_per_box = (prm.TR * Wo) / R;
        const int rowpitch = prm.BW * kCB * 4;          // bytes between box rows
        const uint32_t in_ring_u32 = smem_u32(in_ring);
        int stage = 0, is = 0;
        uint32_t phase = 0, iphase = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait_sleepy(&empty_bar[stage], phase ^ 1);
                const uint32_t a_hi = smem_u32(smem + stage * Cfg::kStageBytes);
                const uint32_t a_lo = a_hi + Cfg::kATile;
#pragma unroll 1
                for (int sb = 0; sb < kBK / kCB; ++sb) {
                    const int c = kb * kBK + sb * kCB + quad * 4;
                    float4 kk[9];
#pragma unroll
                    for (int i = 0; i < 9; ++i) kk[i] = __ldg(reinterpret_cast<const float4*>(prm.dw_w + i * K + c));
                    const float4 bdw = __ldg(reinterpret_cast<const float4*>(prm.dw_b + c));
                    const uint32_t chunk = static_cast<uint32_t>(sb * 4 + (quad >> 1));
#pragma unroll 1
                    for (int part = 0; part < parts; ++part) {
                        mbar_wait_sleepy(&in_full[is], iphase);
                        const uint32_t tile = in_ring_u32 + static_cast<uint32_t>(is * prm.in_stride + quad * 16);
#pragma unroll 1
                        for (int strip = strip0; strip < strips_per_box; strip += kF3ProdThreads / 8) {
                            const int px = strip * R;                     // box-local output pixel
                            const int oh_l = px >> wo_bits, ow0 = px & (Wo - 1);
                            const uint32_t base = tile + static_cast<uint32_t>((oh_l * STRIDE * prm.BW + ow0 * STRIDE) * kCB * 4);
                            float4 acc[R];
#pragma unroll
                            for (int r = 0; r < R; ++r) acc[r] = bdw;
#pragma unroll
                            for (int kh = 0; kh < 3; ++kh) {
                                const uint32_t rowp = base + static_cast<uint32_t>(kh * rowpitch);
                                float4 v[NC];
#pragma unroll
                                for (int j = 0; j < NC; ++j) v[j] = lds128(rowp + j * kCB * 4);
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
                            const int row0 = part * prm.TR * Wo + px;     // tile-local A row of the strip's first pixel
#pragma unroll
                            for (int r = 0; r < R; ++r) {
                                float4 a = acc[r];
                                a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f);
                                const uint32_t row = static_cast<uint32_t>(row0 + r);
                                const uint32_t off = (row >> 3) * 1024u + (row & 7u) * 128u + ((chunk ^ (row & 7u)) << 4) +
                                                     (static_cast<uint32_t>(quad & 1) << 3);
                                const __half h0 = __float2half_rn(a.x), h1 = __float2half_rn(a.y);
                                const __half h2 = __float2half_rn(a.z), h3 = __float2half_rn(a.w);
                                __half2 hp[2] = {__halves2half2(h0, h1), __halves2half2(h2, h3)};
                                sts64(a_hi + off, reinterpret_cast<uint32_t*>(hp)[0], reinterpret_cast<uint32_t*>(hp)[1]);
                                if (NSPLIT > 1) {
                                    __half2 lp[2] = {__halves2half2(__float2half_rn(a.x - __half2float(h0)),
                                                                    __float2half_rn(a.y - __half2float(h1))),
                                                     __halves2half2(__float2half_rn(a.z - __half2float(h2)),
                                                                    __float2half_rn(a.w - __half2float(h3)))};
                                    sts64(a_lo + off, reinterpret_cast<uint32_t*>(lp)[0], reinterpret_cast<uint32_t*>(lp)[1]);
                                }
                            }
                        }
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&in_empty[is]);           // box consumed (release orders the reads)
                        if (++is == in_stages) { is = 0; iphase ^= 1; }
                    }
                }
                fence_proxy_async_smem();            // generic-proxy smem writes -> visible to the tensor-core proxy
                __syncwarp();
                if (lane == 0) mbar_arrive(&full_bar[stage]);
                if (++stage == kABStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ================================================================= epilogue
        const int q = warp & 3;
        const uint32_t stg = smem_u32(epi_base) + static_cast<uint32_t>(q * kEpiBufBytes);
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
            const int m_tile = t / n_tiles, n_blk = t - m_tile * n_tiles;
            const int p = m_tile / prm.tiles_per_patch;
            const int oh0 = (m_tile - p * prm.tiles_per_patch) * prm.TRt;
            const int row0 = (p * Ho + oh0) * Wo + q * 32;               // < 2^31: checked by the launcher
            mbar_wait_sleepy(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const int n0 = n_blk * BN;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN);
            if (q * 32 < valid_rows) {                                   // valid_rows is a multiple of 32
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(taddr + static_cast<uint32_t>(c0), r);
                    if (lane == 0) tma_store_wait_read<0>();          // previous block's store has read the buffer
                    __syncwarp();
                    tmem_ld_wait();
                    const float* bp = biasp.v + n0 + c0;
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
                        tma_store_2d(&map_c, stg, n0 + c0, row0);
                        tma_store_commit();
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&tmem_empty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (lane == 0) tma_store_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
}

bool encode_4d_f32(CUtensorMap* map, const float* ptr, int C, int W, int H, int P, int box_c, int box_w, int box_h) {
    TensorMapEncodeFn fn = tensor_map_encode_fn();
    if (fn == nullptr) return false;
    cuuint64_t gdim[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                          static_cast<cuuint64_t>(P)};
    cuuint64_t gstride[3] = {static_cast<cuuint64_t>(C) * 4, static_cast<cuuint64_t>(W) * C * 4,
                             static_cast<cuuint64_t>(H) * W * C * 4};
    cuuint32_t box[4] = {static_cast<cuuint32_t>(box_c), static_cast<cuuint32_t>(box_w), static_cast<cuuint32_t>(box_h), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(ptr), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

template <int BN, int NSPLIT, int STRIDE, int R>
cudaError_t launch_f3_t(const CUtensorMap& map_in, const CUtensorMap& map_c, const BiasParam& bp, const PwGemmPlan& p,
                        const F3Params& prm, int smem_bytes, int grid, cudaStream_t stream) {
    sep_fused3_kernel<BN, NSPLIT, STRIDE, R><<<grid, kF3Threads, smem_bytes, stream>>>(map_in, p.b_hi, p.b_lo, map_c, bp, prm);
    return cudaGetLastError();
}

template <int BN, int NSPLIT>
cudaError_t set_attr_f3() {
    cudaError_t e;
#define BD_F3_ATTR(S, RR)                                                                                  \
    if ((e = cudaFuncSetAttribute(sep_fused3_kernel<BN, NSPLIT, S, RR>,                                    \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax)) != cudaSuccess)  \
        return e;
    BD_F3_ATTR(1, 4) BD_F3_ATTR(1, 2) BD_F3_ATTR(2, 4) BD_F3_ATTR(2, 2)
#undef BD_F3_ATTR
    return cudaSuccess;
}

}  // namespace

cudaError_t sep_fused3_init_device() {
    cudaError_t e;
    if ((e = set_attr_f3<64, 1>()) != cudaSuccess) return e;
    if ((e = set_attr_f3<128, 1>()) != cudaSuccess) return e;
    if ((e = set_attr_f3<64, 3>()) != cudaSuccess) return e;
    return set_attr_f3<128, 3>();
}

bool sep_fused3_supported(int K, int N, int H, int W, int stride) {
    if (K % kBK != 0 || N % 64 != 0 || N > kMaxN) return false;
    if (stride != 1 && stride != 2) return false;
    if ((H % stride) || (W % stride)) return false;
    const int Ho = H / stride, Wo = W / stride;
    if (Wo < 4 || (Wo & (Wo - 1)) || Wo > kBM) return false;
    const int trt = Ho < kBM / Wo ? Ho : kBM / Wo;
    if (trt < 1 || Ho % trt || (trt * Wo) % 32) return false;       // epilogue stores whole 32-row blocks
    return true;
}

cudaError_t launch_sep_fused3(const PwGemmPlan& p, const float* X, const float* dw_w, const float* dw_b,
                              const float* bias_host, float* C, int P, int H, int W, int stride, int num_sms,
                              cudaStream_t stream) {
    if (P <= 0) return cudaSuccess;
    if (!sep_fused3_supported(p.K, p.N, H, W, stride) || bias_host == nullptr) return cudaErrorInvalidValue;
    if (p.block_n != 64 && p.block_n != 128) return cudaErrorInvalidValue;
    const int Ho = H / stride, Wo = W / stride;
    F3Params prm;
    prm.dw_w = dw_w; prm.dw_b = dw_b;
    prm.P = P; prm.K = p.K; prm.N = p.N; prm.Ho = Ho; prm.Wo = Wo;
    prm.TRt = Ho < kBM / Wo ? Ho : kBM / Wo;
    prm.tiles_per_patch = Ho / prm.TRt;
    prm.BW = (Wo - 1) * stride + 3;
    // rows per input box: the largest divisor of TRt whose box stays under ~40 KB
    int tr = prm.TRt;
    while (tr > 1 && (kCB * prm.BW * ((tr - 1) * stride + 3) * 4 > 40 * 1024 || prm.TRt % tr)) --tr;
    prm.TR = tr;
    prm.BH = (tr - 1) * stride + 3;
    const int box_bytes = kCB * prm.BW * prm.BH * 4;
    prm.in_stride = (box_bytes + 127) & ~127;
    prm.out_scale = p.out_scale;
    const int planes = p.nsplit == 1 ? 1 : 2;
    const int stage_bytes = planes * (kBM * kBK * 2 + p.block_n * kBK * 2);
    const int fixed = 1024 + kABStages * stage_bytes + kEpiBytes + kBarBytes;
    int in_stages = (kSmemMax - fixed) / prm.in_stride;
    if (in_stages > kMaxInStages) in_stages = kMaxInStages;
    if (in_stages < 2) return cudaErrorInvalidValue;
    prm.in_stages = in_stages;
    const int smem_bytes = fixed + in_stages * prm.in_stride;
    // strip length: 4 output pixels when that still gives most stencil threads a strip, else 2
    const int items4 = (kCB / 4) * (prm.TR * Wo / 4);
    const int R = (Wo % 4 == 0 && items4 >= 192) ? 4 : 2;
    const long long tiles = static_cast<long long>(P) * prm.tiles_per_patch * (p.N / p.block_n);
    if (tiles >= (1LL << 31)) return cudaErrorInvalidValue;
    const int grid = static_cast<int>(tiles < num_sms ? tiles : num_sms);
    CUtensorMap map_in, map_c;
    if (!encode_4d_f32(&map_in, X, p.K, W, H, P, kCB, prm.BW, prm.BH)) return cudaErrorUnknown;
    const long long m_rows = static_cast<long long>(P) * Ho * Wo;
    if (m_rows >= (1LL << 31) || !encode_store_map_f32(&map_c, C, m_rows, p.N)) return cudaErrorUnknown;
    BiasParam bp;
    for (int i = 0; i < kMaxN; ++i) bp.v[i] = i < p.N ? bias_host[i] : 0.f;
#define BD_F3(BN, NS)                                                                                              \
    do {                                                                                                           \
        if (stride == 1) return R == 4 ? launch_f3_t<BN, NS, 1, 4>(map_in, map_c, bp, p, prm, smem_bytes, grid, stream)       \
                                       : launch_f3_t<BN, NS, 1, 2>(map_in, map_c, bp, p, prm, smem_bytes, grid, stream);      \
        return R == 4 ? launch_f3_t<BN, NS, 2, 4>(map_in, map_c, bp, p, prm, smem_bytes, grid, stream)                        \
                      : launch_f3_t<BN, NS, 2, 2>(map_in, map_c, bp, p, prm, smem_bytes, grid, stream);                       \
    } while (0)
    if (p.nsplit == 1) {
        if (p.block_n == 64) BD_F3(64, 1);
        BD_F3(128, 1);
    }
    if (p.block_n == 64) BD_F3(64, 3);
    BD_F3(128, 3);
#undef BD_F3
}

}  // namespace bd
