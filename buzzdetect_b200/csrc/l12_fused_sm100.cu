// Layers 1 + 2 in one warp-specialised kernel (sm_100a), version 2:
//     log-mel patch -> conv 3x3/2 (1->32) -> depthwise 3x3 (32) -> pointwise 32->64      (all with folded BN + ReLU)
// Reference ops: _conv + the first _separable_conv, embedders/yamnet/yamnet.py:36-74, layer table :77-80.
//
// l12_fused_kernel (pw_gemm_sm100.cu) runs its five phases back to back behind CTA-wide barriers and is latency
// bound.  Here each phase has its own warps and the phases of consecutive tiles overlap through mbarrier rings:
//
//   warps 8-11   conv1:   13 log-mel rows per tile, fetched three tiles ahead with cp.async (4-buffer ring) -> layer-1
//                         tile 6 x 34 x 32 fp32 with a one-pixel halo (zero outside the image = the depthwise SAME
//                         pad) in a 3-deep shared-memory ring
//   warps 12-15  stencil: depthwise 3x3 from the ring -> hi/lo fp16 straight into the SWIZZLE_128B A tile (2 stages)
//   warp 1       tcgen05.mma issuer: K = 32 -> two k-steps x 3 products (fp16x3) against the resident weight tile
//   warps 4-7    epilogue: tcgen05.ld -> scale + bias (constant bank) + ReLU -> swizzled smem block -> TMA store
//   warp 0       loads the 64 x 32 weight tile once (TMA); warp 2 owns TMEM (2 accumulator stages x 64 columns)
//
// A tile = 4 image rows x 32 columns of one patch (128 output pixels); per tile the kernel reads 13 log-mel rows
// (3.3 KB) and writes 32 KB of layer-2 output: the layer-1 activation (196 KB/patch) and the depthwise planes
// (2 x 98 KB/patch) never leave the SM.  Tap weights of both stencils live in registers for the whole kernel.
#include "bd_common.cuh"
#include "bd_kernels.cuh"

namespace bd {

namespace {

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kThreads = 512;
constexpr int kRows = 4;                                   // image rows per tile
constexpr int kTileH = kRows + 2, kTileW = 34;
constexpr int kC1Bytes = kTileH * kTileW * 32 * 4;         // 26,112
constexpr int kC1Slots = 3;
constexpr int kAStages = 2;
constexpr int kATile = kBM * kBK * 2;                      // one fp16 plane, 16 KB (columns 32..63 unused)
constexpr int kBTile = 64 * kBK * 2;                       // 8 KB per plane
constexpr int kEpiBufBytes = 32 * 128;                     // one [32 rows x 32 float] swizzled block
constexpr int kEpiBytes = 4 * 2 * kEpiBufBytes;            // 4 epilogue warps x 2 buffers
constexpr int kConvWarps = 4, kDwWarps = 4;

struct Bias64 { float v[64]; };                            // kernel parameter = constant bank: uniform reads cost no L1 traffic
constexpr int kLmRows = 2 * kTileH + 1;                    // 13 log-mel rows feed one layer-1 tile
constexpr int kLmBytes = kLmRows * kMel * 4;               // 3,328
constexpr int kLmBufs = 4;                                 // cp.async ring: rows are fetched 3 tiles ahead (DRAM latency > tile time)

template <int NSPLIT>
struct L12Cfg {
    static constexpr int kPlanes = NSPLIT == 1 ? 1 : 2;
    static constexpr int kAStageBytes = kPlanes * kATile;
    static constexpr int kSmemBytes = 1024 + kAStages * kAStageBytes + kEpiBytes + kPlanes * kBTile + kC1Slots * kC1Bytes + kLmBufs * kLmBytes + 256;
};

static_assert(L12Cfg<3>::kSmemBytes <= 227 * 1024, "layers-1+2 kernel: shared memory budget");

template <int NSPLIT>
__global__ void __launch_bounds__(kThreads, 1)
l12_fused2_kernel(const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                  const __grid_constant__ CUtensorMap map_c, const __grid_constant__ Bias64 biasp,
                  const float* __restrict__ logmel, int hop_frames, int P, const float* __restrict__ w1,
                  const float* __restrict__ b1, const float* __restrict__ dw_w, const float* __restrict__ dw_b,
                  float out_scale) {
    using Cfg = L12Cfg<NSPLIT>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* a_base = smem;                                            // [2][planes][16 KB]
    unsigned char* epi_base = a_base + kAStages * Cfg::kAStageBytes;         // [4 warps][2][4 KB], 1024-byte aligned
    unsigned char* b_base = epi_base + kEpiBytes;                            // [planes][8 KB]
    unsigned char* c1_base = b_base + Cfg::kPlanes * kBTile;                 // [3][26,112]
    unsigned char* lm_base = c1_base + kC1Slots * kC1Bytes;                  // [4][13][64] fp32 (cp.async ring)
    uint64_t* bars = reinterpret_cast<uint64_t*>(lm_base + kLmBufs * kLmBytes);
    uint64_t* a_full = bars;            // [2]
    uint64_t* a_empty = bars + 2;       // [2]
    uint64_t* c1_full = bars + 4;       // [kC1Slots <= 4]
    uint64_t* c1_empty = bars + 8;      // [kC1Slots <= 4]
    uint64_t* tmem_full = bars + 12;    // [2]
    uint64_t* tmem_empty = bars + 14;   // [2]
    uint64_t* b_bar = bars + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr unsigned tiles_per_patch = 48 / kRows;
    const unsigned num_tiles = static_cast<unsigned>(P) * tiles_per_patch;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_b_hi);
        if (NSPLIT > 1) tma_prefetch_desc(&map_b_lo);
        tma_prefetch_desc(&map_c);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kAStages; ++i) {
            mbar_init(&a_full[i], kDwWarps);
            mbar_init(&a_empty[i], 1);
        }
        for (int i = 0; i < kC1Slots; ++i) {
            mbar_init(&c1_full[i], kConvWarps);
            mbar_init(&c1_empty[i], kDwWarps);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 128);
        }
        mbar_init(b_bar, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<128>(tmem_slot);
    {
        // columns 32..63 of every A row are never written by the stencil and never read by the MMAs (k-steps 0, 1
        // only), but clear the tiles once anyway so no uninitialised bit pattern sits in an operand buffer
        uint4* a = reinterpret_cast<uint4*>(a_base);
        for (int i = threadIdx.x; i < kAStages * Cfg::kAStageBytes / 16; i += kThreads) a[i] = make_uint4(0, 0, 0, 0);
        fence_proxy_async_smem();
        // layer-1 ring: tile columns 0 and 33 lie outside the image for every tile (the depthwise SAME pad); they are
        // zeroed here once and never written again
        uint4* c = reinterpret_cast<uint4*>(c1_base);
        for (int i = threadIdx.x; i < kC1Slots * kC1Bytes / 16; i += kThreads) c[i] = make_uint4(0, 0, 0, 0);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================================================= weights: once per CTA
        if (lane == 0) {
            mbar_arrive_expect_tx(b_bar, Cfg::kPlanes * kBTile);
            tma_load_2d(b_base, &map_b_hi, b_bar, 0, 0);
            if (NSPLIT > 1) tma_load_2d(b_base + kBTile, &map_b_lo, b_bar, 0, 0);
        }
    } else if (warp == 1) {
        // ================================================================= MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(kBM, 64);
            mbar_wait(b_bar, 0);
            tc_fence_after();
            const uint32_t bh = smem_u32(b_base), bl = bh + kBTile;
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (unsigned t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                mbar_wait(&a_full[stage], phase);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * 64);
                const uint32_t ah = smem_u32(a_base + stage * Cfg::kAStageBytes), al = ah + kATile;
#pragma unroll
                for (int k = 0; k < 2; ++k) {                                // K = 32: two k-steps of 16
                    const uint32_t koff = static_cast<uint32_t>(k) * 32u;
                    umma_f16_ss(d_tmem, umma_desc_k128(ah + koff), umma_desc_k128(bh + koff), idesc, k != 0 ? 1u : 0u);
                    if (NSPLIT > 1) {
                        umma_f16_ss(d_tmem, umma_desc_k128(al + koff), umma_desc_k128(bh + koff), idesc, 1u);
                        umma_f16_ss(d_tmem, umma_desc_k128(ah + koff), umma_desc_k128(bl + koff), idesc, 1u);
                    }
                }
                umma_commit(&a_empty[stage]);
                umma_commit(&tmem_full[acc]);
                if (++stage == kAStages) { stage = 0; phase ^= 1; }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 12) {
        // ================================================================= depthwise stencil: c1 ring -> A tile
        const int t = threadIdx.x - 384;                 // 0..127
        const int quad = t & 7;
        float4 kk[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) kk[i] = __ldg(reinterpret_cast<const float4*>(dw_w + i * 32 + quad * 4));
        const float4 bdw = __ldg(reinterpret_cast<const float4*>(dw_b + quad * 4));
        const uint32_t c1_u32 = smem_u32(c1_base) + static_cast<uint32_t>(quad * 16);
        const uint32_t a_u32 = smem_u32(a_base);
        const uint32_t chunk = static_cast<uint32_t>(quad >> 1);
        int stage = 0, slot = 0;
        uint32_t phase = 0, sphase = 0;
        for (unsigned tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            mbar_wait_sleepy(&a_empty[stage], phase ^ 1);
            mbar_wait_sleepy(&c1_full[slot], sphase);
            const uint32_t src = c1_u32 + static_cast<uint32_t>(slot * kC1Bytes);
            const uint32_t a_hi = a_u32 + static_cast<uint32_t>(stage * Cfg::kAStageBytes), a_lo = a_hi + kATile;
            {
                // block of 4 columns x 2 rows of outputs: 4 input rows x 6 columns, 3 LDS.128 per output vector
                const int blk = t >> 3;                  // 0..15
                const int brow = blk >> 3, ws = blk & 7;
                float4 acc[2][4];
#pragma unroll
                for (int o = 0; o < 2; ++o)
#pragma unroll
                    for (int r = 0; r < 4; ++r) acc[o][r] = bdw;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t rowp = src + static_cast<uint32_t>((((2 * brow + i) * kTileW + ws * 4) * 32) * 4);
                    float4 v[6];
#pragma unroll
                    for (int j = 0; j < 6; ++j) v[j] = lds128(rowp + j * 128);
#pragma unroll
                    for (int o = 0; o < 2; ++o) {
                        const int kh = i - o;
                        if (kh < 0 || kh > 2) continue;
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
#pragma unroll
                            for (int kw = 0; kw < 3; ++kw) {
                                const float4 x = v[r + kw];
                                const float4 w4 = kk[kh * 3 + kw];
                                acc[o][r].x = fmaf(x.x, w4.x, acc[o][r].x);
                                acc[o][r].y = fmaf(x.y, w4.y, acc[o][r].y);
                                acc[o][r].z = fmaf(x.z, w4.z, acc[o][r].z);
                                acc[o][r].w = fmaf(x.w, w4.w, acc[o][r].w);
                            }
                        }
                    }
                }
#pragma unroll
                for (int o = 0; o < 2; ++o) {
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        float4 a = acc[o][r];
                        a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f);
                        const uint32_t row = static_cast<uint32_t>((2 * brow + o) * 32 + ws * 4 + r);
                        const uint32_t off = (row << 7) + ((chunk ^ (row & 7u)) << 4) +        // row * 128 = 8-row groups of 1 KB
                                             (static_cast<uint32_t>(quad & 1) << 3);
                        __half2 hp[2] = {half2_sat(a.x, a.y), half2_sat(a.z, a.w)};
                        const __half h0 = __low2half(hp[0]), h1 = __high2half(hp[0]);
                        const __half h2 = __low2half(hp[1]), h3 = __high2half(hp[1]);
                        sts64(a_hi + off, reinterpret_cast<uint32_t*>(hp)[0], reinterpret_cast<uint32_t*>(hp)[1]);
                        if (NSPLIT > 1) {
                            __half2 lp[2] = {half2_sat(a.x - __half2float(h0), a.y - __half2float(h1)),
                                             half2_sat(a.z - __half2float(h2), a.w - __half2float(h3))};
                            sts64(a_lo + off, reinterpret_cast<uint32_t*>(lp)[0], reinterpret_cast<uint32_t*>(lp)[1]);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&c1_empty[slot]);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_full[stage]);
            if (++stage == kAStages) { stage = 0; phase ^= 1; }
            if (++slot == kC1Slots) { slot = 0; sphase ^= 1; }
        }
    } else if (warp >= 8) {
        // ================================================================= conv1: log-mel -> layer-1 tile (with halo)
        const int t = threadIdx.x - 256;                 // 0..127
        const int cg = t & 7;
        float4 w1r[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) w1r[i] = __ldg(reinterpret_cast<const float4*>(w1 + i * 32 + cg * 4));
        const float4 b1r = __ldg(reinterpret_cast<const float4*>(b1 + cg * 4));
        const uint32_t c1_u32 = smem_u32(c1_base) + static_cast<uint32_t>(cg * 16);
        const uint32_t lm_u32 = smem_u32(lm_base);
        // log-mel rows of a tile: 13 x 64 floats = 208 16-byte chunks, fetched with cp.async one tile ahead (rows outside
        // the patch are zero-filled: conv1's SAME pad below row 95, the halo above row 0)
        auto prefetch = [&](unsigned tile, int buf) {
            if (tile >= num_tiles) {                         // keep one commit group per tile so wait_group counts line up
                asm volatile("cp.async.commit_group;" ::: "memory");
                return;
            }
            const long long p = tile / tiles_per_patch;
            const int r0 = static_cast<int>(tile - static_cast<unsigned>(p) * tiles_per_patch) * kRows;
            const float* in = logmel + p * hop_frames * kMel;
            const int lm_row0 = 2 * (r0 - 1);
#pragma unroll
            for (int c = t; c < kLmRows * 16; c += 128) {
                const int rr = c >> 4, c4 = c & 15;
                const int gr = lm_row0 + rr;
                const bool ok = gr >= 0 && gr < kPatchFrames;
                const float* src = in + (ok ? gr : 0) * kMel + c4 * 4;
                const uint32_t dst = lm_u32 + static_cast<uint32_t>(buf * kLmBytes + c * 16);
                const uint32_t nbytes = ok ? 16u : 0u;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        int slot = 0, buf = 0;
        uint32_t sphase = 0;
#pragma unroll
        for (int i = 0; i < kLmBufs - 1; ++i) prefetch(blockIdx.x + i * gridDim.x, i);
        for (unsigned tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const long long p = tile / tiles_per_patch;
            const int r0 = static_cast<int>(tile - static_cast<unsigned>(p) * tiles_per_patch) * kRows;
            asm volatile("cp.async.wait_group %0;" ::"n"(kLmBufs - 2) : "memory");
            asm volatile("bar.sync 1, 128;" ::: "memory");          // tile's rows visible to all conv warps; previous tile consumed
            prefetch(tile + (kLmBufs - 1) * gridDim.x, (buf + kLmBufs - 1) % kLmBufs);
            mbar_wait_sleepy(&c1_empty[slot], sphase ^ 1);
            const uint32_t dst = c1_u32 + static_cast<uint32_t>(slot * kC1Bytes);
            const uint32_t lm = lm_u32 + static_cast<uint32_t>(buf * kLmBytes);
            // thread = (channel quad, pair of interior columns): image columns ic0 = 2*pair, ic0 + 1 of tile row tr; the
            // five log-mel columns 2*ic0 .. 2*ic0+4 of a row come from one LDS.128 + one LDS.32
#pragma unroll 2
            for (int item = t >> 3; item < kTileH * 16; item += 16) {
                const int tr = item >> 4, pair = item & 15;
                const int ir = r0 - 1 + tr, ic0 = 2 * pair;
                float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
                if (ir >= 0 && ir < 48) {
                    a0 = b1r; a1 = b1r;
                    const uint32_t l0 = lm + static_cast<uint32_t>(((2 * tr) * kMel + 2 * ic0) * 4);   // staged row 2*tr <-> log-mel row 2*ir
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh) {
                        const float4 v = lds128(l0 + kh * kMel * 4);
                        float v4;
                        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v4) : "r"(l0 + kh * kMel * 4 + 16));
                        if (pair == 15) v4 = 0.f;                                 // column 64 is conv1's SAME pad
                        const float4 wa = w1r[kh * 3 + 0], wb = w1r[kh * 3 + 1], wc = w1r[kh * 3 + 2];
                        a0.x = fmaf(v.x, wa.x, a0.x); a0.y = fmaf(v.x, wa.y, a0.y); a0.z = fmaf(v.x, wa.z, a0.z); a0.w = fmaf(v.x, wa.w, a0.w);
                        a0.x = fmaf(v.y, wb.x, a0.x); a0.y = fmaf(v.y, wb.y, a0.y); a0.z = fmaf(v.y, wb.z, a0.z); a0.w = fmaf(v.y, wb.w, a0.w);
                        a0.x = fmaf(v.z, wc.x, a0.x); a0.y = fmaf(v.z, wc.y, a0.y); a0.z = fmaf(v.z, wc.z, a0.z); a0.w = fmaf(v.z, wc.w, a0.w);
                        a1.x = fmaf(v.z, wa.x, a1.x); a1.y = fmaf(v.z, wa.y, a1.y); a1.z = fmaf(v.z, wa.z, a1.z); a1.w = fmaf(v.z, wa.w, a1.w);
                        a1.x = fmaf(v.w, wb.x, a1.x); a1.y = fmaf(v.w, wb.y, a1.y); a1.z = fmaf(v.w, wb.z, a1.z); a1.w = fmaf(v.w, wb.w, a1.w);
                        a1.x = fmaf(v4, wc.x, a1.x); a1.y = fmaf(v4, wc.y, a1.y); a1.z = fmaf(v4, wc.z, a1.z); a1.w = fmaf(v4, wc.w, a1.w);
                    }
                    a0.x = fmaxf(a0.x, 0.f); a0.y = fmaxf(a0.y, 0.f); a0.z = fmaxf(a0.z, 0.f); a0.w = fmaxf(a0.w, 0.f);
                    a1.x = fmaxf(a1.x, 0.f); a1.y = fmaxf(a1.y, 0.f); a1.z = fmaxf(a1.z, 0.f); a1.w = fmaxf(a1.w, 0.f);
                }
                const uint32_t o = dst + static_cast<uint32_t>((tr * kTileW + 1 + ic0) * 128);        // tile column = image column + 1
                sts128(o, a0);
                sts128(o + 128, a1);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&c1_full[slot]);
            if (++slot == kC1Slots) { slot = 0; sphase ^= 1; }
            if (++buf == kLmBufs) buf = 0;
        }
    } else if (warp >= 4) {
        // ================================================================= epilogue
        const int q = warp & 3;
        const uint32_t stg0 = smem_u32(epi_base) + static_cast<uint32_t>(q * 2 * kEpiBufBytes);
        int acc = 0;
        uint32_t acc_phase = 0;
        for (unsigned tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            mbar_wait_sleepy(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const int row0 = static_cast<int>(tile) * kBM + q * 32;      // < 2^31: checked by the launcher
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * 64);
#pragma unroll
            for (int cb = 0; cb < 2; ++cb) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(taddr + static_cast<uint32_t>(cb * 32), r);
                const uint32_t stg = stg0 + static_cast<uint32_t>(cb * kEpiBufBytes);
                if (lane == 0) tma_store_wait_read<1>();              // the store that last read this buffer is done
                __syncwarp();
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 o;
                    o.x = fmaxf(fmaf(__uint_as_float(r[4 * j + 0]), out_scale, biasp.v[cb * 32 + 4 * j + 0]), 0.f);
                    o.y = fmaxf(fmaf(__uint_as_float(r[4 * j + 1]), out_scale, biasp.v[cb * 32 + 4 * j + 1]), 0.f);
                    o.z = fmaxf(fmaf(__uint_as_float(r[4 * j + 2]), out_scale, biasp.v[cb * 32 + 4 * j + 2]), 0.f);
                    o.w = fmaxf(fmaf(__uint_as_float(r[4 * j + 3]), out_scale, biasp.v[cb * 32 + 4 * j + 3]), 0.f);
                    sts128(epi_swz_addr(stg, lane, j), o);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&map_c, stg, cb * 32, row0);
                    tma_store_commit();
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
        tmem_dealloc<128>(tmem_base);
    }
}

}  // namespace

cudaError_t l12_fused2_init_device() {
    cudaError_t e = cudaFuncSetAttribute(l12_fused2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         L12Cfg<1>::kSmemBytes);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(l12_fused2_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, L12Cfg<3>::kSmemBytes);
}

cudaError_t launch_l12_fused2(const PwGemmPlan& p, const float* logmel, int hop_frames, int P, const float* w1,
                              const float* b1, const float* dw_w, const float* dw_b, const float* bias_host, float* C,
                              int num_sms, cudaStream_t stream) {
    if (P <= 0) return cudaSuccess;
    if (p.N != 64 || p.K != 32 || p.block_n != 64 || bias_host == nullptr) return cudaErrorInvalidValue;
    const long long tiles = static_cast<long long>(P) * (48 / kRows);
    if (tiles * kBM >= (1LL << 31)) return cudaErrorInvalidValue;
    const int grid = static_cast<int>(tiles < num_sms ? tiles : num_sms);
    CUtensorMap map_c;
    if (!encode_store_map_f32(&map_c, C, tiles * kBM, 64, 32)) return cudaErrorUnknown;
    Bias64 bp;
    for (int i = 0; i < 64; ++i) bp.v[i] = bias_host[i];
    if (p.nsplit == 1)
        l12_fused2_kernel<1><<<grid, kThreads, L12Cfg<1>::kSmemBytes, stream>>>(p.b_hi, p.b_lo, map_c, bp, logmel, hop_frames, P,
                                                                                w1, b1, dw_w, dw_b, p.out_scale);
    else
        l12_fused2_kernel<3><<<grid, kThreads, L12Cfg<3>::kSmemBytes, stream>>>(p.b_hi, p.b_lo, map_c, bp, logmel, hop_frames, P,
                                                                                w1, b1, dw_w, dw_b, p.out_scale);
    return cudaGetLastError();
}

}  // namespace bd
