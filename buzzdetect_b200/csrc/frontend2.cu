// Fused YAMNet frontend for sm_100a, version 2: frame(400/160) * periodic Hann -> zero-pad to 512 -> real FFT -> |.|
// -> mel (64 bands) -> log(x + 0.001), for one or several independent audio SEGMENTS per launch.
//
// Reference semantics: embedders/yamnet/features.py:22-79 (waveform_to_log_mel_spectrogram_patches) and the op list of
// embedders/yamnet_k2/models/yamnet_wholehop/saved_model.pb: tf.signal.frame -> * hann_window(periodic) ->
// Pad[[0,0],[0,112]] -> RFFT(512) -> ComplexAbs -> MatMul(mel) -> AddV2(0.001) -> Log.  The tail padding of
// pad_waveform (features.py:82-108) is virtual: samples at index >= n_valid read as 0.
//
// Why a rewrite (profiles/r1_summary.md): version 1 ran one frame per warp with two shared-memory FFT exchanges, a
// split step through shared memory and a gather-style sparse mel; ncu showed 85 % of the L1/shared data pipe and 8 % of
// DRAM.  Here
//   * the 512-point real FFT is a 256-point complex FFT as 16 x 16: sixteen lanes hold sixteen points each, so a warp
//     transforms TWO frames at a time with ONE conflict-free transpose through shared memory (32 wavefronts per frame
//     instead of 64 + the split step's 40);
//   * the split step takes its mirror operand Z[256-k] from the partner lane by shuffle (both values of a pair are in
//     registers already), and the magnitudes go to a frame-major tile mag[32][257];
//   * the mel product runs with lane = FRAME over that tile: every lane walks the same (bin, weight) list, so the
//     weights are uniform (constant bank) and each shared-memory read is one conflict-free wavefront for 32 frames
//     (about 14 wavefronts per frame instead of ~200 for the per-frame gather);
//   * the input tile arrives by one TMA bulk copy (prefetched behind the mel phase of the previous tile) and the
//     log-mel tile leaves by TMA store from a swizzled staging block.
// What bounds it now is instruction issue: a 512-point float32 FFT + split + mel costs about 450 warp instructions per
// frame against 896 bytes of HBM traffic, which puts the kernel at the FP32-issue side of the ridge.
//
// A launch covers a table of segments (independent chunks of audio, each framed and padded on its own, exactly as the
// reference treats chunks): segment s writes log-mel rows [row_begin, row_begin + n_rows) of one shared buffer.
#include <cstdlib>
#include <cstring>

#include "bd_common.cuh"
#include "bd_kernels.cuh"

namespace bd {

namespace {

constexpr int kTileFrames = 32;
constexpr int kFeWarps = 8;
constexpr int kFeThreads = kFeWarps * 32;
constexpr int kTileSamples = (kTileFrames - 1) * kHop + kWin;   // 5360 floats = 21,440 B
constexpr int kMagStride = 260;                                 // 16-byte rows: the mel phase reads four bins per LDS.128
constexpr int kMelGroupsMax = 256;                              // 4-bin groups over all bands (shipped matrix: ~170)
constexpr int kExRow = 17;                                      // float2 per transpose row (16 + 1 pad)
constexpr int kExHalf = 16 * kExRow;                            // one frame's 16 x 16 transpose buffer
constexpr int kStageBytes = 2 * 4096;                           // [32 frames x 64 bands] as two swizzled 32-column blocks

constexpr int kOffStage = 0;                                    // 2 staging buffers, 1024-byte aligned
constexpr int kOffSamples = kOffStage + 2 * kStageBytes;
constexpr int kOffMag = kOffSamples + kTileSamples * 4;
constexpr int kOffEx = kOffMag + kTileFrames * kMagStride * 4;
constexpr int kOffBar = kOffEx + kFeWarps * 2 * kExHalf * 8;
constexpr int kFe2Smem = kOffBar + 16 + 1024;                   // + alignment slack
static_assert(kOffSamples % 16 == 0 && kOffEx % 8 == 0 && kOffBar % 8 == 0, "frontend smem layout");
static_assert(2 * (kFe2Smem + 1024) <= 228 * 1024, "two CTAs per SM");

struct FeMel {                     // kernel parameter (constant bank): uniform reads cost no L1/shared traffic
    float4 w4[kMelGroupsMax];      // 0.5 * mel weights in groups of four consecutive bins (zero outside the band); the
                                   // factor 0.5 undoes the split step's scale of 2 (exact)
    int gbin[kMel];                // first bin of the band's first group (a multiple of 4)
    int glen[kMel];                // number of groups
    int goff[kMel];                // index of the band's first group in w4
    int warp_band[kFeWarps + 1];   // warp w computes bands [warp_band[w], warp_band[w+1]) (balanced by work)
    int static_ok;                 // the matrix has the sparsity structure mel_layout.inc was generated for
};

struct FeSeg {
    const float* x;
    long long n_valid;
    long long frame_begin;         // first STFT frame of this launch within the segment
    int row_begin;                 // first log-mel row written
    int n_rows;                    // frames to compute
    int tile_begin;                // prefix sum of tiles
    int fmt;                       // 0 float32 samples, 1 int16 PCM
};

struct FeSegs {
    int n;
    int total_tiles;
    FeSeg s[kMaxLogmelSegs];
};

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cmulc(float2 a, float cr, float ci) {
    return make_float2(a.x * cr - a.y * ci, a.x * ci + a.y * cr);
}
__device__ __forceinline__ float2 mul_neg_i(float2 a) { return make_float2(a.y, -a.x); }

// out[k] = sum_n c[n] * exp(-2 pi i n k / 4), in place, natural order
__device__ __forceinline__ void dft4(float2& c0, float2& c1, float2& c2, float2& c3) {
    const float2 e0 = cadd(c0, c2), e1 = csub(c0, c2);
    const float2 f0 = cadd(c1, c3), f1 = mul_neg_i(csub(c1, c3));
    c0 = cadd(e0, f0);
    c1 = cadd(e1, f1);
    c2 = csub(e0, f0);
    c3 = csub(e1, f1);
}

// 16-point DFT in place (4 x 4): X[k] = sum_n v[n] exp(-2 pi i n k / 16) ends up at v[P16(k)]
__host__ __device__ constexpr int P16(int k) { return 4 * (k & 3) + (k >> 2); }

__device__ __forceinline__ void dft16(float2 (&v)[16]) {
    constexpr float c1 = 0.92387953251128674f, s1 = 0.38268343236508978f, h = 0.70710678118654752f;
#pragma unroll
    for (int b = 0; b < 4; ++b) dft4(v[b], v[4 + b], v[8 + b], v[12 + b]);      // v[4c + b] = t[b][c]
    // t[b][c] *= W16^(b c)
    v[4 * 1 + 1] = cmulc(v[4 * 1 + 1], c1, -s1);                                    // W^1
    v[4 * 2 + 1] = make_float2((v[4 * 2 + 1].x + v[4 * 2 + 1].y) * h, (v[4 * 2 + 1].y - v[4 * 2 + 1].x) * h);   // W^2
    v[4 * 3 + 1] = cmulc(v[4 * 3 + 1], s1, -c1);                                    // W^3
    v[4 * 1 + 2] = make_float2((v[4 * 1 + 2].x + v[4 * 1 + 2].y) * h, (v[4 * 1 + 2].y - v[4 * 1 + 2].x) * h);   // W^2
    v[4 * 2 + 2] = mul_neg_i(v[4 * 2 + 2]);                                         // W^4
    v[4 * 3 + 2] = make_float2((v[4 * 3 + 2].y - v[4 * 3 + 2].x) * h, -(v[4 * 3 + 2].x + v[4 * 3 + 2].y) * h);  // W^6
    v[4 * 1 + 3] = cmulc(v[4 * 1 + 3], s1, -c1);                                    // W^3
    v[4 * 2 + 3] = make_float2((v[4 * 2 + 3].y - v[4 * 2 + 3].x) * h, -(v[4 * 2 + 3].x + v[4 * 2 + 3].y) * h);  // W^6
    v[4 * 3 + 3] = cmulc(v[4 * 3 + 3], -c1, s1);                                    // W^9
#pragma unroll
    for (int c = 0; c < 4; ++c) dft4(v[4 * c + 0], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);   // v[4c + d] = X[c + 4d]
}

// exp(-2 pi i k2 / 32), k2 = 0..15: W512^(l + 16 k2) = W512^l * W32^k2
__device__ __forceinline__ float2 w32(int k2) {
    constexpr float t[16][2] = {
        {1.f, -0.f},
        {0.98078528040323043f, -0.19509032201612825f},
        {0.92387953251128674f, -0.38268343236508978f},
        {0.83146961230254524f, -0.55557023301960218f},
        {0.70710678118654757f, -0.70710678118654746f},
        {0.55557023301960229f, -0.83146961230254524f},
        {0.38268343236508984f, -0.92387953251128674f},
        {0.19509032201612833f, -0.98078528040323043f},
        {0.f, -1.f},
        {-0.19509032201612819f, -0.98078528040323043f},
        {-0.38268343236508973f, -0.92387953251128674f},
        {-0.55557023301960196f, -0.83146961230254546f},
        {-0.70710678118654746f, -0.70710678118654757f},
        {-0.83146961230254535f, -0.55557023301960218f},
        {-0.92387953251128674f, -0.38268343236508989f},
        {-0.98078528040323043f, -0.19509032201612861f}};
    return make_float2(t[k2][0], t[k2][1]);
}

__device__ __forceinline__ float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// log(x) = log2(x) * ln 2 with the hardware approximation: absolute error <= 2^-21.4 for x in [0.5, 2], 3 ulp elsewhere
// (CUDA C++ Programming Guide, intrinsic table) -- below 2e-6 over the log-mel range, against ~30 instructions for logf
__device__ __forceinline__ float log_fast(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y * 0.69314718055994531f;
}

__device__ __forceinline__ void bulk_load_1d(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst),
                 "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// explicit shared-window accesses: the buffers are carved out of an aligned byte array, so ptxas cannot prove the
// address space from the pointers and would fall back to generic LD / ST
__device__ __forceinline__ float2 lds64(uint32_t a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts64f(uint32_t a, float2 v) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ uint32_t lds32u(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ float lds32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts32f(uint32_t a, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}

struct TileInfo {
    const float* x;
    long long n_valid;
    long long g0;          // first sample of the tile
    int row0;              // first log-mel row of the tile
    int rows_valid;        // 32, or n_rows of a segment shorter than one tile
    bool tma;
    bool pcm16;            // the tile holds int16 PCM: the FFT runs on the integer values, the mel sums are scaled by 2^-15
};

// ---- the mel phase unrolled for YAMNet's mel structure (tools/gen_mel_layout.py): no loop control, no table look-ups,
// weights as immediate constant-bank operands of the FMAs, zero-weight lanes skipped
#define MEL_ARGS const FeMel& mel, uint32_t mp, uint32_t stg, int lane, float in_scale
#define MEL_LDS(off) lds128(mp + (off))
#define MEL_W(g) mel.w4[(g)]
#define MEL_EMIT(m, acc)                                                                                              \
    do {                                                                                                              \
        const float val_ = log_fast(fmaf((acc), in_scale, 0.001f));                                                   \
        constexpr int col_ = (m) & 31;                                                                                \
        const uint32_t addr_ = stg + static_cast<uint32_t>(((m) >> 5) * 4096 + lane * 128 +                           \
                                                           ((((col_ >> 2) ^ (lane & 7))) << 4) + (col_ & 3) * 4);      \
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr_), "f"(val_) : "memory");                                   \
    } while (0)
template <int W> __device__ __forceinline__ void mel_static_warp(MEL_ARGS);
#include "mel_layout.inc"
#undef MEL_EMIT
#undef MEL_W
#undef MEL_LDS
#undef MEL_ARGS

template <bool STATIC_MEL>
__global__ void __launch_bounds__(kFeThreads, 2)
logmel2_kernel(const __grid_constant__ FeSegs segs, const __grid_constant__ FeMel mel,
               const __grid_constant__ CUtensorMap map_out, const float* __restrict__ window,
               float* __restrict__ logmel) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float* samp = reinterpret_cast<float*>(smem + kOffSamples);
    float* mag = reinterpret_cast<float*>(smem + kOffMag);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kOffBar);
    const uint32_t stage_u32 = smem_u32(smem + kOffStage);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int h = lane >> 4, l = lane & 15;
    const uint32_t samp_u32 = smem_u32(samp), mag_u32 = smem_u32(mag);
    const uint32_t ex_u32 = smem_u32(smem + kOffEx) + static_cast<uint32_t>((warp * 2 + h) * kExHalf * 8);

    if (tid == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
        tma_prefetch_desc(&map_out);
    }

    // ---- per-lane constants: window taps of this lane's 13 sample pairs, pass-1 twiddles W256^(l k1), W512^l
    float2 wreg[13];
#pragma unroll
    for (int n1 = 0; n1 < 13; ++n1) {
        const int idx = 32 * n1 + 2 * l;
        wreg[n1] = idx < kWin ? make_float2(__ldg(window + idx), __ldg(window + idx + 1)) : make_float2(0.f, 0.f);
    }
    float2 tw1[16];
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) {
        float sn, cs;
        sincospif(-static_cast<float>(l * k1) / 128.0f, &sn, &cs);
        tw1[k1] = make_float2(cs, sn);
    }
    float2 wl;
    {
        float sn, cs;
        sincospif(-static_cast<float>(l) / 256.0f, &sn, &cs);
        wl = make_float2(cs, sn);
    }
    const int src_lane = ((16 - l) & 15) + 16 * h;            // holder of Z[256 - k] for this lane's bins
    // the padding bins 257..259 of every magnitude row meet zero weights in the mel phase: keep them finite
    for (int i = tid; i < kTileFrames * 3; i += kFeThreads) mag[(i / 3) * kMagStride + 257 + i % 3] = 0.f;

    auto tile_info = [&](int t) {
        TileInfo ti;
        int si = 0;
        while (si + 1 < segs.n && t >= segs.s[si + 1].tile_begin) ++si;
        const FeSeg& sg = segs.s[si];
        int f0 = (t - sg.tile_begin) * kTileFrames;
        ti.rows_valid = sg.n_rows < kTileFrames ? sg.n_rows : kTileFrames;
        if (f0 + kTileFrames > sg.n_rows && sg.n_rows >= kTileFrames) f0 = sg.n_rows - kTileFrames;   // last tile: shifted back
        ti.x = sg.x;
        ti.n_valid = sg.n_valid;
        ti.g0 = (sg.frame_begin + f0) * static_cast<long long>(kHop);
        ti.row0 = sg.row_begin + f0;
        ti.tma = (reinterpret_cast<uintptr_t>(sg.x) & 15u) == 0 && ti.g0 + kTileSamples <= sg.n_valid;
        ti.pcm16 = sg.fmt == 1;
        return ti;
    };
    // the tile's 5360 samples -> shared memory: one TMA bulk copy when the tile lies inside the segment (and the base is
    // 16-byte aligned), else guarded loads with zero fill (segment tails, unaligned callers)
    auto issue_load = [&](const TileInfo& ti) {
        if (ti.tma) {
            if (tid == 0) {
                fence_proxy_async_smem();                           // earlier generic-proxy accesses to the buffer are ordered
                const uint32_t bytes = ti.pcm16 ? kTileSamples * 2 : kTileSamples * 4;
                const void* src = ti.pcm16 ? static_cast<const void*>(reinterpret_cast<const short*>(ti.x) + ti.g0)
                                           : static_cast<const void*>(ti.x + ti.g0);
                mbar_arrive_expect_tx(bar, bytes);
                bulk_load_1d(smem_u32(samp), src, bytes, bar);
            }
        } else if (ti.pcm16) {
            const short* xs16 = reinterpret_cast<const short*>(ti.x);
            short* s16 = reinterpret_cast<short*>(samp);
            for (int i = tid; i < kTileSamples; i += kFeThreads) {
                const long long g = ti.g0 + i;
                s16[i] = g < ti.n_valid ? __ldg(xs16 + g) : static_cast<short>(0);
            }
        } else {
            for (int i = tid; i < kTileSamples; i += kFeThreads) {
                const long long g = ti.g0 + i;
                samp[i] = g < ti.n_valid ? __ldg(ti.x + g) : 0.f;
            }
        }
    };

    __syncthreads();                                               // barrier initialised
    int t = blockIdx.x;
    TileInfo cur{};
    if (t < segs.total_tiles) {
        cur = tile_info(t);
        issue_load(cur);
    }
    __syncthreads();
    uint32_t phase = 0;
    int iter = 0;
    for (; t < segs.total_tiles; t += gridDim.x, ++iter) {
        if (cur.tma) {
            mbar_wait(bar, phase);
            phase ^= 1;
        }
        // ================================================================= FFT: warp = frames (fi, fi + 16), two rounds
#pragma unroll 1
        for (int r = 0; r < 2; ++r) {
            // this half-warp's frame within the tile; the two halves are 4 frames apart so that their magnitude rows
            // (260 floats each) fall into disjoint bank halves
            const int f = (warp & 3) + 4 * h + 8 * (warp >> 2) + 16 * r;
            float2 v[16];
            if (cur.pcm16) {
                // int16 PCM: the transform is linear and a power-of-two scale commutes with every rounding in it, so the
                // FFT runs on the integer sample values and the 2^-15 of soundfile's float32 conversion is applied to the
                // mel sums -- bit-identical to converting first, without the conversion pass or its 6 bytes per sample
                const uint32_t xs = samp_u32 + static_cast<uint32_t>((f * kHop + 2 * l) * 2);
#pragma unroll
                for (int n1 = 0; n1 < 13; ++n1) {
                    const uint32_t u = lds32u(n1 < 12 || l < 8 ? xs + 32 * 2 * n1 : xs);
                    const float s0 = static_cast<float>(static_cast<short>(u & 0xFFFFu));
                    const float s1 = static_cast<float>(static_cast<int>(u) >> 16);
                    v[n1] = make_float2(s0 * wreg[n1].x, s1 * wreg[n1].y);
                }
            } else {
                const uint32_t xs = samp_u32 + static_cast<uint32_t>((f * kHop + 2 * l) * 4);
#pragma unroll
                for (int n1 = 0; n1 < 12; ++n1) {
                    const float2 sv = lds64(xs + 32 * 4 * n1);
                    v[n1] = make_float2(sv.x * wreg[n1].x, sv.y * wreg[n1].y);
                }
                // samples 384..399 (lanes l < 8); beyond them the 512-point frame is zero padding.  The other lanes read
                // a clamped address and multiply by their zero window taps.
                const float2 sv = lds64(l < 8 ? xs + 32 * 4 * 12 : xs);
                v[12] = make_float2(sv.x * wreg[12].x, sv.y * wreg[12].y);
            }
            v[13] = v[14] = v[15] = make_float2(0.f, 0.f);
            dft16(v);                                              // over n1; result k1 at v[P16(k1)]
#pragma unroll
            for (int k1 = 0; k1 < 16; ++k1)
                sts64f(ex_u32 + static_cast<uint32_t>((k1 * kExRow + l) * 8), k1 == 0 ? v[P16(0)] : cmul(v[P16(k1)], tw1[k1]));
            __syncwarp();
#pragma unroll
            for (int n2 = 0; n2 < 16; ++n2) v[n2] = lds64(ex_u32 + static_cast<uint32_t>((l * kExRow + n2) * 8));
            __syncwarp();
            dft16(v);                                              // over n2; Z[l + 16 k2] at v[P16(k2)]
            // split step (spectrum scaled by 2): with xe2 = Z[k] + conj Z[256-k], xo2 = -i (Z[k] - conj Z[256-k]) and
            // t = W512^k xo2:   X2[k] = xe2 + t   and   X2[256-k] = conj(xe2 - t).
            // A lane takes its own bins with k2 = 0..7 and produces BOTH magnitudes of each pair; the partner lane
            // (16 - l) does the same from its side, which covers this lane's bins with k2 = 8..15.
            const uint32_t mrow = mag_u32 + static_cast<uint32_t>((f * kMagStride + l) * 4);
            const uint32_t mrow_c = mag_u32 + static_cast<uint32_t>((f * kMagStride + 256 - l) * 4);
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) {
                const float2 zk = v[P16(k2)];
                const float2 pub = v[P16(15 - k2)];                // what the partner lane needs from this lane
                float2 zc;
                zc.x = __shfl_sync(0xFFFFFFFFu, pub.x, src_lane);
                zc.y = __shfl_sync(0xFFFFFFFFu, pub.y, src_lane);
                if (l == 0) zc = v[P16((16 - k2) & 15)];           // bins 16 k2 mirror inside lane 0 (bin 0 pairs with 256)
                const float ex_ = zk.x + zc.x, ey = zk.y - zc.y;
                const float ox = zk.y + zc.y, oy = zc.x - zk.x;
                const float2 w = k2 == 0 ? wl : cmul(wl, w32(k2));
                const float tx = w.x * ox - w.y * oy, ty = w.x * oy + w.y * ox;
                const float re = ex_ + tx, im = ey + ty, rc = ex_ - tx, ic = ey - ty;
                sts32f(mrow + 64 * k2, sqrt_approx(re * re + im * im));
                sts32f(mrow_c - 64 * k2, sqrt_approx(rc * rc + ic * ic));
            }
            if (l == 0) {                                           // bin 128 pairs with itself: |X2[128]| = 2 |Z[128]|
                const float2 z = v[P16(8)];
                sts32f(mrow + 64 * 8, 2.f * sqrt_approx(z.x * z.x + z.y * z.y));
            }
        }
        if (tid == 0) tma_store_wait_read<1>();                     // the staging buffer of two tiles ago is free again
        __syncthreads();                                            // magnitudes complete; sample buffer free
        const int t_next = t + gridDim.x;
        TileInfo nxt{};
        if (t_next < segs.total_tiles) {
            nxt = tile_info(t_next);
            issue_load(nxt);                                        // overlaps the mel phase
        }
        // ================================================================= mel + log: lane = frame
        const uint32_t stg = stage_u32 + static_cast<uint32_t>((iter & 1) * kStageBytes);
        if (STATIC_MEL) {
            const float in_scale = cur.pcm16 ? 3.0517578125e-05f : 1.0f;     // 2^-15
            const uint32_t mp = mag_u32 + static_cast<uint32_t>(lane * kMagStride * 4);
            switch (warp) {
                case 0: mel_static_warp<0>(mel, mp, stg, lane, in_scale); break;
                case 1: mel_static_warp<1>(mel, mp, stg, lane, in_scale); break;
                case 2: mel_static_warp<2>(mel, mp, stg, lane, in_scale); break;
                case 3: mel_static_warp<3>(mel, mp, stg, lane, in_scale); break;
                case 4: mel_static_warp<4>(mel, mp, stg, lane, in_scale); break;
                case 5: mel_static_warp<5>(mel, mp, stg, lane, in_scale); break;
                case 6: mel_static_warp<6>(mel, mp, stg, lane, in_scale); break;
                default: mel_static_warp<7>(mel, mp, stg, lane, in_scale); break;
            }
        } else {
            const float in_scale = cur.pcm16 ? 3.0517578125e-05f : 1.0f;     // 2^-15
            const uint32_t mp = mag_u32 + static_cast<uint32_t>(lane * kMagStride * 4);
            const int m_end = mel.warp_band[warp + 1];
            for (int m = mel.warp_band[warp]; m < m_end; ++m) {
                const int gl = mel.glen[m], go = mel.goff[m];
                const uint32_t mb = mp + static_cast<uint32_t>(mel.gbin[m] * 4);
                float acc = 0.f;
                for (int g = 0; g < gl; ++g) {
                    const float4 x = lds128(mb + 16 * g);
                    const float4 w = mel.w4[go + g];
                    acc = fmaf(x.x, w.x, acc);
                    acc = fmaf(x.y, w.y, acc);
                    acc = fmaf(x.z, w.z, acc);
                    acc = fmaf(x.w, w.w, acc);
                }
                const float val = log_fast(fmaf(acc, in_scale, 0.001f));
                const int col = m & 31;
                const uint32_t addr = stg + static_cast<uint32_t>((m >> 5) * 4096 + lane * 128 +
                                                                 ((((col >> 2) ^ (lane & 7))) << 4) + (col & 3) * 4);
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(val) : "memory");
            }
        }
        fence_proxy_async_smem();
        __syncthreads();                                            // staging complete; magnitude tile free
        if (cur.rows_valid == kTileFrames) {
            if (tid == 0) {
                tma_store_2d(&map_out, stg, 0, cur.row0);
                tma_store_2d(&map_out, stg + 4096, 32, cur.row0);
                tma_store_commit();
            }
        } else {
            // a segment shorter than one tile (test entry points only): copy the live rows by hand
            for (int i = tid; i < cur.rows_valid * kMel; i += kFeThreads) {
                const int row = i >> 6, m = i & 63, col = m & 31;
                const uint32_t addr = stg + static_cast<uint32_t>((m >> 5) * 4096 + row * 128 +
                                                                 ((((col >> 2) ^ (row & 7))) << 4) + (col & 3) * 4);
                float val;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(val) : "r"(addr));
                logmel[static_cast<long long>(cur.row0 + row) * kMel + m] = val;
            }
            __syncthreads();
        }
        cur = nxt;
    }
    if (tid == 0) tma_store_wait_all();
}

}  // namespace

cudaError_t frontend2_init_device() {
    cudaError_t e = cudaFuncSetAttribute(logmel2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFe2Smem);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(logmel2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFe2Smem);
}

bool frontend2_build_mel(const FrontendTables& tab, FrontendMelParam* out_raw) {
    static_assert(sizeof(FrontendMelParam) >= sizeof(FeMel), "FrontendMelParam too small");
    FeMel& m = *reinterpret_cast<FeMel*>(out_raw);
    std::memset(&m, 0, sizeof(m));
    int groups = 0, total = 0;
    int cost[kMel];
    for (int b = 0; b < kMel; ++b) {
        const int st = tab.mel_start[b], ln = tab.mel_len[b];
        const int g0 = (st / 4) * 4;
        const int gl = ln > 0 ? (st + ln - g0 + 3) / 4 : 0;
        if (groups + gl > kMelGroupsMax || g0 + 4 * gl > kMagStride) return false;
        m.gbin[b] = g0;
        m.glen[b] = gl;
        m.goff[b] = groups;
        for (int j = 0; j < ln; ++j) {
            const int k = st + j - g0;
            reinterpret_cast<float*>(&m.w4[groups + k / 4])[k % 4] = 0.5f * tab.mel_w[tab.mel_off[b] + j];
        }
        groups += gl;
        cost[b] = 6 * gl + 16;                                     // instructions per band: groups + log / store
        total += cost[b];
    }
    // contiguous band ranges per warp, balanced by work
    int b = 0, acc = 0;
    m.warp_band[0] = 0;
    for (int w = 1; w < kFeWarps; ++w) {
        const int target = static_cast<int>(static_cast<long long>(total) * w / kFeWarps);
        while (b < kMel && acc + cost[b] / 2 < target) { acc += cost[b]; ++b; }
        m.warp_band[w] = b;
    }
    m.warp_band[kFeWarps] = kMel;
    // does the matrix have the structure the unrolled mel phase was generated for?  (bins per band, and which lanes of
    // every four-bin group carry weight; a zero INSIDE a band's run would only cost an FMA by zero, never a wrong sum)
    static_assert(kFeWarps == 8, "mel_layout.inc is generated for eight warps");
    bool same = groups == kMlGroups;
    for (int b = 0; same && b < kMel; ++b) {
        same = m.gbin[b] == kMlGbin[b] && m.glen[b] == kMlGlen[b] && m.goff[b] == kMlGoff[b];
        for (int g = 0; same && g < m.glen[b]; ++g) {
            int mask = 0;
            for (int c = 0; c < 4; ++c) {
                const int k = m.gbin[b] + 4 * g + c;
                if (k >= tab.mel_start[b] && k < tab.mel_start[b] + tab.mel_len[b]) mask |= 1 << c;
            }
            same = mask == kMlMask[m.goff[b] + g];
        }
    }
    static const int static_env = [] { const char* e = getenv("BD_FE_STATIC_MEL"); return e ? atoi(e) : 1; }();
    m.static_ok = (same && static_env != 0) ? 1 : 0;
    return true;
}

cudaError_t launch_logmel_segs(const LogmelSeg* segs, int n_segs, const FrontendMelParam& mel_raw, const float* window,
                               float* logmel, long long logmel_rows, int num_sms, cudaStream_t stream) {
    if (n_segs <= 0) return cudaSuccess;
    if (n_segs > kMaxLogmelSegs) return cudaErrorInvalidValue;
    FeSegs fs;
    fs.n = 0;
    int tiles = 0;
    for (int i = 0; i < n_segs; ++i) {
        if (segs[i].n_rows <= 0) continue;
        if (segs[i].row_begin < 0 || segs[i].row_begin + static_cast<long long>(segs[i].n_rows) > logmel_rows)
            return cudaErrorInvalidValue;
        FeSeg& s = fs.s[fs.n++];
        s.x = segs[i].x;
        s.n_valid = segs[i].n_valid;
        s.frame_begin = segs[i].frame_begin;
        s.row_begin = segs[i].row_begin;
        s.n_rows = segs[i].n_rows;
        s.tile_begin = tiles;
        s.fmt = segs[i].fmt;
        tiles += (segs[i].n_rows + kTileFrames - 1) / kTileFrames;
    }
    if (fs.n == 0) return cudaSuccess;
    for (int i = fs.n; i < kMaxLogmelSegs; ++i) fs.s[i] = FeSeg{nullptr, 0, 0, 0, 0, tiles, 0};
    fs.total_tiles = tiles;
    CUtensorMap map_out;
    if (!encode_store_map_f32(&map_out, logmel, logmel_rows < 32 ? 32 : logmel_rows, kMel, 32)) return cudaErrorUnknown;
    const int grid = tiles < 2 * num_sms ? tiles : 2 * num_sms;
    const FeMel& melp = *reinterpret_cast<const FeMel*>(&mel_raw);
    if (melp.static_ok) logmel2_kernel<true><<<grid, kFeThreads, kFe2Smem, stream>>>(fs, melp, map_out, window, logmel);
    else logmel2_kernel<false><<<grid, kFeThreads, kFe2Smem, stream>>>(fs, melp, map_out, window, logmel);
    return cudaGetLastError();
}

}  // namespace bd
