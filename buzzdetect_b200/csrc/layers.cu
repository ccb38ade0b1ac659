// CUDA-core kernels of the MobileNet-v1 YAMNet stack (sm_100a): first conv, depthwise 3x3, global-average-pool +
// dense head, and a float32 SIMT pointwise GEMM used as the on-device reference precision mode.
//
// Reference semantics: embedders/yamnet/yamnet.py:36-74 (_conv / _separable_conv: conv -> BN(no scale, eps 1e-4)
// -> ReLU), layer table :77-93, GlobalAveragePooling2D :103; models/model_general_v3/model.py:29 (Dense 1024->13).
// TensorFlow SAME padding: stride 1 pads (1,1); stride 2 on even sizes pads (0 before, 1 after) -- SURVEY.md 2a.
// BatchNorm is folded into weights/bias on the host (buzzdetect_b200/weights.py:fold_yamnet).
// Layout: NHWC, H = time, W = mel; activations are [P*H*W, C] row-major so a pointwise conv is a plain GEMM.
#include <cstdlib>

#include "bd_common.cuh"
#include "bd_kernels.cuh"

namespace bd {

namespace {

// ------------------------------------------------------------------------------------------ conv1
// thread = (output pixel, 4-channel group): 9 broadcast input loads, one float4 store.
__global__ void __launch_bounds__(256) conv1_kernel(const float* __restrict__ logmel, int hop_frames, int P,
                                                    const float* __restrict__ w, const float* __restrict__ b,
                                                    float* __restrict__ out) {
    __shared__ float sw[9 * 32];
    __shared__ float sb[32];
    for (int i = threadIdx.x; i < 9 * 32; i += blockDim.x) sw[i] = w[i];
    if (threadIdx.x < 32) sb[threadIdx.x] = b[threadIdx.x];
    __syncthreads();
    // one block = one output row (32 pixels x 8 channel groups) of one patch: no per-thread division
    const int cg = threadIdx.x & 7;
    const int ow = threadIdx.x >> 3;
    for (unsigned bid = blockIdx.x; bid < static_cast<unsigned>(P) * 48u; bid += gridDim.x) {
        const long long p = bid / 48u;
        const int oh = static_cast<int>(bid - static_cast<unsigned>(p) * 48u);
        const long long pix = (p * 48 + oh) * 32 + ow;
        const float* in = logmel + p * hop_frames * kMel;
        float4 acc = make_float4(sb[cg * 4 + 0], sb[cg * 4 + 1], sb[cg * 4 + 2], sb[cg * 4 + 3]);
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            const int ih = 2 * oh + kh;
            if (ih >= kPatchFrames) continue;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int iw = 2 * ow + kw;
                if (iw >= kMel) continue;
                const float v = __ldg(in + ih * kMel + iw);
                const float* wk = sw + (kh * 3 + kw) * 32 + cg * 4;
                acc.x = fmaf(v, wk[0], acc.x);
                acc.y = fmaf(v, wk[1], acc.y);
                acc.z = fmaf(v, wk[2], acc.z);
                acc.w = fmaf(v, wk[3], acc.w);
            }
        }
        acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f);
        *reinterpret_cast<float4*>(out + pix * 32 + cg * 4) = acc;
    }
}

// ------------------------------------------------------------------------------------------ conv1 + depthwise(layer 2)
// Layer 1 (3x3/2 conv, 1 -> 32) and the depthwise half of layer 2 in one kernel: the [48,32,32] layer-1 activation
// (196 KB per patch, the largest tensor of the net after layer-2's output) is never written to or read from HBM.
// One CTA = 8 output rows x 32 columns x 32 channels of one patch:
//   1. stage the 21 x 64 log-mel rows the tile needs in shared memory (zero beyond row 95 / column 63: conv1's SAME pad)
//   2. compute the 10 x 34 x 32 layer-1 tile (1-pixel halo; ZERO outside the 48 x 32 image: the depthwise SAME pad)
//   3. depthwise 3x3 stride 1 from shared memory, +bias, ReLU, emit fp32 / fp16 hi / hi+lo planes
constexpr int kC1Rows = 8;                                   // output rows per CTA (48 = 6 tiles)
constexpr int kC1TileH = kC1Rows + 2, kC1TileW = 34;
constexpr int kC1LmRows = 2 * kC1TileH + 1, kC1LmStride = 66;
constexpr int kC1LmFloats = (kC1LmRows * kC1LmStride + 3) & ~3;          // keeps the float4 weight table 16-byte aligned
constexpr int kC1SmemBytes = (kC1TileH * kC1TileW * 32 + kC1LmFloats + 9 * 32 + 32) * 4;

template <int OUT_MODE>
__global__ void __launch_bounds__(256) conv1_dw2_kernel(const float* __restrict__ logmel, int hop_frames, int P,
                                                        const float* __restrict__ w1, const float* __restrict__ b1,
                                                        const float* __restrict__ dw_w, const float* __restrict__ dw_b,
                                                        float* __restrict__ out_f32, __half* __restrict__ out_hi,
                                                        __half* __restrict__ out_lo) {
    extern __shared__ __align__(16) float c1_smem[];
    float* c1 = c1_smem;                                     // [10][34][32]
    float* lm = c1 + kC1TileH * kC1TileW * 32;               // [21][66]
    float* sw = lm + kC1LmFloats;                            // [9][32]
    float* sb = sw + 9 * 32;                                 // [32]
    const int tid = threadIdx.x;
    for (int i = tid; i < 9 * 32; i += 256) sw[i] = w1[i];
    if (tid < 32) sb[tid] = b1[tid];
    const int cg = tid & 7;                                  // channel quad (both phases)
    const float4 bdw = __ldg(reinterpret_cast<const float4*>(dw_b + cg * 4));

    const unsigned tiles_per_patch = 48 / kC1Rows;
    for (unsigned bid = blockIdx.x; bid < static_cast<unsigned>(P) * tiles_per_patch; bid += gridDim.x) {
        const long long p = bid / tiles_per_patch;
        const int r0 = static_cast<int>(bid - static_cast<unsigned>(p) * tiles_per_patch) * kC1Rows;
        const float* in = logmel + p * hop_frames * kMel;
        __syncthreads();                                     // previous tile fully consumed (and sw/sb visible)
        // ---- 1. log-mel rows 2*(r0-1) .. 2*(r0-1)+20, columns 0..64 (column 64 and rows >= 96 are padding)
        const int lm_row0 = 2 * (r0 - 1);
        for (int i = tid; i < kC1LmRows * 65; i += 256) {
            const int rr = i / 65, cc = i - rr * 65;
            const int gr = lm_row0 + rr;
            lm[rr * kC1LmStride + cc] = (gr >= 0 && gr < kPatchFrames && cc < kMel) ? __ldg(in + gr * kMel + cc) : 0.f;
        }
        __syncthreads();
        // ---- 2. layer-1 tile with halo: local (tr, tc) <-> image (r0 - 1 + tr, tc - 1); tap weights in registers
        float4 w1r[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) w1r[i] = *reinterpret_cast<const float4*>(sw + i * 32 + cg * 4);
        const float4 b1r = *reinterpret_cast<const float4*>(sb + cg * 4);
        for (int i = tid; i < kC1TileH * kC1TileW * 8; i += 256) {
            const int px = i >> 3;                           // cg == i & 7 == tid & 7 (256 % 8 == 0)
            const int tr = px / kC1TileW, tc = px - tr * kC1TileW;
            const int ir = r0 - 1 + tr, ic = tc - 1;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ir >= 0 && ir < 48 && ic >= 0 && ic < 32) {
                a = b1r;
                const float* l0 = lm + (2 * tr) * kC1LmStride + 2 * ic;      // log-mel row 2*ir - lm_row0 = 2*tr
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        const float v = l0[kh * kC1LmStride + kw];
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
        __syncthreads();
        // ---- 3. depthwise 3x3 stride 1: thread = (row, strip of 4 columns, channel quad); 512 items, 2 per thread
        float4 kk[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) kk[i] = __ldg(reinterpret_cast<const float4*>(dw_w + i * 32 + cg * 4));
#pragma unroll 1
        for (int it = 0; it < 2; ++it) {
            const int item = tid + 256 * it;
            const int ws = (item >> 3) & 7, orow = item >> 6;          // strip 0..7, local output row 0..7
            float4 acc[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[r] = bdw;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const float* rowp = c1 + ((orow + kh) * kC1TileW + ws * 4) * 32 + cg * 4;   // tile col = image col + 1 - 1 + kw
                float4 v[6];
#pragma unroll
                for (int j = 0; j < 6; ++j) v[j] = *reinterpret_cast<const float4*>(rowp + j * 32);
#pragma unroll
                for (int r = 0; r < 4; ++r) {
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        const float4 x = v[r + kw];
                        const float4 w4 = kk[kh * 3 + kw];
                        acc[r].x = fmaf(x.x, w4.x, acc[r].x);
                        acc[r].y = fmaf(x.y, w4.y, acc[r].y);
                        acc[r].z = fmaf(x.z, w4.z, acc[r].z);
                        acc[r].w = fmaf(x.w, w4.w, acc[r].w);
                    }
                }
            }
            const long long pix0 = (p * 48 + r0 + orow) * 32 + ws * 4;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                float4 a = acc[r];
                a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f);
                const long long o = (pix0 + r) * 32 + cg * 4;
                if (OUT_MODE == 0) {
                    *reinterpret_cast<float4*>(out_f32 + o) = a;
                } else {
                    __half2 hp[2] = {half2_sat(a.x, a.y), half2_sat(a.z, a.w)};
                    const __half h0 = __low2half(hp[0]), h1 = __high2half(hp[0]);
                    const __half h2 = __low2half(hp[1]), h3 = __high2half(hp[1]);
                    *reinterpret_cast<uint2*>(out_hi + o) = *reinterpret_cast<uint2*>(hp);
                    if (OUT_MODE == 2) {
                        __half2 lp[2] = {half2_sat(a.x - __half2float(h0), a.y - __half2float(h1)),
                                         half2_sat(a.z - __half2float(h2), a.w - __half2float(h3))};
                        *reinterpret_cast<uint2*>(out_lo + o) = *reinterpret_cast<uint2*>(lp);
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ depthwise
// thread = (strip of R consecutive output pixels along W, 4 channels).  The 3 x ((R-1)*STRIDE+3) input window and
// the 9 weight vectors live in registers, so each input float4 is loaded once per strip instead of once per tap
// (the per-tap version was L1-bandwidth bound: 18 LDG.128 per output vector; this one issues 6.75 for R = 4).
// Channel-contiguous float4 accesses: consecutive threads take consecutive channel quads of the same strip, so
// every warp-level load/store touches whole 128-byte lines.
// HH / WW / CH > 0 fix the layer's extent at compile time (the three stand-alone layers of YAMNet: 12x8x256, 6x4x512,
// 3x2x1024): every load offset becomes an immediate and the bounds tests fold away -- the generic instantiation spends
// a third of its instructions on 64-bit address arithmetic (179 IMAD + 99 IADD3 + 63 LEA against 144 FFMA).
template <int STRIDE, int R, int OUT_MODE, int HH = 0, int WW = 0, int CH = 0>
__global__ void __launch_bounds__(256) depthwise_kernel(const float* __restrict__ in, int P, int H_, int W_, int C_,
                                                        const float* __restrict__ w, const float* __restrict__ b,
                                                        float* __restrict__ out_f32, __half* __restrict__ out_hi,
                                                        __half* __restrict__ out_lo) {
    const int H = HH > 0 ? HH : H_, W = WW > 0 ? WW : W_, C = CH > 0 ? CH : C_;
    // C/4 and Wo/R are powers of two for every YAMNet layer, so (channel quad, strip, row) come from shifts and masks;
    // the only division is one 32-bit block-uniform one.  (64-bit div/mod per thread cost more than the 36 FMAs.)
    const int Ho = H / STRIDE, Wo = W / STRIDE;
    constexpr int PB = STRIDE == 1 ? 1 : 0;
    constexpr int NC = (R - 1) * STRIDE + 3;
    const unsigned c4_bits = 31u - __clz(static_cast<unsigned>(C >> 2));
    const unsigned ws_bits = 31u - __clz(static_cast<unsigned>(Wo / R));
    const unsigned items = static_cast<unsigned>(Ho) << (ws_bits + c4_bits);       // per patch
    const unsigned bpp = (items + blockDim.x - 1) / blockDim.x;                     // blocks per patch
    for (unsigned bid = blockIdx.x; bid < static_cast<unsigned>(P) * bpp; bid += gridDim.x) {
        const unsigned pi = bid / bpp;
        const unsigned item = (bid - pi * bpp) * blockDim.x + threadIdx.x;
        if (item >= items) continue;
        const int c4 = static_cast<int>(item & ((1u << c4_bits) - 1u));
        const int ws = static_cast<int>((item >> c4_bits) & ((1u << ws_bits) - 1u));
        const int oh = static_cast<int>(item >> (c4_bits + ws_bits));
        const long long p = pi;
        const int ow0 = ws * R;
        const float* inp = in + p * H * W * C + c4 * 4;
        float4 k[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) k[i] = __ldg(reinterpret_cast<const float4*>(w + i * C + c4 * 4));
        const float4 bias = __ldg(reinterpret_cast<const float4*>(b + c4 * 4));
        float4 acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = bias;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            const int ih = oh * STRIDE + kh - PB;
            if (ih < 0 || ih >= H) continue;
            const float* rowp = inp + static_cast<long long>(ih) * W * C;
            float4 v[NC];
#pragma unroll
            for (int j = 0; j < NC; ++j) {
                const int iw = ow0 * STRIDE - PB + j;
                v[j] = (iw >= 0 && iw < W) ? __ldg(reinterpret_cast<const float4*>(rowp + static_cast<long long>(iw) * C))
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const float4 x = v[r * STRIDE + kw];
                    const float4 kk = k[kh * 3 + kw];
                    acc[r].x = fmaf(x.x, kk.x, acc[r].x);
                    acc[r].y = fmaf(x.y, kk.y, acc[r].y);
                    acc[r].z = fmaf(x.z, kk.z, acc[r].z);
                    acc[r].w = fmaf(x.w, kk.w, acc[r].w);
                }
            }
        }
        const long long pix0 = (p * Ho + oh) * Wo + ow0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float4 a = acc[r];
            a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f);
            const long long o = (pix0 + r) * C + c4 * 4;
            if (OUT_MODE == 0) {
                *reinterpret_cast<float4*>(out_f32 + o) = a;
            } else {
                __half2 hp[2] = {half2_sat(a.x, a.y), half2_sat(a.z, a.w)};
                const __half h0 = __low2half(hp[0]), h1 = __high2half(hp[0]);
                const __half h2 = __low2half(hp[1]), h3 = __high2half(hp[1]);
                *reinterpret_cast<uint2*>(out_hi + o) = *reinterpret_cast<uint2*>(hp);
                if (OUT_MODE == 2) {
                    __half2 lp[2] = {
                        half2_sat(a.x - __half2float(h0), a.y - __half2float(h1)),
                        half2_sat(a.z - __half2float(h2), a.w - __half2float(h3))};
                    *reinterpret_cast<uint2*>(out_lo + o) = *reinterpret_cast<uint2*>(lp);
                }
                if (OUT_MODE == 3) {
                    // fp16 + fp8 plan: second plane = e5m2 bytes [M, 2C], per 64-channel k-block [lo * 2^11 | hi]
                    const float f0 = __half2float(h0), f1 = __half2float(h1), f2 = __half2float(h2), f3 = __half2float(h3);
                    const int c = c4 * 4;
                    unsigned char* row = reinterpret_cast<unsigned char*>(out_lo) + (pix0 + r) * 2 * C + (c >> 6) * 128 + (c & 63);
                    *reinterpret_cast<uint32_t*>(row) = pack_e5m2x4((a.x - f0) * kF8LoScale, (a.y - f1) * kF8LoScale,
                                                                    (a.z - f2) * kF8LoScale, (a.w - f3) * kF8LoScale);
                    *reinterpret_cast<uint32_t*>(row + 64) = pack_e5m2x4(f0, f1, f2, f3);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ depthwise, 2-row blocks
// Same op, thread = (block of R x 2 output pixels, 4 channels): the (2*STRIDE+1 | 4) x NC input window is walked row by
// row, every row feeding both output rows, so an output vector costs 3 (stride 1) instead of 4.5 global loads and the
// nine tap vectors are fetched once per 2R instead of once per R pixels.  Used when the output height is even; the
// L1/texture data pipe (one 128-byte wavefront per clock per SM), not HBM, is what bounds these kernels.
template <int STRIDE, int R, int OUT_MODE>
__global__ void __launch_bounds__(256) depthwise2_kernel(const float* __restrict__ in, int P, int H, int W, int C,
                                                         const float* __restrict__ w, const float* __restrict__ b,
                                                         float* __restrict__ out_f32, __half* __restrict__ out_hi,
                                                         __half* __restrict__ out_lo) {
    const int Ho = H / STRIDE, Wo = W / STRIDE;
    constexpr int PB = STRIDE == 1 ? 1 : 0;
    constexpr int NC = (R - 1) * STRIDE + 3;
    constexpr int NR = STRIDE + 3;                                                  // input rows per block
    const unsigned c4_bits = 31u - __clz(static_cast<unsigned>(C >> 2));
    const unsigned ws_bits = 31u - __clz(static_cast<unsigned>(Wo / R));
    const unsigned items = static_cast<unsigned>(Ho >> 1) << (ws_bits + c4_bits);  // per patch
    const unsigned bpp = (items + blockDim.x - 1) / blockDim.x;
    for (unsigned bid = blockIdx.x; bid < static_cast<unsigned>(P) * bpp; bid += gridDim.x) {
        const unsigned pi = bid / bpp;
        const unsigned item = (bid - pi * bpp) * blockDim.x + threadIdx.x;
        if (item >= items) continue;
        const int c4 = static_cast<int>(item & ((1u << c4_bits) - 1u));
        const int ws = static_cast<int>((item >> c4_bits) & ((1u << ws_bits) - 1u));
        const int oh0 = static_cast<int>(item >> (c4_bits + ws_bits)) * 2;
        const long long p = pi;
        const int ow0 = ws * R;
        const float* inp = in + p * H * W * C + c4 * 4;
        float4 k[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) k[i] = __ldg(reinterpret_cast<const float4*>(w + i * C + c4 * 4));
        const float4 bias = __ldg(reinterpret_cast<const float4*>(b + c4 * 4));
        float4 acc[2][R];
#pragma unroll
        for (int o = 0; o < 2; ++o)
#pragma unroll
            for (int r = 0; r < R; ++r) acc[o][r] = bias;
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            const int ih = oh0 * STRIDE + i - PB;
            if (ih < 0 || ih >= H) continue;
            const float* rowp = inp + static_cast<long long>(ih) * W * C;
            float4 v[NC];
#pragma unroll
            for (int j = 0; j < NC; ++j) {
                const int iw = ow0 * STRIDE - PB + j;
                v[j] = (iw >= 0 && iw < W) ? __ldg(reinterpret_cast<const float4*>(rowp + static_cast<long long>(iw) * C))
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int o = 0; o < 2; ++o) {
                const int kh = i - o * STRIDE;
                if (kh < 0 || kh > 2) continue;
#pragma unroll
                for (int r = 0; r < R; ++r) {
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        const float4 x = v[r * STRIDE + kw];
                        const float4 kk = k[kh * 3 + kw];
                        acc[o][r].x = fmaf(x.x, kk.x, acc[o][r].x);
                        acc[o][r].y = fmaf(x.y, kk.y, acc[o][r].y);
                        acc[o][r].z = fmaf(x.z, kk.z, acc[o][r].z);
                        acc[o][r].w = fmaf(x.w, kk.w, acc[o][r].w);
                    }
                }
            }
        }
#pragma unroll
        for (int o = 0; o < 2; ++o) {
            const long long pix0 = (p * Ho + oh0 + o) * Wo + ow0;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float4 a = acc[o][r];
                a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f);
                const long long oo = (pix0 + r) * C + c4 * 4;
                if (OUT_MODE == 0) {
                    *reinterpret_cast<float4*>(out_f32 + oo) = a;
                } else {
                    __half2 hp[2] = {half2_sat(a.x, a.y), half2_sat(a.z, a.w)};
                    const __half h0 = __low2half(hp[0]), h1 = __high2half(hp[0]);
                    const __half h2 = __low2half(hp[1]), h3 = __high2half(hp[1]);
                    *reinterpret_cast<uint2*>(out_hi + oo) = *reinterpret_cast<uint2*>(hp);
                    if (OUT_MODE == 2) {
                        __half2 lp[2] = {
                            half2_sat(a.x - __half2float(h0), a.y - __half2float(h1)),
                            half2_sat(a.z - __half2float(h2), a.w - __half2float(h3))};
                        *reinterpret_cast<uint2*>(out_lo + oo) = *reinterpret_cast<uint2*>(lp);
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ SIMT fp32 GEMM
// C[M,N] = relu(A[M,K] * Bt[N,K]^T + bias[N]); 64x64 tile, 16-wide K slabs, 4x4 outputs per thread.
__global__ void __launch_bounds__(256) pw_simt_kernel(const float* __restrict__ A, const float* __restrict__ Bt,
                                                      const float* __restrict__ bias, float* __restrict__ C, int M,
                                                      int N, int K) {
    __shared__ float sA[16][64 + 4];
    __shared__ float sB[16][64 + 4];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += 16) {
        for (int i = threadIdx.x; i < 64 * 16; i += 256) {
            const int r = i >> 4, kk = i & 15;
            const int m = m0 + r, n = n0 + r;
            sA[kk][r] = (m < M && k0 + kk < K) ? A[static_cast<long long>(m) * K + k0 + kk] : 0.f;
            sB[kk][r] = (n < N && k0 + kk < K) ? Bt[static_cast<long long>(n) * K + k0 + kk] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float a[4], bb[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = sA[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bb[j] = sB[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < N) C[static_cast<long long>(m) * N + n] = fmaxf(acc[i][j] + bias[n], 0.f);
        }
    }
}

// ------------------------------------------------------------------------------------------ pool + head
// Global average pool over the rows of a patch (sum in row order, then / rows: tf.reduce_mean of the graph) and the
// Dense(n_classes) head, yamnet.py:96-101 + models/model_general_v3/model.py:16,29.
// Persistent CTAs of 256 threads x 4 channels = 1024 embedding dims walk the patches: the thread's slice of the head
// matrix (4 x n_classes weights) is loaded once into registers, and the rows of the NEXT patch are in flight while the
// current one is reduced (one CTA per patch re-read the strided head weights for every patch and sat at 1.3 TB/s).
template <int NC>
__global__ void __launch_bounds__(256) pool_head_kernel(const float* __restrict__ y, int P, int rows, const float* __restrict__ Wh,
                                                        const float* __restrict__ bh, int n_classes,
                                                        float* __restrict__ emb, float* __restrict__ act) {
    constexpr int kMaxRows = 8;
    __shared__ float red[2][8][NC];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    float w[4][NC];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NC; ++j) w[i][j] = j < n_classes ? __ldg(Wh + static_cast<long long>(t * 4 + i) * n_classes + j) : 0.f;
    const float bias = t < n_classes ? __ldg(bh + t) : 0.f;
    const float d = static_cast<float>(rows);
    float4 v[kMaxRows];
    auto load = [&](long long p) {
        const float* yp = y + p * rows * kEmb + t * 4;
#pragma unroll
        for (int r = 0; r < kMaxRows; ++r)
            if (r < rows) v[r] = __ldg(reinterpret_cast<const float4*>(yp + static_cast<long long>(r) * kEmb));
    };
    long long p = blockIdx.x;
    if (p < P) load(p);
    int buf = 0;
    for (; p < P; p += gridDim.x, buf ^= 1) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < kMaxRows; ++r)
            if (r < rows) { s.x += v[r].x; s.y += v[r].y; s.z += v[r].z; s.w += v[r].w; }
        if (p + gridDim.x < P) load(p + gridDim.x);                 // next patch's rows fly behind the reduction
        s.x /= d; s.y /= d; s.z /= d; s.w /= d;
        if (emb != nullptr) *reinterpret_cast<float4*>(emb + p * kEmb + t * 4) = s;
#pragma unroll
        for (int j = 0; j < NC; ++j) {
            float a = s.x * w[0][j];
            a = fmaf(s.y, w[1][j], a);
            a = fmaf(s.z, w[2][j], a);
            a = fmaf(s.w, w[3][j], a);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (lane == 0) red[buf][warp][j] = a;
        }
        __syncthreads();                                            // red[buf] complete; red[buf ^ 1] is free again
        if (t < n_classes) {
            float a = 0.f;
#pragma unroll
            for (int wi = 0; wi < 8; ++wi) a += red[buf][wi][t];
            act[p * n_classes + t] = a + bias;
        }
    }
}

// experiment switch: BD_DW_TWO_ROWS=1 selects the two-row blocks.  Measured on B200: 102 registers cost a resident CTA
// per SM and the 6x4 layers get SLOWER (0.094 -> 0.119 ms per audio-hour), so one-row strips stay the default.
inline bool dw_two_rows_enabled() {
    static const int v = [] { const char* e = getenv("BD_DW_TWO_ROWS"); return e ? atoi(e) : 0; }();
    return v != 0;
}

inline int grid_for(long long total, int block, int cap) {
    long long g = (total + block - 1) / block;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return static_cast<int>(g);
}

}  // namespace

cudaError_t launch_conv1(const float* logmel, int hop_frames, int P, const float* w, const float* b, float* out,
                         cudaStream_t stream) {
    if (P <= 0) return cudaSuccess;
    const long long blocks = static_cast<long long>(P) * 48;
    if (blocks >= (1LL << 31)) return cudaErrorInvalidValue;
    conv1_kernel<<<static_cast<int>(blocks < 148LL * 64 ? blocks : 148LL * 64), 256, 0, stream>>>(logmel, hop_frames, P,
                                                                                                  w, b, out);
    return cudaGetLastError();
}

cudaError_t layers_init_device() {
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(conv1_dw2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kC1SmemBytes)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(conv1_dw2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kC1SmemBytes)) != cudaSuccess) return e;
    return cudaFuncSetAttribute(conv1_dw2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kC1SmemBytes);
}

cudaError_t launch_conv1_dw2(const float* logmel, int hop_frames, int P, const float* w1, const float* b1,
                             const float* dw_w, const float* dw_b, int out_mode, float* out_f32, __half* out_hi,
                             __half* out_lo, cudaStream_t stream) {
    if (P <= 0) return cudaSuccess;
    if (out_mode < 0 || out_mode > 2) return cudaErrorInvalidValue;
    const long long blocks = static_cast<long long>(P) * (48 / kC1Rows);
    if (blocks >= (1LL << 31)) return cudaErrorInvalidValue;
    const int grid = static_cast<int>(blocks < 148LL * 32 ? blocks : 148LL * 32);
    if (out_mode == 0)
        conv1_dw2_kernel<0><<<grid, 256, kC1SmemBytes, stream>>>(logmel, hop_frames, P, w1, b1, dw_w, dw_b, out_f32, out_hi, out_lo);
    else if (out_mode == 1)
        conv1_dw2_kernel<1><<<grid, 256, kC1SmemBytes, stream>>>(logmel, hop_frames, P, w1, b1, dw_w, dw_b, out_f32, out_hi, out_lo);
    else
        conv1_dw2_kernel<2><<<grid, 256, kC1SmemBytes, stream>>>(logmel, hop_frames, P, w1, b1, dw_w, dw_b, out_f32, out_hi, out_lo);
    return cudaGetLastError();
}

cudaError_t launch_depthwise(const float* in, int P, int H, int W, int C, int stride, const float* w, const float* b,
                             int out_mode, float* out_f32, __half* out_hi, __half* out_lo, cudaStream_t stream) {
    if (P <= 0) return cudaSuccess;
    if ((C & 3) || (stride != 1 && stride != 2) || (stride == 2 && ((H | W) & 1))) return cudaErrorInvalidValue;
    if (out_mode < 0 || out_mode > 3 || (out_mode == 3 && C % 64 != 0)) return cudaErrorInvalidValue;
    const int Wo = W / stride;
    const int R = (Wo % 4 == 0) ? 4 : ((Wo % 2 == 0) ? 2 : 1);
    const unsigned c4 = static_cast<unsigned>(C / 4), wsn = static_cast<unsigned>(Wo / R);
    if ((c4 & (c4 - 1)) || (wsn & (wsn - 1))) return cudaErrorInvalidValue;      // shift/mask indexing (see kernel)
    const bool two_rows = ((H / stride) % 2 == 0) && (R == 4 || R == 2) && dw_two_rows_enabled() && out_mode != 3;
    const long long items = static_cast<long long>(H / stride) / (two_rows ? 2 : 1) * wsn * c4;
    const long long blocks = static_cast<long long>(P) * ((items + 255) / 256);
    if (blocks >= (1LL << 31)) return cudaErrorInvalidValue;
    const int grid = static_cast<int>(blocks < 148LL * 64 ? blocks : 148LL * 64);
    if (two_rows) {
#define BD_DW2(S, RR, MODE) \
    depthwise2_kernel<S, RR, MODE><<<grid, 256, 0, stream>>>(in, P, H, W, C, w, b, out_f32, out_hi, out_lo)
#define BD_DW2_MODE(S, RR)                          \
    do {                                            \
        if (out_mode == 0) BD_DW2(S, RR, 0);        \
        else if (out_mode == 1) BD_DW2(S, RR, 1);   \
        else BD_DW2(S, RR, 2);                      \
    } while (0)
        if (stride == 1) { if (R == 4) BD_DW2_MODE(1, 4); else BD_DW2_MODE(1, 2); }
        else { if (R == 4) BD_DW2_MODE(2, 4); else BD_DW2_MODE(2, 2); }
#undef BD_DW2_MODE
#undef BD_DW2
        return cudaGetLastError();
    }
    // the three stand-alone layers of the default plan, with compile-time extents
#define BD_DW_FIXED(S, RR, HH, WW, CH)                                                                                   \
    if (stride == S && R == RR && H == HH && W == WW && C == CH) {                                                       \
        if (out_mode == 0) depthwise_kernel<S, RR, 0, HH, WW, CH><<<grid, 256, 0, stream>>>(in, P, H, W, C, w, b, out_f32, out_hi, out_lo); \
        else if (out_mode == 1) depthwise_kernel<S, RR, 1, HH, WW, CH><<<grid, 256, 0, stream>>>(in, P, H, W, C, w, b, out_f32, out_hi, out_lo); \
        else if (out_mode == 3) depthwise_kernel<S, RR, 3, HH, WW, CH><<<grid, 256, 0, stream>>>(in, P, H, W, C, w, b, out_f32, out_hi, out_lo); \
        else depthwise_kernel<S, RR, 2, HH, WW, CH><<<grid, 256, 0, stream>>>(in, P, H, W, C, w, b, out_f32, out_hi, out_lo); \
        return cudaGetLastError();                                                                                       \
    }
    static const int fixed_env = [] { const char* e = getenv("BD_DW_FIXED"); return e ? atoi(e) : 1; }();
    if (fixed_env != 0) {
        BD_DW_FIXED(2, 4, 12, 8, 256)
        BD_DW_FIXED(2, 2, 6, 4, 512)
        BD_DW_FIXED(1, 2, 3, 2, 1024)
    }
#undef BD_DW_FIXED
#define BD_DW(S, RR, MODE) \
    depthwise_kernel<S, RR, MODE><<<grid, 256, 0, stream>>>(in, P, H, W, C, w, b, out_f32, out_hi, out_lo)
#define BD_DW_MODE(S, RR)                          \
    do {                                           \
        if (out_mode == 0) BD_DW(S, RR, 0);        \
        else if (out_mode == 1) BD_DW(S, RR, 1);   \
        else if (out_mode == 3) BD_DW(S, RR, 3);   \
        else BD_DW(S, RR, 2);                      \
    } while (0)
    if (stride == 1) {
        if (R == 4) BD_DW_MODE(1, 4); else if (R == 2) BD_DW_MODE(1, 2); else BD_DW_MODE(1, 1);
    } else {
        if (R == 4) BD_DW_MODE(2, 4); else if (R == 2) BD_DW_MODE(2, 2); else BD_DW_MODE(2, 1);
    }
#undef BD_DW_MODE
#undef BD_DW
    return cudaGetLastError();
}

cudaError_t launch_pw_simt(const float* A, const float* Bt, const float* bias, float* C, int M, int N, int K,
                           cudaStream_t stream) {
    if (M <= 0) return cudaSuccess;
    dim3 grid((M + 63) / 64, (N + 63) / 64);
    pw_simt_kernel<<<grid, 256, 0, stream>>>(A, Bt, bias, C, M, N, K);
    return cudaGetLastError();
}

cudaError_t launch_pool_head(const float* y, int P, int rows_per_patch, const float* Wh, const float* bh,
                             int n_classes, float* emb, float* act, cudaStream_t stream) {
    if (P <= 0) return cudaSuccess;
    if (n_classes > kMaxClasses || n_classes < 1) return cudaErrorInvalidValue;
    if (rows_per_patch > 8) return cudaErrorInvalidValue;
    static const int num_sms = [] { int dev = 0, n = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); return n; }();
    const int grid = P < num_sms * 2 ? P : num_sms * 2;
    if (n_classes <= 16) pool_head_kernel<16><<<grid, 256, 0, stream>>>(y, P, rows_per_patch, Wh, bh, n_classes, emb, act);
    else pool_head_kernel<kMaxClasses><<<grid, 256, 0, stream>>>(y, P, rows_per_patch, Wh, bh, n_classes, emb, act);
    return cudaGetLastError();
}

}  // namespace bd
