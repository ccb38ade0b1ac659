// Fused YAMNet frontend for sm_100a: frame(400/160) * periodic Hann -> zero-pad to 512 -> real FFT -> |.| ->
// sparse mel (64 bands) -> log(x + 0.001), one kernel, audio read once from HBM, log-mel written once.
//
// Reference semantics: embedders/yamnet/features.py:22-79 (waveform_to_log_mel_spectrogram_patches) and the
// op list of embedders/yamnet_k2/models/yamnet_wholehop/saved_model.pb:
// tf.signal.frame -> * hann_window(periodic) -> Pad[[0,0],[0,112]] -> RFFT(512) -> ComplexAbs -> MatMul(mel)
// -> AddV2(0.001) -> Log.  The tail zero padding of pad_waveform (features.py:82-108) is virtual here: samples at
// index >= n_valid read as 0, nothing is materialised.
//
// Mapping: one CTA stages a tile of 32 consecutive STFT frames (5360 samples, 21 KB) in shared memory with
// coalesced float4 loads; each of its 8 warps then transforms 4 frames.  A 512-point real FFT is a 256-point
// complex FFT of z[n] = x[2n] + i x[2n+1] plus a split step; the complex FFT runs as radix 8 x 4 x 8 with the 256
// points held 8 per lane, two padded shared-memory exchanges between the passes (conflict-free strides 40 / 9).
#include "bd_kernels.cuh"

namespace bd {

namespace {

constexpr int kTileFrames = 32;
constexpr int kWarps = 8;
constexpr int kTileSamples = (kTileFrames - 1) * kHop + kWin;   // 5360
constexpr int kS1 = 8 * 40;                                     // exchange 1 / Z buffer (float2)
constexpr int kS2 = 32 * 9;                                     // exchange 2 / magnitude buffer (float2)

struct __align__(16) FeSmem {
    float samples[kTileSamples];          // 21440 B
    float win[kWin];                      // 1600 B
    float2 tw512[kBins + 1];              // exp(-2 pi i k / 512), k = 0..256 (+1 pad)
    int mel_start[kMel];
    int mel_len[kMel];
    int mel_off[kMel];
    float mel_w[kMelNnzMax];
    float2 scratch[kWarps][kS1 + kS2];
};

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 mul_neg_i(float2 a) { return make_float2(a.y, -a.x); }

// out[k] = sum_n c[n] * exp(-2 pi i n k / 4)
__device__ __forceinline__ void dft4(float2& c0, float2& c1, float2& c2, float2& c3) {
    float2 e0 = cadd(c0, c2), e1 = csub(c0, c2);
    float2 f0 = cadd(c1, c3), f1 = mul_neg_i(csub(c1, c3));
    c0 = cadd(e0, f0);
    c1 = cadd(e1, f1);
    c2 = csub(e0, f0);
    c3 = csub(e1, f1);
}

// in place: v[k] = sum_n v[n] * exp(-2 pi i n k / 8)   (decimation in frequency, natural-order output)
__device__ __forceinline__ void dft8(float2 (&v)[8]) {
    const float h = 0.70710678118654752f;
    float2 a0 = cadd(v[0], v[4]), a1 = cadd(v[1], v[5]), a2 = cadd(v[2], v[6]), a3 = cadd(v[3], v[7]);
    float2 b0 = csub(v[0], v[4]), b1 = csub(v[1], v[5]), b2 = csub(v[2], v[6]), b3 = csub(v[3], v[7]);
    b1 = make_float2((b1.x + b1.y) * h, (b1.y - b1.x) * h);      // * (1 - i)/sqrt2
    b2 = mul_neg_i(b2);                                          // * (-i)
    b3 = make_float2((b3.y - b3.x) * h, -(b3.x + b3.y) * h);     // * (-1 - i)/sqrt2
    dft4(a0, a1, a2, a3);
    dft4(b0, b1, b2, b3);
    v[0] = a0; v[2] = a1; v[4] = a2; v[6] = a3;
    v[1] = b0; v[3] = b1; v[5] = b2; v[7] = b3;
}

__global__ void __launch_bounds__(kWarps * 32, 3) logmel_kernel(const float* __restrict__ x, long long n_valid,
                                                             long long frame_begin, int n_frames,
                                                             const FrontendTables* __restrict__ tab,
                                                             float* __restrict__ logmel) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FeSmem& s = *reinterpret_cast<FeSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- per-CTA tables
    for (int i = tid; i < kWin; i += blockDim.x) s.win[i] = tab->window[i];
    for (int i = tid; i <= kBins; i += blockDim.x) {
        float sn, cs;
        sincospif(-static_cast<float>(i) / 256.0f, &sn, &cs);
        s.tw512[i] = make_float2(cs, sn);
    }
    for (int i = tid; i < kMel; i += blockDim.x) {
        s.mel_start[i] = tab->mel_start[i];
        s.mel_len[i] = tab->mel_len[i];
        s.mel_off[i] = tab->mel_off[i];
    }
    for (int i = tid; i < kMelNnzMax; i += blockDim.x) s.mel_w[i] = tab->mel_w[i];

    // ---- per-lane twiddles (registers)
    float2 tw1[8];   // exp(-2 pi i lane*k1/256)
#pragma unroll
    for (int k1 = 0; k1 < 8; ++k1) {
        float sn, cs;
        sincospif(-static_cast<float>(lane * k1) / 128.0f, &sn, &cs);
        tw1[k1] = make_float2(cs, sn);
    }
    const int m2 = lane & 7;
    float2 tw2[4];   // exp(-2 pi i m2*q1/32)
#pragma unroll
    for (int q1 = 0; q1 < 4; ++q1) {
        float sn, cs;
        sincospif(-static_cast<float>(m2 * q1) / 16.0f, &sn, &cs);
        tw2[q1] = make_float2(cs, sn);
    }

    __syncthreads();                                                     // s.win / s.tw512 visible
    // this lane's window taps (elements 2c, 2c+1 with c = 32 j + lane; zero beyond sample 399) and split-step
    // twiddles W512^(lane + 32 j): registers instead of 15 shared-memory loads per frame (the kernel is bound by
    // shared-memory wavefronts, profiles/r1_summary.md)
    float2 wreg[7], twreg[8];
#pragma unroll
    for (int j = 0; j < 7; ++j) {
        const int c = 32 * j + lane;
        wreg[j] = (j < 6 || lane < 8) ? *reinterpret_cast<const float2*>(s.win + 2 * c) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) twreg[j] = s.tw512[lane + 32 * j];

    float2* s1 = s.scratch[warp];
    float2* s2 = s1 + kS1;
    float* mag = reinterpret_cast<float*>(s2);
    const bool aligned16 = (reinterpret_cast<uintptr_t>(x) & 15u) == 0;
    const int n_tiles = (n_frames + kTileFrames - 1) / kTileFrames;

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int f0 = tile * kTileFrames;                               // local frame index
        const long long g0 = (frame_begin + f0) * static_cast<long long>(kHop);   // first sample of the tile
        __syncthreads();                                                 // previous tile fully consumed
        if (aligned16) {
            for (int i = tid * 4; i < kTileSamples; i += blockDim.x * 4) {
                const long long g = g0 + i;
                float4 v;
                if (g + 3 < n_valid) {
                    v = __ldg(reinterpret_cast<const float4*>(x + g));
                } else {
                    v.x = (g + 0 < n_valid) ? x[g + 0] : 0.f;
                    v.y = (g + 1 < n_valid) ? x[g + 1] : 0.f;
                    v.z = (g + 2 < n_valid) ? x[g + 2] : 0.f;
                    v.w = (g + 3 < n_valid) ? x[g + 3] : 0.f;
                }
                *reinterpret_cast<float4*>(&s.samples[i]) = v;
            }
        } else {
            for (int i = tid; i < kTileSamples; i += blockDim.x) {
                const long long g = g0 + i;
                s.samples[i] = (g < n_valid) ? x[g] : 0.f;
            }
        }
        __syncthreads();

        for (int fi = warp; fi < kTileFrames; fi += kWarps) {
            const int f = f0 + fi;
            if (f >= n_frames) break;                                   // warp-uniform
            const float* xs = s.samples + fi * kHop;

            // ---- pass 1: radix 8 over n1 (element 32*n1 + lane), twiddle by W256^(lane*k1)
            float2 v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = 32 * j + lane;
                if (j < 6 || (j == 6 && lane < 8)) {
                    const float2 sv = *reinterpret_cast<const float2*>(xs + 2 * c);
                    v[j] = make_float2(sv.x * wreg[j].x, sv.y * wreg[j].y);
                } else {
                    v[j] = make_float2(0.f, 0.f);                        // zero padding 400..511
                }
            }
            dft8(v);
#pragma unroll
            for (int k1 = 0; k1 < 8; ++k1) s1[k1 * 40 + lane] = (k1 == 0) ? v[0] : cmul(v[k1], tw1[k1]);
            __syncwarp();

            // ---- pass 2: radix 4 over m1 (n2 = 8*m1 + m2), twiddle by W32^(m2*q1); two items per lane
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int k1 = (lane >> 3) + 4 * half;
                float2 u0 = s1[k1 * 40 + 0 + m2], u1 = s1[k1 * 40 + 8 + m2];
                float2 u2 = s1[k1 * 40 + 16 + m2], u3 = s1[k1 * 40 + 24 + m2];
                dft4(u0, u1, u2, u3);
                s2[(k1 * 4 + 0) * 9 + m2] = u0;
                s2[(k1 * 4 + 1) * 9 + m2] = cmul(u1, tw2[1]);
                s2[(k1 * 4 + 2) * 9 + m2] = cmul(u2, tw2[2]);
                s2[(k1 * 4 + 3) * 9 + m2] = cmul(u3, tw2[3]);
            }
            __syncwarp();

            // ---- pass 3: radix 8 over m2; lane = 4*k1 + q1 produces Z[k1 + 8*q1 + 32*q2]
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = s2[lane * 9 + j];
            dft8(v);
            {
                const int b = (lane >> 2) + 8 * (lane & 3);
#pragma unroll
                for (int q2 = 0; q2 < 8; ++q2) s1[b + 32 * q2] = v[q2];
            }
            __syncwarp();

            // ---- split step: X[k] = Xe[k] + W512^k Xo[k]; magnitude
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int k = lane + 32 * j;
                const float2 zk = s1[k];
                const float2 zc = s1[(256 - k) & 255];
                const float2 xe = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y - zc.y));
                const float2 xo = make_float2(0.5f * (zk.y + zc.y), -0.5f * (zk.x - zc.x));
                const float2 w = twreg[j];
                const float re = xe.x + w.x * xo.x - w.y * xo.y;
                const float im = xe.y + w.x * xo.y + w.y * xo.x;
                mag[k] = sqrtf(re * re + im * im);
                if (k == 0) mag[256] = fabsf(zk.x - zk.y);
            }
            __syncwarp();

            // ---- sparse mel + log: lane handles bands (lane, 63 - lane) for balance
            float* out = logmel + static_cast<long long>(f) * kMel;
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int band = t == 0 ? lane : 63 - lane;
                const int st = s.mel_start[band], ln = s.mel_len[band], off = s.mel_off[band];
                float acc = 0.f;
                for (int j = 0; j < ln; ++j) acc = fmaf(mag[st + j], s.mel_w[off + j], acc);
                out[band] = logf(acc + 0.001f);
            }
            __syncwarp();
        }
    }
}

}  // namespace

size_t frontend_smem_bytes() { return sizeof(FeSmem); }

// Called once per engine on its own device (function attributes are per device).
cudaError_t frontend_init_device() {
    return cudaFuncSetAttribute(logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(sizeof(FeSmem)));
}

cudaError_t launch_logmel(const float* x, long long n_valid, long long frame_begin, int n_frames,
                          const FrontendTables* tab, float* logmel, int num_sms, cudaStream_t stream) {
    if (n_frames <= 0) return cudaSuccess;
    const int n_tiles = (n_frames + kTileFrames - 1) / kTileFrames;
    int grid = n_tiles < num_sms * 3 ? n_tiles : num_sms * 3;
    logmel_kernel<<<grid, kWarps * 32, sizeof(FeSmem), stream>>>(x, n_valid, frame_begin, n_frames, tab, logmel);
    return cudaGetLastError();
}

}  // namespace bd
