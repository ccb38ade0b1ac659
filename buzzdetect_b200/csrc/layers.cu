// CUDA-core kernels of the MobileNet-v1 YAMNet stack (sm_100a): first conv, depthwise 3x3, global-average-pool +
// dense head, and a float32 SIMT pointwise GEMM used as the on-device reference precision mode.
//
// Reference semantics: embedders/yamnet/yamnet.py:36-74 (_conv / _separable_conv: conv -> BN(no scale, eps 1e-4)
// -> ReLU), layer table :77-93, GlobalAveragePooling2D :103; models/model_general_v3/model.py:29 (Dense 1024->13).
// TensorFlow SAME padding: stride 1 pads (1,1); stride 2 on even sizes pads (0 before, 1 after) -- SURVEY.md 2a.
// BatchNorm is folded into weights/bias on the host (buzzdetect_b200/weights.py:fold_yamnet).
// Layout: NHWC, H = time, W = mel; activations are [P*H*W, C] row-major so a pointwise conv is a plain GEMM.
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
    const long long total = static_cast<long long>(P) * 48 * 32 * 8;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int cg = static_cast<int>(idx & 7);
        const long long pix = idx >> 3;
        const int ow = static_cast<int>(pix & 31);
        const int oh = static_cast<int>((pix >> 5) % 48);
        const long long p = pix / (48 * 32);
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

// ------------------------------------------------------------------------------------------ depthwise
// thread = (strip of R consecutive output pixels along W, 4 channels).  The 3 x ((R-1)*STRIDE+3) input window and
// the 9 weight vectors live in registers, so each input float4 is loaded once per strip instead of once per tap
// (the per-tap version was L1-bandwidth bound: 18 LDG.128 per output vector; this one issues 6.75 for R = 4).
// Channel-contiguous float4 accesses: consecutive threads take consecutive channel quads of the same strip, so
// every warp-level load/store touches whole 128-byte lines.
template <int STRIDE, int R, int OUT_MODE>
__global__ void __launch_bounds__(256) depthwise_kernel(const float* __restrict__ in, int P, int H, int W, int C,
                                                        const float* __restrict__ w, const float* __restrict__ b,
                                                        float* __restrict__ out_f32, __half* __restrict__ out_hi,
                                                        __half* __restrict__ out_lo) {
    const int Ho = H / STRIDE, Wo = W / STRIDE, C4 = C >> 2, WS = Wo / R;
    constexpr int PB = STRIDE == 1 ? 1 : 0;
    constexpr int NC = (R - 1) * STRIDE + 3;
    const long long total = static_cast<long long>(P) * Ho * WS * C4;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c4 = static_cast<int>(idx % C4);
        long long t = idx / C4;
        const int ws = static_cast<int>(t % WS);
        t /= WS;
        const int oh = static_cast<int>(t % Ho);
        const long long p = t / Ho;
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
                const __half h0 = __float2half_rn(a.x), h1 = __float2half_rn(a.y);
                const __half h2 = __float2half_rn(a.z), h3 = __float2half_rn(a.w);
                __half2 hp[2] = {__halves2half2(h0, h1), __halves2half2(h2, h3)};
                *reinterpret_cast<uint2*>(out_hi + o) = *reinterpret_cast<uint2*>(hp);
                if (OUT_MODE == 2) {
                    __half2 lp[2] = {
                        __halves2half2(__float2half_rn(a.x - __half2float(h0)), __float2half_rn(a.y - __half2float(h1))),
                        __halves2half2(__float2half_rn(a.z - __half2float(h2)), __float2half_rn(a.w - __half2float(h3)))};
                    *reinterpret_cast<uint2*>(out_lo + o) = *reinterpret_cast<uint2*>(lp);
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
// one CTA per patch, 256 threads x 4 channels = 1024 embedding dims.
__global__ void __launch_bounds__(256) pool_head_kernel(const float* __restrict__ y, int rows, const float* __restrict__ Wh,
                                                        const float* __restrict__ bh, int n_classes,
                                                        float* __restrict__ emb, float* __restrict__ act) {
    __shared__ float red[8][kMaxClasses];
    const long long p = blockIdx.x;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const float* yp = y + p * rows * kEmb + t * 4;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < rows; ++r) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(yp + static_cast<long long>(r) * kEmb));
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    const float d = static_cast<float>(rows);
    s.x /= d; s.y /= d; s.z /= d; s.w /= d;
    if (emb != nullptr) *reinterpret_cast<float4*>(emb + p * kEmb + t * 4) = s;
    for (int j = 0; j < n_classes; ++j) {
        const float* wj = Wh + static_cast<long long>(t) * 4 * n_classes + j;
        float a = s.x * __ldg(wj);
        a = fmaf(s.y, __ldg(wj + n_classes), a);
        a = fmaf(s.z, __ldg(wj + 2 * n_classes), a);
        a = fmaf(s.w, __ldg(wj + 3 * n_classes), a);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) red[warp][j] = a;
    }
    __syncthreads();
    if (t < n_classes) {
        float a = 0.f;
#pragma unroll
        for (int wi = 0; wi < 8; ++wi) a += red[wi][t];
        act[p * n_classes + t] = a + bh[t];
    }
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
    const long long total = static_cast<long long>(P) * 48 * 32 * 8;
    conv1_kernel<<<grid_for(total, 256, 148 * 64), 256, 0, stream>>>(logmel, hop_frames, P, w, b, out);
    return cudaGetLastError();
}

cudaError_t launch_depthwise(const float* in, int P, int H, int W, int C, int stride, const float* w, const float* b,
                             int out_mode, float* out_f32, __half* out_hi, __half* out_lo, cudaStream_t stream) {
    if (P <= 0) return cudaSuccess;
    if ((C & 3) || (stride != 1 && stride != 2) || (stride == 2 && ((H | W) & 1))) return cudaErrorInvalidValue;
    if (out_mode < 0 || out_mode > 2) return cudaErrorInvalidValue;
    const int Wo = W / stride;
    const int R = (Wo % 4 == 0) ? 4 : ((Wo % 2 == 0) ? 2 : 1);
    const long long total = static_cast<long long>(P) * (H / stride) * (Wo / R) * (C / 4);
    const int grid = grid_for(total, 256, 148 * 64);
#define BD_DW(S, RR, MODE) \
    depthwise_kernel<S, RR, MODE><<<grid, 256, 0, stream>>>(in, P, H, W, C, w, b, out_f32, out_hi, out_lo)
#define BD_DW_MODE(S, RR)                          \
    do {                                           \
        if (out_mode == 0) BD_DW(S, RR, 0);        \
        else if (out_mode == 1) BD_DW(S, RR, 1);   \
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
    pool_head_kernel<<<P, 256, 0, stream>>>(y, rows_per_patch, Wh, bh, n_classes, emb, act);
    return cudaGetLastError();
}

}  // namespace bd
