// Launch-function declarations shared by the translation units of libbuzzdetect_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <cstddef>

namespace bd {

constexpr int kWin = 400;          // 25 ms @ 16 kHz         (embedders/yamnet/params.py:26)
constexpr int kHop = 160;          // 10 ms                  (params.py:27)
constexpr int kFft = 512;
constexpr int kBins = 257;
constexpr int kMel = 64;           // params.py:28
constexpr int kPatchFrames = 96;   // 0.96 s                 (params.py:32)
constexpr int kMinSamples = 15600; // Const_5 of the SavedModel graphs
constexpr int kMelNnzMax = 512;    // 461 non-zeros in the shipped mel matrix
constexpr int kEmb = 1024;
constexpr int kMaxClasses = 32;

struct FrontendTables {            // device-resident, built once per engine
    float window[kWin];            // periodic Hann exactly as the graph computes it in float32
    int mel_start[kMel];           // first non-zero spectrogram bin of each mel band
    int mel_len[kMel];             // number of consecutive bins
    int mel_off[kMel];             // offset of the band's weights in mel_w
    float mel_w[kMelNnzMax];
};

// ---- frontend.cu
size_t frontend_smem_bytes();
cudaError_t frontend_init_device();
cudaError_t launch_logmel(const float* x, long long n_valid, long long frame_begin, int n_frames,
                          const FrontendTables* tab, float* logmel, int num_sms, cudaStream_t stream);

// ---- frontend2.cu  (version 2: 16 x 16 FFT across half-warps, lane = frame mel, TMA in/out, segment table)
constexpr int kMaxLogmelSegs = 64;
struct LogmelSeg {                 // one independent chunk of 16 kHz audio inside a batched launch
    const float* x;                // device samples (fmt 0: float32; fmt 1: int16 PCM, value = s / 32768 as soundfile reads it)
    long long n_valid;             // samples that exist (reads beyond are the virtual zero padding of pad_waveform)
    long long frame_begin;         // first STFT frame to compute (frame f covers samples 160 f .. 160 f + 399)
    int row_begin;                 // first row of the shared log-mel buffer this segment writes
    int n_rows;                    // number of frames
    int fmt = 0;                   // 0 float32, 1 int16 (16 kHz mono PCM straight from the decoder: no conversion pass)
};
struct alignas(16) FrontendMelParam { unsigned char raw[4912]; };   // opaque image of the kernel's mel parameter block
cudaError_t frontend2_init_device();
bool frontend2_build_mel(const FrontendTables& tab, FrontendMelParam* out);   // false: mel matrix too dense
cudaError_t launch_logmel_segs(const LogmelSeg* segs, int n_segs, const FrontendMelParam& mel, const float* window,
                               float* logmel, long long logmel_rows, int num_sms, cudaStream_t stream);

// ---- layers.cu
// conv 3x3 stride 2 SAME (pad 0 before / 1 after), 1 -> 32 channels, folded BN + ReLU.  in: log-mel rows, patch p
// starts at row p*hop_frames of `logmel` (patches are views, never copied).  out: [P,48,32,32] float32 NHWC.
cudaError_t launch_conv1(const float* logmel, int hop_frames, int P, const float* w9x32, const float* b32,
                         float* out, cudaStream_t stream);
// layer 1 + the depthwise half of layer 2 in one kernel (the [48,32,32] layer-1 activation stays in shared memory).
// Output = layer-2 depthwise activation [P*48*32, 32] in the same formats as launch_depthwise.
cudaError_t layers_init_device();
cudaError_t launch_conv1_dw2(const float* logmel, int hop_frames, int P, const float* w1, const float* b1,
                             const float* dw_w, const float* dw_b, int out_mode, float* out_f32, __half* out_hi,
                             __half* out_lo, cudaStream_t stream);
// depthwise 3x3 (stride 1: pad 1/1, stride 2: pad 0/1), folded BN + ReLU.  in [P,H,W,C] float32 NHWC.
// out_mode 0: float32 plane `out_f32` [P*Ho*Wo, C]
// out_mode 1: fp16 hi plane only            (single-pass tensor-core GEMM operand)
// out_mode 2: fp16 hi + lo planes, x ~= hi + lo  (3-product split GEMM operand)
// out_mode 3: fp16 hi plane + e5m2 correction plane at out_lo (unsigned char [M, 2C]; fp16 + fp8 plan, C % 64 == 0)
cudaError_t launch_depthwise(const float* in, int P, int H, int W, int C, int stride, const float* w9xC,
                             const float* bC, int out_mode, float* out_f32, __half* out_hi, __half* out_lo,
                             cudaStream_t stream);
// reference-precision pointwise conv on CUDA cores: C[M,N] = relu(A[M,K] * Bt[N,K]^T + bias)
cudaError_t launch_pw_simt(const float* A, const float* Bt, const float* bias, float* C, int M, int N, int K,
                           cudaStream_t stream);
// global average pool over `rows_per_patch` rows + dense head: y [P*rows, 1024] -> emb [P,1024] (optional),
// act [P, n_classes] = emb @ Wh[1024, n_classes] + bh
cudaError_t launch_pool_head(const float* y, int P, int rows_per_patch, const float* Wh, const float* bh,
                             int n_classes, float* emb, float* act, cudaStream_t stream);

// ---- pw_gemm_sm100.cu  (tcgen05 / TMEM / TMA)
struct PwGemmPlan {
    CUtensorMap a_hi, a_lo, b_hi, b_lo;
    CUtensorMap b64_hi, b64_lo;   // the same weight planes in 64-row boxes (sep_fused3_kernel's 16 KB ring slots)
    CUtensorMap a_c8, b_c8;       // nsplit == 2 ("fp16f8"): e5m2 correction planes [rows, 2K] bytes, per 64-channel k-block
                                  // A: [a_lo 2^11 (64) | a_hi (64)], W: [w_hi 2^-11 (64) | w_lo (64)]
    int M_max, N, K, block_n, nsplit;   // nsplit: 1 = fp16, 3 = fp16 hi/lo x3, 2 = fp16 + fp8 corrections
    float out_scale;   // accumulators are multiplied by this before the bias (weights are stored pre-scaled by 1/out_scale)
};
cudaError_t pw_gemm_init_device();
// Build TMA descriptors for A planes [M_max, K] (row stride lda halfs) and weight planes [N, K].
// block_n: 64/128/256, or 0 to choose automatically.
cudaError_t pw_gemm_make_plan(PwGemmPlan* plan, const __half* a_hi, const __half* a_lo, int M_max, int K,
                              const __half* b_hi, const __half* b_lo, int N, int nsplit, int block_n,
                              float out_scale, const char** err);
// nsplit == 2: a_lo / b_lo point at the e5m2 planes (unsigned char [rows, 2K]); K must be a multiple of 64.
// The weight side of the fp16 + fp8 plan: hi = fp16(w * scale), c8 row = per k-block [e5m2(hi 2^-11) | e5m2(w*scale - hi)].
float split_weights_f16f8(const float* w, size_t rows, size_t K, __half* hi, unsigned char* c8);
bool encode_kmajor_u8_map(CUtensorMap* map, const void* ptr, int rows, int row_bytes, int box_rows);
// Split float32 weights into fp16 hi/lo planes after scaling by a power of two chosen so that max|w| lands in
// [512,1024): the lo plane then stays in fp16's NORMAL range (unscaled 1x1 weights are ~0.05, their lo parts would be
// subnormal and carry only ~1e-6 relative precision).  Returns the inverse scale for PwGemmPlan::out_scale.
float split_weights_f16(const float* w, size_t n, __half* hi, __half* lo);
cudaError_t launch_pw_gemm(const PwGemmPlan& plan, const float* bias, float* C, int M, int num_sms,
                           cudaStream_t stream);
// Layers 1 + 2 (conv1 -> depthwise -> pointwise 32->64) in one kernel; only the plan's weight maps are used.
// C: [P*48*32, 64] float32.
cudaError_t launch_l12_fused(const PwGemmPlan& plan, const float* logmel, int hop_frames, int P, const float* w1,
                             const float* b1, const float* dw_w, const float* dw_b, const float* bias, float* C,
                             int num_sms, cudaStream_t stream);
// Fused separable block: depthwise 3x3 (+bias, ReLU) computed by producer warps straight into the GEMM's A tile.
// X: [P,H,W,K] float32 NHWC; C: [P*(H/stride)*(W/stride), N].  Only the plan's weight maps (b_hi/b_lo) are used.
// Requires (W/stride) % 4 == 0.
cudaError_t launch_sep_fused(const PwGemmPlan& plan, const float* X, const float* dw_w, const float* dw_b,
                             const float* bias, float* C, int P, int H, int W, int stride, int num_sms,
                             cudaStream_t stream);

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TensorMapEncodeFn tensor_map_encode_fn();
// Store map for a row-major float32 [rows, cols] output written in [box_rows x 32] SWIZZLE_128B blocks (epi_swz_addr);
// box_rows = 32, or 8 for the partial lane quad of a tile whose row count is not a multiple of 32.
bool encode_store_map_f32(CUtensorMap* map, float* ptr, long long rows, int cols, int box_rows);

// ---- sep_fused_sm100.cu  (fused separable block v3: TMA-staged depthwise input, see the file header)
cudaError_t sep_fused3_init_device();
bool sep_fused3_supported(int K, int N, int H, int W, int stride);
// X: [P,H,W,K] float32 NHWC (16-byte aligned); C: [P*(H/stride)*(W/stride), N].  Uses the plan's weight maps.
// bias_host: the N pointwise biases in HOST memory (kernel parameter = constant bank).
cudaError_t launch_sep_fused3(const PwGemmPlan& plan, const float* X, const float* dw_w, const float* dw_b,
                              const float* bias_host, float* C, int P, int H, int W, int stride, int num_sms,
                              cudaStream_t stream, bool cta_pairs = false);

// ---- l12_fused_sm100.cu  (layers 1 + 2, warp-specialised: conv1 warps -> stencil warps -> tcgen05 -> epilogue)
cudaError_t l12_fused2_init_device();
// bias_host: the 64 pointwise biases in HOST memory (they travel as a kernel parameter = constant bank).
cudaError_t launch_l12_fused2(const PwGemmPlan& plan, const float* logmel, int hop_frames, int P, const float* w1,
                              const float* b1, const float* dw_w, const float* dw_b, const float* bias_host, float* C,
                              int num_sms, cudaStream_t stream);

// K-major fp16 operand map (SWIZZLE_128B, [box_rows x 64] boxes) over a row-major [rows, cols] plane
bool encode_kmajor_f16_map(CUtensorMap* map, const void* ptr, int rows, int cols, int box_rows);

// ---- resample.cu  (tap-by-tap CUDA-core evaluation: equal-rate downmix, chunk tails, fallback)
// Computes out[m] for m in [m_begin, n_out).
cudaError_t launch_resample(const void* in, int in_fmt /*0 f32, 1 s16*/, int channels, long long n_in_frames,
                            int up, int down, const float* taps, int taps_per_phase, float* out, long long n_out,
                            cudaStream_t stream, long long m_begin = 0);

// ---- resample_tc_sm100.cu  (the same filter as a tcgen05 GEMM over blocks of NB outputs; see the file header)
struct ResampleTcPlan {
    int up, down, T;          // rational ratio, taps per phase
    int NB;                   // outputs per block (a multiple of up and of 32)
    int ntile, n_tiles;       // columns per pass (<= 160), passes per row tile
    int S;                    // input samples between blocks
    int K;                    // window length padded to a multiple of 64
    float out_scale;          // inverse of the power-of-two scale applied to H before the fp16 split
    CUtensorMap b_hi, b_lo;   // H planes [NB, K]
};
cudaError_t resample_tc_init_device();
bool resample_tc_geometry(int up, int down, int taps_per_phase, ResampleTcPlan* plan);
void resample_tc_build_matrix(const ResampleTcPlan& plan, const float* taps /*[T][up]*/, float* H /*[NB][K]*/);
// Writes out[0 .. n_done) with n_done = (n_out / NB) * NB; the caller finishes [n_done, n_out) with launch_resample.
cudaError_t launch_resample_tc(const ResampleTcPlan& plan, const void* in, int in_fmt, int channels,
                               long long n_in_frames, float* out, long long n_out, int num_sms, cudaStream_t stream,
                               long long* n_done);

}  // namespace bd
