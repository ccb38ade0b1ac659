/* buzzdetect_b200 -- C ABI of the B200-native buzzdetect inference hot path (libbuzzdetect_b200.so).
 *
 * Plain pointers and sizes only; no torch / TensorFlow types.  Every entry point replaces a call the reference
 * makes into TensorFlow / librosa from its embedder + model plugins (paths relative to the reference checkout):
 *
 *   bd_frames_for        embedders/yamnet/features.py:82-108 (pad_waveform) + :66-76 (patch framing); host only
 *   bd_engine_create     models/model_general_v3/model.py:11-16 (ModelGeneralV3.initialize: load embedder + head)
 *                        embedders/yamnet_k2/embedder.py:14-24, embedders/yamnet/embedder.py:25-31
 *   bd_predict_host      models/model_general_v3/model.py:18-31 (predict = head(embedder.embed(samples)))
 *                        called from src/inference/worker.py:72
 *   bd_predict_device    same, for audio already resident in HBM (bench "value", chunk pipelines)
 *   bd_submit_host /     same as bd_predict_host but split so a caller can keep several chunks in flight
 *   bd_wait              (H2D of chunk i+1 overlaps compute of chunk i); replaces N analyzer threads
 *   bd_submit_pcm_host   the same for a chunk still in its decoded form (int16 / float32 PCM at the file's rate):
 *                        WorkerStreamer.queue_chunk's downmix + resample (src/stream/worker.py:110-129) run on the
 *                        device in front of predict
 *   bd_resample_*        librosa.resample call in src/stream/worker.py:128 (+ np.mean downmix :116-117)
 *   bd_profile_device    per-stage device times (the reference only has wall-clock `rate`, worker.py:54-65)
 *   bd_debug_*           test hooks: individual kernels against the oracle
 *
 * Error model: functions return 0 on success, non-zero on failure; the message is kept per
 * engine (bd_last_error) -- the Python wrapper raises RuntimeError.  Nothing here aborts the process.
 * Threading: every entry point takes the engine's lock, so one thread may submit chunks (the inferer) while another
 * waits for results (the writer); different engines are independent (one per inferer thread / GPU).
 *
 * Chunk coalescing: bd_submit_* queues the chunk's host->device copy at once but may defer its compute: when the GPU
 * is idle the chunk is launched immediately, otherwise it waits and is launched TOGETHER with the chunks submitted
 * behind it (one pass of the CNN over up to late_patches patches), because a 200 s chunk on its own cannot fill the
 * GPU.  bd_wait / bd_flush / bd_synchronize launch whatever is still pending.  Results are identical either way: every
 * chunk is framed and padded on its own (src/stream/worker.py:109-135 + features.py:82-108).
 */
#ifndef BUZZDETECT_B200_H
#define BUZZDETECT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BD_ABI_VERSION 1

#define BD_N_LAYERS 14
#define BD_EMBED_DIM 1024

/* precision of the pointwise (1x1) convolutions */
#define BD_PRECISION_FP32_SIMT 0   /* float32 FMA on CUDA cores (on-device reference mode)              */
#define BD_PRECISION_FP16X1 1      /* tcgen05, fp16 operands, fp32 accumulate: 1 MMA  (~5e-4 relative)   */
#define BD_FUSE_CONV1_DW2 (1 << 16)   /* fuse_mask bit: layer 1 + layer-2 depthwise in one kernel              */
#define BD_FUSE_L12 (1 << 17)         /* fuse_mask bit: layers 1 and 2 entirely in one kernel (tensor-core modes) */
#define BD_FUSE_V3 (1 << 18)          /* fuse_mask bit: fused layers use sep_fused3_kernel (TMA-staged stencil input) where it applies */
#define BD_FUSE_L12V2 (1 << 19)       /* fuse_mask bit: layers 1+2 in one warp-specialised kernel (l12_fused2_kernel) */
#define BD_FUSE_PAIR (1 << 20)        /* fuse_mask bit: sep_fused3 issues cta_group::2 MMAs from CTA pairs (N >= 256 layers) */
#define BD_FUSE_NO_TC_RESAMPLE (1 << 21) /* fuse_mask bit: resample tap by tap on CUDA cores instead of the tcgen05 GEMM */
#define BD_PRECISION_FP16X3 3      /* tcgen05, hi/lo fp16 split, 3 MMAs: float32-equivalent (default)    */
#define BD_PRECISION_FP16F8 2      /* tcgen05, A_hi W_hi in fp16 + both correction products as ONE e5m2 (FP8)
                                      contraction: 2 MMA-equivalents, ~5e-5 on the logits (layers with K % 64 == 0
                                      in the fused / GEMM kernels; the other layers keep the x3 split)          */

typedef struct bd_engine bd_engine;

typedef struct bd_layer_desc {
    int32_t kind;          /* 0 = 3x3 conv (layer 1), 1 = depthwise-separable block                         */
    int32_t stride, cin, cout, h_in, w_in;
    int64_t dw_w, dw_b;    /* float offsets into `folded`: depthwise [9,cin] and bias [cin]; -1 for kind 0  */
    int64_t w, b;          /* conv: [9,cout]; separable: pointwise [cout,cin] (K-major); bias [cout]        */
} bd_layer_desc;

typedef struct bd_weights {
    const float* folded;           /* BN-folded YAMNet parameters, one flat float32 array                   */
    int64_t folded_len;
    bd_layer_desc layers[BD_N_LAYERS];
    const float* mel;              /* [257*64] mel matrix as carried by the SavedModel graph (Const_1)      */
    const float* window;           /* [400] periodic Hann, float32                                          */
    const float* head_kernel;      /* [1024*n_classes] row-major (Dense kernel)                             */
    const float* head_bias;        /* [n_classes]                                                           */
    int32_t n_classes;
} bd_weights;

typedef struct bd_config {
    int32_t device;                /* CUDA device ordinal                                                   */
    int32_t precision;             /* BD_PRECISION_*                                                        */
    int32_t early_patches;         /* patches per sub-batch for frontend..layer 7 depthwise (0 = 4096)      */
    int32_t late_patches;          /* patches per sub-batch for layer 7 pointwise..head     (0 = 4096)      */
    int32_t use_graph;             /* 1 = capture and replay CUDA graphs per (n_samples, hop)               */
    int32_t n_slots;               /* in-flight host chunks for bd_submit_host (1..64, 0 = 2)               */
    int32_t fuse_mask;             /* bit (L-2): run separable layer L as ONE fused depthwise+pointwise kernel
                                      (ignored in BD_PRECISION_FP32_SIMT); BD_FUSE_CONV1_DW2: layer 1 + layer-2
                                      depthwise in one kernel; BD_FUSE_L12: layers 1+2 in one kernel.
                                      BD_FUSE_V3 / BD_FUSE_L12V2: use the TMA-staged / warp-specialised variants.
                                      -1 = default (BD_FUSE_L12V2 | BD_FUSE_V3 | BD_FUSE_CONV1_DW2 | layers 3..6, 8..12).       */
} bd_config;

int32_t bd_abi_version(void);

/* Host-only framing math.  hop_frames = 96 (framehop_prop 1) or 48 (0.5), any value in [1,96] is accepted.
 * Returns 0; n_patches may be 0.  The hop count uses the same float32 division + ceil as the reference graph. */
int32_t bd_frames_for(int64_t n_samples, int32_t hop_frames, int64_t* n_padded, int64_t* n_stft_frames,
                      int64_t* n_patches);

int32_t bd_engine_create(const bd_config* cfg, const bd_weights* w, bd_engine** out, char* err, size_t err_len);
void bd_engine_destroy(bd_engine* e);
const char* bd_last_error(const bd_engine* e);

/* samples: n float32 at 16 kHz mono.  act: [n_patches, n_classes]; emb: [n_patches, 1024] or NULL. */
int32_t bd_predict_host(bd_engine* e, const float* samples, int64_t n, int32_t hop_frames, float* act, float* emb,
                        int64_t* n_patches);
int32_t bd_predict_device(bd_engine* e, const float* d_samples, int64_t n, int32_t hop_frames, float* d_act,
                          float* d_emb, int64_t* n_patches);
int32_t bd_submit_host(bd_engine* e, int32_t slot, const float* samples, int64_t n, int32_t hop_frames, float* act,
                       float* emb, int64_t* n_patches);
/* Same as bd_submit_host, for a chunk still in its decoded form: interleaved PCM [n_frames, channels] at src_rate
 * (fmt 0 float32, 1 int16).  Downmix + resample run on the device, so only the raw PCM crosses PCIe
 * (src/stream/worker.py:110-129 + src/inference/worker.py:72 in one call). */
int32_t bd_submit_pcm_host(bd_engine* e, int32_t slot, const void* pcm, int32_t fmt, int32_t channels,
                           int64_t n_frames, int32_t src_rate, int32_t hop_frames, float* act, float* emb,
                           int64_t* n_patches);
int32_t bd_wait(bd_engine* e, int32_t slot);
int32_t bd_flush(bd_engine* e);                  /* launch every pending chunk now (does not wait)                      */
int32_t bd_synchronize(bd_engine* e);
int32_t bd_reserve_slots(bd_engine* e, int64_t n_samples, int64_t pcm_bytes, int32_t hop_frames);
                                                 /* pre-size every slot (device + pinned buffers) for chunks of this size */
int32_t bd_debug_stats(bd_engine* e, char* buf, size_t len);   /* dispatcher counters, human readable               */
int32_t bd_trace(bd_engine* e, int32_t on, char* buf, size_t len);   /* timeline of the slot API: 1 start, 0 stop + dump
                                                 ("kind a b host_ms device_ms" per line; kinds in csrc/engine.cu)      */
int32_t bd_set_auto_flush(bd_engine* e, int32_t on);   /* 0: submitted chunks wait for bd_wait / bd_flush (default 1) */
int32_t bd_slot_state(bd_engine* e, int32_t slot);   /* 0 free, 1 pending, 2 launched; -1 bad argument                 */
int32_t bd_batch_stats(bd_engine* e, int64_t* batches, int64_t* chunks);   /* CNN passes launched / chunks they carried */

/* Pinned (page-locked) host memory for the streamer's chunk ring and for result buffers: what the new
 * src/stream/worker.py:109-135 fills instead of fresh numpy arrays.  With pinned buffers bd_submit_* returns as soon
 * as the copies are queued; with pageable input the host->device copy is staged synchronously by the driver, and
 * pageable outputs are delivered through the slot's own pinned staging at bd_wait. */
int32_t bd_host_alloc(size_t bytes, int32_t write_combined, void** out);
void bd_host_free(void* p);

/* Per-kernel-class device time of one un-graphed pass over device-resident audio (CUDA events around every
 * launch).  ms[BD_PROFILE_SLOTS] / launches[BD_PROFILE_SLOTS]: 0 frontend, 1 conv1, 2+i depthwise of layer i+2,
 * 15+i pointwise of layer i+2 (i = 0..12), 28 pool+head. */
#define BD_PROFILE_SLOTS 29
int32_t bd_profile_device(bd_engine* e, const float* d_samples, int64_t n, int32_t hop_frames, float* ms,
                          int64_t* launches);
int64_t bd_launch_count(const bd_engine* e);     /* kernels launched (or replayed from graphs) so far */
/* `steps` back-to-back passes over the same device-resident chunk, timed with CUDA events recorded on the
 * engine's own compute stream (torch.cuda.Event would only see torch's current stream). */
int32_t bd_bench_device(bd_engine* e, const float* d_samples, int64_t n, int32_t hop_frames, float* d_act,
                        int32_t steps, float* ms_total);

/* Downmix + polyphase resample of one decoded chunk to 16 kHz (src/stream/worker.py:116-128).
 * in: interleaved [n_frames, channels], fmt 0 = float32, 1 = int16 (scaled by 1/32768 like soundfile).
 * out: ceil(n_frames * 16000 / src_rate) float32 samples. */
int64_t bd_resample_out_len(int64_t n_frames, int32_t src_rate);
int32_t bd_resample_host(bd_engine* e, const void* in, int32_t fmt, int32_t channels, int64_t n_frames,
                         int32_t src_rate, float* out, int64_t out_capacity, int64_t* n_out);
int32_t bd_resample_device(bd_engine* e, const void* d_in, int32_t fmt, int32_t channels, int64_t n_frames,
                           int32_t src_rate, float* d_out, int64_t out_capacity, int64_t* n_out);

/* ---- test hooks (host buffers) ---- */
int32_t bd_debug_logmel(bd_engine* e, const float* samples, int64_t n, int64_t n_frames, float* logmel);
int32_t bd_debug_pw_gemm(bd_engine* e, const float* A, const float* W, const float* bias, int32_t M, int32_t N,
                         int32_t K, int32_t precision, int32_t block_n, float* C);
/* Run the stack on the first min(P, early_patches) patches and return the activation after `stage`:
 * 0 log-mel [F,64]; 1 layer-1 output; 2(L-1) depthwise output of layer L; 2(L-1)+1 pointwise output of layer L
 * (L = 2..14); values NHWC float32.  n_out receives the element count written. */
int32_t bd_debug_stage(bd_engine* e, const float* samples, int64_t n, int32_t hop_frames, int32_t stage, float* out,
                       int64_t out_capacity, int64_t* n_out);

#ifdef __cplusplus
}
#endif
#endif /* BUZZDETECT_B200_H */
