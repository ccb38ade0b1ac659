// bd_engine: device buffers, weight upload, the chunk pipeline and the C ABI (include/buzzdetect_b200.h).
//
// One engine = one inferer (reference: one model instance per WorkerInferer thread, src/inference/worker.py:21).
// A chunk of 16 kHz audio goes through
//     frontend -> layers 1+2 (one kernel) -> fused depthwise+pointwise kernels (layers 3-6, 8-12) and
//     depthwise / GEMM pairs (layers 7, 13, 14) -> mean(H,W) -> dense head
// in sub-batches that bound the working set independently of the chunk length:
//   * "early" phase (frontend .. layer-7 depthwise, up to 384 KB of activations per patch): early_patches at a time
//   * "late"  phase (layer-7 pointwise .. head, <= 48 KB per patch): late_patches at a time
// Both default to 4096 patches (1.09 audio-hours): measured, large sub-batches beat L2-resident small ones because
// kernel tails and launches cost more than the HBM round trips they would save (profiles/r1_summary.md).
// The whole chunk is captured into a CUDA graph per (buffer, length, hop) and replayed.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <map>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <string>
#include <tuple>
#include <vector>

#include <cuda_fp8.h>

#include "../../include/buzzdetect_b200.h"
#include "bd_kernels.cuh"

using namespace bd;

namespace {

struct LayerDev {
    bd_layer_desc d;
    int h_out, w_out;
    const float* dw_w = nullptr;    // device pointers into folded blob
    const float* dw_b = nullptr;
    const float* w = nullptr;       // conv: [9,32]; sep: [cout,cin] float32
    const float* b = nullptr;
    __half* w_hi = nullptr;         // [cout,cin] fp16 planes (tensor-core modes)
    __half* w_lo = nullptr;
    PwGemmPlan plan;
    int nsplit = 3;                 // operand plan of this layer's pointwise product: 1 fp16, 3 fp16 hi/lo x3, 2 fp16 + fp8
    bool late = false;              // pointwise runs in the late phase
    bool fused = false;             // depthwise computed inside the pointwise GEMM (sep_fused_kernel)
    bool fused_v3 = false;          // ... by sep_fused3_kernel (TMA-staged depthwise input)
};

struct Slot {
    float* d_in = nullptr;
    int64_t in_cap = 0;
    unsigned char* d_pcm = nullptr;  // raw decoded chunk (bd_submit_pcm_host)
    int64_t pcm_cap = 0;
    float* d_act = nullptr;
    float* d_emb = nullptr;
    int64_t out_cap = 0;            // patches
    // results leave the device through pinned staging owned by the slot (a device->host copy into pageable memory
    // would block the submitting thread until the chunk's compute has finished); bd_wait copies them to the caller
    float* h_act = nullptr;
    float* h_emb = nullptr;
    int64_t h_act_cap = 0, h_emb_cap = 0;   // floats
    float* user_act = nullptr;      // where bd_wait delivers (nullptr: the copy went straight to the caller's pinned buffer)
    float* user_emb = nullptr;
    float* dst_act = nullptr;       // destination of the device->host copies (caller's pinned buffer or the staging)
    float* dst_emb = nullptr;
    int64_t n = 0, P = 0;
    int hop = 0;
    bool want_emb = false;
    // decoded-PCM chunks: downmix + resample run on the compute stream at the head of the pass that takes the chunk
    bool has_pcm = false;
    int pcm_fmt = 0, pcm_channels = 1, pcm_rate = 16000;
    int64_t pcm_frames = 0;
    cudaEvent_t ev_in = nullptr, ev_comp = nullptr, ev_out = nullptr;
    int state = 0;                  // 0 free, 1 pending (input on its way, compute not enqueued), 2 launched
    bool arrived = false;           // dispatcher: the input copy has completed
};

struct GraphKey {
    const float* x; int64_t n; int hop; float* act; float* emb;
    bool operator<(const GraphKey& o) const {
        return std::tie(x, n, hop, act, emb) < std::tie(o.x, o.n, o.hop, o.act, o.emb);
    }
};

}  // namespace

struct bd_engine {
    int device = 0, num_sms = 148, precision = BD_PRECISION_FP16X3;
    int S1 = 4096, S2 = 4096, n_classes = 13;
    bool use_graph = true;
    cudaStream_t s_compute = nullptr, s_in = nullptr, s_out = nullptr;
    float* d_folded = nullptr;
    std::vector<float> h_folded;          // host copy: biases travel to the fused kernels as kernel parameters
    FrontendTables* d_tab = nullptr;
    float* d_headW = nullptr;
    float* d_headB = nullptr;
    std::vector<LayerDev> layers;
    // activations
    float* d_logmel = nullptr;
    float* d_F_early = nullptr;
    unsigned char* d_H_early = nullptr;
    size_t H_early_plane_bytes = 0;      // offset of the lo plane
    float* d_F_late = nullptr;
    unsigned char* d_H_late = nullptr;
    size_t H_late_plane_bytes = 0;
    float* d_F_early2 = nullptr;          // ping-pong partners for fused separable blocks (cannot run in place)
    float* d_F_late2 = nullptr;
    int first_late = 6;                   // index of the layer whose pointwise output starts the late phase (layer 7)
    bool fuse_conv1 = true;               // layer 1 + layer-2 depthwise in one kernel (conv1_dw2_kernel)
    bool fuse_l12 = true;                 // layers 1 + 2 entirely in one kernel (l12_fused_kernel, tensor-core modes)
    bool l12_v2 = false;                  // ... using the warp-specialised l12_fused2_kernel
    bool cta_pairs = false;               // sep_fused3 with cta_group::2 MMAs where it applies (BD_FUSE_PAIR)
    const void* dbg_ptr = nullptr;        // bd_debug_stage: where the requested stage's output lives
    int dbg_planes = 0;                   // 0: float32; else the producing layer's operand plan (1, 2, 3)
    size_t dbg_plane_off = 0;
    std::vector<Slot> slots;
    std::vector<int> pending;             // slots submitted but not yet launched, in submission order
    std::recursive_mutex mu;                       // the slot API may be driven by two threads (inferer submits, writer waits)
    // coalesced batches (several small chunks in ONE pass of the CNN): outputs of the whole batch, double-buffered
    float* d_act_batch[2] = {nullptr, nullptr};
    float* d_emb_batch[2] = {nullptr, nullptr};
    cudaEvent_t ev_batch_out[2] = {nullptr, nullptr};
    int batch_set = 0;
    // dispatcher: one thread per engine launches what is pending as soon as the GPU has room (at most two passes in the
    // stream) and the first pending chunk has arrived -- work-conserving dynamic batching, independent of how the
    // caller's threads are scheduled (a Python writer thread may sit behind the GIL for milliseconds)
    std::thread dispatcher;
    std::condition_variable_any cv_work, cv_launched;
    bool stop = false, flush_req = false;
    cudaEvent_t ev_pass[2] = {nullptr, nullptr};   // end of compute of pass i (i & 1)
    int64_t pass_seq = 0;
    // diagnostics (bd_debug_stats)
    int64_t st_sleep_full = 0, st_sleep_small = 0, st_sleep_noarr = 0, st_allocs = 0;
    double st_launch_s = 0.0, st_submit_s = 0.0, st_alloc_s = 0.0;
    int64_t batches = 0, batched_chunks = 0;
    bool auto_flush = true;               // bd_set_auto_flush(0): chunks wait until bd_wait / bd_flush (tests, batch drivers)
    int64_t coalesce_target = 3072;       // pending patches that trigger a launch even while the GPU is busy
    FrontendMelParam mel_param;
    bool frontend_v1 = false;             // BD_FRONTEND_V1=1: the round-1 kernel (A/B measurements)
    std::map<GraphKey, cudaGraphExec_t> graphs;
    std::map<GraphKey, int64_t> graph_launches;
    int64_t launch_count = 0;
    // timeline trace (bd_trace): host clock + timing events at every hand-over of the slot API
    struct TraceRec { int kind, a, b; double host_ms; cudaEvent_t ev; };
    bool tracing = false;
    std::vector<TraceRec> trace;
    cudaEvent_t trace_base = nullptr;
    std::chrono::steady_clock::time_point trace_t0;
    // profiling hooks (bd_profile_device)
    bool profiling = false;
    std::vector<cudaEvent_t> prof_events;
    std::vector<int> prof_cat;
    // resampler filter cache: (src_rate) -> device taps
    struct Resampler {
        int up, down, taps_per_phase;
        float* d_taps;
        bool tc_ok = false;               // tensor-core formulation available (resample_tc_sm100.cu)
        ResampleTcPlan tc;
        __half* d_h_hi = nullptr;
        __half* d_h_lo = nullptr;
    };
    bool tc_resample = true;              // BD_FUSE_NO_TC_RESAMPLE clears it
    std::map<int, Resampler> resamplers;
    std::string last_error;
};

extern "C" {
static int get_resampler(bd_engine* e, int src_rate, bd_engine::Resampler** out);
static int run_resample(bd_engine* e, const bd_engine::Resampler* r, const void* d_in, int fmt, int channels,
                        long long n_frames, float* d_out, long long no, cudaStream_t st);
}

namespace {

thread_local std::string g_create_error;
constexpr float kActScale = 0.0625f;      // power of two applied to the depthwise outputs in the tensor-core modes

#define BD_CHECK(e, expr)                                                                       \
    do {                                                                                        \
        cudaError_t _err = (expr);                                                              \
        if (_err != cudaSuccess) {                                                              \
            (e)->last_error = std::string(#expr) + ": " + cudaGetErrorString(_err);             \
            return 1;                                                                           \
        }                                                                                       \
    } while (0)

int fail(bd_engine* e, const std::string& msg) {
    e->last_error = msg;
    return 1;
}

// ------------------------------------------------------------------------------------------ framing math
// embedders/yamnet/features.py:82-108 -- the hop count is ceil(float32(after) / float32(hop)).
void frames_for(int64_t n, int hop_frames, int64_t* n_padded, int64_t* n_frames, int64_t* n_patches) {
    const int64_t hop_samples = static_cast<int64_t>(hop_frames) * kHop;
    int64_t pad = std::max<int64_t>(0, kMinSamples - n);
    const int64_t n2 = std::max<int64_t>(n, kMinSamples);
    const int64_t after = n2 - kMinSamples;
    const float q = static_cast<float>(after) / static_cast<float>(hop_samples);
    const int64_t hops = static_cast<int64_t>(std::ceil(q));
    pad += hop_samples * hops - after;
    const int64_t np = n + pad;
    const int64_t f = np >= kWin ? 1 + (np - kWin) / kHop : 0;
    const int64_t p = f >= kPatchFrames ? 1 + (f - kPatchFrames) / hop_frames : 0;
    if (n_padded) *n_padded = np;
    if (n_frames) *n_frames = f;
    if (n_patches) *n_patches = p;
}

// ------------------------------------------------------------------------------------------ profiling hooks
// profiling categories: 0 frontend, 1 conv1, 2+i depthwise of layer i+2, 15+i pointwise of layer i+2, 28 pool+head
enum Cat { CAT_FRONTEND = 0, CAT_CONV1 = 1, CAT_DW = 2, CAT_PW = 15, CAT_POOL = 28, CAT_COUNT = 29 };

void mark(bd_engine* e, int cat, cudaStream_t st) {
    e->launch_count++;
    if (!e->profiling) return;
    cudaEvent_t ev;
    cudaEventCreate(&ev);
    cudaEventRecord(ev, st);
    e->prof_events.push_back(ev);
    e->prof_cat.push_back(cat);
}

// Timeline trace: kinds 0 submit call (a = slot), 1 input copy done (a = slot), 2 pass begins on the device
// (a = pass, b = chunks), 3 pass ends, 4 result copy done (a = slot), 5 bd_wait returns (a = slot), 6 host side of the
// pass launch finished (a = pass).  st == nullptr: host time only.
void trace_point(bd_engine* e, int kind, int a, int b, cudaStream_t st) {
    if (!e->tracing) return;
    bd_engine::TraceRec r{kind, a, b, 0.0, nullptr};
    r.host_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - e->trace_t0).count();
    if (st) {
        cudaEventCreate(&r.ev);
        cudaEventRecord(r.ev, st);
    }
    e->trace.push_back(r);
}

// ------------------------------------------------------------------------------------------ the pipeline
// Enqueue the whole chunk on `st`.  x: device audio, n samples.  Outputs are device pointers.
// stop_stage >= 0 (debug): stop after that stage of the FIRST early sub-batch (see bd_debug_stage); the location of
// the stage's output is left in e->dbg_*.
// What the frontend reads in one pass: ONE chunk (n_segs == 0: x / n; the chunk may span several late batches), or a
// coalesced batch of whole chunks (n_segs > 0; P = patch slots of the whole batch <= late_patches).  In a batch, chunk c
// owns the patch slots [g_c, g_c + P_c) and the log-mel rows from g_c * hop on; with overlapping patches (hop < 96) the
// slot after a chunk's last patch straddles two chunks and its output is never read.
struct FrontJob {
    const float* x = nullptr;
    int64_t n = 0;
    int n_segs = 0;
    LogmelSeg segs[kMaxLogmelSegs];
};

int run_frontend(bd_engine* e, const FrontJob& job, int64_t big, int nb, int hop_frames, cudaStream_t st) {
    const int n_fr = (nb - 1) * hop_frames + kPatchFrames;
    const long long cap_rows = static_cast<long long>(e->S2) * kPatchFrames;
    if (job.n_segs == 0) {
        if (e->frontend_v1) {
            BD_CHECK(e, launch_logmel(job.x, job.n, big * hop_frames, n_fr, e->d_tab, e->d_logmel, e->num_sms, st));
        } else {
            LogmelSeg sg{job.x, job.n, big * hop_frames, 0, n_fr};
            BD_CHECK(e, launch_logmel_segs(&sg, 1, e->mel_param, e->d_tab->window, e->d_logmel, cap_rows, e->num_sms, st));
        }
        return 0;
    }
    if (big != 0) return fail(e, "internal: a coalesced batch must fit one late batch");
    if (e->frontend_v1) {
        for (int i = 0; i < job.n_segs; ++i) {
            const LogmelSeg& sg = job.segs[i];
            BD_CHECK(e, launch_logmel(sg.x, sg.n_valid, sg.frame_begin, sg.n_rows, e->d_tab,
                                      e->d_logmel + static_cast<size_t>(sg.row_begin) * kMel, e->num_sms, st));
        }
    } else {
        BD_CHECK(e, launch_logmel_segs(job.segs, job.n_segs, e->mel_param, e->d_tab->window, e->d_logmel, cap_rows,
                                       e->num_sms, st));
    }
    return 0;
}

int enqueue_chunk(bd_engine* e, const FrontJob& job, int hop_frames, float* d_act, float* d_emb,
                  int64_t P, cudaStream_t st, int stop_stage = -1) {
    const int prec = e->precision;
    const int dw_mode = prec == BD_PRECISION_FP32_SIMT ? 0 : (prec == BD_PRECISION_FP16X1 ? 1 : 2);
    const int first_late = e->first_late;
    auto dw_mode_of = [&](int L) {                      // output format of the depthwise kernel feeding layer L's GEMM
        if (prec == BD_PRECISION_FP32_SIMT) return 0;
        const int ns = e->layers[L].nsplit;
        return ns == 1 ? 1 : (ns == 2 ? 3 : 2);
    };
    auto stop_at = [&](int stage, const void* ptr, int planes, size_t plane_off) {
        if (stop_stage != stage) return false;
        e->dbg_ptr = ptr; e->dbg_planes = planes; e->dbg_plane_off = plane_off;
        return true;
    };
    // separable block L (index, >= 1) without fusion: depthwise in -> H planes, pointwise H -> out
    auto unfused = [&](int L, const float* in, int np, unsigned char* H, size_t plane, int64_t row_off, float* out,
                       bool do_dw, bool do_pw) -> int {
        const LayerDev& l = e->layers[L];
        float* o32 = reinterpret_cast<float*>(H) + row_off * l.d.cin;
        __half* ohi = reinterpret_cast<__half*>(H) + row_off * l.d.cin;
        __half* olo = reinterpret_cast<__half*>(H + plane) + row_off * l.d.cin;
        if (do_dw) {
            const int dwm = dw_mode_of(L);
            BD_CHECK(e, launch_depthwise(in, np, l.d.h_in, l.d.w_in, l.d.cin, l.d.stride, l.dw_w, l.dw_b, dwm, o32,
                                         ohi, olo, st));
            mark(e, CAT_DW + L - 1, st);
            if (stop_at(2 * L, H + (dwm == 0 ? 4 : 2) * row_off * l.d.cin, dwm == 0 ? 0 : l.nsplit, plane)) return 2;
        }
        if (do_pw) {
            const int M = np * l.h_out * l.w_out;
            if (prec == BD_PRECISION_FP32_SIMT) {
                BD_CHECK(e, launch_pw_simt(reinterpret_cast<const float*>(H), l.w, l.b, out, M, l.d.cout, l.d.cin, st));
            } else {
                BD_CHECK(e, launch_pw_gemm(l.plan, l.b, out, M, e->num_sms, st));
            }
            mark(e, CAT_PW + L - 1, st);
            if (stop_at(2 * L + 1, out, 0, 0)) return 2;
        }
        return 0;
    };
    auto fused = [&](int L, const float* in, int np, float* out) -> int {
        const LayerDev& l = e->layers[L];
        if (l.fused_v3)
            BD_CHECK(e, launch_sep_fused3(l.plan, in, l.dw_w, l.dw_b, e->h_folded.data() + l.d.b, out, np, l.d.h_in, l.d.w_in, l.d.stride,
                                          e->num_sms, st, e->cta_pairs));
        else
            BD_CHECK(e, launch_sep_fused(l.plan, in, l.dw_w, l.dw_b, l.b, out, np, l.d.h_in, l.d.w_in, l.d.stride,
                                         e->num_sms, st));
        mark(e, CAT_PW + L - 1, st);
        if (stop_stage == 2 * L) { e->last_error = "stage is fused away (depthwise output stays in shared memory)"; return 1; }
        if (stop_at(2 * L + 1, out, 0, 0)) return 2;
        return 0;
    };
    for (int64_t big = 0; big < P; big += e->S2) {
        const int nb = static_cast<int>(std::min<int64_t>(e->S2, P - big));
        // ---------------- frontend: log-mel of the whole late batch in one launch (24.6 KB per patch)
        {
            if (run_frontend(e, job, big, nb, hop_frames, st)) return 1;
            mark(e, CAT_FRONTEND, st);
            if (stop_at(0, e->d_logmel, 0, 0)) return 0;
        }
        const LayerDev& lb = e->layers[first_late];          // boundary layer: its pointwise output lives in F_late
        // ---------------- early phase
        for (int small = 0; small < nb; small += e->S1) {
            const int ns = std::min(e->S1, nb - small);
            const LayerDev& l1 = e->layers[0];
            float* cur = e->d_F_early;
            float* alt = e->d_F_early2;
            const float* lm0 = e->d_logmel + static_cast<int64_t>(small) * hop_frames * kMel;
            int L0 = 1;
            if (e->fuse_l12) {
                // layers 1 + 2 (conv1 -> depthwise -> pointwise) in one kernel: log-mel in, layer-2 output out
                const LayerDev& l2 = e->layers[1];
                if (e->l12_v2)
                    BD_CHECK(e, launch_l12_fused2(l2.plan, lm0, hop_frames, ns, l1.w, l1.b, l2.dw_w, l2.dw_b,
                                                  e->h_folded.data() + l2.d.b, cur, e->num_sms, st));
                else
                    BD_CHECK(e, launch_l12_fused(l2.plan, lm0, hop_frames, ns, l1.w, l1.b, l2.dw_w, l2.dw_b, l2.b, cur,
                                                 e->num_sms, st));
                mark(e, CAT_CONV1, st);
                if (stop_stage == 1 || stop_stage == 2) {
                    e->last_error = "stage is fused away (layers 1-2 run as one kernel)";
                    return 1;
                }
                if (stop_at(3, cur, 0, 0)) return 0;
                L0 = 2;
            } else if (e->fuse_conv1 && !e->layers[1].fused) {
                // layer 1 + layer-2 depthwise in one kernel, then layer-2 pointwise
                const LayerDev& l2 = e->layers[1];
                BD_CHECK(e, launch_conv1_dw2(lm0, hop_frames, ns, l1.w, l1.b, l2.dw_w, l2.dw_b, dw_mode,
                                             reinterpret_cast<float*>(e->d_H_early), reinterpret_cast<__half*>(e->d_H_early),
                                             reinterpret_cast<__half*>(e->d_H_early + e->H_early_plane_bytes), st));
                mark(e, CAT_CONV1, st);
                if (stop_stage == 1) { e->last_error = "stage is fused away (layer-1 output stays in shared memory)"; return 1; }
                if (stop_at(2, e->d_H_early, dw_mode == 0 ? 0 : l2.nsplit, e->H_early_plane_bytes)) return 0;
                const int rc2 = unfused(1, nullptr, ns, e->d_H_early, e->H_early_plane_bytes, 0, cur, false, true);
                if (rc2) return rc2 == 2 ? 0 : rc2;
                L0 = 2;
            } else {
                BD_CHECK(e, launch_conv1(lm0, hop_frames, ns, l1.w, l1.b, cur, st));
                mark(e, CAT_CONV1, st);
                if (stop_at(1, cur, 0, 0)) return 0;
            }
            for (int L = L0; L < first_late; ++L) {
                int rc;
                if (e->layers[L].fused) {
                    rc = fused(L, cur, ns, alt);
                    std::swap(cur, alt);
                } else {
                    rc = unfused(L, cur, ns, e->d_H_early, e->H_early_plane_bytes, 0, cur, true, true);
                }
                if (rc) return rc == 2 ? 0 : rc;
            }
            const int64_t row_off = static_cast<int64_t>(small) * lb.h_out * lb.w_out;
            int rc;
            if (lb.fused) rc = fused(first_late, cur, ns, e->d_F_late + row_off * lb.d.cout);
            else rc = unfused(first_late, cur, ns, e->d_H_late, e->H_late_plane_bytes, row_off, nullptr, true, false);
            if (rc) return rc == 2 ? 0 : rc;
        }
        // ---------------- late phase
        float* cur = e->d_F_late;
        float* alt = e->d_F_late2;
        for (int L = first_late; L < BD_N_LAYERS; ++L) {
            int rc = 0;
            if (L == first_late) {
                if (!lb.fused) rc = unfused(L, nullptr, nb, e->d_H_late, e->H_late_plane_bytes, 0, cur, false, true);
            } else if (e->layers[L].fused) {
                rc = fused(L, cur, nb, alt);
                std::swap(cur, alt);
            } else {
                rc = unfused(L, cur, nb, e->d_H_late, e->H_late_plane_bytes, 0, cur, true, true);
            }
            if (rc) return rc == 2 ? 0 : rc;
        }
        const LayerDev& last = e->layers[BD_N_LAYERS - 1];
        BD_CHECK(e, launch_pool_head(cur, nb, last.h_out * last.w_out, e->d_headW, e->d_headB, e->n_classes,
                                     d_emb ? d_emb + big * kEmb : nullptr, d_act + big * e->n_classes, st));
        mark(e, CAT_POOL, st);
    }
    return 0;
}

// Run (or replay) the chunk on s_compute.
int run_chunk(bd_engine* e, const float* x, int64_t n, int hop_frames, float* d_act, float* d_emb, int64_t P,
              bool allow_graph = true) {
    if (P <= 0) return 0;
    FrontJob job;
    job.x = x;
    job.n = n;
    if (!e->use_graph || e->profiling || !allow_graph) return enqueue_chunk(e, job, hop_frames, d_act, d_emb, P, e->s_compute);
    GraphKey key{x, n, hop_frames, d_act, d_emb};
    auto it = e->graphs.find(key);
    if (it == e->graphs.end()) {
        if (e->graphs.size() >= 32) {           // bounded cache
            for (auto& kv : e->graphs) cudaGraphExecDestroy(kv.second);
            e->graphs.clear();
            e->graph_launches.clear();
        }
        const int64_t before = e->launch_count;
        BD_CHECK(e, cudaStreamBeginCapture(e->s_compute, cudaStreamCaptureModeThreadLocal));
        const int rc = enqueue_chunk(e, job, hop_frames, d_act, d_emb, P, e->s_compute);
        cudaGraph_t g = nullptr;
        cudaError_t ce = cudaStreamEndCapture(e->s_compute, &g);
        const int64_t per_graph = e->launch_count - before;
        e->launch_count = before;
        if (rc != 0) { if (g) cudaGraphDestroy(g); return rc; }
        BD_CHECK(e, ce);
        cudaGraphExec_t ge = nullptr;
        cudaError_t ie = cudaGraphInstantiate(&ge, g, 0);
        cudaGraphDestroy(g);
        BD_CHECK(e, ie);
        it = e->graphs.emplace(key, ge).first;
        e->graph_launches[key] = per_graph;
    }
    BD_CHECK(e, cudaGraphLaunch(it->second, e->s_compute));
    e->launch_count += e->graph_launches[key];
    return 0;
}

struct StopWatch {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    double s() const { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};

int ensure_slot(bd_engine* e, Slot& s, int64_t n, int64_t P) {
    StopWatch sw;
    struct Acc { bd_engine* e; StopWatch& sw; bool on = false; ~Acc() { if (on) { e->st_alloc_s += sw.s(); e->st_allocs++; } } } acc{e, sw};
    if (n > s.in_cap || P > s.out_cap) acc.on = true;
    if (n > s.in_cap) {
        if (s.d_in) cudaFree(s.d_in);
        s.d_in = nullptr;
        s.in_cap = 0;
        const int64_t cap = ((n + 4095) / 4096) * 4096 + 64;
        BD_CHECK(e, cudaMalloc(&s.d_in, cap * sizeof(float)));
        s.in_cap = cap;
    }
    if (P > s.out_cap) {
        if (s.d_act) cudaFree(s.d_act);
        if (s.d_emb) cudaFree(s.d_emb);
        s.d_act = s.d_emb = nullptr;
        s.out_cap = 0;
        const int64_t cap = ((P + 255) / 256) * 256;
        BD_CHECK(e, cudaMalloc(&s.d_act, cap * e->n_classes * sizeof(float)));
        BD_CHECK(e, cudaMalloc(&s.d_emb, cap * kEmb * sizeof(float)));
        s.out_cap = cap;
    }
    return 0;
}

bool is_pinned_host(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

// pinned staging for the results of a slot (grown on demand)
int ensure_staging(bd_engine* e, Slot& s, int64_t act_floats, int64_t emb_floats) {
    if (act_floats > s.h_act_cap) {
        if (s.h_act) cudaFreeHost(s.h_act);
        s.h_act = nullptr; s.h_act_cap = 0;
        const int64_t cap = ((act_floats + 4095) / 4096) * 4096;
        BD_CHECK(e, cudaHostAlloc(reinterpret_cast<void**>(&s.h_act), cap * sizeof(float), cudaHostAllocDefault));
        s.h_act_cap = cap;
    }
    if (emb_floats > s.h_emb_cap) {
        if (s.h_emb) cudaFreeHost(s.h_emb);
        s.h_emb = nullptr; s.h_emb_cap = 0;
        const int64_t cap = ((emb_floats + 65535) / 65536) * 65536;
        BD_CHECK(e, cudaHostAlloc(reinterpret_cast<void**>(&s.h_emb), cap * sizeof(float), cudaHostAllocDefault));
        s.h_emb_cap = cap;
    }
    return 0;
}

// Where the device->host copies of a slot go: straight to the caller's buffer when that is pinned, else to the slot's
// pinned staging (bd_wait then finishes with a host memcpy).
int route_outputs(bd_engine* e, Slot& s, float* act, float* emb) {
    const int64_t na = s.P * e->n_classes, ne = emb ? s.P * kEmb : 0;
    const bool act_direct = is_pinned_host(act), emb_direct = emb == nullptr || is_pinned_host(emb);
    if (ensure_staging(e, s, act_direct ? 0 : na, emb_direct ? 0 : ne)) return 1;
    s.user_act = act_direct ? nullptr : act;
    s.user_emb = emb_direct ? nullptr : emb;
    s.dst_act = act_direct ? act : s.h_act;
    s.dst_emb = emb == nullptr ? nullptr : (emb_direct ? emb : s.h_emb);
    return 0;
}

struct SlotOut { float* act; float* emb; };

// ---- launching what is pending ------------------------------------------------------------------------------------
// One chunk alone: the chunk pipeline as before (CUDA graph for long chunks).  Several chunks: ONE pass of the CNN over
// all of them (a 208-patch chunk on its own leaves most of the 148 persistent CTAs of every kernel idle).
int launch_group(bd_engine* e, const std::vector<int>& group, const std::vector<SlotOut>& outs) {
    const int hop = e->slots[group[0]].hop;
    cudaEvent_t ev_pass = e->ev_pass[e->pass_seq & 1];
    e->pass_seq++;
    // 16 kHz mono int16 PCM needs no resampler: the frontend reads it as it lies in the slot (LogmelSeg::fmt = 1), which
    // saves a conversion kernel per chunk and 6 bytes of HBM traffic per sample
    auto direct16 = [&](const Slot& s) {
        return s.has_pcm && s.pcm_fmt == 1 && s.pcm_channels == 1 && s.pcm_rate == 16000 && !e->frontend_v1 && s.P <= e->S2;
    };
    for (int si : group) {
        Slot& s = e->slots[si];
        BD_CHECK(e, cudaStreamWaitEvent(e->s_compute, s.ev_in, 0));
        if (s.has_pcm && !direct16(s)) {
            bd_engine::Resampler ident{1, 1, 1, nullptr};
            bd_engine::Resampler* r = &ident;
            if (s.pcm_rate != 16000 && get_resampler(e, s.pcm_rate, &r)) return 1;
            if (run_resample(e, r, s.d_pcm, s.pcm_fmt, s.pcm_channels, s.pcm_frames, s.d_in, s.n, e->s_compute)) return 1;
        }
    }
    const int pass_id = static_cast<int>(e->pass_seq - 1);
    trace_point(e, 2, pass_id, static_cast<int>(group.size()), e->s_compute);
    if (group.size() == 1 && !direct16(e->slots[group[0]])) {
        Slot& s = e->slots[group[0]];
        if (run_chunk(e, s.d_in, s.n, hop, s.d_act, s.want_emb ? s.d_emb : nullptr, s.P, s.P >= 512)) return 1;
        BD_CHECK(e, cudaEventRecord(s.ev_comp, e->s_compute));
        BD_CHECK(e, cudaEventRecord(ev_pass, e->s_compute));
        trace_point(e, 3, pass_id, 1, e->s_compute);
        BD_CHECK(e, cudaStreamWaitEvent(e->s_out, s.ev_comp, 0));
        BD_CHECK(e, cudaMemcpyAsync(outs[0].act, s.d_act, s.P * e->n_classes * sizeof(float), cudaMemcpyDeviceToHost, e->s_out));
        if (s.want_emb)
            BD_CHECK(e, cudaMemcpyAsync(outs[0].emb, s.d_emb, s.P * kEmb * sizeof(float), cudaMemcpyDeviceToHost, e->s_out));
        BD_CHECK(e, cudaEventRecord(s.ev_out, e->s_out));
        trace_point(e, 4, group[0], 0, e->s_out);
        trace_point(e, 6, pass_id, 0, nullptr);
        e->batches++; e->batched_chunks++;
        return 0;
    }
    FrontJob job;
    const int tail_slots = (kPatchFrames + hop - 1) / hop - 1;      // slots between two chunks (0 at hop 96, 1 at hop 48)
    int64_t g = 0;
    bool any_emb = false;
    std::vector<int64_t> g0(group.size());
    for (size_t i = 0; i < group.size(); ++i) {
        Slot& s = e->slots[group[i]];
        const bool last = i + 1 == group.size();
        g0[i] = g;
        LogmelSeg& sg = job.segs[job.n_segs++];
        sg.fmt = direct16(s) ? 1 : 0;
        sg.x = sg.fmt ? reinterpret_cast<const float*>(s.d_pcm) : s.d_in;
        sg.n_valid = s.n;
        sg.frame_begin = 0;
        sg.row_begin = static_cast<int>(g * hop);
        const int64_t g_next = g + s.P + tail_slots;
        sg.n_rows = static_cast<int>(last ? (s.P - 1) * hop + kPatchFrames : (g_next - g) * hop);
        g = last ? g + s.P : g_next;
        any_emb = any_emb || s.want_emb;
    }
    const int64_t P_total = g;
    if (P_total > e->S2) return fail(e, "internal: coalesced batch larger than one late batch");
    const int set = e->batch_set;
    e->batch_set ^= 1;
    if (!e->d_act_batch[set]) BD_CHECK(e, cudaMalloc(&e->d_act_batch[set], static_cast<size_t>(e->S2) * e->n_classes * sizeof(float)));
    if (any_emb && !e->d_emb_batch[set]) BD_CHECK(e, cudaMalloc(&e->d_emb_batch[set], static_cast<size_t>(e->S2) * kEmb * sizeof(float)));
    BD_CHECK(e, cudaStreamWaitEvent(e->s_compute, e->ev_batch_out[set], 0));   // the set's previous results have left
    if (enqueue_chunk(e, job, hop, e->d_act_batch[set], any_emb ? e->d_emb_batch[set] : nullptr, P_total, e->s_compute))
        return 1;
    Slot& s0 = e->slots[group[0]];
    BD_CHECK(e, cudaEventRecord(s0.ev_comp, e->s_compute));
    BD_CHECK(e, cudaEventRecord(ev_pass, e->s_compute));
    trace_point(e, 3, pass_id, static_cast<int>(group.size()), e->s_compute);
    BD_CHECK(e, cudaStreamWaitEvent(e->s_out, s0.ev_comp, 0));
    for (size_t i = 0; i < group.size(); ++i) {
        Slot& s = e->slots[group[i]];
        BD_CHECK(e, cudaMemcpyAsync(outs[i].act, e->d_act_batch[set] + g0[i] * e->n_classes,
                                    s.P * e->n_classes * sizeof(float), cudaMemcpyDeviceToHost, e->s_out));
        if (s.want_emb)
            BD_CHECK(e, cudaMemcpyAsync(outs[i].emb, e->d_emb_batch[set] + g0[i] * kEmb, s.P * kEmb * sizeof(float),
                                        cudaMemcpyDeviceToHost, e->s_out));
        BD_CHECK(e, cudaEventRecord(s.ev_out, e->s_out));
        trace_point(e, 4, group[i], 0, e->s_out);
    }
    trace_point(e, 6, pass_id, 0, nullptr);
    BD_CHECK(e, cudaEventRecord(e->ev_batch_out[set], e->s_out));
    e->batches++; e->batched_chunks += static_cast<int64_t>(group.size());
    return 0;
}

// Launch pending chunks, grouped into passes of chunks with the same hop that fit one late batch.  only_arrived: stop at
// the first chunk whose input is still on its way and launch ONE pass (dispatcher); otherwise everything (explicit flush).
// Caller holds e->mu.  On failure the affected slots are released (state 0) so the engine stays usable.
int launch_pending(bd_engine* e, bool only_arrived) {
    size_t pos = 0;
    int rc = 0;
    while (pos < e->pending.size() && rc == 0) {
        std::vector<int> group;
        std::vector<SlotOut> outs;
        const int hop = e->slots[e->pending[pos]].hop;
        const int tail_slots = (kPatchFrames + hop - 1) / hop - 1;
        int64_t g = 0;
        while (pos < e->pending.size() && static_cast<int>(group.size()) < kMaxLogmelSegs) {
            Slot& s = e->slots[e->pending[pos]];
            if (s.hop != hop) break;
            if (!group.empty() && g + tail_slots + s.P > e->S2) break;
            if (only_arrived && !s.arrived) break;
            g += (group.empty() ? 0 : tail_slots) + s.P;
            group.push_back(e->pending[pos]);
            outs.push_back(SlotOut{s.dst_act, s.want_emb ? s.dst_emb : nullptr});
            ++pos;
            if (g >= e->S2) break;
        }
        if (only_arrived && group.size() >= 6) {
            // Quantisation-aware pass size: the 6x4 layers (8-12, 30 % of a pass) run one 5-patch tile per SM and round,
            // so a pass whose tile count sits just above a multiple of the SM count pays a whole extra round for a
            // few tiles.  Up to two trailing chunks are left for the next pass when that fills the rounds better
            // (they are at the head of the queue then, so nothing is deferred twice).
            auto fill = [&](int64_t patches) {
                const int64_t tiles = (patches + 4) / 5;
                const int64_t rounds = (tiles + e->num_sms - 1) / e->num_sms;
                return static_cast<double>(tiles) / static_cast<double>(rounds * e->num_sms);
            };
            int64_t pk = g;
            double best = fill(g);
            int drop = 0;
            for (int k = 1; k <= 2; ++k) {
                pk -= e->slots[group[group.size() - k]].P + tail_slots;
                const double f = fill(pk);
                if (f > best + 0.03) { best = f; drop = k; }
            }
            for (int k = 0; k < drop; ++k) { group.pop_back(); outs.pop_back(); --pos; }
        }
        rc = launch_group(e, group, outs);
        for (int si : group) e->slots[si].state = rc == 0 ? 2 : 0;
        if (only_arrived) break;
    }
    if (rc != 0)
        for (; pos < e->pending.size(); ++pos) e->slots[e->pending[pos]].state = 0;
    e->pending.erase(e->pending.begin(), e->pending.begin() + static_cast<long>(std::min(pos, e->pending.size())));
    if (rc != 0) e->pending.clear();
    e->cv_launched.notify_all();
    return rc;
}

int flush_pending(bd_engine* e) { return launch_pending(e, false); }

// Dispatcher policy (auto mode).  `arrived` = pending chunks, in order, whose input has reached the device.
//   * nothing in the stream            -> launch what has arrived at once (a lone chunk is not kept waiting);
//   * one pass running                 -> queue the next pass behind it only when enough has arrived to fill the GPU
//                                         (coalesce_target patches), otherwise keep collecting until the pass ends;
//   * two passes in the stream         -> wait.
// So an input-bound caller gets every chunk computed as soon as it lands, and a compute-bound one gets passes that
// grow towards late_patches -- without either having to call anything.
void dispatcher_main(bd_engine* e) {
    cudaSetDevice(e->device);
    std::unique_lock<std::recursive_mutex> lk(e->mu);
    while (true) {
        e->cv_work.wait(lk, [&] { return e->stop || (!e->pending.empty() && (e->auto_flush || e->flush_req)); });
        if (e->stop) return;
        if (e->flush_req) {                         // explicit flush: everything, whatever has or has not arrived
            e->flush_req = false;
            launch_pending(e, false);
            continue;
        }
        int n_arr = 0;
        int64_t arrived = 0;
        for (int si : e->pending) {
            Slot& s = e->slots[si];
            if (!s.arrived) {
                if (cudaEventQuery(s.ev_in) != cudaSuccess) { cudaGetLastError(); break; }
                s.arrived = true;
            }
            ++n_arr;
            arrived += s.P;
        }
        int inflight = 0;
        for (int back = 1; back <= 2 && back <= e->pass_seq; ++back)
            if (cudaEventQuery(e->ev_pass[(e->pass_seq - back) & 1]) == cudaErrorNotReady) ++inflight;
        cudaGetLastError();
        const bool go = n_arr > 0 && (inflight == 0 || (inflight == 1 && arrived >= e->coalesce_target));
        if (go) {
            StopWatch sw;
            launch_pending(e, true);
            e->st_launch_s += sw.s();
        } else {
            if (inflight == 2) e->st_sleep_full++;
            else if (n_arr == 0) e->st_sleep_noarr++;
            else e->st_sleep_small++;
            lk.unlock();
            std::this_thread::sleep_for(std::chrono::microseconds(15));
            lk.lock();
        }
    }
}

// A chunk has been queued: wake the dispatcher.
int maybe_flush(bd_engine* e) {
    e->cv_work.notify_one();
    return 0;
}

// fp32 -> hi/lo fp16 planes on the host (weights only; done once)
void split_f16(const float* src, size_t n, std::vector<__half>& hi, std::vector<__half>& lo) {
    hi.resize(n);
    lo.resize(n);
    for (size_t i = 0; i < n; ++i) {
        const __half h = __float2half_rn(src[i]);
        hi[i] = h;
        lo[i] = __float2half_rn(src[i] - __half2float(h));
    }
}

// ------------------------------------------------------------------------------------------ resampler design
double bessel_i0(double x) {
    double sum = 1.0, term = 1.0;
    const double q = x * x / 4.0;
    for (int k = 1; k < 200; ++k) {
        term *= q / (static_cast<double>(k) * k);
        sum += term;
        if (term < 1e-18 * sum) break;
    }
    return sum;
}

int64_t gcd64(int64_t a, int64_t b) { while (b) { int64_t t = a % b; a = b; b = t; } return a; }

}  // namespace

// ============================================================================================ C ABI
extern "C" {

int32_t bd_abi_version(void) { return BD_ABI_VERSION; }

int32_t bd_frames_for(int64_t n_samples, int32_t hop_frames, int64_t* n_padded, int64_t* n_stft_frames,
                      int64_t* n_patches) {
    if (n_samples < 0 || hop_frames < 1 || hop_frames > kPatchFrames) return 1;
    frames_for(n_samples, hop_frames, n_padded, n_stft_frames, n_patches);
    return 0;
}

const char* bd_last_error(const bd_engine* e) { return e ? e->last_error.c_str() : g_create_error.c_str(); }

int64_t bd_launch_count(const bd_engine* e) { return e ? e->launch_count : 0; }

void bd_engine_destroy(bd_engine* e) {
    if (!e) return;
    if (e->dispatcher.joinable()) {
        {
            std::lock_guard<std::recursive_mutex> lk(e->mu);
            e->stop = true;
        }
        e->cv_work.notify_all();
        e->dispatcher.join();
    }
    cudaSetDevice(e->device);
    cudaDeviceSynchronize();
    for (auto& kv : e->graphs) cudaGraphExecDestroy(kv.second);
    for (int i = 0; i < 2; ++i) {
        cudaFree(e->d_act_batch[i]); cudaFree(e->d_emb_batch[i]);
        if (e->ev_batch_out[i]) cudaEventDestroy(e->ev_batch_out[i]);
    }
    for (int i = 0; i < 2; ++i) if (e->ev_pass[i]) cudaEventDestroy(e->ev_pass[i]);
    for (auto& s : e->slots) {
        cudaFree(s.d_in); cudaFree(s.d_pcm); cudaFree(s.d_act); cudaFree(s.d_emb);
        if (s.h_act) cudaFreeHost(s.h_act);
        if (s.h_emb) cudaFreeHost(s.h_emb);
        if (s.ev_in) cudaEventDestroy(s.ev_in);
        if (s.ev_comp) cudaEventDestroy(s.ev_comp);
        if (s.ev_out) cudaEventDestroy(s.ev_out);
    }
    for (auto& l : e->layers) { cudaFree(l.w_hi); cudaFree(l.w_lo); }
    for (auto& kv : e->resamplers) { cudaFree(kv.second.d_taps); cudaFree(kv.second.d_h_hi); cudaFree(kv.second.d_h_lo); }
    cudaFree(e->d_folded); cudaFree(e->d_tab); cudaFree(e->d_headW); cudaFree(e->d_headB);
    cudaFree(e->d_logmel); cudaFree(e->d_F_early2); cudaFree(e->d_F_late2); cudaFree(e->d_F_early); cudaFree(e->d_H_early); cudaFree(e->d_F_late); cudaFree(e->d_H_late);
    if (e->s_compute) cudaStreamDestroy(e->s_compute);
    if (e->s_in) cudaStreamDestroy(e->s_in);
    if (e->s_out) cudaStreamDestroy(e->s_out);
    delete e;
}

int32_t bd_engine_create(const bd_config* cfg, const bd_weights* w, bd_engine** out, char* err, size_t err_len) {
    auto set_err = [&](const std::string& m) {
        g_create_error = m;
        if (err && err_len) { std::strncpy(err, m.c_str(), err_len - 1); err[err_len - 1] = 0; }
        return 1;
    };
    if (!cfg || !w || !out) return set_err("null argument");
    *out = nullptr;
    if (cfg->precision != BD_PRECISION_FP32_SIMT && cfg->precision != BD_PRECISION_FP16X1 &&
        cfg->precision != BD_PRECISION_FP16X3 && cfg->precision != BD_PRECISION_FP16F8)
        return set_err("precision must be 0 (fp32 SIMT), 1 (fp16), 2 (fp16 + fp8 corrections) or 3 (fp16 x3 split)");
    if (w->n_classes < 1 || w->n_classes > kMaxClasses) return set_err("n_classes out of range");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return set_err(std::string("no CUDA device: ") + cudaGetErrorString(ce) + " (this library has no CPU path)");
    if (cfg->device < 0 || cfg->device >= ndev) return set_err("device ordinal out of range");
    if (cudaSetDevice(cfg->device) != cudaSuccess) return set_err("cudaSetDevice failed");
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, cfg->device);
    if (prop.major != 10) return set_err(std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) +
                                          "; this library is built for sm_100a (B200) only");

    bd_engine* e = new bd_engine();
    e->device = cfg->device;
    e->num_sms = prop.multiProcessorCount;
    e->precision = cfg->precision;
    // Measured on B200 (profiles/r1_summary.md): with persistent fused kernels the step gets faster all the way up to
    // one sub-batch per audio-hour (3.48 ms at 1024/4096, 3.34 ms at 4096/4096): tails and launches cost more than L2
    // residency of the inter-layer activations would win (148..888-patch sub-batches: 3.57-3.80 ms).
    e->S1 = cfg->early_patches > 0 ? cfg->early_patches : 4096;
    e->S2 = cfg->late_patches > 0 ? cfg->late_patches : 4096;
    e->S2 = std::max(e->S1, (e->S2 / e->S1) * e->S1);           // late batch = whole number of early batches
    e->use_graph = cfg->use_graph != 0;
    e->n_classes = w->n_classes;
    auto bail = [&](const std::string& m) { std::string mm = m; bd_engine_destroy(e); return set_err(mm); };
#define BD_CREATE(expr)                                                                      \
    do {                                                                                     \
        cudaError_t _err = (expr);                                                           \
        if (_err != cudaSuccess) return bail(std::string(#expr) + ": " + cudaGetErrorString(_err)); \
    } while (0)

    BD_CREATE(cudaStreamCreateWithFlags(&e->s_compute, cudaStreamNonBlocking));
    BD_CREATE(cudaStreamCreateWithFlags(&e->s_in, cudaStreamNonBlocking));
    BD_CREATE(cudaStreamCreateWithFlags(&e->s_out, cudaStreamNonBlocking));
    BD_CREATE(frontend_init_device());
    BD_CREATE(frontend2_init_device());
    {
        const char* v1 = getenv("BD_FRONTEND_V1");
        e->frontend_v1 = v1 && atoi(v1) != 0;
    }
    BD_CREATE(layers_init_device());
    if (e->precision != BD_PRECISION_FP32_SIMT) BD_CREATE(pw_gemm_init_device());
    if (e->precision != BD_PRECISION_FP32_SIMT) BD_CREATE(sep_fused3_init_device());
    if (e->precision != BD_PRECISION_FP32_SIMT) BD_CREATE(l12_fused2_init_device());
    BD_CREATE(resample_tc_init_device());

    // ---- weights
    e->h_folded.assign(w->folded, w->folded + w->folded_len);
    if (e->precision != BD_PRECISION_FP32_SIMT) {
        // fp16 operand headroom: the depthwise taps and biases are scaled by 2^-4 (exact), so every depthwise output
        // reaches the fp16 hi/lo split 16 times smaller and the pointwise epilogue multiplies it back (out_scale).
        // Activations up to ~1e6 then keep their full 22 bits (beyond that half_sat() saturates instead of producing
        // inf); post-ReLU values below ~1e-3 fall into fp16's subnormal range, an absolute error < 1e-6.
        for (int L = 1; L < BD_N_LAYERS; ++L) {
            const bd_layer_desc& d = w->layers[L];
            if (d.dw_w < 0 || d.dw_b < 0) continue;
            for (int i = 0; i < 9 * d.cin; ++i) e->h_folded[d.dw_w + i] *= kActScale;
            for (int i = 0; i < d.cin; ++i) e->h_folded[d.dw_b + i] *= kActScale;
        }
    }
    BD_CREATE(cudaMalloc(&e->d_folded, w->folded_len * sizeof(float)));
    BD_CREATE(cudaMemcpy(e->d_folded, e->h_folded.data(), w->folded_len * sizeof(float), cudaMemcpyHostToDevice));
    BD_CREATE(cudaMalloc(&e->d_headW, sizeof(float) * kEmb * w->n_classes));
    BD_CREATE(cudaMemcpy(e->d_headW, w->head_kernel, sizeof(float) * kEmb * w->n_classes, cudaMemcpyHostToDevice));
    BD_CREATE(cudaMalloc(&e->d_headB, sizeof(float) * w->n_classes));
    BD_CREATE(cudaMemcpy(e->d_headB, w->head_bias, sizeof(float) * w->n_classes, cudaMemcpyHostToDevice));

    // ---- frontend tables: window as given; mel -> per-band contiguous non-zero runs
    {
        FrontendTables t;
        std::memset(&t, 0, sizeof(t));
        std::memcpy(t.window, w->window, sizeof(float) * kWin);
        int off = 0;
        for (int m = 0; m < kMel; ++m) {
            int first = -1, lastnz = -1;
            for (int k = 0; k < kBins; ++k)
                if (w->mel[k * kMel + m] != 0.f) { if (first < 0) first = k; lastnz = k; }
            if (first < 0) { first = 0; lastnz = -1; }
            const int len = lastnz - first + 1;
            if (off + len > kMelNnzMax) return bail("mel matrix has too many non-zeros for the sparse table");
            t.mel_start[m] = first; t.mel_len[m] = len; t.mel_off[m] = off;
            for (int j = 0; j < len; ++j) t.mel_w[off + j] = w->mel[(first + j) * kMel + m];
            off += len;
        }
        if (!frontend2_build_mel(t, &e->mel_param)) return bail("mel matrix does not fit the frontend's band tables");
        BD_CREATE(cudaMalloc(&e->d_tab, sizeof(FrontendTables)));
        BD_CREATE(cudaMemcpy(e->d_tab, &t, sizeof(t), cudaMemcpyHostToDevice));
    }

    // ---- layers + buffer plan
    size_t f_early = 0, h_early = 0, f_late = 0, h_late = 0;    // elements per patch
    e->layers.resize(BD_N_LAYERS);
    const int first_late = e->first_late;                        // layer 7 (index 6): its depthwise output feeds the late phase
    // which separable blocks run fused: bit (L-2) for layer L.  Measured on B200 (profiles/fusion_r1.md): with the
    // current register-fed producers only layer 3 (stride 2, K=64) beats the two-kernel path, so that is the default.
    // Default (fuse_mask < 0), measured on B200 (profiles/r1_summary.md): layers 1+2 in the warp-specialised
    // l12_fused2_kernel, layers 3..6 and 8..12 in sep_fused3_kernel (TMA-staged stencil input, whole-patch tiles for
    // the 6x4 layers); layer 7 (stride 2 into 6x4: five small boxes per tile) and the 3x2 layers 13-14 stay as
    // depthwise + GEMM pairs.
    const int cfg_mask = cfg->fuse_mask < 0 ? (BD_FUSE_L12V2 | BD_FUSE_V3 | BD_FUSE_CONV1_DW2 | 0x7DE) : cfg->fuse_mask;
    const int fuse_mask = e->precision == BD_PRECISION_FP32_SIMT ? 0 : (cfg_mask & 0x1FFF);
    const bool use_v3 = (cfg_mask & BD_FUSE_V3) != 0;
    e->fuse_conv1 = (cfg_mask & BD_FUSE_CONV1_DW2) != 0;
    e->l12_v2 = e->precision != BD_PRECISION_FP32_SIMT && (cfg_mask & BD_FUSE_L12V2) != 0;
    e->cta_pairs = (cfg_mask & BD_FUSE_PAIR) != 0;
    e->tc_resample = (cfg_mask & BD_FUSE_NO_TC_RESAMPLE) == 0;
    e->fuse_l12 = e->precision != BD_PRECISION_FP32_SIMT && (cfg_mask & (BD_FUSE_L12 | BD_FUSE_L12V2)) != 0;
    for (int L = 0; L < BD_N_LAYERS; ++L) {
        LayerDev& l = e->layers[L];
        l.d = w->layers[L];
        if ((L == 0) != (l.d.kind == 0)) return bail("layer table: layer 1 must be the conv, the rest separable");
        l.h_out = l.d.h_in / l.d.stride;
        l.w_out = l.d.w_in / l.d.stride;
        l.late = L >= first_late;
        l.fused = L >= 1 && ((fuse_mask >> (L - 1)) & 1) && (l.d.w_in / l.d.stride) % 4 == 0 && l.d.cin % 4 == 0;
        l.fused_v3 = l.fused && use_v3 && l.d.cout % 128 == 0 &&
                     sep_fused3_supported(l.d.cin, l.d.cout, l.d.h_in, l.d.w_in, l.d.stride);
        auto ptr = [&](int64_t o) -> const float* { return o >= 0 ? e->d_folded + o : nullptr; };
        l.dw_w = ptr(l.d.dw_w); l.dw_b = ptr(l.d.dw_b); l.w = ptr(l.d.w); l.b = ptr(l.d.b);
        const size_t out_px = static_cast<size_t>(l.h_out) * l.w_out;
        if (L == 0) {
            if (l.d.cin != 1 || l.d.cout != 32 || l.d.h_in != 96 || l.d.w_in != 64 || l.d.stride != 2)
                return bail("layer 1 must be 3x3/2 conv 1->32 on a 96x64 patch");
            f_early = std::max(f_early, out_px * l.d.cout);
        } else if (!l.late) {
            h_early = std::max(h_early, out_px * l.d.cin);
            f_early = std::max(f_early, out_px * l.d.cout);
        } else {
            h_late = std::max(h_late, out_px * l.d.cin);
            f_late = std::max(f_late, out_px * l.d.cout);
        }
    }
    if (e->layers.back().d.cout != kEmb) return bail("last layer must have 1024 channels");
    BD_CREATE(cudaMalloc(&e->d_logmel, sizeof(float) * kMel * (static_cast<size_t>(e->S2) * kPatchFrames)));
    // experiment knob: extra bytes between the hi and lo planes (decorrelates the DRAM mapping of the two lock-step
    // read streams of the pointwise GEMMs)
    size_t plane_pad = 0;
    if (const char* pp = getenv("BD_PLANE_PAD")) plane_pad = static_cast<size_t>(atoll(pp)) & ~size_t(1023);
    BD_CREATE(cudaMalloc(&e->d_F_early, sizeof(float) * f_early * e->S1));
    e->H_early_plane_bytes = sizeof(__half) * h_early * e->S1 + plane_pad;
    BD_CREATE(cudaMalloc(&e->d_H_early, sizeof(float) * h_early * e->S1 + plane_pad));
    BD_CREATE(cudaMemset(e->d_H_early, 0, sizeof(float) * h_early * e->S1 + plane_pad));
    BD_CREATE(cudaMalloc(&e->d_F_late, sizeof(float) * f_late * e->S2));
    e->H_late_plane_bytes = sizeof(__half) * h_late * e->S2 + plane_pad;
    BD_CREATE(cudaMalloc(&e->d_H_late, sizeof(float) * h_late * e->S2 + plane_pad));
    BD_CREATE(cudaMemset(e->d_H_late, 0, sizeof(float) * h_late * e->S2 + plane_pad));

    // ping-pong partners, only touched by fused separable blocks
    BD_CREATE(cudaMalloc(&e->d_F_early2, sizeof(float) * f_early * e->S1));
    BD_CREATE(cudaMalloc(&e->d_F_late2, sizeof(float) * f_late * e->S2));

    // ---- tensor-core operands: weight planes + TMA descriptors
    if (e->precision != BD_PRECISION_FP32_SIMT) {
        for (int L = 1; L < BD_N_LAYERS; ++L) {
            LayerDev& l = e->layers[L];
            // operand plan per layer.  "fp16f8": the fp16 + fp8 plan where a kernel implements it (sep_fused3 and the
            // depthwise + GEMM pairs, K a multiple of 64); the HBM-bound early layers keep the x3 split
            int nsplit = e->precision == BD_PRECISION_FP16X1 ? 1 : 3;
            if (e->precision == BD_PRECISION_FP16F8 && l.d.cin % 64 == 0 && (l.fused_v3 || !l.fused)) nsplit = 2;
            l.nsplit = nsplit;
            const size_t nw = static_cast<size_t>(l.d.cout) * l.d.cin;
            std::vector<__half> hi(nw), lo(nw);
            float out_scale;
            if (nsplit == 2)
                out_scale = split_weights_f16f8(w->folded + l.d.w, l.d.cout, l.d.cin, hi.data(),
                                                reinterpret_cast<unsigned char*>(lo.data())) / kActScale;
            else
                out_scale = split_weights_f16(w->folded + l.d.w, nw, hi.data(), lo.data()) / kActScale;
            BD_CREATE(cudaMalloc(&l.w_hi, nw * sizeof(__half)));
            BD_CREATE(cudaMalloc(&l.w_lo, nw * sizeof(__half)));
            BD_CREATE(cudaMemcpy(l.w_hi, hi.data(), nw * sizeof(__half), cudaMemcpyHostToDevice));
            BD_CREATE(cudaMemcpy(l.w_lo, lo.data(), nw * sizeof(__half), cudaMemcpyHostToDevice));
            unsigned char* H = l.late ? e->d_H_late : e->d_H_early;
            const size_t plane = l.late ? e->H_late_plane_bytes : e->H_early_plane_bytes;
            const int S = l.late ? e->S2 : e->S1;
            const int M_max = S * l.h_out * l.w_out;
            const char* perr = nullptr;
            // stand-alone GEMMs of wide layers use 128x256 tiles (an MMA's operand fetch is ~ (4 KB + 32 B x N) / 64 B per
            // clock: N = 256 amortises the A tile twice as far); the fused kernel wants 128-row weight boxes
            int bn = 0;
            {
                static const int bn_env = [] { const char* v = getenv("BD_PW_BN256"); return v ? atoi(v) : 1; }();
                if (bn_env && !l.fused_v3 && !l.fused && l.d.cout % 256 == 0 && l.d.cin >= 512) bn = 256;
            }
            cudaError_t pe = pw_gemm_make_plan(&l.plan, reinterpret_cast<__half*>(H), reinterpret_cast<__half*>(H + plane),
                                               M_max, l.d.cin, l.w_hi, l.w_lo, l.d.cout, nsplit, bn, out_scale, &perr);
            if (pe != cudaSuccess) return bail(std::string("pointwise plan for layer ") + std::to_string(L + 1) + ": " +
                                               (perr ? perr : cudaGetErrorString(pe)));
        }
    }

    // ---- host-chunk slots
    int ns = cfg->n_slots <= 0 ? 2 : std::min(cfg->n_slots, 64);
    e->slots.resize(ns);
    for (int i = 0; i < 2; ++i) BD_CREATE(cudaEventCreateWithFlags(&e->ev_batch_out[i], cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) BD_CREATE(cudaEventCreateWithFlags(&e->ev_pass[i], cudaEventDisableTiming));
    e->coalesce_target = std::max<int64_t>(1, static_cast<int64_t>(e->S2) / 2);
    if (const char* ct = getenv("BD_COALESCE_PATCHES")) e->coalesce_target = std::max<int64_t>(1, atoll(ct));
    for (auto& s : e->slots) {
        BD_CREATE(cudaEventCreateWithFlags(&s.ev_in, cudaEventDisableTiming));
        BD_CREATE(cudaEventCreateWithFlags(&s.ev_comp, cudaEventDisableTiming));
        BD_CREATE(cudaEventCreateWithFlags(&s.ev_out, cudaEventDisableTiming));
    }
    BD_CREATE(cudaDeviceSynchronize());
#undef BD_CREATE
    e->dispatcher = std::thread(dispatcher_main, e);
    *out = e;
    return 0;
}

// explicit flush: hand everything pending to the dispatcher and wait until it has been launched
static int flush_and_wait_launched(bd_engine* e, std::unique_lock<std::recursive_mutex>& lk) {
    if (e->pending.empty()) return 0;
    e->flush_req = true;
    e->cv_work.notify_one();
    e->cv_launched.wait(lk, [&] { return e->pending.empty() || e->stop; });
    return 0;
}

int32_t bd_flush(bd_engine* e) {
    if (!e) return 1;
    std::unique_lock<std::recursive_mutex> lk(e->mu);
    return flush_and_wait_launched(e, lk);
}

int32_t bd_synchronize(bd_engine* e) {
    if (!e) return 1;
    {
        std::unique_lock<std::recursive_mutex> lk(e->mu);
        flush_and_wait_launched(e, lk);
    }
    BD_CHECK(e, cudaSetDevice(e->device));
    BD_CHECK(e, cudaStreamSynchronize(e->s_in));
    BD_CHECK(e, cudaStreamSynchronize(e->s_compute));
    BD_CHECK(e, cudaStreamSynchronize(e->s_out));
    return 0;
}

int32_t bd_predict_device(bd_engine* e, const float* d_samples, int64_t n, int32_t hop_frames, float* d_act,
                          float* d_emb, int64_t* n_patches) {
    if (!e) return 1;
    std::lock_guard<std::recursive_mutex> lk(e->mu);
    if (n < 0 || hop_frames < 1 || hop_frames > kPatchFrames) return fail(e, "bad n / hop_frames");
    BD_CHECK(e, cudaSetDevice(e->device));
    int64_t P = 0;
    frames_for(n, hop_frames, nullptr, nullptr, &P);
    if (n_patches) *n_patches = P;
    if (P > 0 && (!d_samples || !d_act)) return fail(e, "null device buffer");
    if (run_chunk(e, d_samples, n, hop_frames, d_act, d_emb, P)) return 1;
    BD_CHECK(e, cudaStreamSynchronize(e->s_compute));
    return 0;
}

int32_t bd_submit_host(bd_engine* e, int32_t slot, const float* samples, int64_t n, int32_t hop_frames, float* act,
                       float* emb, int64_t* n_patches) {
    if (!e) return 1;
    std::lock_guard<std::recursive_mutex> lk(e->mu);
    if (slot < 0 || slot >= static_cast<int>(e->slots.size())) return fail(e, "slot out of range");
    if (n < 0 || hop_frames < 1 || hop_frames > kPatchFrames) return fail(e, "bad n / hop_frames");
    BD_CHECK(e, cudaSetDevice(e->device));
    Slot& s = e->slots[slot];
    if (s.state != 0) return fail(e, "slot still in flight: call bd_wait first");
    int64_t P = 0;
    frames_for(n, hop_frames, nullptr, nullptr, &P);
    if (n_patches) *n_patches = P;
    if (P == 0) return 0;
    if (!samples || !act) return fail(e, "null host buffer");
    if (ensure_slot(e, s, n, P)) return 1;
    s.n = n; s.P = P; s.hop = hop_frames; s.want_emb = emb != nullptr; s.has_pcm = false; s.arrived = false;
    if (route_outputs(e, s, act, emb)) return 1;
    trace_point(e, 0, slot, 0, nullptr);
    BD_CHECK(e, cudaMemcpyAsync(s.d_in, samples, n * sizeof(float), cudaMemcpyHostToDevice, e->s_in));
    BD_CHECK(e, cudaEventRecord(s.ev_in, e->s_in));
    trace_point(e, 1, slot, 0, e->s_in);
    s.state = 1;
    e->pending.push_back(slot);
    return maybe_flush(e);
}

int32_t bd_wait(bd_engine* e, int32_t slot) {
    if (!e) return 1;
    cudaEvent_t ev = nullptr;
    {
        std::unique_lock<std::recursive_mutex> lk(e->mu);
        if (slot < 0 || slot >= static_cast<int>(e->slots.size())) return fail(e, "slot out of range");
        Slot& s = e->slots[slot];
        if (s.state == 0) return 0;
        BD_CHECK(e, cudaSetDevice(e->device));
        if (s.state == 1) {
            // not launched yet: the dispatcher takes it as soon as the GPU has room; without auto-flush, ask for it
            if (!e->auto_flush) { e->flush_req = true; e->cv_work.notify_one(); }
            e->cv_launched.wait(lk, [&] { return e->slots[slot].state != 1 || e->stop; });
        }
        if (s.state != 2) return fail(e, e->last_error.empty() ? "slot was released by a failed launch" : e->last_error);
        ev = s.ev_out;
    }
    const cudaError_t we = cudaEventSynchronize(ev);        // outside the lock: the other thread keeps submitting
    std::lock_guard<std::recursive_mutex> lk(e->mu);
    Slot& s = e->slots[slot];
    s.state = 0;
    BD_CHECK(e, we);
    if (s.user_act) std::memcpy(s.user_act, s.h_act, static_cast<size_t>(s.P) * e->n_classes * sizeof(float));
    if (s.user_emb) std::memcpy(s.user_emb, s.h_emb, static_cast<size_t>(s.P) * kEmb * sizeof(float));
    s.user_act = s.user_emb = nullptr;
    trace_point(e, 5, slot, 0, nullptr);
    return 0;
}

/* on = 1: start recording (drains the streams first).  on = 0: stop, drain, and write one line per record
 * "kind a b host_ms device_ms" (device_ms = -1 for host-only records; both clocks start at the bd_trace(1) call). */
int32_t bd_trace(bd_engine* e, int32_t on, char* buf, size_t len) {
    if (!e) return 1;
    if (bd_synchronize(e)) return 1;
    std::lock_guard<std::recursive_mutex> lk(e->mu);
    BD_CHECK(e, cudaSetDevice(e->device));
    if (on) {
        for (auto& r : e->trace) if (r.ev) cudaEventDestroy(r.ev);
        e->trace.clear();
        if (!e->trace_base) BD_CHECK(e, cudaEventCreate(&e->trace_base));
        BD_CHECK(e, cudaEventRecord(e->trace_base, e->s_compute));
        BD_CHECK(e, cudaEventSynchronize(e->trace_base));
        e->trace_t0 = std::chrono::steady_clock::now();
        e->tracing = true;
        return 0;
    }
    e->tracing = false;
    size_t pos = 0;
    if (buf && len) buf[0] = 0;
    for (auto& r : e->trace) {
        float ms = -1.f;
        if (r.ev) {
            cudaEventSynchronize(r.ev);
            cudaEventElapsedTime(&ms, e->trace_base, r.ev);
            cudaEventDestroy(r.ev);
            r.ev = nullptr;
        }
        if (buf && pos + 64 < len)
            pos += static_cast<size_t>(snprintf(buf + pos, len - pos, "%d %d %d %.4f %.4f\n", r.kind, r.a, r.b, r.host_ms, ms));
    }
    e->trace.clear();
    return 0;
}

int32_t bd_debug_stats(bd_engine* e, char* buf, size_t len) {
    if (!e || !buf || !len) return 1;
    std::lock_guard<std::recursive_mutex> lk(e->mu);
    snprintf(buf, len, "passes=%lld chunks=%lld launch_s=%.4f alloc_s=%.4f allocs=%lld polls{two_in_flight=%lld, waiting_for_running_pass=%lld, "
                       "nothing_arrived=%lld}",
             (long long)e->batches, (long long)e->batched_chunks, e->st_launch_s, e->st_alloc_s, (long long)e->st_allocs,
             (long long)e->st_sleep_full, (long long)e->st_sleep_small, (long long)e->st_sleep_noarr);
    return 0;
}

/* Pre-size every slot for chunks of up to n_samples 16 kHz samples / pcm_bytes of decoded PCM (0: none), so that no
 * device or pinned allocation happens while chunks are in flight (cudaMalloc and cudaHostAlloc synchronise). */
int32_t bd_reserve_slots(bd_engine* e, int64_t n_samples, int64_t pcm_bytes, int32_t hop_frames) {
    if (!e) return 1;
    std::lock_guard<std::recursive_mutex> lk(e->mu);
    if (n_samples < 0 || pcm_bytes < 0 || hop_frames < 1 || hop_frames > kPatchFrames) return fail(e, "bad arguments");
    BD_CHECK(e, cudaSetDevice(e->device));
    int64_t P = 0;
    frames_for(n_samples, hop_frames, nullptr, nullptr, &P);
    for (auto& s : e->slots) {
        if (s.state != 0) continue;
        if (ensure_slot(e, s, n_samples, P)) return 1;
        if (ensure_staging(e, s, P * e->n_classes, 0)) return 1;
        if (pcm_bytes > s.pcm_cap) {
            if (s.d_pcm) cudaFree(s.d_pcm);
            s.d_pcm = nullptr;
            s.pcm_cap = 0;
            const int64_t cap = ((pcm_bytes + 65535) / 65536) * 65536;
            BD_CHECK(e, cudaMalloc(&s.d_pcm, cap));
            s.pcm_cap = cap;
        }
    }
    for (int set = 0; set < 2; ++set)
        if (!e->d_act_batch[set])
            BD_CHECK(e, cudaMalloc(&e->d_act_batch[set], static_cast<size_t>(e->S2) * e->n_classes * sizeof(float)));
    return 0;
}

int32_t bd_set_auto_flush(bd_engine* e, int32_t on) {
    if (!e) return 1;
    std::lock_guard<std::recursive_mutex> lk(e->mu);
    e->auto_flush = on != 0;
    return 0;
}

int32_t bd_slot_state(bd_engine* e, int32_t slot) {
    if (!e) return -1;
    std::lock_guard<std::recursive_mutex> lk(e->mu);
    if (slot < 0 || slot >= static_cast<int>(e->slots.size())) return -1;
    return e->slots[slot].state;
}

int32_t bd_batch_stats(bd_engine* e, int64_t* batches, int64_t* chunks) {
    if (!e) return 1;
    std::lock_guard<std::recursive_mutex> lk(e->mu);
    if (batches) *batches = e->batches;
    if (chunks) *chunks = e->batched_chunks;
    return 0;
}

int32_t bd_predict_host(bd_engine* e, const float* samples, int64_t n, int32_t hop_frames, float* act, float* emb,
                        int64_t* n_patches) {
    if (!e) return 1;
    if (bd_wait(e, 0)) return 1;
    if (bd_submit_host(e, 0, samples, n, hop_frames, act, emb, n_patches)) return 1;
    return bd_wait(e, 0);
}

/* pinned (page-locked) host memory for the streamer's chunk ring and for result buffers: DMA reads / writes it
 * directly, so bd_submit_* returns as soon as the copies are queued */
int32_t bd_host_alloc(size_t bytes, int32_t write_combined, void** out) {
    if (!out) return 1;
    *out = nullptr;
    unsigned flags = cudaHostAllocPortable | (write_combined ? cudaHostAllocWriteCombined : 0u);
    const cudaError_t ce = cudaHostAlloc(out, bytes ? bytes : 1, flags);
    if (ce != cudaSuccess) {
        g_create_error = std::string("cudaHostAlloc: ") + cudaGetErrorString(ce);
        *out = nullptr;
        return 1;
    }
    return 0;
}

void bd_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int32_t bd_profile_device(bd_engine* e, const float* d_samples, int64_t n, int32_t hop_frames, float* ms,
                          int64_t* launches) {
    if (!e) return 1;
    std::lock_guard<std::recursive_mutex> lk(e->mu);
    if (n < 0 || hop_frames < 1 || hop_frames > kPatchFrames) return fail(e, "bad n / hop_frames");
    BD_CHECK(e, cudaSetDevice(e->device));
    int64_t P = 0;
    frames_for(n, hop_frames, nullptr, nullptr, &P);
    for (int i = 0; i < CAT_COUNT; ++i) { ms[i] = 0.f; launches[i] = 0; }
    if (P == 0) return 0;
    float* d_act = nullptr;
    BD_CHECK(e, cudaMalloc(&d_act, P * e->n_classes * sizeof(float)));
    BD_CHECK(e, cudaStreamSynchronize(e->s_compute));
    e->profiling = true;
    e->prof_events.clear();
    e->prof_cat.clear();
    cudaEvent_t start;
    cudaEventCreate(&start);
    cudaEventRecord(start, e->s_compute);
    const int64_t before = e->launch_count;
    FrontJob pjob;
    pjob.x = d_samples;
    pjob.n = n;
    const int rc = enqueue_chunk(e, pjob, hop_frames, d_act, nullptr, P, e->s_compute);
    e->profiling = false;
    e->launch_count = before;
    cudaError_t se = cudaStreamSynchronize(e->s_compute);
    cudaEvent_t prev = start;
    for (size_t i = 0; i < e->prof_events.size(); ++i) {
        float t = 0.f;
        if (rc == 0 && se == cudaSuccess && cudaEventElapsedTime(&t, prev, e->prof_events[i]) == cudaSuccess) {
            ms[e->prof_cat[i]] += t;
            launches[e->prof_cat[i]] += 1;
        }
        prev = e->prof_events[i];
    }
    cudaEventDestroy(start);
    for (auto ev : e->prof_events) cudaEventDestroy(ev);
    e->prof_events.clear();
    e->prof_cat.clear();
    cudaFree(d_act);
    if (rc) return rc;
    BD_CHECK(e, se);
    return 0;
}

int32_t bd_bench_device(bd_engine* e, const float* d_samples, int64_t n, int32_t hop_frames, float* d_act,
                        int32_t steps, float* ms_total) {
    if (!e) return 1;
    std::lock_guard<std::recursive_mutex> lk(e->mu);
    if (n < 0 || hop_frames < 1 || hop_frames > kPatchFrames || steps < 1) return fail(e, "bad arguments");
    BD_CHECK(e, cudaSetDevice(e->device));
    int64_t P = 0;
    frames_for(n, hop_frames, nullptr, nullptr, &P);
    if (P == 0 || !d_samples || !d_act) return fail(e, "nothing to run");
    cudaEvent_t e0, e1;
    BD_CHECK(e, cudaEventCreate(&e0));
    BD_CHECK(e, cudaEventCreate(&e1));
    BD_CHECK(e, cudaStreamSynchronize(e->s_compute));
    BD_CHECK(e, cudaEventRecord(e0, e->s_compute));
    int rc = 0;
    for (int i = 0; i < steps && rc == 0; ++i) rc = run_chunk(e, d_samples, n, hop_frames, d_act, nullptr, P);
    cudaEventRecord(e1, e->s_compute);
    cudaError_t se = cudaStreamSynchronize(e->s_compute);
    float ms = 0.f;
    if (rc == 0 && se == cudaSuccess) se = cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (rc) return rc;
    BD_CHECK(e, se);
    if (ms_total) *ms_total = ms;
    return 0;
}

// ------------------------------------------------------------------------------------------ resampler
int64_t bd_resample_out_len(int64_t n_frames, int32_t src_rate) {
    if (n_frames <= 0 || src_rate <= 0) return 0;
    if (src_rate == 16000) return n_frames;
    // librosa: int(np.ceil(n * ratio)) with ratio = float(target)/orig (python floats)
    const double ratio = 16000.0 / static_cast<double>(src_rate);
    return static_cast<int64_t>(std::ceil(static_cast<double>(n_frames) * ratio));
}

static int get_resampler(bd_engine* e, int src_rate, bd_engine::Resampler** out) {
    auto it = e->resamplers.find(src_rate);
    if (it == e->resamplers.end()) {
        const int64_t g = gcd64(16000, src_rate);
        const int up = static_cast<int>(16000 / g), down = static_cast<int>(src_rate / g);
        if (up > 4096) return fail(e, "unsupported sample-rate ratio (interpolation factor > 4096)");
        // Kaiser-windowed sinc at the virtual rate up*src_rate: pass-band edge 0.9136*Nyq, stop-band at Nyq of the
        // lower rate, ~125 dB (soxr "HQ": 20-bit).  See oracle/resample_oracle.py for the same design in numpy.
        const double f_low = 0.5 * std::min(16000, src_rate);
        const double fpass = 0.9136 * f_low, fstop = f_low;
        const double fc = 0.5 * (fpass + fstop);
        const double att = 125.0;
        const double beta = 0.1102 * (att - 8.7);
        const double fs_v = static_cast<double>(up) * src_rate;
        const double dw = 2.0 * M_PI * (fstop - fpass) / fs_v;
        int64_t half = static_cast<int64_t>(std::ceil((att - 8.0) / (2.285 * dw) / 2.0));
        const int tpp = 2 * (static_cast<int>((half + up - 1) / up) + 1);   // taps per phase: covers |t| <= half
        const int64_t centre = static_cast<int64_t>(tpp / 2) * up;       // h index of t = 0
        std::vector<float> taps(static_cast<size_t>(tpp) * up, 0.f);
        const double i0b = bessel_i0(beta);
        for (int j = 0; j < tpp; ++j) {
            for (int ph = 0; ph < up; ++ph) {
                const int64_t idx = static_cast<int64_t>(j) * up + ph;   // prototype index
                const double t = static_cast<double>(idx - centre);      // in virtual samples
                const double r = t / static_cast<double>(half);
                double v = 0.0;
                if (std::fabs(r) <= 1.0) {
                    const double arg = 2.0 * fc / fs_v * t;
                    const double sinc = arg == 0.0 ? 1.0 : std::sin(M_PI * arg) / (M_PI * arg);
                    const double win = bessel_i0(beta * std::sqrt(1.0 - r * r)) / i0b;
                    v = 2.0 * fc / fs_v * sinc * win * up;
                }
                taps[static_cast<size_t>(j) * up + ph] = static_cast<float>(v);   // layout [tap][phase]
            }
        }
        bd_engine::Resampler r;
        r.up = up; r.down = down; r.taps_per_phase = tpp; r.d_taps = nullptr;
        BD_CHECK(e, cudaMalloc(&r.d_taps, taps.size() * sizeof(float)));
        BD_CHECK(e, cudaMemcpy(r.d_taps, taps.data(), taps.size() * sizeof(float), cudaMemcpyHostToDevice));
        // the same filter as a [blocks x K] x [K x NB] GEMM: H holds, for each of the NB outputs of a block, its taps laid
        // along the block's input window (resample_tc_sm100.cu)
        if (resample_tc_geometry(up, down, tpp, &r.tc)) {
            const size_t nh = static_cast<size_t>(r.tc.NB) * r.tc.K;
            std::vector<float> H(nh);
            resample_tc_build_matrix(r.tc, taps.data(), H.data());
            std::vector<__half> hi(nh), lo(nh);
            r.tc.out_scale = split_weights_f16(H.data(), nh, hi.data(), lo.data());
            BD_CHECK(e, cudaMalloc(&r.d_h_hi, nh * sizeof(__half)));
            BD_CHECK(e, cudaMalloc(&r.d_h_lo, nh * sizeof(__half)));
            BD_CHECK(e, cudaMemcpy(r.d_h_hi, hi.data(), nh * sizeof(__half), cudaMemcpyHostToDevice));
            BD_CHECK(e, cudaMemcpy(r.d_h_lo, lo.data(), nh * sizeof(__half), cudaMemcpyHostToDevice));
            r.tc_ok = encode_kmajor_f16_map(&r.tc.b_hi, r.d_h_hi, r.tc.NB, r.tc.K, r.tc.ntile) &&
                      encode_kmajor_f16_map(&r.tc.b_lo, r.d_h_lo, r.tc.NB, r.tc.K, r.tc.ntile);
        }
        BD_CHECK(e, cudaDeviceSynchronize());   // the tap upload is not stream-ordered with s_compute
        it = e->resamplers.emplace(src_rate, r).first;
    }
    *out = &it->second;
    return 0;
}

// downmix + resample of one chunk on `st`: tensor-core GEMM for the whole blocks, tap-by-tap kernel for the tail (and for
// equal rates, where only the downmix / int16 conversion is left)
static int run_resample(bd_engine* e, const bd_engine::Resampler* r, const void* d_in, int fmt, int channels,
                        long long n_frames, float* d_out, long long no, cudaStream_t st) {
    long long done = 0;
    if (r->tc_ok && e->tc_resample && r->d_taps != nullptr) {
        BD_CHECK(e, launch_resample_tc(r->tc, d_in, fmt, channels, n_frames, d_out, no, e->num_sms, st, &done));
        if (done > 0) e->launch_count++;
    }
    if (done < no) {
        BD_CHECK(e, launch_resample(d_in, fmt, channels, n_frames, r->up, r->down, r->d_taps, r->taps_per_phase, d_out, no,
                                    st, done));
        e->launch_count++;
    }
    return 0;
}

int32_t bd_resample_device(bd_engine* e, const void* d_in, int32_t fmt, int32_t channels, int64_t n_frames,
                           int32_t src_rate, float* d_out, int64_t out_capacity, int64_t* n_out) {
    if (!e) return 1;
    std::lock_guard<std::recursive_mutex> lk(e->mu);
    if ((fmt != 0 && fmt != 1) || channels < 1 || channels > 8 || src_rate < 1000 || n_frames < 0)
        return fail(e, "bad resample arguments");
    BD_CHECK(e, cudaSetDevice(e->device));
    const int64_t no = bd_resample_out_len(n_frames, src_rate);
    if (n_out) *n_out = no;
    if (no > out_capacity) return fail(e, "resample output buffer too small");
    if (no == 0) return 0;
    bd_engine::Resampler ident{1, 1, 1, nullptr};
    bd_engine::Resampler* r = &ident;
    if (src_rate != 16000 && get_resampler(e, src_rate, &r)) return 1;
    if (run_resample(e, r, d_in, fmt, channels, n_frames, d_out, no, e->s_compute)) return 1;
    BD_CHECK(e, cudaStreamSynchronize(e->s_compute));
    return 0;
}

int32_t bd_submit_pcm_host(bd_engine* e, int32_t slot, const void* pcm, int32_t fmt, int32_t channels,
                           int64_t n_frames, int32_t src_rate, int32_t hop_frames, float* act, float* emb,
                           int64_t* n_patches) {
    if (!e) return 1;
    std::lock_guard<std::recursive_mutex> lk(e->mu);
    if (slot < 0 || slot >= static_cast<int>(e->slots.size())) return fail(e, "slot out of range");
    if ((fmt != 0 && fmt != 1) || channels < 1 || channels > 8 || src_rate < 1000 || n_frames < 0)
        return fail(e, "bad PCM arguments");
    if (hop_frames < 1 || hop_frames > kPatchFrames) return fail(e, "bad hop_frames");
    BD_CHECK(e, cudaSetDevice(e->device));
    Slot& s = e->slots[slot];
    if (s.state != 0) return fail(e, "slot still in flight: call bd_wait first");
    const int64_t n = bd_resample_out_len(n_frames, src_rate);
    int64_t P = 0;
    frames_for(n, hop_frames, nullptr, nullptr, &P);
    if (n_patches) *n_patches = P;
    if (P == 0) return 0;
    if ((n_frames > 0 && !pcm) || !act) return fail(e, "null host buffer");
    if (ensure_slot(e, s, n, P)) return 1;
    const int64_t pcm_bytes = n_frames * channels * (fmt == 0 ? 4 : 2);
    if (pcm_bytes > s.pcm_cap) {
        if (s.d_pcm) cudaFree(s.d_pcm);
        s.d_pcm = nullptr;
        s.pcm_cap = 0;
        const int64_t cap = ((pcm_bytes + 65535) / 65536) * 65536;
        BD_CHECK(e, cudaMalloc(&s.d_pcm, cap));
        s.pcm_cap = cap;
    }
    bd_engine::Resampler ident{1, 1, 1, nullptr};
    bd_engine::Resampler* r = &ident;
    if (src_rate != 16000 && get_resampler(e, src_rate, &r)) return 1;
    s.n = n; s.P = P; s.hop = hop_frames; s.want_emb = emb != nullptr; s.arrived = false;
    if (route_outputs(e, s, act, emb)) return 1;
    trace_point(e, 0, slot, 0, nullptr);
    if (n_frames > 0) BD_CHECK(e, cudaMemcpyAsync(s.d_pcm, pcm, pcm_bytes, cudaMemcpyHostToDevice, e->s_in));
    BD_CHECK(e, cudaEventRecord(s.ev_in, e->s_in));
    trace_point(e, 1, slot, 0, e->s_in);
    // downmix + resample to the slot's 16 kHz buffer run at the head of the pass that takes the chunk (launch_group)
    s.has_pcm = true; s.pcm_fmt = fmt; s.pcm_channels = channels; s.pcm_rate = src_rate; s.pcm_frames = n_frames;
    s.state = 1;
    e->pending.push_back(slot);
    return maybe_flush(e);
}

int32_t bd_resample_host(bd_engine* e, const void* in, int32_t fmt, int32_t channels, int64_t n_frames,
                         int32_t src_rate, float* out, int64_t out_capacity, int64_t* n_out) {
    if (!e) return 1;
    std::lock_guard<std::recursive_mutex> lk(e->mu);
    if ((fmt != 0 && fmt != 1) || channels < 1 || channels > 8 || src_rate < 1000 || n_frames < 0)
        return fail(e, "bad resample arguments");
    BD_CHECK(e, cudaSetDevice(e->device));
    const int64_t no = bd_resample_out_len(n_frames, src_rate);
    if (n_out) *n_out = no;
    if (no > out_capacity) return fail(e, "resample output buffer too small");
    if (no == 0) return 0;
    const size_t in_bytes = static_cast<size_t>(n_frames) * channels * (fmt == 0 ? 4 : 2);
    void* d_in = nullptr;
    float* d_out = nullptr;
    BD_CHECK(e, cudaMalloc(&d_in, in_bytes));
    cudaError_t ce = cudaMalloc(&d_out, no * sizeof(float));
    if (ce != cudaSuccess) { cudaFree(d_in); BD_CHECK(e, ce); }
    int rc = 0;
    ce = cudaMemcpyAsync(d_in, in, in_bytes, cudaMemcpyHostToDevice, e->s_compute);
    if (ce == cudaSuccess) {
        rc = bd_resample_device(e, d_in, fmt, channels, n_frames, src_rate, d_out, no, nullptr);
        if (rc == 0) ce = cudaMemcpy(out, d_out, no * sizeof(float), cudaMemcpyDeviceToHost);
    }
    cudaFree(d_in);
    cudaFree(d_out);
    if (rc) return rc;
    BD_CHECK(e, ce);
    return 0;
}

// ------------------------------------------------------------------------------------------ test hooks
int32_t bd_debug_logmel(bd_engine* e, const float* samples, int64_t n, int64_t n_frames, float* logmel) {
    if (!e) return 1;
    std::lock_guard<std::recursive_mutex> lk(e->mu);
    BD_CHECK(e, cudaSetDevice(e->device));
    if (n_frames <= 0) return 0;
    float *d_x = nullptr, *d_lm = nullptr;
    BD_CHECK(e, cudaMalloc(&d_x, std::max<int64_t>(n, 1) * sizeof(float)));
    BD_CHECK(e, cudaMalloc(&d_lm, n_frames * kMel * sizeof(float)));
    BD_CHECK(e, cudaMemcpyAsync(d_x, samples, n * sizeof(float), cudaMemcpyHostToDevice, e->s_compute));
    if (e->frontend_v1) {
        BD_CHECK(e, launch_logmel(d_x, n, 0, static_cast<int>(n_frames), e->d_tab, d_lm, e->num_sms, e->s_compute));
    } else {
        LogmelSeg sg{d_x, n, 0, 0, static_cast<int>(n_frames)};
        BD_CHECK(e, launch_logmel_segs(&sg, 1, e->mel_param, e->d_tab->window, d_lm, n_frames, e->num_sms, e->s_compute));
    }
    e->launch_count++;
    BD_CHECK(e, cudaStreamSynchronize(e->s_compute));
    BD_CHECK(e, cudaMemcpy(logmel, d_lm, n_frames * kMel * sizeof(float), cudaMemcpyDeviceToHost));
    cudaFree(d_x);
    cudaFree(d_lm);
    return 0;
}

int32_t bd_debug_pw_gemm(bd_engine* e, const float* A, const float* W, const float* bias, int32_t M, int32_t N,
                         int32_t K, int32_t precision, int32_t block_n, float* C) {
    if (!e) return 1;
    std::lock_guard<std::recursive_mutex> lk(e->mu);
    BD_CHECK(e, cudaSetDevice(e->device));
    const size_t na = static_cast<size_t>(M) * K, nw = static_cast<size_t>(N) * K, nc = static_cast<size_t>(M) * N;
    float *dA = nullptr, *dW = nullptr, *dB = nullptr, *dC = nullptr;
    __half *a_hi = nullptr, *a_lo = nullptr, *w_hi = nullptr, *w_lo = nullptr;
    BD_CHECK(e, cudaMalloc(&dB, N * sizeof(float)));
    BD_CHECK(e, cudaMalloc(&dC, nc * sizeof(float)));
    BD_CHECK(e, cudaMemcpyAsync(dB, bias, N * sizeof(float), cudaMemcpyHostToDevice, e->s_compute));
    int rc = 0;
    if (precision == BD_PRECISION_FP32_SIMT) {
        BD_CHECK(e, cudaMalloc(&dA, na * sizeof(float)));
        BD_CHECK(e, cudaMalloc(&dW, nw * sizeof(float)));
        BD_CHECK(e, cudaMemcpyAsync(dA, A, na * sizeof(float), cudaMemcpyHostToDevice, e->s_compute));
        BD_CHECK(e, cudaMemcpyAsync(dW, W, nw * sizeof(float), cudaMemcpyHostToDevice, e->s_compute));
        BD_CHECK(e, launch_pw_simt(dA, dW, dB, dC, M, N, K, e->s_compute));
    } else {
        BD_CHECK(e, pw_gemm_init_device());
        std::vector<__half> hi, lo;
        split_f16(A, na, hi, lo);
        if (precision == BD_PRECISION_FP16F8) {
            // second plane of A for the fp16 + fp8 plan: e5m2 bytes [M, 2K], per 64-channel k-block [lo * 2^11 | hi]
            if (K % 64 != 0) return fail(e, "fp16f8 needs K % 64 == 0");
            unsigned char* c8 = reinterpret_cast<unsigned char*>(lo.data());
            std::vector<unsigned char> tmp(na * 2);
            for (int64_t r = 0; r < M; ++r)
                for (int64_t k = 0; k < K; ++k) {
                    const float h = __half2float(hi[r * K + k]);
                    unsigned char* row = tmp.data() + r * 2 * K + (k / 64) * 128 + (k % 64);
                    row[0] = static_cast<unsigned char>(__nv_cvt_float_to_fp8((A[r * K + k] - h) * 2048.f, __NV_SATFINITE, __NV_E5M2));
                    row[64] = static_cast<unsigned char>(__nv_cvt_float_to_fp8(h, __NV_SATFINITE, __NV_E5M2));
                }
            std::memcpy(c8, tmp.data(), na * 2);
        }
        BD_CHECK(e, cudaMalloc(&a_hi, na * sizeof(__half)));
        BD_CHECK(e, cudaMalloc(&a_lo, na * sizeof(__half)));
        BD_CHECK(e, cudaMemcpyAsync(a_hi, hi.data(), na * sizeof(__half), cudaMemcpyHostToDevice, e->s_compute));
        BD_CHECK(e, cudaStreamSynchronize(e->s_compute));
        BD_CHECK(e, cudaMemcpyAsync(a_lo, lo.data(), na * sizeof(__half), cudaMemcpyHostToDevice, e->s_compute));
        BD_CHECK(e, cudaStreamSynchronize(e->s_compute));
        hi.resize(nw); lo.resize(nw);
        const float out_scale = precision == BD_PRECISION_FP16F8
            ? split_weights_f16f8(W, N, K, hi.data(), reinterpret_cast<unsigned char*>(lo.data()))
            : split_weights_f16(W, nw, hi.data(), lo.data());
        BD_CHECK(e, cudaMalloc(&w_hi, nw * sizeof(__half)));
        BD_CHECK(e, cudaMalloc(&w_lo, nw * sizeof(__half)));
        BD_CHECK(e, cudaMemcpyAsync(w_hi, hi.data(), nw * sizeof(__half), cudaMemcpyHostToDevice, e->s_compute));
        BD_CHECK(e, cudaStreamSynchronize(e->s_compute));
        BD_CHECK(e, cudaMemcpyAsync(w_lo, lo.data(), nw * sizeof(__half), cudaMemcpyHostToDevice, e->s_compute));
        BD_CHECK(e, cudaStreamSynchronize(e->s_compute));
        PwGemmPlan plan;
        const char* perr = nullptr;
        cudaError_t pe = pw_gemm_make_plan(&plan, a_hi, a_lo, M, K, w_hi, w_lo, N,
                                           precision == BD_PRECISION_FP16X3 ? 3 : (precision == BD_PRECISION_FP16F8 ? 2 : 1),
                                           block_n, out_scale, &perr);
        if (pe != cudaSuccess) rc = fail(e, std::string("plan: ") + (perr ? perr : cudaGetErrorString(pe)));
        if (rc == 0) {
            cudaError_t le = launch_pw_gemm(plan, dB, dC, M, e->num_sms, e->s_compute);
            if (le != cudaSuccess) rc = fail(e, std::string("launch_pw_gemm: ") + cudaGetErrorString(le));
        }
    }
    e->launch_count++;
    if (rc == 0) {
        cudaError_t se = cudaStreamSynchronize(e->s_compute);
        if (se != cudaSuccess) rc = fail(e, std::string("pw gemm execution: ") + cudaGetErrorString(se));
    }
    if (rc == 0) {
        cudaError_t me = cudaMemcpy(C, dC, nc * sizeof(float), cudaMemcpyDeviceToHost);
        if (me != cudaSuccess) rc = fail(e, std::string("copy back: ") + cudaGetErrorString(me));
    }
    cudaFree(dA); cudaFree(dW); cudaFree(dB); cudaFree(dC);
    cudaFree(a_hi); cudaFree(a_lo); cudaFree(w_hi); cudaFree(w_lo);
    return rc;
}

int32_t bd_debug_stage(bd_engine* e, const float* samples, int64_t n, int32_t hop_frames, int32_t stage, float* out,
                       int64_t out_capacity, int64_t* n_out) {
    if (!e) return 1;
    std::lock_guard<std::recursive_mutex> lk(e->mu);
    if (stage < 0 || stage > 2 * (BD_N_LAYERS - 1) + 1) return fail(e, "stage out of range");
    if (n < 0 || hop_frames < 1 || hop_frames > kPatchFrames) return fail(e, "bad n / hop_frames");
    BD_CHECK(e, cudaSetDevice(e->device));
    int64_t P = 0;
    frames_for(n, hop_frames, nullptr, nullptr, &P);
    P = std::min<int64_t>(P, e->S1);
    if (n_out) *n_out = 0;
    if (P == 0) return 0;
    float* d_x = nullptr;
    float* d_act = nullptr;
    BD_CHECK(e, cudaMalloc(&d_x, std::max<int64_t>(n, 1) * sizeof(float)));
    BD_CHECK(e, cudaMalloc(&d_act, P * e->n_classes * sizeof(float)));
    BD_CHECK(e, cudaMemcpyAsync(d_x, samples, n * sizeof(float), cudaMemcpyHostToDevice, e->s_compute));
    e->dbg_ptr = nullptr;
    FrontJob djob;
    djob.x = d_x;
    djob.n = n;
    int rc = enqueue_chunk(e, djob, hop_frames, d_act, nullptr, P, e->s_compute, stage);
    if (rc == 0) {
        cudaError_t se = cudaStreamSynchronize(e->s_compute);
        if (se != cudaSuccess) rc = fail(e, std::string("stage execution: ") + cudaGetErrorString(se));
    }
    if (rc == 0) {
        int64_t count = 0;
        const void* src = e->dbg_ptr;
        const int planes = e->dbg_planes;
        const size_t plane_off = e->dbg_plane_off;
        if (stage == 0) {
            count = (static_cast<int64_t>(P - 1) * hop_frames + kPatchFrames) * kMel;
        } else if (stage == 1) {
            count = P * 48 * 32 * 32;
        } else {
            const LayerDev& l = e->layers[stage / 2];   // stage 2L = depthwise out, 2L+1 = pointwise out of layer index L
            count = P * l.h_out * l.w_out * ((stage % 2) == 0 ? l.d.cin : l.d.cout);
        }
        if (src == nullptr) rc = fail(e, "debug stage was not reached");
        if (rc == 0 && count > out_capacity) {
            rc = fail(e, "debug stage output buffer too small");
        } else if (rc == 0 && planes == 0) {
            cudaError_t me = cudaMemcpy(out, src, count * sizeof(float), cudaMemcpyDeviceToHost);
            if (me != cudaSuccess) rc = fail(e, cudaGetErrorString(me));
        } else if (rc == 0) {
            std::vector<__half> hi(count), lo(count);
            cudaMemcpy(hi.data(), src, count * sizeof(__half), cudaMemcpyDeviceToHost);
            if (planes != 1)
                cudaMemcpy(lo.data(), static_cast<const unsigned char*>(src) + plane_off, count * sizeof(__half),
                           cudaMemcpyDeviceToHost);
            const LayerDev& ld = e->layers[stage / 2];
            const unsigned char* c8 = reinterpret_cast<const unsigned char*>(lo.data());
            for (int64_t i = 0; i < count; ++i) {    // undo the 2^-4 operand scale (see bd_engine_create)
                float v = __half2float(hi[i]);
                if (planes == 3) v += __half2float(lo[i]);
                if (planes == 2) {                   // e5m2 plane: per 64-channel k-block [lo * 2^11 | hi]
                    const int64_t row = i / ld.d.cin, c = i % ld.d.cin;
                    const __half_raw hr = __nv_cvt_fp8_to_halfraw(c8[row * 2 * ld.d.cin + (c / 64) * 128 + (c % 64)], __NV_E5M2);
                    v += __half2float(__half(hr)) / 2048.f;
                }
                out[i] = v / kActScale;
            }
        }
        if (rc == 0 && n_out) *n_out = count;
    }
    cudaFree(d_x);
    cudaFree(d_act);
    return rc;
}

}  // extern "C"
