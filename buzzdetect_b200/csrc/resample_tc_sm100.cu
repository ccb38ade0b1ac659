// Downmix + polyphase resampling to 16 kHz as a tcgen05 GEMM (sm_100a).
//
// Reference: np.mean over channels + librosa.resample(y, sr, 16000) in WorkerStreamer.queue_chunk,
// src/stream/worker.py:116-128 (soxr "HQ"; pinned to its published spec, see oracle/resample_oracle.py).
//
// The HQ filter is long: 522 taps per output sample at 44.1 kHz (83 k prototype taps over 160 phases).  Evaluated tap
// by tap on CUDA cores that is ~3000 instructions per output sample and 37-80 ms per audio-hour -- ten times the whole
// CNN.  It is, however, a dense contraction: with up/down = U/D, a block of NB = U*c consecutive outputs starting at
// output k*NB reads the input window  w_k[i] = x[k*S - T/2 + 1 + i],  S = D*c,  i in [0, S + T - 1),  and
//     out[k*NB + r] = sum_i w_k[i] * H[r][i],     H[r][i] = taps[q_r + T - 1 - i][(r*D) mod U],  q_r = floor(r*D/U)
// with H the SAME matrix for every block.  So   OUT[k][r] = W[k][:] . H[r][:]   is a [blocks x K] x [K x NB] GEMM whose
// C matrix, row-major, IS the output signal, and whose A matrix is a strided (overlapping) view of the input.
//
// Kernel (512 threads, one CTA per SM, persistent over 128-block tiles):
//   warps 8-15  A producers: downmix / int16->float on the fly, 8 consecutive window samples per item -> hi/lo fp16 ->
//               one STS.128 per plane into the SWIZZLE_128B K-major A tile (2 stages of 128 rows x 64 samples)
//   warp 0      TMA loads of H tiles (hi/lo planes, [ntile rows x 64] boxes) into a 3-slot ring
//   warp 1      warp-uniform tcgen05.mma issue loop: 4 k-steps x 3 products (fp16 hi/lo split = float32-equivalent)
//   warp 2      TMEM allocator (2 accumulator stages);   warps 4-7  epilogue: TMEM -> scale -> swizzled block -> TMA store
// The < NB output samples after the last whole block are left to resample_kernel (resample.cu).
#include "bd_common.cuh"
#include "bd_kernels.cuh"

namespace bd {

namespace {

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kThreads = 512;
constexpr int kATile = kBM * kBK * 2;                 // one fp16 plane, 16 KB
constexpr int kAStageBytes = 2 * kATile;              // hi + lo
constexpr int kAStages = 2;
constexpr int kBSlots = 3;
constexpr int kNtMax = 160;                           // columns (outputs of a block) per pass
constexpr int kBSlotPlane = kNtMax * kBK * 2;         // 20 KB
constexpr int kBSlotBytes = 2 * kBSlotPlane;
constexpr int kEpiBufBytes = 32 * 128;
constexpr int kEpiBytes = 4 * kEpiBufBytes;
constexpr int kSmemBytes = 1024 + kAStages * kAStageBytes + kEpiBytes + kBSlots * kBSlotBytes + 256;

struct RsParams {
    const void* in;
    int channels;
    long long n_in;
    int S;                   // input samples between consecutive blocks
    int lead;                // T/2 - 1: window start relative to k*S
    int num_kb;              // K / 64
    int ntile, n_tiles;      // columns per pass, passes per row tile
    long long rows;          // whole blocks
    int m_tiles;
    float out_scale;
};

// one sample, any channel count, with the chunk's zero state outside [0, n_in)
template <int FMT>
__device__ __forceinline__ float load_mono_checked(const void* in, int channels, long long n_in, long long idx) {
    if (idx < 0 || idx >= n_in) return 0.f;
    float s = 0.f;
    if (FMT == 0) {
        const float* p = static_cast<const float*>(in) + idx * channels;
        if (channels == 1) return __ldg(p);
        for (int c = 0; c < channels; ++c) s += __ldg(p + c);
    } else {
        const short* p = static_cast<const short*>(in) + idx * channels;
        const float k = 1.0f / 32768.0f;
        if (channels == 1) return static_cast<float>(__ldg(p)) * k;
        for (int c = 0; c < channels; ++c) s += static_cast<float>(__ldg(p + c)) * k;
    }
    return s / static_cast<float>(channels);
}

// eight consecutive mono samples starting at s0.  CH = 1 / 2: straight-line code for the interior (the common case: no
// branches between the loads, so a thread keeps all of them in flight); CH = 0: any channel count.
template <int FMT, int CH>
__device__ __forceinline__ void load8(const void* in, int channels, long long n_in, long long s0, float (&x)[8]) {
    if (CH != 0 && s0 >= 0 && s0 + 8 <= n_in) {
        const float k = 1.0f / 32768.0f;
        if (FMT == 1 && CH == 2) {
            const int* p = static_cast<const int*>(in) + s0;                 // one int = (left, right) int16
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int v = __ldg(p + j);
                const float l = static_cast<float>(static_cast<short>(v & 0xFFFF)) * k;
                const float r = static_cast<float>(static_cast<short>(v >> 16)) * k;
                x[j] = (l + r) / 2.0f;
            }
        } else if (FMT == 1) {
            const short* p = static_cast<const short*>(in) + s0;
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = static_cast<float>(__ldg(p + j)) * k;
        } else if (CH == 2) {
            const float2* p = static_cast<const float2*>(in) + s0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float2 v = __ldg(p + j);
                x[j] = (v.x + v.y) / 2.0f;
            }
        } else {
            const float* p = static_cast<const float*>(in) + s0;
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = __ldg(p + j);
        }
        return;
    }
#pragma unroll 1
    for (int j = 0; j < 8; ++j) x[j] = load_mono_checked<FMT>(in, channels, n_in, s0 + j);
}

template <int FMT, int CH>
__global__ void __launch_bounds__(kThreads, 1)
resample_tc_kernel(const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                   const __grid_constant__ CUtensorMap map_c, const RsParams prm) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* a_base = smem;
    unsigned char* epi_base = a_base + kAStages * kAStageBytes;
    unsigned char* b_base = epi_base + kEpiBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_base + kBSlots * kBSlotBytes);
    uint64_t* a_full = bars;            // [2]
    uint64_t* a_empty = bars + 2;       // [2]
    uint64_t* b_full = bars + 4;        // [3]
    uint64_t* b_empty = bars + 7;       // [3]
    uint64_t* tmem_full = bars + 10;    // [2]
    uint64_t* tmem_empty = bars + 12;   // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_pass = prm.m_tiles * prm.n_tiles;
    const int num_kb = prm.num_kb;
    const int ntile = prm.ntile;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_b_hi);
        tma_prefetch_desc(&map_b_lo);
        tma_prefetch_desc(&map_c);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kAStages; ++i) {
            mbar_init(&a_full[i], 4);                   // the four warps of the producer group that owns the stage
            mbar_init(&a_empty[i], 1);
        }
        for (int i = 0; i < kBSlots; ++i) {
            mbar_init(&b_full[i], 1);
            mbar_init(&b_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 128);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================================================= H tiles (TMA)
        if (lane == 0) {
            int bs = 0;
            uint32_t bphase = 0;
            const uint32_t bytes = static_cast<uint32_t>(2 * ntile * kBK * 2);
            for (int ps = blockIdx.x; ps < num_pass; ps += gridDim.x) {
                const int nt = ps % prm.n_tiles;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&b_empty[bs], bphase ^ 1);
                    unsigned char* dst = b_base + bs * kBSlotBytes;
                    mbar_arrive_expect_tx(&b_full[bs], bytes);
                    tma_load_2d(dst, &map_b_hi, &b_full[bs], kb * kBK, nt * ntile);
                    tma_load_2d(dst + kBSlotPlane, &map_b_lo, &b_full[bs], kb * kBK, nt * ntile);
                    if (++bs == kBSlots) { bs = 0; bphase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================================================= MMA issuer (warp-uniform loop)
        const uint32_t idesc = umma_idesc_f16(kBM, static_cast<uint32_t>(ntile));
        int stage = 0, bs = 0, acc = 0;
        uint32_t phase = 0, bphase = 0, acc_phase = 0;
        for (int ps = blockIdx.x; ps < num_pass; ps += gridDim.x) {
            mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * 256);
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&a_full[stage], phase);
                mbar_wait(&b_full[bs], bphase);
                tc_fence_after();
                const uint32_t a_hi = smem_u32(a_base + stage * kAStageBytes);
                const uint32_t b_hi = smem_u32(b_base + bs * kBSlotBytes);
                const uint64_t da_hi = umma_desc_k128(a_hi), da_lo = umma_desc_k128(a_hi + kATile);
                const uint64_t db_hi = umma_desc_k128(b_hi), db_lo = umma_desc_k128(b_hi + kBSlotPlane);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k) {
                        const uint64_t koff = static_cast<uint64_t>(k) * 2u;
                        umma_f16_ss(d_tmem, da_hi + koff, db_hi + koff, idesc, (kb | k) != 0 ? 1u : 0u);
                        umma_f16_ss(d_tmem, da_lo + koff, db_hi + koff, idesc, 1u);
                        umma_f16_ss(d_tmem, da_hi + koff, db_lo + koff, idesc, 1u);
                    }
                    umma_commit(&b_empty[bs]);
                    umma_commit(&a_empty[stage]);
                    if (kb == num_kb - 1) umma_commit(&tmem_full[acc]);
                }
                __syncwarp();
                if (++bs == kBSlots) { bs = 0; bphase ^= 1; }
                if (++stage == kAStages) { stage = 0; phase ^= 1; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= 8) {
        // ================================================================= A producers: input windows -> A tile
        // Two groups of four warps take alternate k-blocks (= alternate A stages), and a thread issues the 32 loads of
        // four items before it converts any of them: the stage is bound by load latency, not by instruction count.
        const int g = (warp - 8) >> 2;
        const int tp = threadIdx.x - 256 - g * 128;
        const uint32_t a_u32 = smem_u32(a_base);
        uint32_t phase = 0;
        unsigned gk = 0;                                 // k-blocks issued so far by this CTA (all passes): stage = gk & 1
        for (int ps = blockIdx.x; ps < num_pass; ps += gridDim.x) {
            const long long row_base = static_cast<long long>(ps / prm.n_tiles) * kBM;
            for (int kb = 0; kb < num_kb; ++kb, ++gk) {
                const int stage = static_cast<int>(gk & 1u);
                if (stage != g) continue;
                mbar_wait_sleepy(&a_empty[stage], phase ^ 1);
                const uint32_t a_hi = a_u32 + static_cast<uint32_t>(stage * kAStageBytes), a_lo = a_hi + kATile;
#pragma unroll 1
                for (int batch = 0; batch < 2; ++batch) {
                    float x[4][8];
#pragma unroll
                    for (int it = 0; it < 4; ++it) {
                        const int item = tp + 128 * (batch * 4 + it);
                        const int kk = item >> 3, ch = item & 7;             // tile row, 8-sample chunk of the k-block
                        const long long s0 = (row_base + kk) * prm.S - prm.lead + kb * kBK + ch * 8;
                        load8<FMT, CH>(prm.in, prm.channels, prm.n_in, s0, x[it]);
                    }
#pragma unroll
                    for (int it = 0; it < 4; ++it) {
                        const int item = tp + 128 * (batch * 4 + it);
                        const int kk = item >> 3, ch = item & 7;
                        __half2 hi[4], lo[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float x0 = x[it][2 * j], x1 = x[it][2 * j + 1];
                            const __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
                            hi[j] = __halves2half2(h0, h1);
                            lo[j] = __halves2half2(__float2half_rn(x0 - __half2float(h0)), __float2half_rn(x1 - __half2float(h1)));
                        }
                        const uint32_t off = static_cast<uint32_t>((kk >> 3) * 1024 + (kk & 7) * 128 + ((ch ^ (kk & 7)) << 4));
                        const uint32_t* ph = reinterpret_cast<const uint32_t*>(hi);
                        const uint32_t* pl = reinterpret_cast<const uint32_t*>(lo);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_hi + off), "r"(ph[0]), "r"(ph[1]), "r"(ph[2]), "r"(ph[3]) : "memory");
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_lo + off), "r"(pl[0]), "r"(pl[1]), "r"(pl[2]), "r"(pl[3]) : "memory");
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_full[stage]);
                phase ^= 1;
            }
        }
    } else if (warp >= 4) {
        // ================================================================= epilogue
        const int q = warp & 3;
        const uint32_t stg = smem_u32(epi_base) + static_cast<uint32_t>(q * kEpiBufBytes);
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int ps = blockIdx.x; ps < num_pass; ps += gridDim.x) {
            const int m_tile = ps / prm.n_tiles, nt = ps - m_tile * prm.n_tiles;
            const int row0 = m_tile * kBM + q * 32;
            mbar_wait_sleepy(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * 256);
            if (row0 < prm.rows) {
#pragma unroll 1
                for (int c0 = 0; c0 < ntile; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(taddr + static_cast<uint32_t>(c0), r);
                    if (lane == 0) tma_store_wait_read<0>();
                    __syncwarp();
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 o;
                        o.x = __uint_as_float(r[4 * j + 0]) * prm.out_scale;
                        o.y = __uint_as_float(r[4 * j + 1]) * prm.out_scale;
                        o.z = __uint_as_float(r[4 * j + 2]) * prm.out_scale;
                        o.w = __uint_as_float(r[4 * j + 3]) * prm.out_scale;
                        sts128(epi_swz_addr(stg, lane, j), o);
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&map_c, stg, nt * ntile + c0, row0);        // rows >= prm.rows are clipped by the map
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
        tmem_dealloc<512>(tmem_base);
    }
}

long long gcd_ll(long long a, long long b) { while (b) { long long t = a % b; a = b; b = t; } return a; }

}  // namespace

cudaError_t resample_tc_init_device() {
    cudaError_t e;
#define BD_RS_ATTR(F, C)                                                                                          \
    if ((e = cudaFuncSetAttribute(resample_tc_kernel<F, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes)) != \
        cudaSuccess)                                                                                              \
        return e;
    BD_RS_ATTR(0, 0) BD_RS_ATTR(0, 1) BD_RS_ATTR(0, 2) BD_RS_ATTR(1, 0) BD_RS_ATTR(1, 1) BD_RS_ATTR(1, 2)
#undef BD_RS_ATTR
    return cudaSuccess;
}

bool resample_tc_geometry(int up, int down, int taps_per_phase, ResampleTcPlan* g) {
    if (up < 1 || down < 1 || taps_per_phase < 2 || (taps_per_phase & 1)) return false;
    const long long base = static_cast<long long>(up) / gcd_ll(up, 32) * 32;           // lcm(up, 32)
    if (base > 4096) return false;
    long long nb = base <= kNtMax ? base * (kNtMax / base) : base;
    int ntile = 0;
    if (nb <= kNtMax) ntile = static_cast<int>(nb);
    else for (int cand : {160, 128, 96, 64, 32}) if (nb % cand == 0) { ntile = cand; break; }
    if (ntile == 0) return false;
    const long long c = nb / up;
    const long long S = static_cast<long long>(down) * c;
    const long long L = S + taps_per_phase - 1;
    if (S > (1 << 20) || L > 8192) return false;
    g->NB = static_cast<int>(nb);
    g->ntile = ntile;
    g->n_tiles = static_cast<int>(nb / ntile);
    g->S = static_cast<int>(S);
    g->K = static_cast<int>((L + kBK - 1) / kBK * kBK);
    g->T = taps_per_phase;
    g->up = up;
    g->down = down;
    return true;
}

void resample_tc_build_matrix(const ResampleTcPlan& g, const float* taps /*[T][up]*/, float* H /*[NB][K]*/) {
    for (int r = 0; r < g.NB; ++r) {
        const long long v = static_cast<long long>(r) * g.down;
        const int q = static_cast<int>(v / g.up), ph = static_cast<int>(v % g.up);
        for (int i = 0; i < g.K; ++i) {
            const int j = q + g.T - 1 - i;
            H[static_cast<size_t>(r) * g.K + i] = (j >= 0 && j < g.T) ? taps[static_cast<size_t>(j) * g.up + ph] : 0.f;
        }
    }
}

cudaError_t launch_resample_tc(const ResampleTcPlan& g, const void* in, int in_fmt, int channels, long long n_in_frames,
                               float* out, long long n_out, int num_sms, cudaStream_t stream, long long* n_done) {
    *n_done = 0;
    const long long rows = n_out / g.NB;
    if (rows < 1) return cudaSuccess;
    if (rows >= (1LL << 31) - kBM) return cudaErrorInvalidValue;
    RsParams prm;
    prm.in = in; prm.channels = channels; prm.n_in = n_in_frames;
    prm.S = g.S; prm.lead = g.T / 2 - 1; prm.num_kb = g.K / kBK;
    prm.ntile = g.ntile; prm.n_tiles = g.n_tiles; prm.rows = rows;
    prm.m_tiles = static_cast<int>((rows + kBM - 1) / kBM);
    prm.out_scale = g.out_scale;
    CUtensorMap map_c;
    if (!encode_store_map_f32(&map_c, out, rows, g.NB, 32)) return cudaErrorUnknown;
    const long long passes = static_cast<long long>(prm.m_tiles) * prm.n_tiles;
    const int grid = static_cast<int>(passes < num_sms ? passes : num_sms);
#define BD_RS_LAUNCH(F, C) resample_tc_kernel<F, C><<<grid, kThreads, kSmemBytes, stream>>>(g.b_hi, g.b_lo, map_c, prm)
    const int chsel = channels == 1 ? 1 : (channels == 2 ? 2 : 0);
    if (in_fmt == 0) {
        if (chsel == 1) BD_RS_LAUNCH(0, 1); else if (chsel == 2) BD_RS_LAUNCH(0, 2); else BD_RS_LAUNCH(0, 0);
    } else {
        if (chsel == 1) BD_RS_LAUNCH(1, 1); else if (chsel == 2) BD_RS_LAUNCH(1, 2); else BD_RS_LAUNCH(1, 0);
    }
#undef BD_RS_LAUNCH
    *n_done = rows * g.NB;
    return cudaGetLastError();
}

}  // namespace bd
