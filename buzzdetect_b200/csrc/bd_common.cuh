// Shared device/host helpers for the buzzdetect_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdio>

#ifndef BD_SLEEP_HINT_NS
#define BD_SLEEP_HINT_NS 20000u
#endif
#ifndef BD_SPIN_LIMIT
// Bounded mbarrier spins: a mis-programmed pipeline traps instead of hanging the GPU box.
#define BD_SPIN_LIMIT (1u << 26)
#endif

namespace bd {

// ------------------------------------------------------------------------------------------ mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// try_wait with a suspend-time hint (ns): the thread sleeps in hardware until the phase completes or the hint expires,
// instead of burning issue slots in a polling loop (matters when a dozen warps of a CTA wait at once).
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_sleepy(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait_hint(bar, parity, BD_SLEEP_HINT_NS)) {
        if (++spins > (1u << 22)) {
            printf("bd: mbarrier timeout block %d thread %d\n", (int)blockIdx.x, (int)threadIdx.x);
            __trap();
        }
    }
}

// ------------------------------------------------------------------------------------------ explicit shared-memory access
// (32-bit shared-window addresses: keeps ptxas on LDS/STS when the pointer provenance is not provable)
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
// predicated form: zeros when `pred` is false (no branch: the load itself is predicated)
__device__ __forceinline__ float4 lds128_pred(uint32_t addr, bool pred) {
    float4 v;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.f32 %0, 0f00000000;\n\tmov.f32 %1, 0f00000000;\n\tmov.f32 %2, 0f00000000;\n\tmov.f32 %3, 0f00000000;\n\t"
        "@p ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n\t}"
        : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
        : "r"(addr), "r"(static_cast<uint32_t>(pred)));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > BD_SPIN_LIMIT) {
            printf("bd: mbarrier timeout block %d thread %d\n", (int)blockIdx.x, (int)threadIdx.x);
            __trap();
        }
    }
}

// ------------------------------------------------------------------------------------------ TMA
__device__ __forceinline__ void tma_prefetch_desc(const void* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// 2-D tiled load: coordinates are (inner = element index along the contiguous dim, outer = row).
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* map, uint64_t* bar, int32_t c_inner,
                                            int32_t c_outer) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer)
        : "memory");
}

// 256-bit global store (sm_100: STG.256): one full 32-byte sector per lane
__device__ __forceinline__ void stg256(float* p, float a, float b, float c, float d, float e, float f, float g, float h) {
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d),
                 "f"(e), "f"(f), "f"(g), "f"(h)
                 : "memory");
}

// 2-D tiled store shared -> global (bulk async-group completion)
__device__ __forceinline__ void tma_store_2d(const void* map, uint32_t smem_src, int32_t c_inner, int32_t c_outer) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(map)),
                 "r"(smem_src), "r"(c_inner), "r"(c_outer)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {      // <= N groups may still be reading their smem source
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// Epilogue staging for a TMA store of a [32 rows x 32 float] block with CU_TENSOR_MAP_SWIZZLE_128B: thread = row,
// 16-byte chunk j of the row lives at chunk (j ^ (row & 7)) -- the 8 lanes of an STS.128 phase hit 8 different
// chunks, so the transposing write is bank-conflict free.  `stg` must be 1024-byte aligned.
__device__ __forceinline__ uint32_t epi_swz_addr(uint32_t stg, int row, int j) {
    return stg + static_cast<uint32_t>(row * 128 + ((j ^ (row & 7)) << 4));
}

// ------------------------------------------------------------------------------------------ tcgen05
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst) {   // whole warp, .sync.aligned
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "n"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {     // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f16 (fp16/bf16 inputs, fp32 accumulate).
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f8f6f4 (8-bit float inputs, fp32 accumulate), K = 32 per instruction.
__device__ __forceinline__ void umma_f8_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// 32 lanes x 32 columns of 32-bit: thread t of the warp gets row (lane base + t), 32 consecutive columns.
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte-swizzled shared-memory operand descriptor (cute::UMMA::SmemDescriptor, sm_100 version 1):
//   [0,14) start>>4 | [16,30) LBO>>4 (ignored for swizzled K-major) | [32,46) SBO>>4 = 1024 B between 8-row groups
//   [46,48) version = 1 | [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
// cute::UMMA::InstrDescriptor for kind::f16: fp16 A/B (format 0), fp32 accumulate, both K-major.
__host__ __device__ constexpr uint32_t umma_idesc_f16(uint32_t M, uint32_t N) {
    return (1u << 4)            // c_format = F32
           | (0u << 7)          // a_format = F16
           | (0u << 10)         // b_format = F16
           | (0u << 15)         // a_major = K
           | (0u << 16)         // b_major = K
           | ((N >> 3) << 17)   // n_dim
           | ((M >> 4) << 24);  // m_dim
}

// cute::UMMA::InstrDescriptor for kind::f8f6f4 with both operands E5M2 (format 1), fp32 accumulate, both K-major.
__host__ __device__ constexpr uint32_t umma_idesc_e5m2(uint32_t M, uint32_t N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------ fp16 + fp8 operand plan
// "fp16f8": C = A_hi W_hi (fp16) + [A_lo 2^11 | A_hi] . [W_hi 2^-11 | W_lo] (e5m2), i.e. the two correction products of
// the hi/lo split run as ONE fp8 contraction of twice the length -- 2 MMA-equivalents per k-step instead of 3, half the
// operand bytes for the corrections.  A correction term is ~2^-11 of the main product, so e5m2's 2-bit mantissa leaves
// a relative error of ~2^-14 (simulated end to end: tools/precision_ladder.py, 5e-5 max abs on the logits).
// e5m2 has fp16's exponent range, so anything representable in the fp16 planes converts without overflow.
constexpr float kF8LoScale = 2048.f;            // A_lo is stored times 2^11, W_hi times 2^-11
__device__ __forceinline__ uint32_t pack_e5m2x4(float a, float b, float c, float d) {
    uint16_t lo, hi;
    asm("cvt.rn.satfinite.e5m2x2.f32 %0, %1, %2;" : "=h"(lo) : "f"(b), "f"(a));   // first source -> upper byte
    asm("cvt.rn.satfinite.e5m2x2.f32 %0, %1, %2;" : "=h"(hi) : "f"(d), "f"(c));
    return static_cast<uint32_t>(lo) | (static_cast<uint32_t>(hi) << 16);
}
__device__ __forceinline__ void sts32u(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// ------------------------------------------------------------------------------------------ CTA pairs (cta_group::2)
// Two CTAs of a cluster on one TPC drive ONE tcgen05.mma of M = 256: each supplies its own 128 rows of A and half of
// the B tile from its own shared memory (same offsets in both), so per SM the operand fetch of B halves.  Barriers
// that gate the MMA live in the even ("leader") CTA; an address of the CTA's own shared window with bit 24 cleared
// names the same offset in the leader (cute::Sm100MmaPeerBitMask).
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {          // every thread of both CTAs
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive (release at cluster scope) on the barrier at this offset in the LEADER CTA of the pair
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait_cluster(bar, parity)) {
        if (++spins > BD_SPIN_LIMIT) {
            printf("bd: cluster mbarrier timeout block %d thread %d\n", (int)blockIdx.x, (int)threadIdx.x);
            __trap();
        }
    }
}
// 2-D tiled load whose completion is signalled on the LEADER CTA's barrier (destination: this CTA's shared memory)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const void* map, uint64_t* bar, int32_t c_inner,
                                                 int32_t c_outer) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c_inner), "r"(c_outer)
        : "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst) {   // one whole warp in EACH CTA of the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "n"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
// D[tmem of both CTAs] (+)= A * B with M = 256 over the pair; issued by one thread of the leader CTA
__device__ __forceinline__ void umma_f16_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .b16 m;\n\tmov.b16 m, 3;\n\t"
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}"
        ::"r"(smem_u32(bar))
        : "memory");
}

// ------------------------------------------------------------------------------------------ fp16 operand range guard
// Post-ReLU activations become fp16 hi (+ lo) planes for the tensor cores.  They arrive here pre-scaled by 2^-4 (the
// engine folds that factor into the depthwise taps, engine.cu:bd_engine_create), so anything up to ~1e6 keeps its ~22
// significant bits; beyond that hi saturates at fp16's largest finite value (and lo carries what it can) instead of
// turning into inf.  A real checkpoint's activations are O(1)..O(100); the guard costs one FMNMX per conversion.
__device__ __forceinline__ __half half_sat(float x) { return __float2half_rn(fminf(x, 65504.f)); }
// Two values at once: ONE F2FP.SATFINITE.F16.F32.PACK_AB instead of two FMNMX + a packed convert (the hi/lo split was
// costing as many ALU-pipe instructions as the stencil's FMAs cost FMA-pipe instructions).  Same result as half_sat()
// on each element for every finite input: saturate to +-65504, round to nearest even.  .x = first, .y = second.
__device__ __forceinline__ __half2 half2_sat(float first, float second) {
    uint32_t p;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(second), "f"(first));
    return *reinterpret_cast<__half2*>(&p);
}

// ------------------------------------------------------------------------------------------ misc
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xFFFFFFFF;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

}  // namespace bd
