// b200_stream_common.cuh -- device helpers shared by the streaming decode kernels (b200_gemv_stream.cu: one mul_mat per
// launch; b200_plan.cu: a whole dependent sequence of decode mul_mats as one persistent launch): mbarrier / bulk-copy
// wrappers, 32-bit shared-window accessors, the exact per-block int32 dot (src/ggml-quants.c:3858-3869, :5010-5015) and the
// tagged 8-byte ("LL") element exchange.
#pragma once
#include "b200_internal.cuh"

namespace b200s {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// shared-memory accessors on 32-bit shared-window addresses (keeps all ring/row arithmetic in 32-bit registers)
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float lds_h2f(uint32_t addr) {
    unsigned short h;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(addr));
    return __half2float(__ushort_as_half(h));
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ int lds_s32(uint32_t addr) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar_addr, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(bar_addr), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar_addr) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_addr) : "memory");
}

// unsigned-byte x signed-byte 4-way dot with int32 accumulate (SASS IDP.4A.U8.S8)
__device__ __forceinline__ int dp4a_u8s8(uint32_t a, int b, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// exact int32 dot of one weight block against one activation block held as two uint4 (elements 0..15, 16..31).
//   Q4_0: low nibbles are masked in place; HIGH nibbles stay in place too (w & 0xF0F0F0F0 == 16 * nib as an unsigned
//   byte) and go through the unsigned x signed dp4a, so their partial is exactly 16 * sum and is shifted back.
//   (nib - 8) . q == nib . q - 8 * sum(q): s8 = 8 * sum(q) comes precomputed with the activations.
template <int TYPE>
__device__ __forceinline__ int block_dot(const uint4 &w0, const uint4 &w1, const uint4 &alo, const uint4 &ahi, int s8) {
    if (TYPE == B200_TYPE_Q4_0) {
        int lo = -s8;
        lo = __dp4a((int)(w0.x & 0x0F0F0F0Fu), (int)alo.x, lo);
        lo = __dp4a((int)(w0.y & 0x0F0F0F0Fu), (int)alo.y, lo);
        lo = __dp4a((int)(w0.z & 0x0F0F0F0Fu), (int)alo.z, lo);
        lo = __dp4a((int)(w0.w & 0x0F0F0F0Fu), (int)alo.w, lo);
        int hi = 0;
        hi = dp4a_u8s8(w0.x & 0xF0F0F0F0u, (int)ahi.x, hi);
        hi = dp4a_u8s8(w0.y & 0xF0F0F0F0u, (int)ahi.y, hi);
        hi = dp4a_u8s8(w0.z & 0xF0F0F0F0u, (int)ahi.z, hi);
        hi = dp4a_u8s8(w0.w & 0xF0F0F0F0u, (int)ahi.w, hi);
        return lo + (hi >> 4);                             // hi is an exact multiple of 16
    } else {
        int sumi = __dp4a((int)w0.x, (int)alo.x, 0);
        sumi = __dp4a((int)w0.y, (int)alo.y, sumi);
        sumi = __dp4a((int)w0.z, (int)alo.z, sumi);
        sumi = __dp4a((int)w0.w, (int)alo.w, sumi);
        sumi = __dp4a((int)w1.x, (int)ahi.x, sumi);
        sumi = __dp4a((int)w1.y, (int)ahi.y, sumi);
        sumi = __dp4a((int)w1.z, (int)ahi.z, sumi);
        sumi = __dp4a((int)w1.w, (int)ahi.w, sumi);
        return sumi;
    }
}
// ---- fused all-gather ("LL" elements: {fp32 value, u32 tag} in one 8-byte word) --------------------------------
// 16 consecutive LL elements (128 bytes) -> 16 floats, re-reading until every tag is the expected one.  Volatile loads go
// to L2, which is where peer stores land; 8-byte stores are delivered atomically, so a matching tag implies the value.
__device__ __forceinline__ void ll_load16(const void *src, uint32_t tag, float4 (&out)[4]) {
    const uint4 *p = reinterpret_cast<const uint4 *>(src);
    unsigned backoff = 64;
    for (;;) {
        uint4 w[8];
#pragma unroll
        for (int j = 0; j < 8; j++)
            asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w[j].x), "=r"(w[j].y), "=r"(w[j].z), "=r"(w[j].w) : "l"(p + j));
        bool ok = true;
#pragma unroll
        for (int j = 0; j < 8; j++) ok = ok && w[j].y == tag && w[j].w == tag;
        if (ok) {
#pragma unroll
            for (int j = 0; j < 4; j++)
                out[j] = make_float4(__uint_as_float(w[2 * j].x), __uint_as_float(w[2 * j].z), __uint_as_float(w[2 * j + 1].x), __uint_as_float(w[2 * j + 1].z));
            return;
        }
        __nanosleep(backoff);
        if (backoff < 512) backoff += 64;
    }
}
// Bounded waiting: every wait on a peer's tagged stores looks at the context's abort word and at %globaltimer every
// 256th failed poll; on expiry it raises the abort word (device copy for the other waits, host copy for b200_synchronize, which then
// returns B200_ERR_CUDA) and gives up with whatever it has.  A dead or diverged peer rank no longer hangs the GPU.
struct LLWait {
    uint32_t *abort_dev, *abort_host;
    unsigned long long timeout_ns;
};
__device__ __forceinline__ bool ll_wait_expired(const LLWait &w, unsigned &polls, unsigned long long &t0, uint32_t what) {
    if ((++polls & 255u) != 0u || w.abort_dev == nullptr) return false;
    if (*reinterpret_cast<volatile uint32_t *>(w.abort_dev) != 0u) return true;
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    if (t0 == 0ull) {
        t0 = now;
        return false;
    }
    if (now - t0 > w.timeout_ns) {
        if (atomicCAS(w.abort_dev, 0u, what | 0x80000000u) == 0u && w.abort_host) *reinterpret_cast<volatile uint32_t *>(w.abort_host) = what | 0x80000000u;
        return true;
    }
    return false;
}
// Warp-cooperative form: the warp's 32 lane-tasks are one contiguous 4 KB run of LL elements.  It is read with fully
// coalesced 128-bit volatile loads (lane stride 16 B; volatile loads bypass L1, so the per-lane form above costs 32 sector
// requests per instruction), the tags are verified warp-wide, the values are parked in a 2 KB per-warp staging area and
// each lane picks up its 16 consecutive floats.  nvalid = number of live lane-tasks of this warp (0..32).
__device__ __forceinline__ void ll_load16_warp(const char *wbase, int nvalid, uint32_t tag, float *wstage, int lane, float4 (&out)[4], const LLWait &lw) {
    uint4 w[8];
    const int nv8 = nvalid * 8;
    unsigned polls = 0;
    unsigned long long t0 = 0;
    for (;;) {
        bool ok = true;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            // lanes past the end re-read the run's first chunk (always valid) so that no load is conditional
            const int idx = j * 32 + lane < nv8 ? j * 32 + lane : 0;
            asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w[j].x), "=r"(w[j].y), "=r"(w[j].z), "=r"(w[j].w) : "l"(wbase + (size_t)idx * 16));
            ok = ok && w[j].y == tag && w[j].w == tag;
        }
        if (__all_sync(0xffffffffu, ok)) break;
        __nanosleep(64);
        if (__shfl_sync(0xffffffffu, (int)ll_wait_expired(lw, polls, t0, 5u | (tag << 8)), 0) != 0) break;      // (lane 0 decides for the warp)
    }
#pragma unroll
    for (int j = 0; j < 8; j++)
        if (j * 32 + lane < nv8) *reinterpret_cast<float2 *>(wstage + 2 * (j * 32 + lane)) = make_float2(__uint_as_float(w[j].x), __uint_as_float(w[j].z));
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; j++) out[j] = *reinterpret_cast<const float4 *>(wstage + lane * 16 + j * 4);
    __syncwarp();
}
// cheap readiness probe: spin on ONE element (a warp-uniform address -> one 32-byte sector per warp per poll) until it
// carries the tag; the full loads that follow still verify every element, so this only has to be a good predictor
__device__ __forceinline__ void ll_probe(const void *elem, uint32_t tag, const LLWait &lw) {
    uint32_t v, t;
    unsigned polls = 0;
    unsigned long long t0 = 0;
    for (;;) {
        asm volatile("ld.volatile.global.v2.u32 {%0,%1}, [%2];" : "=r"(v), "=r"(t) : "l"(elem));
        if (t == tag) return;
        __nanosleep(32);
        if (ll_wait_expired(lw, polls, t0, 5u | (tag << 8))) return;
    }
}
__device__ __forceinline__ void ll_store(void *vec, int64_t idx, float v, uint32_t tag) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(reinterpret_cast<uint2 *>(vec) + idx), "r"(__float_as_uint(v)), "r"(tag) : "memory");
}

}  // namespace b200s
