// b200_plan.cu -- a whole dependent sequence of decode mul_mats (n == 1) as ONE persistent launch.
//
// Stands in for ggml_backend_graph_plan_create / _compute (src/ggml-backend-impl.h:94-99; the reference's CUDA backend has
// the opt-in CUDA-graph replay of a cgraph, src/ggml-cuda.cu:2461-2709) for graphs whose nodes are
// GGML_OP_MUL_MAT{Q4_0|Q8_0} x F32 with one activation column.  Arithmetic per mul_mat is exactly that of
// gemv_stream_kernel (fused quantize_row_q8_0, src/ggml-quants.c:535-618; exact per-block int32 dots scaled by d_w*d_x and
// accumulated in fp32, src/ggml-quants.c:3858-3869 / :5010-5015; same lane -> block mapping and summation order, so results
// are bit-identical to the one-launch-per-mul_mat path).
//
// What a decode token looks like to the memory system: ~170 matrices of 9-150 MB, each streamed once.  Launch boundaries
// cost ~2.6 us each of idle HBM (profiles/r01_stream_bw_microbench.txt), a third of the step.  Here there are none:
//   * grid = one persistent CTA per SM (cooperative launch: all CTAs resident or none); every CTA walks the op list; in op i
//     it owns a contiguous run of rows of W_i.  12 warps:
//   * warp 8, the producer: ONE elected thread streams this CTA's byte ranges of op 0, 1, 2, ... back to back into a
//     shared-memory ring (cp.async.bulk + mbarrier complete_tx), 8 slots x 18 KB in flight per SM.  Weights never depend on
//     activations, so the producer runs arbitrarily far ahead of the consumers;
//   * warp 11, the L2 prefetcher: ONE thread walks the same byte ranges a window of ring slots AHEAD of the producer and asks
//     L2 for them (cp.async.bulk.prefetch.L2).  The ring is only as deep as the latency x bandwidth product of the stream, so
//     a hand-off stall of H us used to idle HBM for H us; now HBM keeps filling L2 during the stall and the ring refills from
//     L2 (lower latency, higher bandwidth) afterwards.  A dedicated thread, because a bulk-class instruction costs the issuing
//     thread ~75 ns and the producer has none to spare (round 1 measured the same idea inside the producer as a loss);
//   * warps 0-7, the consumers: a ring slot (8 rows of a 4096-wide k-segment) always belongs to the same team of warps (one
//     warp for k <= 4096), which turns the whole slot into results: up to four rows' loads in flight, exact dp4a block dots
//     against the quantized src1 in shared memory, four row sums reduced together by a transposed butterfly;
//   * dependencies between ops travel as tagged 8-byte elements {fp32 value, u32 tag} ("LL" vectors, one per op, in an
//     arena): the GEMV epilogue stores them, the next op's activation loads re-read until every tag matches.  No grid
//     barrier, no flags, no fences; with world > 1 the same stores go to every rank's arena over NVLink, which makes the
//     all-gather of a row-split mul_mat part of the epilogue;
//   * warp 9, the publisher, and warp 10, the fetcher: an in-plan src1 that was produced at least two ops back ("published"
//     vectors: GPT-J's o <- v and fc_out <- fc_in) is quantized ONCE per GPU -- every CTA's publisher does 1/grid of its
//     blocks as soon as their fp32 values arrive and leaves them in the arena in the consumers' activation layout, then bumps
//     an arrival counter -- and the fetcher brings the complete vector in with ONE bulk copy into a free activation buffer
//     (three rotate) while the consumers are still busy with earlier ops.  When they reach the op they wait on an mbarrier
//     that has normally completed long ago: no L2 round trip, no quantization, no CTA barrier on the critical path;
//   * consecutive ops that read the same vector (fc_in, v, q, k) share one activation buffer.
// Every wait on global memory is bounded (%globaltimer): on expiry the kernel raises the context's abort flag, every other
// wait gives up too, the launch drains, and b200_synchronize reports B200_ERR_CUDA instead of hanging.
#include "b200_stream_common.cuh"

#include <stdlib.h>
#include <vector>

using namespace b200s;

namespace {

constexpr int kRowSplit = 1;                 // (a 16-consumer-warp variant lost on every hand-off: profiles/r01_plan_experiments.md)
constexpr int kCW = 8;                       // consumer warps
constexpr int kCT = kCW * 32;                // consumer threads
constexpr int kPlanThreads = (kCW + 4) * 32; // consumers, producer, publisher, fetcher, L2 prefetcher
constexpr int kSegBlocks = 128;              // blocks of k per warp-segment (4 per lane)
constexpr int kMaxSlots = 24;
constexpr int kMaxAct = 3;                   // activation buffers that rotate
constexpr int kPartFloats = 4096;            // k-split partials parked per CTA per op: rows_per_cta * G (aliases the LL staging)
constexpr int kMaxOps = 1023;
constexpr int kDescCap = 64;                 // op descriptors staged in shared memory per window
constexpr int kPfGroup = 2;                  // ring slots per L2 prefetch instruction pair

// CDesc.flags: bits 0..7 below, 8..15 rows per ring slot, 16..19 log2 of the k-split, 20..21 activation buffer,
// 22 wait for the buffer to be released first, 23 parity of that wait, 24 parity of the published vector's arrival
enum : int { OPF_SAME_INPUT = 1, OPF_WRITE_LL = 2, OPF_EXPORT = 4, OPF_SRC_PUB = 16 };

struct __align__(16) PDesc {                 // what the producer needs of an op
    const uint8_t *qs;       // qs plane, first row of this rank's slice
    const __half *d;         // d plane, same
    int k;
    int rows_q, rows_rem;    // CTA c owns local rows [c*rows_q + min(c, rows_rem), + rows_q + (c < rows_rem))
    int rs;                  // rows per ring slot (computed on the host: the producer thread must not divide)
};
static_assert(sizeof(PDesc) == 32, "PDesc layout");
struct __align__(16) CDesc {                 // what the consumers need (staged in shared memory, kDescCap at a time)
    float *dst_plain;        // local plain fp32 vector [m_total] or null
    const float *src_plain;  // src1 when it comes from outside the plan (src_op < 0)
    int ll_dst, ll_src;      // element offsets in the arena of this op's / its producer's LL vector
    int k, flags;
    int rows_q, rows_rem;
    int src_op, row0;        // producing op or -1; global row of local row 0
};
static_assert(sizeof(CDesc) == 48, "CDesc layout");
struct ExportDesc {
    float *dst;
    int ll, m_total, op, pad;
};
struct PubDesc {             // an in-plan src1 vector that is quantized once per GPU (for the op `op` that reads it)
    int ll_src, k, src_op, op;
    long long pub;           // arena element offset of the published planes
    int buf, free_wait, free_par, pad;   // activation buffer the fetcher fills; wait for its release first (parity)
};

struct PlanGeom {
    int slot_bytes, nslots;
    int nact;                // activation buffers (2 or 3)
    int l2_window;           // ring slots the L2 prefetcher may run ahead of the producer (0 = off)
    int evict_first;         // ring copies carry an L2 evict-first policy: streamed-once weights must not push prefetched lines out
    int ring_off, act_off, act_stride, ll_off, desc_off, bar_off, pub_off, total;
};

struct PlanArgs {
    const PDesc *pdesc;
    const CDesc *cdesc;
    const ExportDesc *exports;
    int nops, nexports;
    int world, rank;
    void *arena[B200_MAX_RANKS];   // [rank] local; LL vectors live at the same offsets on every rank
    uint32_t *state;               // {arrived CTAs, completed launches}
    unsigned long long *trace;     // optional: [nops][gridDim.x][4] globaltimer stamps
    const PubDesc *pub;            // the publisher's / fetcher's work list, in op order
    int npub;
    uint32_t *pub_count;           // per work-list entry, CTAs that have published (never reset: grid per launch)
    uint32_t *abort_flag;          // the context's abort word in DEVICE memory: what the waits poll (a poll of host memory costs a PCIe round trip)
    uint32_t *abort_host;          // the same word in pinned host memory (device-mapped): written once on expiry, read by b200_synchronize
    unsigned long long timeout_ns; // bound of every wait on global memory
};

__device__ __forceinline__ void cbar() { asm volatile("bar.sync 1, %0;" ::"n"(kCT) : "memory"); }
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void mbar_arrive_cnt(uint32_t bar_addr, uint32_t count) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr), "r"(count) : "memory");
}

// Bounded waiting on global memory.  Called after a FAILED poll: every 64th call looks at the abort word (another wait
// already gave up) and at the clock.  `what` identifies the wait in the error message (b200_synchronize).
struct Spin {
    unsigned long long t0;
    unsigned n;
};
__device__ __forceinline__ bool spin_expired(Spin &s, const PlanArgs &pa, uint32_t what) {
    if ((++s.n & 63u) != 0u) return false;
    if (*reinterpret_cast<volatile uint32_t *>(pa.abort_flag) != 0u) return true;
    const unsigned long long now = gtime();
    if (s.t0 == 0ull) {
        s.t0 = now;
        return false;
    }
    if (now - s.t0 > pa.timeout_ns) {
        if (atomicCAS(pa.abort_flag, 0u, what | 0x80000000u) == 0u) {
            *reinterpret_cast<volatile uint32_t *>(pa.abort_host) = what | 0x80000000u;
            __threadfence_system();
        }
        return true;
    }
    return false;
}

// everything a warp needs to turn rows of a ring slot into results
template <int TYPE>
struct RowCtx {
    // shared-window addresses of the lane's first block (b0 + lane; blocks i = 1..3 are 32 blocks further each) in the planes
    // of the current quantized src1: int8 elements 0..15 / 16..31 of every block, fp32 scale, 8 * sum(q).  (Reading them per
    // row instead of keeping 40 registers resident leaves room for four rows' loads in flight: measured 5.9 vs 5.4 TB/s.)
    uint32_t a_lo, a_hi, a_d, a_s;
    bool blive[4];
    uint32_t woff0, soff0;    // lane offsets inside a row / inside the scale area
    int row_qs, row_sc;
};

// NR (4, 2 or 1) rows r .. r+NR-1 of the stage at stage_a (rows past `rows` are clamped and dropped): two rows' loads are in
// flight together, the NR row sums share one transposed butterfly.  Returns with lane (32/NR)*u holding row r+u.
template <int TYPE, int NR>
__device__ __forceinline__ float chunk_rows(const RowCtx<TYPE> &c, uint32_t stage_a, int r, int rows, int lane) {
    constexpr int QSB = TYPE == B200_TYPE_Q4_0 ? 16 : 32;
    constexpr int NU = (NR >= 4 && kRowSplit == 1) ? 4 : (NR >= 2 ? 2 : 1);       // rows whose loads are in flight together
    float acc[NR];
#pragma unroll
    for (int h = 0; h < NR; h += NU) {
        uint4 w0[NU][4], w1[NU][4];
        unsigned short sc[NU][4];
#pragma unroll
        for (int u = 0; u < NU; u++) {
            const int rr = min(r + h + u, rows - 1);
            const uint32_t wbase = stage_a + (uint32_t)(rr * c.row_qs) + c.woff0;
            const uint32_t sbase = stage_a + (uint32_t)(rr * c.row_sc) + c.soff0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                w0[u][i] = lds128(wbase + (uint32_t)(i * 32 * QSB));
                if (TYPE == B200_TYPE_Q8_0) w1[u][i] = lds128(wbase + (uint32_t)(i * 32 * QSB + 16));
                else w1[u][i] = make_uint4(0, 0, 0, 0);
                asm volatile("ld.shared.u16 %0, [%1];" : "=h"(sc[u][i]) : "r"(sbase + (uint32_t)(i * 64)));
            }
        }
        // block by block: the activation block (two 16-byte planes, scale, 8 * sum) is read ONCE and used for all NU rows;
        // every row still accumulates its blocks in the order i = 0..3 (the summation order of gemv_stream_kernel)
#pragma unroll
        for (int u = 0; u < NU; u++) acc[h + u] = 0.0f;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint4 alo = lds128(c.a_lo + (uint32_t)(i * 512)), ahi = lds128(c.a_hi + (uint32_t)(i * 512));
            const float da = lds_f32(c.a_d + (uint32_t)(i * 128));
            const int s8 = TYPE == B200_TYPE_Q4_0 ? lds_s32(c.a_s + (uint32_t)(i * 128)) : 0;
#pragma unroll
            for (int u = 0; u < NU; u++) {
                const int sumi = block_dot<TYPE>(w0[u][i], w1[u][i], alo, ahi, s8);
                const float dw = __half2float(__ushort_as_half(sc[u][i]));
                if (c.blive[i]) acc[h + u] = fmaf((float)sumi, dw * da, acc[h + u]);
            }
        }
    }
    // transposed butterfly: the same pairings -- hence the same bits -- as acc += shfl_xor(acc, 16, 8, 4, 2, 1) per row
    float kk;
    if (NR == 4) {
        const bool up16 = (lane & 16) != 0, up8 = (lane & 8) != 0;
        float k0 = up16 ? acc[2] : acc[0], k1 = up16 ? acc[NR - 1] : acc[1];
        const float s0 = up16 ? acc[0] : acc[2], s1 = up16 ? acc[1] : acc[NR - 1];
        k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
        k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
        kk = up8 ? k1 : k0;
        const float ss = up8 ? k0 : k1;
        kk += __shfl_xor_sync(0xffffffffu, ss, 8);
    } else if (NR == 2) {
        const bool up16 = (lane & 16) != 0;
        kk = up16 ? acc[NR - 1] : acc[0];
        const float ss = up16 ? acc[0] : acc[NR - 1];
        kk += __shfl_xor_sync(0xffffffffu, ss, 16);
        kk += __shfl_xor_sync(0xffffffffu, kk, 8);
    } else {
        kk = acc[0];
        kk += __shfl_xor_sync(0xffffffffu, kk, 16);
        kk += __shfl_xor_sync(0xffffffffu, kk, 8);
    }
    kk += __shfl_xor_sync(0xffffffffu, kk, 4);
    kk += __shfl_xor_sync(0xffffffffu, kk, 2);
    kk += __shfl_xor_sync(0xffffffffu, kk, 1);
    return kk;
}

// 32 lane-tasks (16 LL elements = 128 bytes each) of a tagged vector -> 16 floats per lane.  Coalesced 128-bit volatile loads
// (L2 is where peer stores land), tags verified warp-wide, values transposed through a 2 KB per-warp staging area.  The
// first attempt goes straight for the data (one L2 round trip when the vector is already complete); after a miss the warp
// spins on its run's LAST element only (one sector per poll) before trying again.  Bounded: gives up (with whatever it has)
// once the launch is aborted.
__device__ __forceinline__ void ll_fetch_warp(const PlanArgs &pa, const char *wbase, int nvalid, uint32_t tag, float *wstage, int lane, float4 (&out)[4]) {
    uint4 w[8];
    const int nv8 = nvalid * 8;
    Spin sp = {0ull, 0u};
    for (;;) {
        bool ok = true;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int idx = j * 32 + lane < nv8 ? j * 32 + lane : 0;
            asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w[j].x), "=r"(w[j].y), "=r"(w[j].z), "=r"(w[j].w) : "l"(wbase + (size_t)idx * 16));
            ok = ok && w[j].y == tag && w[j].w == tag;
        }
        if (__all_sync(0xffffffffu, ok)) break;
        // probe: a warp-uniform address, so every lane sees the same word and the loop stays converged
        const char *probe = wbase + (size_t)nvalid * 128 - 8;
        bool dead = false;
        for (;;) {
            uint32_t v, t;
            asm volatile("ld.volatile.global.v2.u32 {%0,%1}, [%2];" : "=r"(v), "=r"(t) : "l"(probe));
            if (t == tag) break;
            __nanosleep(32);
            dead = __shfl_sync(0xffffffffu, (int)spin_expired(sp, pa, 1u | (tag << 8)), 0) != 0;
            if (dead) break;
        }
        if (dead) break;
    }
#pragma unroll
    for (int j = 0; j < 8; j++)
        if (j * 32 + lane < nv8) *reinterpret_cast<float2 *>(wstage + 2 * (j * 32 + lane)) = make_float2(__uint_as_float(w[j].x), __uint_as_float(w[j].z));
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; j++) out[j] = *reinterpret_cast<const float4 *>(wstage + lane * 16 + j * 4);
    __syncwarp();
}

// quantize_row_q8_0 of one lane-task (half a block: 16 consecutive floats; the two lanes of a block pair up by shuffle):
// src/ggml-quants.c:535-618 with explicitly rounded IEEE operations.  Returns the packed int8 in pk, the block's amax and
// the sum of the block's quants.
__device__ __forceinline__ void quantize_half_block(const float4 (&v)[4], uint32_t (&pk)[4], float &amax, int &sq) {
    amax = 0.0f;
#pragma unroll
    for (int j = 0; j < 4; j++)
        amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v[j].x), fabsf(v[j].y)), fmaxf(fabsf(v[j].z), fabsf(v[j].w))));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
    const float id = amax != 0.0f ? __fdiv_rn(127.f, amax) : 0.0f;
    sq = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int q0 = __float2int_rn(__fmul_rn(v[j].x, id)), q1 = __float2int_rn(__fmul_rn(v[j].y, id));
        const int q2 = __float2int_rn(__fmul_rn(v[j].z, id)), q3 = __float2int_rn(__fmul_rn(v[j].w, id));
        sq += q0 + q1 + q2 + q3;
        pk[j] = (uint32_t)(q0 & 0xff) | ((uint32_t)(q1 & 0xff) << 8) | ((uint32_t)(q2 & 0xff) << 16) | ((uint32_t)(q3 & 0xff) << 24);
    }
    sq += __shfl_xor_sync(0xffffffffu, sq, 1);
}

template <int TYPE>
__global__ void __launch_bounds__(kPlanThreads, 1) plan_kernel(const __grid_constant__ PlanArgs pa, const __grid_constant__ PlanGeom pg) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int QSB = TYPE == B200_TYPE_Q4_0 ? 16 : 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cta = blockIdx.x;

    unsigned char *ring = smem + pg.ring_off;
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + pg.bar_off);
    uint64_t *empty_bar = full_bar + kMaxSlots;
    uint64_t *actfull_bar = empty_bar + kMaxSlots;       // [kMaxAct] the fetcher's bulk copy of a published vector has landed
    uint64_t *free_bar = actfull_bar + kMaxAct;          // [kMaxAct] all consumer warps have left the buffer's current vector
    uint32_t *s_epoch = reinterpret_cast<uint32_t *>(free_bar + kMaxAct);
    uint32_t *s_prod_seq = s_epoch + 1;                  // ring slots the producer has issued so far (read by the L2 prefetcher)

    if (threadIdx.x == 0) {
        for (int s = 0; s < pg.nslots; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kCW);       // a slot's consumers (one team of G warps) arrive with 8/G each
        }
        for (int b = 0; b < kMaxAct; b++) {
            mbar_init(&actfull_bar[b], 1);
            mbar_init(&free_bar[b], kCW);
        }
        *s_epoch = pa.state[1] + 1u;    // every CTA reads it before any CTA can finish (the bump needs all of them)
        *s_prod_seq = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t epoch = *s_epoch;
    const int nslots = pg.nslots;
    const char *arena_local = reinterpret_cast<const char *>(pa.arena[0]);
#pragma unroll
    for (int r = 1; r < B200_MAX_RANKS; r++)
        if (r == pa.rank) arena_local = reinterpret_cast<const char *>(pa.arena[r]);

    if (warp == kCW) {
        // ===== producer: ONE thread streams this CTA's rows of every op, in op order, through the ring.  The next op's
        // descriptor is loaded while this op's copies are issued (the op's own copies take at least its HBM time, longer
        // than an L2 round trip, so the load never stalls a busy stream). =====
        if (lane == 0) {
            int st = 0;
            uint32_t par = 0, seq = 0;
            unsigned long long prod_blocked = 0;
            const uint32_t ring_a = smem_u32(ring);
            const uint32_t seq_a = smem_u32(s_prod_seq);
            const bool tell = pg.l2_window > 0, hint = pg.evict_first != 0;
            unsigned long long policy = 0;
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
            PDesc cur = pa.pdesc[0];
#pragma unroll 1
            for (int op = 0; op < pa.nops; op++) {
                PDesc nxt = cur;
                if (op + 1 < pa.nops) nxt = pa.pdesc[op + 1];
                const int nb = cur.k >> 5, row_qs = nb * QSB, row_sc = nb * 2, rs_c = cur.rs;
                const long long r_begin = (long long)cta * cur.rows_q + min(cta, cur.rows_rem);
                const int nrows = cur.rows_q + (cta < cur.rows_rem ? 1 : 0);
                const uint8_t *gq = cur.qs + r_begin * row_qs;
                const uint8_t *gs = reinterpret_cast<const uint8_t *>(cur.d) + r_begin * row_sc;
                const uint32_t stage_qs = (uint32_t)(rs_c * row_qs);
                for (int r = 0; r < nrows; r += rs_c) {
                    const int rows = min(rs_c, nrows - r);
                    const uint32_t fb = smem_u32(&full_bar[st]);
                    if (pa.trace) {
                        const unsigned long long t0 = gtime();
                        mbar_wait(&empty_bar[st], par ^ 1u);
                        prod_blocked += gtime() - t0;
                    } else {
                        mbar_wait(&empty_bar[st], par ^ 1u);
                    }
                    const uint32_t dst = ring_a + (uint32_t)(st * pg.slot_bytes);
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"((uint32_t)(rows * (row_qs + row_sc))) : "memory");
                    if (hint) {
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
                                     "l"(gq + (size_t)r * row_qs), "r"((uint32_t)(rows * row_qs)), "r"(fb), "l"(policy) : "memory");
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst + stage_qs),
                                     "l"(gs + (size_t)r * row_sc), "r"((uint32_t)(rows * row_sc)), "r"(fb), "l"(policy) : "memory");
                    } else {
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                                     "l"(gq + (size_t)r * row_qs), "r"((uint32_t)(rows * row_qs)), "r"(fb) : "memory");
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + stage_qs),
                                     "l"(gs + (size_t)r * row_sc), "r"((uint32_t)(rows * row_sc)), "r"(fb) : "memory");
                    }
                    if (++st == nslots) { st = 0; par ^= 1u; }
                    if (tell) asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(seq_a), "r"(++seq) : "memory");
                }
                cur = nxt;
            }
            if (pa.trace) pa.trace[(size_t)pa.nops * gridDim.x * 4 + (size_t)cta * 4 + 0] = prod_blocked;
        }
    } else if (warp == kCW + 3) {
        // ===== L2 prefetcher: ONE thread walks the producer's byte ranges at most pg.l2_window ring slots ahead of it.
        // While the consumers sit in a hand-off and the ring is full, HBM keeps delivering into L2. =====
        if (lane == 0 && pg.l2_window > 0) {
            const uint32_t seq_a = smem_u32(s_prod_seq);
            uint32_t seq = 0;
#pragma unroll 1
            for (int op = 0; op < pa.nops; op++) {
                const PDesc d = pa.pdesc[op];
                const int nb = d.k >> 5, row_qs = nb * QSB, row_sc = nb * 2;
                const long long r_begin = (long long)cta * d.rows_q + min(cta, d.rows_rem);
                const int nrows = d.rows_q + (cta < d.rows_rem ? 1 : 0);
                const uint8_t *gq = d.qs + r_begin * row_qs;
                const uint8_t *gs = reinterpret_cast<const uint8_t *>(d.d) + r_begin * row_sc;
                const int step = d.rs * kPfGroup;
                for (int r = 0; r < nrows; r += step) {
                    const int rows = min(step, nrows - r);
                    uint32_t ps;
                    for (;;) {
                        asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(ps) : "r"(seq_a));
                        if ((int)(seq - ps) < pg.l2_window) break;
                        __nanosleep(128);
                    }
                    if ((int)(seq - ps) >= 1) {          // (behind the producer: the copy itself is already on its way)
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gq + (size_t)r * row_qs), "r"((uint32_t)(rows * row_qs)) : "memory");
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gs + (size_t)r * row_sc), "r"((uint32_t)(rows * row_sc)) : "memory");
                    }
                    seq += (uint32_t)((rows + d.rs - 1) / d.rs);
                }
            }
        }
    } else if (warp == kCW + 1) {
        // ===== publisher: for every published src1, in op order, this CTA's share of the blocks
        // [cta * nb / grid, (cta + 1) * nb / grid): wait for the fp32 values (tagged vector of the producing op), quantize_row_q8_0,
        // store them in the consumers' activation layout (plain), release, count.  Runs as far ahead of the consumers as the
        // data allows; every rank does the same for its own arena. =====
        float *pstage = reinterpret_cast<float *>(smem + pg.pub_off);
        unsigned char *arena_b = reinterpret_cast<unsigned char *>(const_cast<char *>(arena_local));
#pragma unroll 1
        for (int i = 0; i < pa.npub; i++) {
            const PubDesc d = pa.pub[i];
            const int nb = d.k >> 5;
            const int b_lo = (cta * nb) / (int)gridDim.x, b_hi = ((cta + 1) * nb) / (int)gridDim.x;
            const int nt = 2 * (b_hi - b_lo);
            if (nt <= 0) {
                if (lane == 0) atomicAdd(pa.pub_count + i, 1u);      // nothing to publish, but the count is per CTA
                continue;
            }
            const uint32_t src_tag = (epoch << 10) | (uint32_t)(d.src_op & 1023);
            float4 vv[4];
            ll_fetch_warp(pa, arena_local + (size_t)d.ll_src * 8 + (size_t)(2 * b_lo) * 128, nt, src_tag, pstage, lane, vv);
            const int b = b_lo + (lane >> 1), h = lane & 1;
            uint32_t pk[4];
            float amax;
            int sq;
            quantize_half_block(vv, pk, amax, sq);
            if (lane < nt) {
                // int8 elements 0..15 / 16..31 of every block, fp32 d, 8 * sum(q)
                unsigned char *reg = arena_b + (size_t)d.pub * 8;
                *reinterpret_cast<uint4 *>(reg + (size_t)h * (d.k >> 1) + (size_t)b * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                if (h == 0) {
                    reinterpret_cast<float *>(reg + d.k)[b] = __half2float(__float2half_rn(__fdiv_rn(amax, 127.f)));
                    reinterpret_cast<int *>(reg + d.k + (size_t)nb * 4)[b] = 8 * sq;
                }
            }
            // release: this warp's stores before its arrival, for readers in the generic AND the async proxy
            __threadfence();
            asm volatile("fence.proxy.async.global;" ::: "memory");
            __syncwarp();
            if (lane == 0) atomicAdd(pa.pub_count + i, 1u);
        }
    } else if (warp == kCW + 2) {
        // ===== fetcher: ONE thread.  For every published vector, in op order: all CTAs have published (arrival count),
        // the destination activation buffer has been released by the consumers -> ONE bulk copy brings the whole quantized
        // vector in; the consumers find it behind an mbarrier when they get to the op. =====
        if (lane == 0) {
#pragma unroll 1
            for (int i = 0; i < pa.npub; i++) {
                const PubDesc d = pa.pub[i];
                const uint32_t target = epoch * gridDim.x;        // the counters are never reset: grid arrivals per launch
                const uint32_t *cnt = pa.pub_count + i;
                Spin sp = {0ull, 0u};
                for (;;) {
                    uint32_t seen;
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(cnt) : "memory");
                    if ((int32_t)(seen - target) >= 0) break;
                    __nanosleep(64);
                    if (spin_expired(sp, pa, 2u | ((uint32_t)d.op << 8))) break;
                }
                if (d.free_wait) mbar_wait(&free_bar[d.buf], (uint32_t)d.free_par);
                asm volatile("fence.proxy.async.global;" ::: "memory");
                const uint32_t ab = smem_u32(&actfull_bar[d.buf]), abytes = (uint32_t)(d.k + (d.k >> 5) * 8);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ab), "r"(abytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem + pg.act_off + d.buf * pg.act_stride)),
                             "l"(arena_local + (size_t)d.pub * 8), "r"(abytes), "r"(ab) : "memory");
            }
        }
    } else if (warp < kCW) {
        // ===== consumers =====
        const uint32_t ring_a = smem_u32(ring);
        const uint32_t full_a = smem_u32(full_bar), empty_a = smem_u32(empty_bar);
        const uint32_t actfull_a = smem_u32(actfull_bar), free_a = smem_u32(free_bar);
        CDesc *sdesc = reinterpret_cast<CDesc *>(smem + pg.desc_off);
        int cur_buf = -1;           // activation buffer of the current src1
        float *llstage = reinterpret_cast<float *>(smem + pg.ll_off) + warp * 512;
        float *part = reinterpret_cast<float *>(smem + pg.ll_off);      // same bytes, other phase (bar.sync in between)
        int st0 = 0;                // ring position of the op's first stage
        uint32_t par0 = 0;
        int prevG = 1;
        unsigned long long cons_blocked = 0, quant_time = 0;
        RowCtx<TYPE> c;
        c.a_lo = c.a_hi = c.a_d = c.a_s = smem_u32(smem + pg.act_off);
#pragma unroll
        for (int i = 0; i < 4; i++) c.blive[i] = false;

#pragma unroll 1
        for (int op = 0; op < pa.nops; op++) {
            const int di = op & (kDescCap - 1);
            if (di == 0) {
                // next window of op descriptors -> shared memory: an op boundary must not cost an L2 round trip (~0.7 us
                // under full streaming load)
                cbar();
                const int n3 = min(kDescCap, pa.nops - op) * 3;
                const uint4 *src = reinterpret_cast<const uint4 *>(pa.cdesc + op);
                for (int t = threadIdx.x; t < n3; t += kCT) reinterpret_cast<uint4 *>(sdesc)[t] = src[t];
                cbar();
            }
            const CDesc *o = sdesc + di;
            // geometry comes precomputed in the descriptor
            const int k = o->k, flags = o->flags;
            const int nb = k >> 5;
            const int rs = (flags >> 8) & 0xff, log2g = (flags >> 16) & 0xf;
            const int rows_q = o->rows_q, rows_rem = o->rows_rem;
            const int r_begin = cta * rows_q + min(cta, rows_rem);
            const int nrows = rows_q + (cta < rows_rem ? 1 : 0);
            // 8/G teams; a team = G warps splitting k
            const int G = 1 << log2g, seg = warp & (G - 1);
            const int team = warp >> log2g, team_mask = (8 >> log2g) - 1;
            const int b0 = seg * kSegBlocks;
            const uint32_t tag = (epoch << 10) | (uint32_t)op;
            unsigned long long *tr = pa.trace ? pa.trace + ((size_t)op * gridDim.x + cta) * 4 : nullptr;

            if (G != prevG) {
                // The team size changes: a warp is about to wait on ring slots whose previous fills it has not watched.  After
                // this barrier every earlier fill has been consumed, so the parity waits below cannot be a lap off.
                cbar();
                prevG = G;
            }
            if (!(flags & OPF_SAME_INPUT)) {
                const unsigned long long tq0 = tr ? gtime() : 0ull;
                // ---- a new src1.  This warp is done with the previous one: release its buffer (generic-proxy reads before
                // a later async-proxy write of the same bytes) ----
                if (cur_buf >= 0) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive_a(free_a + 8u * (uint32_t)cur_buf);
                }
                cur_buf = (flags >> 20) & 3;
                unsigned char *act = smem + pg.act_off + cur_buf * pg.act_stride;
                const uint32_t act_a = smem_u32(act);
                if (flags & OPF_SRC_PUB) {
                    // ---- published: the fetcher has brought (or is bringing) the quantized vector into this buffer ----
                    if (tr && threadIdx.x == 0) tr[0] = gtime();
                    mbar_wait_a(actfull_a + 8u * (uint32_t)cur_buf, (uint32_t)(flags >> 24) & 1u);
                } else {
                    // ---- all 8 consumer warps fetch and quantize it (quantize_row_q8_0, bit-exact, same code path as
                    // gemv_stream_kernel) once every warp has left the buffer's previous vector ----
                    if (flags & (1 << 22)) mbar_wait_a(free_a + 8u * (uint32_t)cur_buf, (uint32_t)(flags >> 23) & 1u);
                    const int src_op = o->src_op;
                    const bool ll_in = src_op >= 0;
                    const uint32_t src_tag = (epoch << 10) | (uint32_t)(src_op & 1023);
                    const char *xsrc = ll_in ? arena_local + (size_t)o->ll_src * 8 : reinterpret_cast<const char *>(o->src_plain);
                    constexpr int kQB = 4;
                    const int tpc = nb * 2;      // lane-tasks: (block, half) = 16 consecutive floats
#pragma unroll 1
                    for (int base = 0; base < tpc; base += kCT * kQB) {
                        float4 v[kQB][4];
#pragma unroll
                        for (int u = 0; u < kQB; u++) {
                            const int tb = base + u * kCT + warp * 32;      // first lane-task of this warp
                            if (tb >= tpc) continue;
                            if (ll_in) {
                                ll_fetch_warp(pa, xsrc + (size_t)tb * 128, min(32, tpc - tb), src_tag, llstage, lane, v[u]);
                            } else {
                                const int t = min(tb + lane, tpc - 1);
                                const float4 *src = reinterpret_cast<const float4 *>(xsrc) + (size_t)t * 4;
#pragma unroll
                                for (int j = 0; j < 4; j++) v[u][j] = src[j];
                            }
                        }
                        if (tr && threadIdx.x == 0 && base == 0) tr[0] = gtime();
#pragma unroll
                        for (int u = 0; u < kQB; u++) {
                            const int tb = base + u * kCT + warp * 32;
                            if (tb >= tpc) continue;                        // warp-uniform: the shuffles below stay full
                            const int tt = tb + lane;
                            const bool live = tt < tpc;
                            const int t = live ? tt : tpc - 1;
                            const int b = t >> 1, h = t & 1;
                            uint32_t pk[4];
                            float amax;
                            int sq;
                            quantize_half_block(v[u], pk, amax, sq);
                            if (live) {
                                *reinterpret_cast<uint4 *>(act + (size_t)h * (k >> 1) + (size_t)b * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                                if (h == 0) {
                                    reinterpret_cast<float *>(act + k)[b] = __half2float(__float2half_rn(__fdiv_rn(amax, 127.f)));
                                    if (TYPE == B200_TYPE_Q4_0) reinterpret_cast<int *>(act + k + (size_t)nb * 4)[b] = 8 * sq;
                                }
                            }
                        }
                    }
                    cbar();
                }
                {
                    const uint32_t b = (uint32_t)(b0 + lane);       // (dead lanes read past nb into the next plane: predicated off)
                    c.a_lo = act_a + b * 16;
                    c.a_hi = act_a + (uint32_t)(k >> 1) + b * 16;
                    c.a_d = act_a + (uint32_t)k + b * 4;
                    c.a_s = act_a + (uint32_t)k + (uint32_t)nb * 4 + b * 4;
#pragma unroll
                    for (int i = 0; i < 4; i++) c.blive[i] = b0 + lane + 32 * i < nb;
                }
                if (tr && threadIdx.x == 0) tr[1] = gtime();
                if (tr) quant_time += gtime() - tq0;
            }

            // ---- the CTA's rows: whole ring slots are dealt round-robin to teams of G warps (one warp when k <= 4096) ----
            c.row_qs = nb * QSB;
            c.row_sc = nb * 2;
            c.woff0 = (uint32_t)((b0 + lane) * QSB);
            c.soff0 = (uint32_t)(rs * c.row_qs + (b0 + lane) * 2);
            float *dst_plain = o->dst_plain;
            const int ll_dst = o->ll_dst + o->row0 + r_begin;
            const int prow0 = o->row0 + r_begin;
            const bool write_ll = (flags & OPF_WRITE_LL) != 0;
            if (tr && threadIdx.x == 0) tr[2] = gtime();
            int st = st0;
            uint32_t par = par0;
#pragma unroll 1
            for (int rbase = 0; rbase < nrows; rbase += rs) {
                // A ring slot always belongs to the same team (slot % teams): its warps see every fill of the slot in order, so
                // the parity wait can never be a lap off (fills complete out of order; a warp that skipped a fill could
                // mistake the previous lap's phase for its own).
                if ((st & team_mask) == team) {
                    if (tr && warp == 2) {
                        const unsigned long long t0 = gtime();
                        mbar_wait_a(full_a + 8u * (uint32_t)st, par);
                        cons_blocked += gtime() - t0;
                    } else {
                        mbar_wait_a(full_a + 8u * (uint32_t)st, par);
                    }
                    const int rows = min(rs, nrows - rbase);
                    const uint32_t stage_a = ring_a + (uint32_t)(st * pg.slot_bytes);
#pragma unroll 1
                    for (int r = 0; r < rows;) {
                        float v;
                        int u, step;
                        bool holder;
                        if (rows - r > 2) {
                            v = chunk_rows<TYPE, 4>(c, stage_a, r, rows, lane);
                            u = lane >> 3; holder = (lane & 7) == 0; step = 4;
                        } else if (rows - r == 2) {
                            v = chunk_rows<TYPE, 2>(c, stage_a, r, rows, lane);
                            u = lane >> 4; holder = (lane & 15) == 0; step = 2;
                        } else {
                            v = chunk_rows<TYPE, 1>(c, stage_a, r, rows, lane);
                            u = 0; holder = lane == 0; step = 1;
                        }
                        if (holder && r + u < rows) {
                            const int gr = rbase + r + u;     // row relative to r_begin
                            if (G == 1) {
                                if (dst_plain) dst_plain[prow0 + gr] = v;
                                if (write_ll) {
#pragma unroll
                                    for (int rk = 0; rk < B200_MAX_RANKS; rk++)
                                        if (rk < pa.world) ll_store(pa.arena[rk], ll_dst + gr, v, tag);
                                }
                            } else {
                                part[gr * G + seg] = v;
                            }
                        }
                        r += step;
                    }
                    // every lane's weights of this slot have been consumed by the dots: hand it back to the producer
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cnt(empty_a + 8u * (uint32_t)st, (uint32_t)(8 >> log2g));
                }
                if (++st == nslots) { st = 0; par ^= 1u; }
            }
            st0 = st;
            par0 = par;
            if (G > 1) {
                // k-split rows: combine the segments' partials in segment order (fixed summation order)
                cbar();
                for (int t = threadIdx.x; t < nrows; t += kCT) {
                    float v = 0.0f;
                    for (int sg = 0; sg < G; sg++) v += part[t * G + sg];
                    if (dst_plain) dst_plain[prow0 + t] = v;
                    if (write_ll) {
#pragma unroll
                        for (int rk = 0; rk < B200_MAX_RANKS; rk++)
                            if (rk < pa.world) ll_store(pa.arena[rk], ll_dst + t, v, tag);
                    }
                }
                cbar();
            }
            if (tr && threadIdx.x == 0) tr[3] = gtime();
        }
        if (pa.trace && lane == 0 && warp == 2) {
            pa.trace[(size_t)pa.nops * gridDim.x * 4 + (size_t)cta * 4 + 1] = cons_blocked;
            pa.trace[(size_t)pa.nops * gridDim.x * 4 + (size_t)cta * 4 + 2] = quant_time;
        }
        // ---- row-split plans: ops marked EXPORT leave their COMPLETE vector (all ranks' slices) in the local plain dst ----
        for (int e = 0; e < pa.nexports; e++) {
            const ExportDesc x = pa.exports[e];
            const uint32_t tag = (epoch << 10) | (uint32_t)x.op;
            const uint2 *ll = reinterpret_cast<const uint2 *>(arena_local) + x.ll;
            Spin sp = {0ull, 0u};
            for (int i = cta * kCT + (int)threadIdx.x; i < x.m_total; i += gridDim.x * kCT) {
                uint32_t v, t;
                for (;;) {
                    asm volatile("ld.volatile.global.v2.u32 {%0,%1}, [%2];" : "=r"(v), "=r"(t) : "l"(ll + i));
                    if (t == tag) break;
                    if (spin_expired(sp, pa, 3u | ((uint32_t)x.op << 8))) break;
                }
                x.dst[i] = __uint_as_float(v);
            }
        }
    }

    // bookkeeping: the last CTA to finish publishes the launch count (tags of the next launch)
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t arrived = atomicAdd(&pa.state[0], 1u);
        if (arrived == gridDim.x - 1u) {
            pa.state[0] = 0;
            __threadfence();
            pa.state[1] = epoch;
        }
    }
}

}  // namespace

typedef void (*plan_kernel_fn)(const PlanArgs, const PlanGeom);
static plan_kernel_fn plan_kernel_for(int type) { return type == B200_TYPE_Q4_0 ? plan_kernel<B200_TYPE_Q4_0> : plan_kernel<B200_TYPE_Q8_0>; }

struct b200_plan {
    int type, nops, grid, world, rank, device;
    PlanGeom geom;
    PDesc *pdesc_dev;
    CDesc *cdesc_dev;
    ExportDesc *exports_dev;
    PubDesc *pub_dev;           // the publisher's / fetcher's work list
    uint32_t *pubcnt_dev;       // arrival counters of the published vectors
    void *arena_own;            // allocated here when world == 1
    uint32_t *state_dev;
    unsigned long long *trace_dev;
    PlanArgs args;
    size_t arena_bytes;
};

// Defaults of the two publication knobs (b200_ctx_set_option "plan_pub_min_k" / "plan_pub_dist"): an in-plan src1 is quantized
// once per GPU when it is at least this long and its producing op lies at least this many ops back (behind a producer that
// has only just finished, the publication would be a second exchange on the critical path).
constexpr int kDefaultPubMinK = 4096, kDefaultPubDist = 2;

// arena layout: one LL vector per op, then room for the published planes of every op's src1 (k + k/32*8 bytes) -- sized
// independently of any option, so that every rank of a row-split plan computes the same layout
static size_t plan_arena_elems(const b200_mul_mat_args *args, int count, const b200_plan_split *split, std::vector<long long> *offs,
                               std::vector<long long> *pub_offs = nullptr) {
    size_t total = 0;
    for (int i = 0; i < count; i++) {
        const long long mt = split && split->m_total ? split->m_total[i] : args[i].ne01;
        if (offs) offs->push_back((long long)total);
        total += (size_t)((mt + 15) / 16 * 16);      // 128-byte aligned LL vectors
    }
    for (int i = 0; i < count; i++) {
        if (pub_offs) pub_offs->push_back((long long)total);
        const size_t bytes = (size_t)args[i].ne00 + (size_t)(args[i].ne00 / B200_QK) * 8;
        total += (bytes / 8 + 15) / 16 * 16;
    }
    return total;
}

static bool ranges_overlap(const void *a, size_t an, const void *b, size_t bn) {
    const uintptr_t a0 = (uintptr_t)a, b0 = (uintptr_t)b;
    return a0 < b0 + bn && b0 < a0 + an;
}

// Host-only part of b200_plan_create (no CUDA call, so it is testable without a device): shapes, and ggml's dataflow between
// the ops.  src_op[i] = index of the op whose dst is op i's src1, or -1 for a vector from outside the plan.
// plain_ok[i] (may be null) = whether op i's plain dst may be written by the plan: a dst that a LATER op's plain dst or an
// outside input overlaps (buffer reuse by a graph allocator, ggml_gallocr) is an intermediate whose memory has been handed to
// someone else by the time the graph ends -- inside the plan its value travels as a tagged vector, so the plain store is
// simply dropped.  What the allocator guarantees (a buffer is reused only after its last reader) is what makes that legal.
static int plan_analyze(b200_ctx *ctx, const b200_mul_mat_args *args, int count, const b200_plan_split *split, std::vector<int> *src_op,
                        std::vector<int> *m_total_out, std::vector<char> *plain_ok) {
    B200_REQUIRE(ctx, args && count >= 1, B200_ERR_INVALID);
    B200_REQUIRE(ctx, count <= kMaxOps, B200_ERR_UNSUPPORTED);
    const int world = split ? split->world : 1, rank = split ? split->rank : 0;
    B200_REQUIRE(ctx, world >= 1 && world <= B200_MAX_RANKS && rank >= 0 && rank < world, B200_ERR_INVALID);
    if (split) B200_REQUIRE(ctx, split->row0 && split->m_total, B200_ERR_INVALID);
    const int type = args[0].type;
    B200_REQUIRE(ctx, type == B200_TYPE_Q4_0 || type == B200_TYPE_Q8_0, B200_ERR_UNSUPPORTED);
    std::vector<int> &so = *src_op;
    std::vector<int> &m_total = *m_total_out;
    so.assign((size_t)count, -1);
    m_total.assign((size_t)count, 0);
    for (int i = 0; i < count; i++) {
        const b200_mul_mat_args *a = &args[i];
        // decode shapes the streaming kernels take; anything else is the caller's node-by-node path
        B200_REQUIRE(ctx, a->type == type && !(a->flags & B200_MM_FORCE_GEMM), B200_ERR_UNSUPPORTED);
        B200_REQUIRE(ctx, a->ne11 == 1 && a->ne12 == 1 && a->ne13 == 1 && a->ne02 == 1 && a->ne03 == 1, B200_ERR_UNSUPPORTED);
        B200_REQUIRE(ctx, a->ne00 > 0 && a->ne00 % 256 == 0 && a->ne00 <= 32768, B200_ERR_UNSUPPORTED);
        B200_REQUIRE(ctx, a->ne01 >= 0 && a->ne01 < (1ll << 30), B200_ERR_UNSUPPORTED);
        B200_REQUIRE(ctx, a->src0_dev && a->src1_dev, B200_ERR_INVALID);
        const int64_t nb = a->ne00 / B200_QK;
        B200_REQUIRE(ctx, a->src0_block_off >= 0 && a->src0_block_off + nb * a->ne01 <= a->src0_nblocks_total, B200_ERR_INVALID);
        const int64_t mt = split ? split->m_total[i] : a->ne01;
        const int64_t row0 = split ? split->row0[i] : 0;
        B200_REQUIRE(ctx, row0 >= 0 && row0 + a->ne01 <= mt && mt < (1ll << 30), B200_ERR_INVALID);
        m_total[i] = (int)mt;
        // dataflow: src1 is the dst of the latest earlier op with that address, else an outside vector
        for (int j = i - 1; j >= 0; j--) {
            if (!args[j].dst_dev) continue;
            if (!ranges_overlap(a->src1_dev, (size_t)a->ne00 * 4, args[j].dst_dev, (size_t)m_total[j] * 4)) continue;
            // must be exactly that vector
            B200_REQUIRE(ctx, (const void *)a->src1_dev == (const void *)args[j].dst_dev && m_total[j] == a->ne00, B200_ERR_UNSUPPORTED);
            so[i] = j;
            break;
        }
        if (so[i] < 0) B200_REQUIRE(ctx, ((uintptr_t)a->src1_dev & 15) == 0, B200_ERR_UNSUPPORTED);
    }
    // Hazards sequential execution would hide but dataflow execution does not.
    std::vector<char> store((size_t)count, 1);
    for (int i = 0; i < count; i++) {
        if (!args[i].dst_dev) { store[(size_t)i] = 0; continue; }
        const size_t di = (size_t)m_total[i] * 4;
        // (1) a LATER op's dst reuses op i's memory: op i is a dead intermediate at graph end -> no plain store (a slow CTA's
        //     store of op i could otherwise land on top of a fast CTA's store of the later op)
        for (int j = i + 1; j < count; j++)
            if (args[j].dst_dev && ranges_overlap(args[i].dst_dev, di, args[j].dst_dev, (size_t)m_total[j] * 4)) store[(size_t)i] = 0;
    }
    for (int i = 0; i < count; i++) {
        if (!store[(size_t)i]) continue;
        const size_t di = (size_t)m_total[i] * 4;
        for (int j = 0; j < count; j++) {
            if (so[j] >= 0 || !ranges_overlap(args[i].dst_dev, di, args[j].src1_dev, (size_t)args[j].ne00 * 4)) continue;
            // (2) op i stores where the outside input of op j lives.  Legal only when every CTA has read that input before any
            //     CTA can reach op i: op j lies at or before op i's own producer (whose completion op i waits for).
            if (!(j < i && so[i] >= 0 && j <= so[i])) {
                if (ctx) b200_set_error(ctx, "b200_plan: dst of op %d overlaps the outside input of op %d", i, j);
                return B200_ERR_UNSUPPORTED;
            }
        }
    }
    if (plain_ok) *plain_ok = store;
    return B200_OK;
}

// Host-only: which ops take their src1 from the publisher warps (quantized once per GPU)?  A vector produced inside the plan,
// long enough, not shared with the previous op (that one already brought it in), produced far enough back, and small enough
// per CTA: every CTA quantizes 1/grid of the blocks in one warp's run of at most 16 blocks.
static std::vector<char> plan_published(const b200_mul_mat_args *args, int count, const std::vector<int> &src_op, int grid, int mink, int dist) {
    std::vector<char> pub((size_t)count, 0);
    if (mink <= 0 || grid <= 0) return pub;
    if (dist < 1) dist = 1;
    for (int i = 0; i < count; i++) {
        const int64_t k = args[i].ne00, nb = k / B200_QK;
        const bool same_input = i > 0 && k == args[i - 1].ne00 && src_op[(size_t)i] == src_op[(size_t)i - 1] &&
                                (src_op[(size_t)i] >= 0 || args[i].src1_dev == args[i - 1].src1_dev);
        pub[(size_t)i] = src_op[(size_t)i] >= 0 && !same_input && k >= mink && (nb + grid - 1) / grid + 1 <= 16 && i - src_op[(size_t)i] >= dist;
    }
    return pub;
}

extern "C" {

int b200_plan_published(const b200_mul_mat_args *args, int count, const b200_plan_split *split, int sm_count, int min_k, int dist,
                        int32_t *published_out) {
    std::vector<int> so, mt;
    std::vector<char> ok;
    const int rc = plan_analyze(NULL, args, count, split, &so, &mt, &ok);
    if (rc != B200_OK) return rc;
    if (sm_count <= 0 || !published_out) return B200_ERR_INVALID;
    const std::vector<char> pub = plan_published(args, count, so, sm_count, min_k > 0 ? min_k : kDefaultPubMinK, dist > 0 ? dist : kDefaultPubDist);
    for (int i = 0; i < count; i++) published_out[i] = pub[(size_t)i];
    return B200_OK;
}

size_t b200_plan_arena_bytes(const b200_mul_mat_args *args, int count, const b200_plan_split *split) {
    if (!args || count <= 0) return 0;
    return plan_arena_elems(args, count, split, nullptr) * 8;
}

int b200_plan_analyze(const b200_mul_mat_args *args, int count, const b200_plan_split *split, int32_t *src_op_out) {
    std::vector<int> so, mt;
    std::vector<char> ok;
    const int rc = plan_analyze(NULL, args, count, split, &so, &mt, &ok);
    if (rc == B200_OK && src_op_out)
        for (int i = 0; i < count; i++) src_op_out[i] = so[(size_t)i];
    return rc;
}

int b200_plan_plain_stores(const b200_mul_mat_args *args, int count, const b200_plan_split *split, int32_t *plain_out) {
    std::vector<int> so, mt;
    std::vector<char> ok;
    const int rc = plan_analyze(NULL, args, count, split, &so, &mt, &ok);
    if (rc == B200_OK && plain_out)
        for (int i = 0; i < count; i++) plain_out[i] = ok[(size_t)i] ? 1 : 0;
    return rc;
}

void b200_plan_destroy(b200_plan *p) {
    if (!p) return;
    cudaSetDevice(p->device);
    if (p->pdesc_dev) cudaFree(p->pdesc_dev);
    if (p->cdesc_dev) cudaFree(p->cdesc_dev);
    if (p->exports_dev) cudaFree(p->exports_dev);
    if (p->pub_dev) cudaFree(p->pub_dev);
    if (p->pubcnt_dev) cudaFree(p->pubcnt_dev);
    if (p->arena_own) cudaFree(p->arena_own);
    if (p->state_dev) cudaFree(p->state_dev);
    if (p->trace_dev) cudaFree(p->trace_dev);
    (void)cudaGetLastError();
    free(p);
}

int b200_plan_create(b200_ctx *ctx, const b200_mul_mat_args *args, int count, const b200_plan_split *split, b200_plan **out) {
    B200_REQUIRE(ctx, ctx && args && out && count >= 1, B200_ERR_INVALID);
    *out = NULL;
    std::vector<int> src_op, m_total;
    std::vector<char> plain_ok;
    {
        const int rc = plan_analyze(ctx, args, count, split, &src_op, &m_total, &plain_ok);     // shapes, dataflow, aliasing hazards
        if (rc != B200_OK) return rc;
    }
    const int world = split ? split->world : 1, rank = split ? split->rank : 0;
    if (split)
        for (int r = 0; r < world; r++) B200_REQUIRE(ctx, split->peer_arena[r] != NULL, B200_ERR_INVALID);
    const int type = args[0].type;
    B200_REQUIRE(ctx, type == B200_TYPE_Q4_0 || type == B200_TYPE_Q8_0, B200_ERR_UNSUPPORTED);
    B200_REQUIRE(ctx, ctx->abort_dev != NULL && ctx->abort_host_dev != NULL, B200_ERR_CUDA);
    B200_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int grid = ctx->sm_count;
    const int qsb = b200_qs_bytes(type);
    // ring slot = 8 rows of a 4096-wide k-segment (18432 B for Q4_0 = one row of k = 32768).  The producer thread needs
    // ~250 ns per slot (measured: 4-row slots cap the stream at 5.4 TB/s), so slots must not be small.
    const int slot_rows = 8;
    const int slot_bytes = slot_rows * 128 * (qsb + 2);

    std::vector<long long> ll_off, pub_off;
    const size_t arena_elems = plan_arena_elems(args, count, split, &ll_off, &pub_off);
    std::vector<PubDesc> pubs;
    const std::vector<char> published = plan_published(args, count, src_op, grid, ctx->opt_plan_pub_min_k, ctx->opt_plan_pub_dist);
    B200_REQUIRE(ctx, arena_elems < (1ull << 30), B200_ERR_UNSUPPORTED);
    std::vector<PDesc> pd((size_t)count);
    std::vector<CDesc> cd((size_t)count);
    std::vector<ExportDesc> ex;
    int kmax = 0;
    for (int i = 0; i < count; i++)
        if (args[i].ne00 > kmax) kmax = (int)args[i].ne00;
    // shared-memory geometry: three activation buffers when the ring keeps its 8 slots, else two
    PlanGeom g;
    memset(&g, 0, sizeof(g));
    g.slot_bytes = slot_bytes;
    const int act_bytes = (int)b200_align_up((size_t)kmax + (size_t)(kmax / 32) * 8, 128);   // int8 planes + fp32 scales + 8 * sums
    const int ll_bytes = kCW * 2048 > kPartFloats * 4 ? kCW * 2048 : kPartFloats * 4;       // LL staging, aliased by the k-split partials
    const int desc_bytes = kDescCap * (int)sizeof(CDesc);
    const int bar_bytes = (2 * kMaxSlots + 2 * kMaxAct) * 8 + 64;
    const int max_smem = 227 * 1024;
    int nact = kMaxAct, nslots = 0;
    for (;; nact--) {
        const int fixed = nact * act_bytes + ll_bytes + desc_bytes + bar_bytes + 2048;
        nslots = (max_smem - fixed) / g.slot_bytes;
        if (nslots >= 8 || nact == 2) break;
    }
    if (nslots > kMaxSlots) nslots = kMaxSlots;
    if (nslots >= 8) nslots = nslots / 8 * 8;      // a multiple of the consumer warps when possible (every warp then owns the same number)
    if (ctx->opt_plan_slots >= 2 && ctx->opt_plan_slots <= nslots) nslots = ctx->opt_plan_slots;
    if (nslots < 2) { b200_set_error(ctx, "b200_plan_create: k = %d leaves no room for the weight ring", kmax); return B200_ERR_UNSUPPORTED; }
    g.nslots = nslots;
    g.nact = nact;
    g.l2_window = ctx->opt_plan_l2_window;
    g.evict_first = ctx->opt_plan_evict_first;
    g.ring_off = 0;
    g.act_off = nslots * g.slot_bytes;
    g.act_stride = act_bytes;
    g.ll_off = g.act_off + nact * act_bytes;
    g.desc_off = g.ll_off + ll_bytes;
    g.bar_off = g.desc_off + desc_bytes;
    g.pub_off = g.bar_off + bar_bytes;
    g.total = g.pub_off + 2048;

    int in_idx = -1;
    int pub_use[kMaxAct] = {0, 0, 0};
    for (int i = 0; i < count; i++) {
        const b200_mul_mat_args *a = &args[i];
        const int64_t nb = a->ne00 / B200_QK;
        const int64_t mt = m_total[i];
        const int64_t row0 = split ? split->row0[i] : 0;
        PDesc &p = pd[i];
        CDesc &c = cd[i];
        memset(&p, 0, sizeof(p));
        memset(&c, 0, sizeof(c));
        p.qs = (const uint8_t *)a->src0_dev + a->src0_block_off * qsb;
        p.d = (const __half *)((const uint8_t *)a->src0_dev + a->src0_nblocks_total * qsb) + a->src0_block_off;
        B200_REQUIRE(ctx, ((uintptr_t)p.qs & 15) == 0 && ((uintptr_t)p.d & 15) == 0, B200_ERR_UNSUPPORTED);
        p.k = c.k = (int)a->ne00;
        p.rows_q = c.rows_q = (int)(a->ne01 / grid);
        p.rows_rem = c.rows_rem = (int)(a->ne01 % grid);
        {
            int rs = slot_bytes / (int)(nb * (qsb + 2));             // rows per ring slot
            p.rs = rs > 64 ? 64 : rs;
        }
        c.dst_plain = plain_ok[(size_t)i] ? a->dst_dev : NULL;
        c.row0 = (int)row0;
        c.ll_dst = (int)ll_off[i];
        if (a->flags & B200_MM_EXPORT) {
            c.flags |= OPF_EXPORT;
            if (world > 1 && a->dst_dev) {
                B200_REQUIRE(ctx, plain_ok[(size_t)i] != 0, B200_ERR_UNSUPPORTED);
                c.flags |= OPF_WRITE_LL;
                ExportDesc e;
                e.dst = a->dst_dev; e.ll = c.ll_dst; e.m_total = (int)mt; e.op = i; e.pad = 0;
                ex.push_back(e);
            }
        }
        const int G = (int)((nb + kSegBlocks - 1) / kSegBlocks);
        const int Gp = G <= 1 ? 1 : (G <= 2 ? 2 : (G <= 4 ? 4 : 8));
        c.flags |= (p.rs << 8) | ((Gp == 1 ? 0 : Gp == 2 ? 1 : Gp == 4 ? 2 : 3) << 16);
        B200_REQUIRE(ctx, Gp == 1 || (size_t)(c.rows_q + 1) * Gp <= (size_t)kPartFloats, B200_ERR_UNSUPPORTED);
        // dataflow (plan_analyze): src1 is the dst of an earlier op, else an outside vector
        c.src_op = src_op[i];
        c.src_plain = a->src1_dev;
        if (c.src_op >= 0) {
            c.ll_src = cd[c.src_op].ll_dst;
            c.src_plain = NULL;
            cd[c.src_op].flags |= OPF_WRITE_LL;
        }
        // ops that read the vector the previous op read keep using the quantized activations already in shared memory
        if (i > 0 && c.k == cd[i - 1].k && c.src_op == cd[i - 1].src_op && (c.src_op >= 0 || c.src_plain == cd[i - 1].src_plain))
            c.flags |= OPF_SAME_INPUT;
        if (!(c.flags & OPF_SAME_INPUT)) {
            // a new src1: next activation buffer in rotation; its previous vector must have been released by every warp
            in_idx++;
            const int buf = in_idx % nact, use = in_idx / nact;
            c.flags |= buf << 20;
            if (use > 0) c.flags |= (1 << 22) | (((use - 1) & 1) << 23);
            if (published[(size_t)i]) {
                // quantized once per GPU by the publisher warps, brought in by the fetcher
                c.flags |= OPF_SRC_PUB | (((pub_use[buf]++) & 1) << 24);
                PubDesc pb;
                memset(&pb, 0, sizeof(pb));
                pb.ll_src = c.ll_src; pb.k = c.k; pb.src_op = c.src_op; pb.op = i; pb.pub = pub_off[i];
                pb.buf = buf; pb.free_wait = use > 0; pb.free_par = (use - 1) & 1;
                pubs.push_back(pb);
            }
        }
    }
    b200_plan *p = (b200_plan *)calloc(1, sizeof(b200_plan));
    if (!p) return B200_ERR_ALLOC;
    p->type = type; p->nops = count; p->grid = grid; p->world = world; p->rank = rank; p->device = ctx->device;
    p->arena_bytes = arena_elems * 8;
    p->geom = g;

    cudaError_t e = cudaMalloc((void **)&p->pdesc_dev, sizeof(PDesc) * (size_t)count);
    if (e == cudaSuccess) e = cudaMalloc((void **)&p->cdesc_dev, sizeof(CDesc) * (size_t)count);
    if (e == cudaSuccess && !ex.empty()) e = cudaMalloc((void **)&p->exports_dev, sizeof(ExportDesc) * ex.size());
    if (e == cudaSuccess) e = cudaMalloc((void **)&p->state_dev, 16);
    if (e == cudaSuccess && world == 1) e = cudaMalloc(&p->arena_own, p->arena_bytes);
    if (e == cudaSuccess && ctx->opt_plan_trace) {
        e = cudaMalloc((void **)&p->trace_dev, (size_t)(count + 1) * grid * 4 * 8);
        if (e == cudaSuccess) e = cudaMemset(p->trace_dev, 0, (size_t)(count + 1) * grid * 4 * 8);
    }
    if (e == cudaSuccess) e = cudaMemcpy(p->pdesc_dev, pd.data(), sizeof(PDesc) * (size_t)count, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(p->cdesc_dev, cd.data(), sizeof(CDesc) * (size_t)count, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && !ex.empty()) e = cudaMemcpy(p->exports_dev, ex.data(), sizeof(ExportDesc) * ex.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(p->state_dev, 0, 16);
    if (e == cudaSuccess && world == 1) e = cudaMemset(p->arena_own, 0, p->arena_bytes);
    if (e == cudaSuccess && !pubs.empty()) e = cudaMalloc((void **)&p->pubcnt_dev, sizeof(uint32_t) * pubs.size());
    if (e == cudaSuccess && !pubs.empty()) e = cudaMemset(p->pubcnt_dev, 0, sizeof(uint32_t) * pubs.size());
    if (e == cudaSuccess && !pubs.empty()) e = cudaMalloc((void **)&p->pub_dev, sizeof(PubDesc) * pubs.size());
    if (e == cudaSuccess && !pubs.empty()) e = cudaMemcpy(p->pub_dev, pubs.data(), sizeof(PubDesc) * pubs.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(plan_kernel_for(type), cudaFuncAttributeMaxDynamicSharedMemorySize, g.total);
    int per_sm = 0;
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, plan_kernel_for(type), kPlanThreads, (size_t)g.total);
    if (e != cudaSuccess) {
        b200_set_error(ctx, "b200_plan_create: %s", cudaGetErrorString(e));
        (void)cudaGetLastError();
        b200_plan_destroy(p);
        return e == cudaErrorMemoryAllocation ? B200_ERR_ALLOC : B200_ERR_CUDA;
    }
    if (per_sm < 1) {
        // the CTAs of a plan wait for one another: without room for all of them at once there is no plan
        b200_set_error(ctx, "b200_plan_create: the plan kernel does not fit an SM of this device (%d bytes of shared memory)", g.total);
        b200_plan_destroy(p);
        return B200_ERR_UNSUPPORTED;
    }
    PlanArgs &pa = p->args;
    memset(&pa, 0, sizeof(pa));
    pa.pdesc = p->pdesc_dev;
    pa.cdesc = p->cdesc_dev;
    pa.exports = p->exports_dev;
    pa.nops = count;
    pa.nexports = (int)ex.size();
    pa.world = world;
    pa.rank = rank;
    for (int r = 0; r < world; r++) pa.arena[r] = split ? split->peer_arena[r] : p->arena_own;
    pa.state = p->state_dev;
    pa.trace = p->trace_dev;
    pa.pub = p->pub_dev;
    pa.npub = (int)pubs.size();
    pa.pub_count = p->pubcnt_dev;
    pa.abort_flag = ctx->abort_dev;
    pa.abort_host = ctx->abort_host_dev;
    pa.timeout_ns = (unsigned long long)(ctx->opt_plan_timeout_ms > 0 ? ctx->opt_plan_timeout_ms : (world > 1 ? 120000 : 10000)) * 1000000ull;
    *out = p;
    return B200_OK;
}

// Cooperative launch: the CTAs of a plan spin on one another's stores, so either all of them are resident or the launch
// does not start (another plan, a collective, an MPS partition holding SMs delays the launch instead of deadlocking it).
int b200_plan_launch(b200_ctx *ctx, b200_plan *p) {
    B200_REQUIRE(ctx, ctx && p, B200_ERR_INVALID);
    B200_REQUIRE(ctx, p->device == ctx->device, B200_ERR_INVALID);
    B200_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)p->grid, 1, 1);
    cfg.blockDim = dim3(kPlanThreads, 1, 1);
    cfg.dynamicSmemBytes = (size_t)p->geom.total;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    void *kargs[2] = {(void *)&p->args, (void *)&p->geom};
    const cudaError_t e = cudaLaunchKernelExC(&cfg, (const void *)plan_kernel_for(p->type), kargs);
    if (e == cudaErrorCooperativeLaunchTooLarge) {
        (void)cudaGetLastError();
        b200_set_error(ctx, "b200_plan_launch: %d CTAs cannot be co-resident on this device right now", p->grid);
        return B200_ERR_UNSUPPORTED;
    }
    B200_CUDA_TRY(ctx, e);
    ctx->launches++;
    return B200_OK;
}

int b200_plan_trace(b200_ctx *ctx, b200_plan *p, unsigned long long *out_host, size_t capacity_u64, int *nops, int *grid) {
    B200_REQUIRE(ctx, ctx && p, B200_ERR_INVALID);
    if (nops) *nops = p->nops;
    if (grid) *grid = p->grid;
    if (!p->trace_dev || !out_host) return B200_ERR_UNSUPPORTED;
    const size_t n = (size_t)(p->nops + 1) * p->grid * 4;     // + one row of per-CTA totals
    B200_REQUIRE(ctx, capacity_u64 >= n, B200_ERR_INVALID);
    B200_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    B200_CUDA_TRY(ctx, cudaMemcpy(out_host, p->trace_dev, n * 8, cudaMemcpyDeviceToHost));
    return B200_OK;
}

}  // extern "C"
