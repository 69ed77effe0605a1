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
//   * grid = one persistent CTA per SM; every CTA walks the op list; in op i it owns a contiguous run of rows of W_i;
//   * warp 8 is the producer: ONE elected thread streams the CTA's byte ranges of op 0, 1, 2, ... back to back into a
//     shared-memory ring (cp.async.bulk + mbarrier complete_tx), 8 slots x 18 KB in flight per SM (~22 MB per GPU = 3 us of HBM).
//     Weights never depend on activations, so the producer runs arbitrarily far ahead of the consumers: HBM keeps
//     streaming W_{i+1} while the grid exchanges the result of op i;
//   * dependencies between ops travel as tagged 8-byte elements {fp32 value, u32 tag} ("LL" vectors, one per op, in an
//     arena): the GEMV epilogue stores them, the next op's activation loads re-read until every tag matches.  No grid
//     barrier, no flags, no fences; with world > 1 the same stores go to every rank's arena over NVLink, which makes the
//     all-gather of a row-split mul_mat part of the epilogue;
//   * consumers: 8 warps; a ring slot (8 rows of a 4096-wide k-segment) always belongs to the same team of warps (one warp for
//     k <= 4096), which turns the whole slot into results: up to four rows' loads in flight, exact dp4a block dots against
//     the quantized src1 in shared memory, four row sums reduced together by a transposed butterfly;
//   * consecutive ops that read the same vector (fc_in, v, q, k) share one activation quantization (two shared-memory buffers,
//     so a new src1 can be quantized while rows of the previous one are still running in other warps).
#include "b200_stream_common.cuh"

#include <stdlib.h>
#include <vector>

using namespace b200s;

namespace {

#ifndef B200_PLAN_ROW_SPLIT
#define B200_PLAN_ROW_SPLIT 1                // 2: the rows of a ring slot are shared by two warps (16 consumer warps, 96 registers each)
#endif
constexpr int kRowSplit = B200_PLAN_ROW_SPLIT;
constexpr int kCW = 8 * kRowSplit;           // consumer warps
constexpr int kCT = kCW * 32;                // consumer threads
constexpr int kPlanThreads = (kCW + 1) * 32;         // consumers + the producer warp
constexpr int plan_threads(int mode) { return (mode & 12) ? kPlanThreads + 32 : kPlanThreads; }   // MODE 4 / 8: + the publisher warp
constexpr int kSegBlocks = 128;              // blocks of k per warp-segment (4 per lane)
constexpr int kMaxSlots = 24;
constexpr int kPartFloats = 4096;            // k-split partials parked per CTA per op: rows_per_cta * G (aliases the LL staging)
constexpr int kMaxOps = 1023;
constexpr int kDescCap = 128;                // op descriptors staged in shared memory per window

enum : int { OPF_SAME_INPUT = 1, OPF_WRITE_LL = 2, OPF_EXPORT = 4, OPF_SRC_RING = 8, OPF_SRC_LLQ = 16 };

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
    union {
        const float *src_plain;  // src1 when it comes from outside the plan (src_op < 0)
        long long src_pub;       // OPF_SRC_LLQ: arena element offset of the vector the publisher warps quantize once per GPU
    };
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
struct PubDesc {             // MODE 4: an in-plan src1 vector that is quantized once per GPU (for the op `op` that reads it)
    int ll_src, k, src_op, llq, op, pad[3];      // llq: arena element offset of the published vector
};

struct PlanGeom {
    int slot_bytes, nslots;
    int l2_ahead;            // ops: when the producer starts op i it prefetches its rows of op i + l2_ahead into L2 (0 = off)
    int l2_slots;            // MODE & 2: ring slots the L2 prefetch cursor runs ahead of the copies
    int ring_off, act_off, act_stride, ll_off, desc_off, bar_off, total;   // act: two buffers of act_stride bytes
    int pub_off;             // MODE 4: 2 KB of staging for the publisher warp
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
    const int *p_ll;               // LLRING kernels: per op, arena element offset of a src1 vector the producer feeds through the ring, or -1
    const PubDesc *pub;            // MODE 4 / 8: the publisher warp's work list, in op order
    int npub;
    uint32_t *pub_count;           // MODE 8: per work-list entry, CTAs that have published (never reset: grid per launch)
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
// spins on its run's LAST element only (one sector per poll) before trying again.
__device__ __forceinline__ void ll_fetch_warp(const char *wbase, int nvalid, uint32_t tag, float *wstage, int lane, float4 (&out)[4]) {
    uint4 w[8];
    const int nv8 = nvalid * 8;
    for (;;) {
        bool ok = true;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int idx = j * 32 + lane < nv8 ? j * 32 + lane : 0;
            asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w[j].x), "=r"(w[j].y), "=r"(w[j].z), "=r"(w[j].w) : "l"(wbase + (size_t)idx * 16));
            ok = ok && w[j].y == tag && w[j].w == tag;
        }
        if (__all_sync(0xffffffffu, ok)) break;
        ll_probe(wbase + (size_t)nvalid * 128 - 8, tag);
    }
#pragma unroll
    for (int j = 0; j < 8; j++)
        if (j * 32 + lane < nv8) *reinterpret_cast<float2 *>(wstage + 2 * (j * 32 + lane)) = make_float2(__uint_as_float(w[j].x), __uint_as_float(w[j].z));
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; j++) out[j] = *reinterpret_cast<const float4 *>(wstage + lane * 16 + j * 4);
    __syncwarp();
}

// The same 32 lane-tasks when the producer thread has pushed the tagged vector through the weight ring (a src1 that was complete
// long before its consumer: no L2 round trips, the bytes are already in shared memory).  The vector lies in consecutive ring
// slots starting at position st0; byte0 = offset of the run's first task in the vector.  Returns false (warp-uniform) when a
// tag does not match -- the copy was taken before some CTA had stored its rows -- and the caller falls back to ll_fetch_warp.
__device__ __forceinline__ bool ll_fetch_ring(uint32_t ring_a, uint32_t full_a, int st0, uint32_t par0, int nslots, int slot_bytes, uint32_t byte0,
                                              int nvalid, uint32_t tag, float *wstage, int lane, float4 (&out)[4]) {
    const int j0 = (int)(byte0 / (uint32_t)slot_bytes), j1 = (int)((byte0 + (uint32_t)nvalid * 128u - 1u) / (uint32_t)slot_bytes);
    int p0 = st0 + j0, p1 = st0 + j1;
    uint32_t q0 = par0, q1 = par0;
    if (p0 >= nslots) { p0 -= nslots; q0 ^= 1u; }
    if (p1 >= nslots) { p1 -= nslots; q1 ^= 1u; }
    mbar_wait_a(full_a + 8u * (uint32_t)p0, q0);
    if (j1 != j0) mbar_wait_a(full_a + 8u * (uint32_t)p1, q1);
    const uint32_t bound = (uint32_t)(j0 + 1) * (uint32_t)slot_bytes;
    const uint32_t base0 = ring_a + (uint32_t)p0 * (uint32_t)slot_bytes - (uint32_t)j0 * (uint32_t)slot_bytes;   // + byte offset in the vector
    const uint32_t base1 = ring_a + (uint32_t)p1 * (uint32_t)slot_bytes - (uint32_t)j1 * (uint32_t)slot_bytes;
    uint4 w[8];
    const int nv8 = nvalid * 8;
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int idx = j * 32 + lane < nv8 ? j * 32 + lane : 0;
        const uint32_t b = byte0 + (uint32_t)idx * 16u;
        w[j] = lds128((b < bound ? base0 : base1) + b);
        ok = ok && w[j].y == tag && w[j].w == tag;
    }
    if (!__all_sync(0xffffffffu, ok)) return false;
#pragma unroll
    for (int j = 0; j < 8; j++)
        if (j * 32 + lane < nv8) *reinterpret_cast<float2 *>(wstage + 2 * (j * 32 + lane)) = make_float2(__uint_as_float(w[j].x), __uint_as_float(w[j].z));
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; j++) out[j] = *reinterpret_cast<const float4 *>(wstage + lane * 16 + j * 4);
    __syncwarp();
    return true;
}

// MODE bit 0: ring-fed src1 vectors (B200_PLAN_LL_RING); bit 1: per-slot L2 prefetch ahead of the ring (B200_PLAN_L2_SLOTS).
// bit 2: a long src1 (B200_PLAN_LLQ) is quantized ONCE per GPU -- every CTA does 1/grid of its blocks and publishes them as tagged
// words -- instead of once per CTA; each CTA then fetches 80 bytes per block instead of 256 and does no arithmetic.
// bit 3 (B200_PLAN_PUBQ=1; first GPU run at the very end of round 1: 4 DAG cases pass bit for bit, speed not measured yet): the same once-per-GPU quantization, published as the plain activation planes
// + a per-vector arrival counter (release / acquire) instead of tagged words, so that a CTA takes the whole quantized vector
// with ONE bulk copy straight into its activation buffer -- no registers, no tag checks, 40 bytes per block instead of 80.
// bit 4 (B200_PLAN_NOSPLIT=1, NEVER RUN): ops with k > 4096 are not split over a team of warps; the slot's warp walks the 4096-wide
// segments itself (a butterfly per segment, the partials added in segment order: the same bits), so there is no partials pass,
// no barrier around it, no barrier when the team size changes, and the rows leave slot by slot instead of at the end of the op.
// MODE 0 is the shipped kernel; the others are experiments kept out of its code.
template <int TYPE, int MODE>
__global__ void __launch_bounds__(plan_threads(MODE), 1) plan_kernel(const __grid_constant__ PlanArgs pa, const __grid_constant__ PlanGeom pg) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int QSB = TYPE == B200_TYPE_Q4_0 ? 16 : 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cta = blockIdx.x;
    constexpr bool LLRING = (MODE & 1) != 0, L2SLOTS = (MODE & 2) != 0, LLQ = (MODE & 4) != 0, PUBQ = (MODE & 8) != 0;
    constexpr bool PUBW = LLQ || PUBQ;           // a publisher warp exists
    constexpr bool NOSPLIT = (MODE & 16) != 0;

    unsigned char *ring = smem + pg.ring_off;
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + pg.bar_off);
    uint64_t *empty_bar = full_bar + kMaxSlots;
    uint32_t *s_epoch = reinterpret_cast<uint32_t *>(empty_bar + kMaxSlots);
    uint64_t *act_bar = empty_bar + kMaxSlots + 2;       // MODE 8: completion of the bulk copy of a published vector

    if (threadIdx.x == 0) {
        for (int s = 0; s < pg.nslots; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kCW);       // a slot's consumers (one team of G [x 2] warps) arrive with 8/G each
        }
        *s_epoch = pa.state[1] + 1u;    // every CTA reads it before any CTA can finish (the bump needs all of them)
        if (PUBQ) mbar_init(act_bar, 1);
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
            uint32_t par = 0;
            unsigned long long prod_blocked = 0;
            const uint32_t ring_a = smem_u32(ring);
            PDesc cur = pa.pdesc[0];
            int cur_ll = LLRING ? pa.p_ll[0] : -1;
            // L2SLOTS: a second cursor walks the same rows pg.l2_slots ring slots ahead of the copies and asks L2 for them
            // (cp.async.bulk.prefetch.L2): while the consumers are busy with a hand-off and the ring is full, HBM keeps
            // delivering into L2, and the ring restarts from L2 instead of from DRAM.
            int pf_op = 0, pf_r = 0, pf_nrows = 0, pf_skip = L2SLOTS ? pg.l2_slots : 0;
            PDesc pfd = cur, pfn = cur;
            if (L2SLOTS) {
                pf_nrows = pfd.rows_q + (cta < pfd.rows_rem ? 1 : 0);
                if (pa.nops > 1) pfn = pa.pdesc[1];
            }
#pragma unroll 1
            for (int op = 0; op < pa.nops; op++) {
                PDesc nxt = cur;
                int nxt_ll = -1;
                if (op + 1 < pa.nops) {
                    nxt = pa.pdesc[op + 1];
                    if (LLRING) nxt_ll = pa.p_ll[op + 1];
                }
                if (LLRING && cur_ll >= 0) {
                    // this op's src1 was finished long ago: its tagged vector travels through the ring ahead of the op's weights
                    const char *src = arena_local + (size_t)cur_ll * 8;
                    const int bytes = cur.k * 8;
                    for (int off = 0; off < bytes; off += pg.slot_bytes) {
                        const uint32_t nbytes = (uint32_t)min(pg.slot_bytes, bytes - off);
                        const uint32_t fb = smem_u32(&full_bar[st]);
                        mbar_wait(&empty_bar[st], par ^ 1u);
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"(nbytes) : "memory");
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(ring_a + (uint32_t)(st * pg.slot_bytes)),
                                     "l"(src + off), "r"(nbytes), "r"(fb) : "memory");
                        if (++st == nslots) { st = 0; par ^= 1u; }
                    }
                }
                if (pg.l2_ahead > 0 && op + pg.l2_ahead < pa.nops) {
                    // HBM -> L2 for an op the ring will reach later (optional; off by default)
                    const PDesc pf = pa.pdesc[op + pg.l2_ahead];
                    const int nbp = pf.k >> 5;
                    const long long rb = (long long)cta * pf.rows_q + min(cta, pf.rows_rem);
                    const int nr = pf.rows_q + (cta < pf.rows_rem ? 1 : 0);
                    if (nr > 0) {
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pf.qs + rb * nbp * QSB), "r"((uint32_t)(nr * nbp * QSB)) : "memory");
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const uint8_t *>(pf.d) + rb * nbp * 2), "r"((uint32_t)(nr * nbp * 2)) : "memory");
                    }
                }
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
                    if (L2SLOTS) {
                        // one slot's worth at the prefetch cursor (the first l2_slots steps only move it ahead)
                        while (pf_op < pa.nops && pf_r >= pf_nrows) {
                            pf_op++;
                            pfd = pfn;
                            pf_r = 0;
                            pf_nrows = pfd.rows_q + (cta < pfd.rows_rem ? 1 : 0);
                            if (pf_op + 1 < pa.nops) pfn = pa.pdesc[pf_op + 1];      // consumed at the next crossing
                        }
                        if (pf_op < pa.nops) {
                            if (pf_skip > 0) {
                                pf_skip--;
                            } else {
                                const int pnb = pfd.k >> 5;
                                const long long prow = (long long)cta * pfd.rows_q + min(cta, pfd.rows_rem) + pf_r;
                                const int prows = min(pfd.rs, pf_nrows - pf_r);
                                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pfd.qs + prow * pnb * QSB), "r"((uint32_t)(prows * pnb * QSB)) : "memory");
                                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const uint8_t *>(pfd.d) + prow * pnb * 2), "r"((uint32_t)(prows * pnb * 2)) : "memory");
                            }
                            pf_r += pfd.rs;
                        }
                    }
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"((uint32_t)(rows * (row_qs + row_sc))) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                                 "l"(gq + (size_t)r * row_qs), "r"((uint32_t)(rows * row_qs)), "r"(fb) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + stage_qs),
                                 "l"(gs + (size_t)r * row_sc), "r"((uint32_t)(rows * row_sc)), "r"(fb) : "memory");
                    if (++st == nslots) { st = 0; par ^= 1u; }
                }
                cur = nxt;
                cur_ll = nxt_ll;
            }
            if (pa.trace) pa.trace[(size_t)pa.nops * gridDim.x * 4 + (size_t)cta * 4 + 0] = prod_blocked;
        }
    } else if (PUBW && warp == kCW + 1) {
        // ===== publisher (MODE 4 / 8): for every long in-plan src1, in op order, this CTA's share of the blocks
        // [cta * nb / grid, (cta + 1) * nb / grid): wait for the fp32 values (tagged vector of the producing op), quantize_row_q8_0,
        // publish.  Runs as far ahead of the consumers as the data allows; every rank does the same for its own arena. =====
        float *pstage = reinterpret_cast<float *>(smem + pg.pub_off);
        uint2 *arena_w = reinterpret_cast<uint2 *>(const_cast<char *>(arena_local));
#pragma unroll 1
        for (int i = 0; i < pa.npub; i++) {
            const PubDesc d = pa.pub[i];
            const int nb = d.k >> 5;
            const int b_lo = (cta * nb) / (int)gridDim.x, b_hi = ((cta + 1) * nb) / (int)gridDim.x;
            const int nt = 2 * (b_hi - b_lo);
            if (nt <= 0) {
                if (PUBQ && lane == 0) atomicAdd(pa.pub_count + i, 1u);      // nothing to publish, but the count is per CTA
                continue;
            }
            const uint32_t src_tag = (epoch << 10) | (uint32_t)(d.src_op & 1023), tag = (epoch << 10) | (uint32_t)d.op;
            float4 vv[4];
            ll_fetch_warp(arena_local + (size_t)d.ll_src * 8 + (size_t)(2 * b_lo) * 128, nt, src_tag, pstage, lane, vv);
            const int b = b_lo + (lane >> 1), h = lane & 1;
            float amax = 0.0f;
#pragma unroll
            for (int j = 0; j < 4; j++)
                amax = fmaxf(amax, fmaxf(fmaxf(fabsf(vv[j].x), fabsf(vv[j].y)), fmaxf(fabsf(vv[j].z), fabsf(vv[j].w))));
            amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
            const float id = amax != 0.0f ? __fdiv_rn(127.f, amax) : 0.0f;
            uint32_t pk[4];
            int sq = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int q0 = __float2int_rn(__fmul_rn(vv[j].x, id)), q1 = __float2int_rn(__fmul_rn(vv[j].y, id));
                const int q2 = __float2int_rn(__fmul_rn(vv[j].z, id)), q3 = __float2int_rn(__fmul_rn(vv[j].w, id));
                sq += q0 + q1 + q2 + q3;
                pk[j] = (uint32_t)(q0 & 0xff) | ((uint32_t)(q1 & 0xff) << 8) | ((uint32_t)(q2 & 0xff) << 16) | ((uint32_t)(q3 & 0xff) << 24);
            }
            sq += __shfl_xor_sync(0xffffffffu, sq, 1);
            if (PUBQ) {
                // the consumers' activation-buffer layout, plain: int8 elements 0..15 / 16..31 of every block, fp32 d, 8 * sum(q)
                if (lane < nt) {
                    unsigned char *reg = reinterpret_cast<unsigned char *>(arena_w + (size_t)d.llq);
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
            } else if (lane < nt) {
                uint2 *blk = arena_w + (size_t)d.llq + (size_t)b * 10;
#pragma unroll
                for (int j = 0; j < 4; j++) ll_store(blk, h * 4 + j, __uint_as_float(pk[j]), tag);
                if (h == 0) {
                    ll_store(blk, 8, __half2float(__float2half_rn(__fdiv_rn(amax, 127.f))), tag);
                    ll_store(blk, 9, __int_as_float(8 * sq), tag);
                }
            }
        }
    } else if (!PUBW || warp < kCW) {
        // ===== consumers =====
        const uint32_t ring_a = smem_u32(ring);
        const uint32_t full_a = smem_u32(full_bar), empty_a = smem_u32(empty_bar);
        CDesc *sdesc = reinterpret_cast<CDesc *>(smem + pg.desc_off);
        int act_sel = 1;            // which of the two activation buffers holds the current src1
        float *llstage = reinterpret_cast<float *>(smem + pg.ll_off) + warp * 512;
        float *part = reinterpret_cast<float *>(smem + pg.ll_off);      // same bytes, other phase (bar.sync in between)
        int st0 = 0;                // ring position of the op's first stage
        uint32_t par0 = 0;
        int prevG = 1;
        unsigned long long cons_blocked = 0, quant_time = 0;
        unsigned ring_miss = 0;     // LLRING: runs of a ring-fed vector this warp had to re-fetch from L2
        uint32_t act_par = 0;       // MODE 8: phase of act_bar
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
            // geometry comes precomputed in the descriptor (flags: bits 8..15 rows per slot, 16..19 log2 of the k-split)
            const int k = o->k, flags = o->flags;
            const int nb = k >> 5;
            const int rs = (flags >> 8) & 0xff, log2g_op = (flags >> 16) & 0xf;
            const int log2g = NOSPLIT ? 0 : log2g_op;             // NOSPLIT: every slot belongs to one warp ...
            const int nseg = NOSPLIT ? (1 << log2g_op) : 1;       // ... which walks the op's k-segments itself
            const int rows_q = o->rows_q, rows_rem = o->rows_rem;
            const int r_begin = cta * rows_q + min(cta, rows_rem);
            const int nrows = rows_q + (cta < rows_rem ? 1 : 0);
            // 8/G teams; a team = G warps splitting k (x kRowSplit warps splitting the slot's rows)
            constexpr int kLog2Split = kRowSplit == 2 ? 1 : 0;
            const int G = 1 << log2g, seg = warp & (G - 1), half = (warp >> log2g) & (kRowSplit - 1);
            const int team = warp >> (log2g + kLog2Split), team_mask = (8 >> log2g) - 1;
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
                // ---- a new src1: all 8 consumer warps fetch and quantize it (quantize_row_q8_0, bit-exact, same code path as
                // gemv_stream_kernel) into one of two shared-memory buffers ----
                const unsigned long long tq0 = tr ? gtime() : 0ull;
                // two buffers: rows of the previous input may still be running in other warps; the barrier after THIS
                // quantization guarantees they are done before that buffer is written again two inputs later
                act_sel ^= 1;
                unsigned char *act = smem + pg.act_off + act_sel * pg.act_stride;
                const uint32_t act_a = smem_u32(act);
                const int src_op = o->src_op;
                const bool ll_in = src_op >= 0;
                const uint32_t src_tag = (epoch << 10) | (uint32_t)(src_op & 1023);
                const bool ring_src = LLRING && (flags & OPF_SRC_RING) != 0;
                // ring-fed src1: every warp reads slots it does not own, so all earlier fills must have been consumed first
                // (a parity wait on a slot whose previous fill is still pending would pass a lap early)
                if (ring_src) cbar();
                const char *xsrc = ll_in ? arena_local + (size_t)o->ll_src * 8 : reinterpret_cast<const char *>(o->src_plain);
                constexpr int kQB = 4 / kRowSplit;
                const int tpc = nb * 2;      // lane-tasks: (block, half) = 16 consecutive floats
                const bool llq_src = LLQ && (flags & OPF_SRC_LLQ) != 0;
                if (PUBQ && (flags & OPF_SRC_LLQ) != 0) {
                    // ---- the publisher warps of all CTAs have left (or will leave) the quantized vector in the arena in exactly
                    // the layout of the activation buffer: wait for the arrival count, then ONE bulk copy brings it in.
                    if (threadIdx.x == 0) {
                        const uint32_t target = epoch * gridDim.x;        // the counters are never reset: grid arrivals per launch
                        const uint32_t *cnt = pa.pub_count + o->ll_src;   // (ll_src carries the work-list index for these ops)
                        uint32_t seen;
                        do {
                            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(cnt) : "memory");
                        } while ((int32_t)(seen - target) < 0);
                        asm volatile("fence.proxy.async.global;" ::: "memory");
                        const uint32_t ab = smem_u32(act_bar), abytes = (uint32_t)(k + nb * 8);
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ab), "r"(abytes) : "memory");
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(act_a),
                                     "l"(arena_local + (size_t)o->src_pub * 8), "r"(abytes), "r"(ab) : "memory");
                    }
                    if (tr && threadIdx.x == 0) tr[0] = gtime();
                    mbar_wait_a(smem_u32(act_bar), act_par);
                    act_par ^= 1u;
                } else if (llq_src) {
                    // ---- the vector was quantized once per GPU by the publisher warps (below): 10 tagged words per block
                    // {q[0..31] as 8 words, fp32 d, 8 * sum(q)} in the local arena.
                    uint2 *llq = reinterpret_cast<uint2 *>(const_cast<char *>(arena_local)) + (size_t)o->src_pub;
                    if (tr && threadIdx.x == 0) tr[0] = gtime();
                    // every warp: two blocks per thread and round trip, straight into the activation planes
#pragma unroll 1
                    for (int bb = 0; bb < nb; bb += 2 * kCT) {
                        const int bA = bb + (int)threadIdx.x, bB = bb + kCT + (int)threadIdx.x;
                        const uint4 *pA = reinterpret_cast<const uint4 *>(llq + (size_t)min(bA, nb - 1) * 10);
                        const uint4 *pB = reinterpret_cast<const uint4 *>(llq + (size_t)min(bB, nb - 1) * 10);
                        const int wlast = min(bb + warp * 32 + 31, nb - 1);       // probe target after a miss: this warp's last block
                        uint4 wa[5], wb[5];
                        for (;;) {
                            bool ok = true;
#pragma unroll
                            for (int j = 0; j < 5; j++) {
                                asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(wa[j].x), "=r"(wa[j].y), "=r"(wa[j].z), "=r"(wa[j].w) : "l"(pA + j));
                                asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(wb[j].x), "=r"(wb[j].y), "=r"(wb[j].z), "=r"(wb[j].w) : "l"(pB + j));
                            }
#pragma unroll
                            for (int j = 0; j < 5; j++) ok = ok && wa[j].y == tag && wa[j].w == tag && wb[j].y == tag && wb[j].w == tag;
                            if (__all_sync(0xffffffffu, ok)) break;
                            ll_probe(llq + (size_t)wlast * 10 + 9, tag);
                        }
                        if (bA < nb) {
                            *reinterpret_cast<uint4 *>(act + (size_t)bA * 16) = make_uint4(wa[0].x, wa[0].z, wa[1].x, wa[1].z);
                            *reinterpret_cast<uint4 *>(act + (size_t)(k >> 1) + (size_t)bA * 16) = make_uint4(wa[2].x, wa[2].z, wa[3].x, wa[3].z);
                            reinterpret_cast<uint32_t *>(act + k)[bA] = wa[4].x;
                            if (TYPE == B200_TYPE_Q4_0) reinterpret_cast<uint32_t *>(act + k + (size_t)nb * 4)[bA] = wa[4].z;
                        }
                        if (bB < nb) {
                            *reinterpret_cast<uint4 *>(act + (size_t)bB * 16) = make_uint4(wb[0].x, wb[0].z, wb[1].x, wb[1].z);
                            *reinterpret_cast<uint4 *>(act + (size_t)(k >> 1) + (size_t)bB * 16) = make_uint4(wb[2].x, wb[2].z, wb[3].x, wb[3].z);
                            reinterpret_cast<uint32_t *>(act + k)[bB] = wb[4].x;
                            if (TYPE == B200_TYPE_Q4_0) reinterpret_cast<uint32_t *>(act + k + (size_t)nb * 4)[bB] = wb[4].z;
                        }
                    }
                } else {
#pragma unroll 1
                for (int base = 0; base < tpc; base += kCT * kQB) {
                    float4 v[kQB][4];
#pragma unroll
                    for (int u = 0; u < kQB; u++) {
                        const int tb = base + u * kCT + warp * 32;      // first lane-task of this warp
                        if (tb >= tpc) continue;
                        if (ll_in) {
                            bool got = false;
                            if (ring_src) {
                                got = ll_fetch_ring(ring_a, full_a, st0, par0, nslots, pg.slot_bytes, (uint32_t)tb * 128u, min(32, tpc - tb), src_tag, llstage, lane, v[u]);
                                if (!got) ring_miss++;
                            }
                            if (!got) ll_fetch_warp(xsrc + (size_t)tb * 128, min(32, tpc - tb), src_tag, llstage, lane, v[u]);
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
                        float amax = 0.0f;
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v[u][j].x), fabsf(v[u][j].y)), fmaxf(fabsf(v[u][j].z), fabsf(v[u][j].w))));
                        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
                        const float id = amax != 0.0f ? __fdiv_rn(127.f, amax) : 0.0f;
                        uint32_t pk[4];
                        int sq = 0;
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const int q0 = __float2int_rn(__fmul_rn(v[u][j].x, id)), q1 = __float2int_rn(__fmul_rn(v[u][j].y, id));
                            const int q2 = __float2int_rn(__fmul_rn(v[u][j].z, id)), q3 = __float2int_rn(__fmul_rn(v[u][j].w, id));
                            sq += q0 + q1 + q2 + q3;
                            pk[j] = (uint32_t)(q0 & 0xff) | ((uint32_t)(q1 & 0xff) << 8) | ((uint32_t)(q2 & 0xff) << 16) | ((uint32_t)(q3 & 0xff) << 24);
                        }
                        sq += __shfl_xor_sync(0xffffffffu, sq, 1);
                        if (live) {
                            *reinterpret_cast<uint4 *>(act + (size_t)h * (k >> 1) + (size_t)b * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                            if (h == 0) {
                                reinterpret_cast<float *>(act + k)[b] = __half2float(__float2half_rn(__fdiv_rn(amax, 127.f)));
                                if (TYPE == B200_TYPE_Q4_0) reinterpret_cast<int *>(act + k + (size_t)nb * 4)[b] = 8 * sq;
                            }
                        }
                    }
                }
                }
                if (ring_src) {
                    // this warp is done with the vector's slots (each of the 8 warps arrives once per slot)
                    __syncwarp();
                    const int nls = (k * 8 + pg.slot_bytes - 1) / pg.slot_bytes;
                    for (int j = 0; j < nls; j++) {
                        if (lane == 0) mbar_arrive_cnt(empty_a + 8u * (uint32_t)st0, 1u);
                        if (++st0 == nslots) { st0 = 0; par0 ^= 1u; }
                    }
                }
                cbar();
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
                    const int rows_slot = min(rs, nrows - rbase);
                    // with two warps per slot: the first takes rows [0, split), the second [split, rows_slot)
                    const int split = kRowSplit == 2 ? (rows_slot + 1) >> 1 : rows_slot;
                    const int rows = half == 0 ? split : rows_slot;
                    const uint32_t stage_a = ring_a + (uint32_t)(st * pg.slot_bytes);
#pragma unroll 1
                    for (int r = half == 0 ? 0 : split; r < rows;) {
                        float v;
                        int u, step;
                        bool holder;
                        if (NOSPLIT && nseg > 1) {
                            // all segments of these rows in this warp: v = ((0 + p_0) + p_1) + ... exactly as the partials pass adds them
                            const int nr = rows - r > 2 ? 4 : (rows - r == 2 ? 2 : 1);
                            v = 0.0f;
#pragma unroll 1
                            for (int sg = 0; sg < nseg; sg++) {
                                RowCtx<TYPE> cs = c;
                                cs.a_lo += (uint32_t)(sg * kSegBlocks * 16);
                                cs.a_hi += (uint32_t)(sg * kSegBlocks * 16);
                                cs.a_d += (uint32_t)(sg * kSegBlocks * 4);
                                cs.a_s += (uint32_t)(sg * kSegBlocks * 4);
                                cs.woff0 += (uint32_t)(sg * kSegBlocks * QSB);
                                cs.soff0 += (uint32_t)(sg * kSegBlocks * 2);
#pragma unroll
                                for (int i = 0; i < 4; i++) cs.blive[i] = sg * kSegBlocks + lane + 32 * i < nb;
                                v += nr == 4 ? chunk_rows<TYPE, 4>(cs, stage_a, r, rows, lane)
                                             : (nr == 2 ? chunk_rows<TYPE, 2>(cs, stage_a, r, rows, lane) : chunk_rows<TYPE, 1>(cs, stage_a, r, rows, lane));
                            }
                            u = nr == 4 ? lane >> 3 : (nr == 2 ? lane >> 4 : 0);
                            holder = nr == 4 ? (lane & 7) == 0 : (nr == 2 ? (lane & 15) == 0 : lane == 0);
                            step = nr;
                        } else if (rows - r > 2) {
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
            pa.trace[(size_t)pa.nops * gridDim.x * 4 + (size_t)cta * 4 + 3] = ring_miss;
        }
        // ---- row-split plans: ops marked EXPORT leave their COMPLETE vector (all ranks' slices) in the local plain dst ----
        for (int e = 0; e < pa.nexports; e++) {
            const ExportDesc x = pa.exports[e];
            const uint32_t tag = (epoch << 10) | (uint32_t)x.op;
            const uint2 *ll = reinterpret_cast<const uint2 *>(arena_local) + x.ll;
            for (int i = cta * kCT + (int)threadIdx.x; i < x.m_total; i += gridDim.x * kCT) {
                uint32_t v, t;
                do {
                    asm volatile("ld.volatile.global.v2.u32 {%0,%1}, [%2];" : "=r"(v), "=r"(t) : "l"(ll + i));
                } while (t != tag);
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
static plan_kernel_fn plan_kernel_for(int type, int mode) {
    if (mode & 16) {        // experimental: no k-split teams
        if (type == B200_TYPE_Q4_0) return mode == 20 ? plan_kernel<B200_TYPE_Q4_0, 20> : mode == 24 ? plan_kernel<B200_TYPE_Q4_0, 24> : plan_kernel<B200_TYPE_Q4_0, 16>;
        return mode == 20 ? plan_kernel<B200_TYPE_Q8_0, 20> : mode == 24 ? plan_kernel<B200_TYPE_Q8_0, 24> : plan_kernel<B200_TYPE_Q8_0, 16>;
    }
    if (type == B200_TYPE_Q4_0)
        return mode == 1 ? plan_kernel<B200_TYPE_Q4_0, 1> : mode == 2 ? plan_kernel<B200_TYPE_Q4_0, 2> : mode == 4 ? plan_kernel<B200_TYPE_Q4_0, 4> :
               mode == 8 ? plan_kernel<B200_TYPE_Q4_0, 8> : plan_kernel<B200_TYPE_Q4_0, 0>;
    return mode == 1 ? plan_kernel<B200_TYPE_Q8_0, 1> : mode == 2 ? plan_kernel<B200_TYPE_Q8_0, 2> : mode == 4 ? plan_kernel<B200_TYPE_Q8_0, 4> :
           mode == 8 ? plan_kernel<B200_TYPE_Q8_0, 8> : plan_kernel<B200_TYPE_Q8_0, 0>;
}

struct b200_plan {
    int type, nops, grid, world, rank;
    PlanGeom geom;
    PDesc *pdesc_dev;
    CDesc *cdesc_dev;
    ExportDesc *exports_dev;
    int *pll_dev;               // ring-fed src1 vectors (null: none)
    PubDesc *pub_dev;           // MODE 4 publisher work list
    int mode;                   // kernel variant: 0 = plain; bit 0 ring-fed src1, bit 1 per-slot L2 prefetch, bit 2 / 3 publisher warp
    uint32_t *pubcnt_dev;       // MODE 8 arrival counters
    void *arena_own;            // allocated here when world == 1
    uint32_t *state_dev;
    unsigned long long *trace_dev;
    PlanArgs args;
    size_t arena_bytes;
};

// Smallest k whose in-plan src1 is quantized once per GPU (kernel MODE 4).  Default 8192 on one GPU (measured: GPT-J fc_out,
// k = 16384, 722 vs 741 us per token; at k = 4096 the second exchange costs more than it saves: 811), off for row-split plans
// (not measured there yet); B200_PLAN_LLQ=k overrides, 0 = off.
static int plan_llq_min_k(int world) {
    const char *e = getenv("B200_PLAN_LLQ");
    const int v = e ? atoi(e) : (world == 1 ? 8192 : 0);
    return v > 0 ? v : 0;
}

static size_t plan_arena_elems(const b200_mul_mat_args *args, int count, const b200_plan_split *split, std::vector<long long> *offs,
                               std::vector<long long> *llq_offs = nullptr) {
    size_t total = 0;
    for (int i = 0; i < count; i++) {
        const long long mt = split && split->m_total ? split->m_total[i] : args[i].ne01;
        if (offs) offs->push_back((long long)total);
        total += (size_t)((mt + 15) / 16 * 16);      // 128-byte aligned LL vectors
    }
    // quantized src1 vectors (kernel MODE 4): 10 tagged words per block of 32, room for every op that is long enough
    const int mink = plan_llq_min_k(split ? split->world : 1);
    for (int i = 0; i < count; i++) {
        const bool room = mink > 0 && args[i].ne00 >= mink;
        if (llq_offs) llq_offs->push_back(room ? (long long)total : -1);
        if (room) total += (size_t)((args[i].ne00 / B200_QK * 10 + 15) / 16 * 16);
    }
    return total;
}

static bool ranges_overlap(const void *a, size_t an, const void *b, size_t bn) {
    const uintptr_t a0 = (uintptr_t)a, b0 = (uintptr_t)b;
    return a0 < b0 + bn && b0 < a0 + an;
}

// Host-only part of b200_plan_create (no CUDA call, so it is testable without a device): shapes, and ggml's dataflow between
// the ops.  src_op[i] = index of the op whose dst is op i's src1, or -1 for a vector from outside the plan.
static int plan_analyze(b200_ctx *ctx, const b200_mul_mat_args *args, int count, const b200_plan_split *split, std::vector<int> *src_op,
                        std::vector<int> *m_total_out) {
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
    // Hazards sequential execution would hide but dataflow execution does not: plain dst vectors that alias each other or an
    // outside input (buffer reuse by a graph allocator).  Those graphs stay on the node-by-node path.
    for (int i = 0; i < count; i++) {
        if (!args[i].dst_dev) continue;
        const size_t di = (size_t)m_total[i] * 4;
        for (int j = 0; j < count; j++) {
            if (j > i && args[j].dst_dev && ranges_overlap(args[i].dst_dev, di, args[j].dst_dev, (size_t)m_total[j] * 4)) {
                b200_set_error(ctx, "b200_plan: dst of op %d aliases dst of op %d", i, j);
                return B200_ERR_UNSUPPORTED;
            }
            if (so[j] < 0 && ranges_overlap(args[i].dst_dev, di, args[j].src1_dev, (size_t)args[j].ne00 * 4)) {
                b200_set_error(ctx, "b200_plan: dst of op %d aliases the outside input of op %d", i, j);
                return B200_ERR_UNSUPPORTED;
            }
        }
    }
    return B200_OK;
}

static int plan_llq_dist() {        // B200_PLAN_LLQ_DIST: the producing op must lie at least this many ops back (default 2)
    const char *e = getenv("B200_PLAN_LLQ_DIST");
    const int v = e ? atoi(e) : 2;
    return v >= 1 ? v : 2;
}

// Host-only: which ops take their src1 from the publisher warps (quantized once per GPU, kernel MODE 4 / 8)?  A vector produced
// inside the plan, long enough (B200_PLAN_LLQ), not shared with the previous op (that one already brought it in), produced
// far enough back (behind a producer that has just finished the publication would be a second exchange on the critical
// path), and small enough per CTA: every CTA quantizes 1/grid of the blocks in one warp's run of at most 16 blocks.
static std::vector<char> plan_published(const b200_mul_mat_args *args, int count, const b200_plan_split *split, const std::vector<int> &src_op, int grid) {
    std::vector<char> pub((size_t)count, 0);
    const int mink = plan_llq_min_k(split ? split->world : 1), dist = plan_llq_dist();
    if (mink <= 0 || grid <= 0) return pub;
    for (int i = 0; i < count; i++) {
        const int64_t k = args[i].ne00, nb = k / B200_QK;
        const bool same_input = i > 0 && k == args[i - 1].ne00 && src_op[(size_t)i] == src_op[(size_t)i - 1] &&
                                (src_op[(size_t)i] >= 0 || args[i].src1_dev == args[i - 1].src1_dev);
        pub[(size_t)i] = src_op[(size_t)i] >= 0 && !same_input && k >= mink && (nb + grid - 1) / grid + 1 <= 16 && i - src_op[(size_t)i] >= dist;
    }
    return pub;
}

extern "C" {

int b200_plan_published(const b200_mul_mat_args *args, int count, const b200_plan_split *split, int sm_count, int32_t *published_out) {
    std::vector<int> so, mt;
    const int rc = plan_analyze(NULL, args, count, split, &so, &mt);
    if (rc != B200_OK) return rc;
    if (sm_count <= 0 || !published_out) return B200_ERR_INVALID;
    const std::vector<char> pub = plan_published(args, count, split, so, sm_count);
    for (int i = 0; i < count; i++) published_out[i] = pub[(size_t)i];
    return B200_OK;
}

size_t b200_plan_arena_bytes(const b200_mul_mat_args *args, int count, const b200_plan_split *split) {
    if (!args || count <= 0) return 0;
    return plan_arena_elems(args, count, split, nullptr) * 8;
}

int b200_plan_analyze(const b200_mul_mat_args *args, int count, const b200_plan_split *split, int32_t *src_op_out) {
    std::vector<int> so, mt;
    const int rc = plan_analyze(NULL, args, count, split, &so, &mt);
    if (rc == B200_OK && src_op_out)
        for (int i = 0; i < count; i++) src_op_out[i] = so[(size_t)i];
    return rc;
}

void b200_plan_destroy(b200_plan *p) {
    if (!p) return;
    if (p->pdesc_dev) cudaFree(p->pdesc_dev);
    if (p->cdesc_dev) cudaFree(p->cdesc_dev);
    if (p->exports_dev) cudaFree(p->exports_dev);
    if (p->pll_dev) cudaFree(p->pll_dev);
    if (p->pub_dev) cudaFree(p->pub_dev);
    if (p->pubcnt_dev) cudaFree(p->pubcnt_dev);
    if (p->arena_own) cudaFree(p->arena_own);
    if (p->state_dev) cudaFree(p->state_dev);
    if (p->trace_dev) cudaFree(p->trace_dev);
    free(p);
}

int b200_plan_create(b200_ctx *ctx, const b200_mul_mat_args *args, int count, const b200_plan_split *split, b200_plan **out) {
    B200_REQUIRE(ctx, ctx && args && out && count >= 1, B200_ERR_INVALID);
    *out = NULL;
    std::vector<int> src_op, m_total;
    {
        const int rc = plan_analyze(ctx, args, count, split, &src_op, &m_total);     // shapes, dataflow, aliasing hazards
        if (rc != B200_OK) return rc;
    }
    const int world = split ? split->world : 1, rank = split ? split->rank : 0;
    if (split)
        for (int r = 0; r < world; r++) B200_REQUIRE(ctx, split->peer_arena[r] != NULL, B200_ERR_INVALID);
    const int type = args[0].type;
    B200_REQUIRE(ctx, type == B200_TYPE_Q4_0 || type == B200_TYPE_Q8_0, B200_ERR_UNSUPPORTED);
    B200_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int grid = ctx->sm_count;
    const int qsb = b200_qs_bytes(type);
    // ring slot = `slot_rows` rows of a 4096-wide k-segment (8 -> 18432 B for Q4_0 = one row of k = 32768).  The producer
    // thread needs ~250 ns per slot (measured: 4-row slots cap the stream at 5.4 TB/s), so slots must not be small.
    int slot_rows = 8;
    if (const char *e = getenv("B200_PLAN_SLOT_ROWS")) { const int v = atoi(e); if (v >= 8 && v <= 16) slot_rows = v; }
    const int slot_bytes = slot_rows * 128 * (qsb + 2);

    std::vector<long long> ll_off, llq_off;
    const size_t arena_elems = plan_arena_elems(args, count, split, &ll_off, &llq_off);
    std::vector<PubDesc> pubs;
    // experimental protocol (MODE 8); not combinable with the ring-fed experiment, which needs ll_src as it is
    const bool pubq = getenv("B200_PLAN_PUBQ") && atoi(getenv("B200_PLAN_PUBQ")) != 0 && !getenv("B200_PLAN_LL_RING");
    const std::vector<char> published = plan_published(args, count, split, src_op, grid);
    B200_REQUIRE(ctx, arena_elems < (1ull << 30), B200_ERR_UNSUPPORTED);
    std::vector<PDesc> pd((size_t)count);
    std::vector<CDesc> cd((size_t)count);
    std::vector<ExportDesc> ex;
    int kmax = 0;
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
        c.dst_plain = a->dst_dev;
        c.row0 = (int)row0;
        c.ll_dst = (int)ll_off[i];
        if (a->flags & B200_MM_EXPORT) {
            c.flags |= OPF_EXPORT;
            if (world > 1 && a->dst_dev) {
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
        if (c.k > kmax) kmax = c.k;
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
        // a long in-plan src1 that the publisher warps quantize once per GPU (plan_published)
        if (published[(size_t)i] && llq_off[i] >= 0 && !(c.flags & OPF_SAME_INPUT)) {
            c.flags |= OPF_SRC_LLQ;
            c.src_pub = llq_off[i];
            PubDesc pb;
            memset(&pb, 0, sizeof(pb));
            pb.ll_src = c.ll_src; pb.k = c.k; pb.src_op = c.src_op; pb.llq = (int)llq_off[i]; pb.op = i;
            if (pubq) c.ll_src = (int)pubs.size();      // MODE 8 consumers need the work-list index, not the fp32 vector
            pubs.push_back(pb);
        }
    }
    b200_plan *p = (b200_plan *)calloc(1, sizeof(b200_plan));
    if (!p) return B200_ERR_ALLOC;
    p->type = type; p->nops = count; p->grid = grid; p->world = world; p->rank = rank;
    p->arena_bytes = arena_elems * 8;
    // shared-memory geometry
    PlanGeom &g = p->geom;
    g.slot_bytes = slot_bytes;
    g.l2_ahead = 0;
    if (const char *e = getenv("B200_PLAN_L2_AHEAD")) { const int v = atoi(e); if (v >= 0 && v <= 31) g.l2_ahead = v; }
    g.l2_slots = 0;
    if (const char *e = getenv("B200_PLAN_L2_SLOTS")) { const int v = atoi(e); if (v >= 0 && v <= 256) g.l2_slots = v; }
    const int act_bytes = (int)b200_align_up((size_t)kmax + (size_t)(kmax / 32) * 8, 128);   // int8 planes + fp32 scales + 8 * sums
    const int ll_bytes = kCW * 2048 > kPartFloats * 4 ? kCW * 2048 : kPartFloats * 4;       // LL staging, aliased by the k-split partials
    const int desc_bytes = kDescCap * (int)sizeof(CDesc);
    const int bar_bytes = 2 * kMaxSlots * 8 + 64;
    const int max_smem = 227 * 1024;
    const int fixed = 2 * act_bytes + ll_bytes + desc_bytes + bar_bytes + (pubs.empty() ? 0 : 2048);
    // ring: as many slots as fit, a multiple of the consumer warps when possible (every warp then owns the same number)
    int nslots = (max_smem - fixed) / g.slot_bytes;
    if (nslots > kMaxSlots) nslots = kMaxSlots;
    if (nslots >= 8) nslots = nslots / 8 * 8;
    if (const char *e = getenv("B200_PLAN_SLOTS")) { const int v = atoi(e); if (v >= 2 && v <= kMaxSlots && v * g.slot_bytes <= max_smem - fixed) nslots = v; }
    if (nslots < 2) { free(p); b200_set_error(ctx, "b200_plan_create: k = %d leaves no room for the weight ring", kmax); return B200_ERR_UNSUPPORTED; }
    g.nslots = nslots;
    // Ring-fed src1 (B200_PLAN_LL_RING = n > 0, experimental): a src1 produced by an op that lies at least n ring slots (per
    // CTA) back in the stream is complete on every CTA long before the producer thread reaches its consumer, so the
    // producer copies the tagged vector through the ring like weights and the consumers skip the L2 round trips.  A copy
    // taken too early is caught by its tags (the consumers then fetch from L2 as usual).
    std::vector<int> pll((size_t)count, -1);
    bool any_ring = false;
    if (const char *e = getenv("B200_PLAN_LL_RING")) {
        // n > 0: at least n slots in between; -1: every src1 produced two or more ops back; -2: every src1 produced in the plan
        // (the negative settings exist for the tests: copies taken too early must be caught by their tags)
        const int min_slots = atoi(e);
        for (int i = 0; min_slots != 0 && i < count; i++) {
            CDesc &c = cd[i];
            if (c.src_op < 0 || (c.flags & OPF_SAME_INPUT)) continue;
            const int nls = (c.k * 8 + g.slot_bytes - 1) / g.slot_bytes;
            if (nls > nslots) continue;
            long long between = 0;
            for (int j = c.src_op + 1; j < i; j++) {
                between += cd[j].rows_q / pd[j].rs;
                if (pll[j] >= 0) between += (cd[j].k * 8 + g.slot_bytes - 1) / g.slot_bytes;
            }
            if (min_slots > 0 && between < min_slots) continue;
            if (min_slots == -1 && c.src_op > i - 2) continue;
            pll[i] = c.ll_src;
            c.flags |= OPF_SRC_RING;
            any_ring = true;
        }
    }
    g.ring_off = 0;
    g.act_off = nslots * g.slot_bytes;
    g.act_stride = act_bytes;
    g.ll_off = g.act_off + 2 * act_bytes;
    g.desc_off = g.ll_off + ll_bytes;
    g.bar_off = g.desc_off + desc_bytes;
    g.total = g.bar_off + bar_bytes;
    g.pub_off = g.total;
    if (!pubs.empty()) g.total += 2048;          // the publisher warp's staging (MODE 4)

    cudaError_t e = cudaMalloc((void **)&p->pdesc_dev, sizeof(PDesc) * (size_t)count);
    if (e == cudaSuccess) e = cudaMalloc((void **)&p->cdesc_dev, sizeof(CDesc) * (size_t)count);
    if (e == cudaSuccess && !ex.empty()) e = cudaMalloc((void **)&p->exports_dev, sizeof(ExportDesc) * ex.size());
    if (e == cudaSuccess && any_ring) e = cudaMalloc((void **)&p->pll_dev, sizeof(int) * (size_t)count);
    if (e == cudaSuccess && any_ring) e = cudaMemcpy(p->pll_dev, pll.data(), sizeof(int) * (size_t)count, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc((void **)&p->state_dev, 16);
    if (e == cudaSuccess && world == 1) e = cudaMalloc(&p->arena_own, p->arena_bytes);
    if (e == cudaSuccess && getenv("B200_PLAN_TRACE")) {
        e = cudaMalloc((void **)&p->trace_dev, (size_t)(count + 1) * grid * 4 * 8);
        if (e == cudaSuccess) e = cudaMemset(p->trace_dev, 0, (size_t)(count + 1) * grid * 4 * 8);
    }
    if (e == cudaSuccess) e = cudaMemcpy(p->pdesc_dev, pd.data(), sizeof(PDesc) * (size_t)count, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(p->cdesc_dev, cd.data(), sizeof(CDesc) * (size_t)count, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && !ex.empty()) e = cudaMemcpy(p->exports_dev, ex.data(), sizeof(ExportDesc) * ex.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(p->state_dev, 0, 16);
    if (e == cudaSuccess && world == 1) e = cudaMemset(p->arena_own, 0, p->arena_bytes);
    const bool any_llq = !any_ring && !pubs.empty();
    p->mode = any_ring ? 1 : any_llq ? (pubq ? 8 : 4) : (g.l2_slots > 0 ? 2 : 0);
    if (getenv("B200_PLAN_NOSPLIT") && atoi(getenv("B200_PLAN_NOSPLIT")) != 0 && (p->mode == 0 || p->mode == 4 || p->mode == 8)) p->mode |= 16;
    if (e == cudaSuccess && p->mode == 8) e = cudaMalloc((void **)&p->pubcnt_dev, sizeof(uint32_t) * pubs.size());
    if (e == cudaSuccess && p->mode == 8) e = cudaMemset(p->pubcnt_dev, 0, sizeof(uint32_t) * pubs.size());
    if (e == cudaSuccess && any_llq) e = cudaMalloc((void **)&p->pub_dev, sizeof(PubDesc) * pubs.size());
    if (e == cudaSuccess && any_llq) e = cudaMemcpy(p->pub_dev, pubs.data(), sizeof(PubDesc) * pubs.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(plan_kernel_for(type, p->mode), cudaFuncAttributeMaxDynamicSharedMemorySize, g.total);
    if (e != cudaSuccess) {
        b200_set_error(ctx, "b200_plan_create: %s", cudaGetErrorString(e));
        (void)cudaGetLastError();
        b200_plan_destroy(p);
        return e == cudaErrorMemoryAllocation ? B200_ERR_ALLOC : B200_ERR_CUDA;
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
    pa.p_ll = p->pll_dev;
    pa.pub = p->pub_dev;
    pa.npub = p->pub_dev ? (int)pubs.size() : 0;
    pa.pub_count = p->pubcnt_dev;
    *out = p;
    return B200_OK;
}

int b200_plan_launch(b200_ctx *ctx, b200_plan *p) {
    B200_REQUIRE(ctx, ctx && p, B200_ERR_INVALID);
    B200_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    plan_kernel_for(p->type, p->mode)<<<p->grid, plan_threads(p->mode), p->geom.total, ctx->stream>>>(p->args, p->geom);
    ctx->launches++;
    B200_CUDA_TRY(ctx, cudaGetLastError());
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
