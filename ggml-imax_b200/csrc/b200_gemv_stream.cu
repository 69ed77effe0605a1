// b200_gemv_stream.cu -- the decode GEMV as a streaming kernel: dst[m, n<=8] = W[m,k] x X[k,n], 2-D case.
//
// Same arithmetic as b200_gemv.cu (fused quantize_row_q8_0 of the activations, src/ggml-quants.c:535-618;
// per-block exact int32 dp4a dots scaled by d_w*d_x and accumulated in fp32, src/ggml-quants.c:3858-3869 /
// :5010-5015) but organised around what bounds a chain of microsecond-sized mul_mats on B200: keeping HBM
// busy ACROSS kernel boundaries.
//
//  * one persistent CTA per SM; CTA c owns a contiguous run of weight rows, i.e. one contiguous byte range of
//    the qs plane and one of the fp16 scale plane (the repacked layout makes both 16-byte aligned);
//  * one elected thread streams that range into a shared-memory ring with 1-D bulk async copies
//    (cp.async.bulk ... mbarrier::complete_tx, SASS UBLKCP) -- no registers are tied up by loads in flight,
//    ~100 KB per SM can be in flight;
//  * programmatic dependent launch: the kernel is co-resident with its predecessor (2 x ~100 KB of shared memory
//    per SM), fills its ring BEFORE griddepcontrol.wait because weights never depend on the previous mul_mat,
//    and only then reads the activations.  HBM therefore keeps streaming the next matrix while the current one
//    is being quantized/reduced/stored;
//  * n == 1: every lane keeps the int8 activations of "its" blocks (lane + 32 i) in registers for the whole
//    kernel, so a weight block costs 16 B + 2 B of shared-memory reads; long rows (k > 4096) are split across
//    2/4/8 warps by k-segment and combined through shared memory in a fixed order;
//  * n in 2..8: activations are read from shared memory per block.
// Shapes outside (k % 256 == 0, k <= 32768, no batch dims) use the generic kernel in b200_gemv.cu.
#include "b200_stream_common.cuh"

namespace {

using namespace b200s;

constexpr int kConsumerWarps = 8;
constexpr int kConsumerThreads = kConsumerWarps * 32;
constexpr int kThreads = kConsumerThreads;       // lane 0 of warp 0 doubles as the bulk-copy producer (8 warps -> 128 regs at 2 CTAs/SM)
constexpr int kMaxStages = 8;
constexpr int kSegBlocks = 128;                  // blocks of k handled by one warp (4 per lane)

__device__ __forceinline__ void consumer_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory"); }
__device__ __forceinline__ void stamp(unsigned long long *trace, int slot) {
    if (trace) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        trace[(size_t)blockIdx.x * B200_TRACE_STAMPS + slot] = t;
    }
}

struct StreamGeom {
    int rs;           // rows per stage
    int pr;           // rows whose k-split partials are parked before one combine
    int stages;       // ring depth
    int g;            // warps per row (k-segments)
    int stage_qs;     // bytes of qs per full stage
    int stage_sc;     // bytes of scales per full stage
    int stage_bytes;  // aligned total
    int act_col;      // bytes of activation scratch per column
    int ring_off, act_off, part_off, bar_off, llstage_off, total;
    // division-free bookkeeping, precomputed on the host
    int log2g;        // g == 1 << log2g
    int ppst;         // passes per stage = rs / (8 / g)
};

// Up to 4 independent decode mul_mats that share src1 (same k, same activations: q/k/v/fc_in of a transformer block) run
// as ONE launch: the CTAs are divided among the matrices in proportion to their bytes, every CTA works on one matrix.
constexpr int kMaxBatch = 4;
struct StreamBatch {
    int count;
    int total_ctas;            // host side only: gridDim.x
    struct Sub {
        const uint8_t *qs;
        const __half *d;
        float *dst;
        int m;
        int cta0;              // first CTA of this matrix
        int rows_q, rows_rem;  // CTA c (relative) owns rows [c*rows_q + min(c, rows_rem), +rows_q + (c < rows_rem))
        // fused all-gather only: where this matrix's LL vector sits relative to gather.peer_dst[r], its first global row,
        // its launch slot (tag + execution count) and how many CTAs serve it
        long long ll_off;
        long long row0;
        int slot, ctas;
    } sub[kMaxBatch];
};

__host__ __device__ inline size_t stream_act_col_bytes(int type, int k) {
    const size_t nb = (size_t)(k >> 5);
    const size_t raw = (size_t)k + nb * 4 + (type == B200_TYPE_Q4_0 ? nb * 4 : 0);
    return (raw + 15) & ~(size_t)15;
}



// one CTA: wait for an LL vector to be complete and write it out as plain fp32
__global__ void __launch_bounds__(1024) gather_finish_kernel(const b200_gather gd, const uint2 *ll, float *out, int64_t count, const LLWait lw) {
    __shared__ uint32_t s_epoch;
    uint32_t *st = gd.state + 2 * (size_t)gd.slot;
    if (threadIdx.x == 0) s_epoch = st[1];
    __syncthreads();
    const uint32_t tag = ((s_epoch + 1u) << 10) | (uint32_t)gd.wait_slot;   // what the producing slot stamped this time round
    unsigned polls = 0;
    unsigned long long t0 = 0;
    for (int64_t i = threadIdx.x; i < count; i += blockDim.x) {
        uint32_t v, t;
        for (;;) {
            asm volatile("ld.volatile.global.v2.u32 {%0,%1}, [%2];" : "=r"(v), "=r"(t) : "l"(ll + i));
            if (t == tag || ll_wait_expired(lw, polls, t0, 6u | (tag << 8))) break;
        }
        out[i] = __uint_as_float(v);
    }
    __syncthreads();
    if (threadIdx.x == 0) st[1] = s_epoch + 1u;
}

template <int TYPE, int NCOLS, bool DOTS>
__global__ void __launch_bounds__(kThreads, 2) gemv_stream_kernel(const b200_gemv_params p, const StreamGeom g, const b200_gather gd, const StreamBatch sb) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int QSB = TYPE == B200_TYPE_Q4_0 ? 16 : 32;
    const int k = (int)p.k, nb = k >> 5;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row_qs = nb * QSB, row_sc = nb * 2;
    const LLWait lwait = {p.abort_dev, p.abort_host, p.wait_timeout_ns};

    unsigned char *ring = smem + g.ring_off;
    unsigned char *act = smem + g.act_off;
    float *part = reinterpret_cast<float *>(smem + g.part_off);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + g.bar_off);
    uint64_t *empty_bar = full_bar + kMaxStages;

    // this CTA's contiguous run of rows
    // which matrix of the batch this CTA serves (count == 1: the plain single-matrix launch)
    int si = 0;
#pragma unroll
    for (int j = 1; j < kMaxBatch; j++)
        if (j < sb.count && (int)blockIdx.x >= sb.sub[j].cta0) si = j;
    // (selected with compile-time indices: a runtime index into kernel parameters would force a local-memory copy)
    const uint8_t *m_qs = sb.sub[0].qs;
    const __half *m_d = sb.sub[0].d;
    float *m_dst = sb.sub[0].dst;
    int m_rows32 = sb.sub[0].m, m_cta0 = sb.sub[0].cta0, m_rows_q = sb.sub[0].rows_q, m_rows_rem = sb.sub[0].rows_rem;
    long long m_ll0 = sb.sub[0].ll_off + sb.sub[0].row0;
    int m_slot = sb.sub[0].slot, m_ctas = sb.sub[0].ctas;
#pragma unroll
    for (int j = 1; j < kMaxBatch; j++)
        if (j == si) {
            m_qs = sb.sub[j].qs; m_d = sb.sub[j].d; m_dst = sb.sub[j].dst;
            m_rows32 = sb.sub[j].m; m_cta0 = sb.sub[j].cta0; m_rows_q = sb.sub[j].rows_q; m_rows_rem = sb.sub[j].rows_rem;
            m_ll0 = sb.sub[j].ll_off + sb.sub[j].row0; m_slot = sb.sub[j].slot; m_ctas = sb.sub[j].ctas;
        }
    const int64_t m_rows = m_rows32;
    const int bx = (int)blockIdx.x - m_cta0;
    const int64_t r_begin = (int64_t)bx * m_rows_q + min(bx, m_rows_rem);
    const int nrows = m_rows_q + (bx < m_rows_rem ? 1 : 0);
    int nstage_iters = 0;
    for (int r = 0; r < nrows; r += g.rs) nstage_iters++;   // ceil(nrows / rs) without a division (a handful of iterations)
    // which (row-in-pass, k-segment) this warp serves -- all shifts, computed before the dependency wait
    const int G = g.g;
    const int seg = warp & (G - 1);
    const int row_in_pass = warp >> g.log2g;
    const int rows_per_pass = kConsumerWarps >> g.log2g;
    const int b0 = seg * kSegBlocks;  // first block of the segment

    if (threadIdx.x == 0) {
        for (int s = 0; s < g.stages; s++) {
            mbar_init(&full_bar[s], 1);
            reinterpret_cast<int *>(empty_bar)[s] = 0;   // per-stage "warps that have left" counter
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    pdl_launch_dependents();
    if (threadIdx.x == 0) stamp(p.trace, 0);

    // ===== streaming this CTA's byte ranges into the ring =====
    // issue_stage(f, slot): bulk-copy stage use f (rows f*rs ..) into ring slot `slot`.  The first `stages` uses are issued
    // by thread 0 right here, BEFORE waiting for the previous grid.  Afterwards the ring refills itself: every warp counts
    // itself out of a stage (shared-memory counter), and the warp that leaves last re-issues that slot for use f + stages.
    const uint8_t *gq0 = m_qs + r_begin * row_qs;
    const uint8_t *gs0 = reinterpret_cast<const uint8_t *>(m_d) + r_begin * row_sc;
    int *stage_cnt = reinterpret_cast<int *>(empty_bar);   // [kMaxStages] ints, reusing the (now unused) empty-barrier words
    auto issue_stage = [&](int f, int slot) {
        const int rows = min(g.rs, nrows - f * g.rs);
        unsigned char *dst = ring + (size_t)slot * g.stage_bytes;
        mbar_expect_tx(&full_bar[slot], (uint32_t)(rows * (row_qs + row_sc)));
        bulk_g2s(dst, gq0 + (size_t)f * g.rs * row_qs, (uint32_t)(rows * row_qs), &full_bar[slot]);
        bulk_g2s(dst + g.stage_qs, gs0 + (size_t)f * g.rs * row_sc, (uint32_t)(rows * row_sc), &full_bar[slot]);
    };
    if (threadIdx.x == 0) {
        const int n0 = min(nstage_iters, g.stages);
        for (int f = 0; f < n0; f++) issue_stage(f, f);
    }
    __syncwarp();   // reconverge BEFORE griddepcontrol.wait: a warp parked in the wait takes its diverged lane 0 with it

    // ===== consumers =====
    // activations (and dst, for write-after-read) belong to the previous grid until here -- except in the fused
    // all-gather chain, where src1 arrives as tagged LL elements: the tags ARE the dependency (a matching tag proves the
    // producing CTA, local or remote, is past its own input reads), so the kernel does not wait for grid completion at all
    if (!(gd.world > 0 && gd.wait_slot >= 0)) pdl_wait();
    if (threadIdx.x == 0) stamp(p.trace, 2);
    // fused all-gather: tags are (execution count + 1) << 10 | producing slot -- unique per (launch, replay), so a stale
    // element of the ping-pong buffer (written by slot - 2 in the same replay) can never be mistaken for the new one.
    // Every slot runs once per sequence, so this launch's own execution count is also its producer's.
    uint32_t gd_tag = 0, gd_src_tag = 0;
    if (gd.world > 0) {
        uint32_t *s_tag = reinterpret_cast<uint32_t *>(empty_bar + kMaxStages);
        if (threadIdx.x == 0) *s_tag = gd.state[2 * (size_t)m_slot + 1] + 1u;
        consumer_bar_sync();
        gd_tag = (*s_tag << 10) | (uint32_t)m_slot;
        gd_src_tag = (*s_tag << 10) | (uint32_t)(gd.wait_slot & 1023);
    }
    const bool ll_in = gd.world > 0 && gd.wait_slot >= 0;

    // ---- quantize the activation columns into shared memory: quantize_row_q8_0, bit-exact ----
    // Two lanes per block, 16 consecutive floats (4 x 128-bit loads) per lane: one amax shuffle, ONE 127/amax
    // division per 16 elements, and the lane's 16 int8 are exactly one 16-byte store into the lo/hi plane.
    // All loads of a batch are issued before any use (one L2 round trip per 8192 elements per CTA).
    {
        constexpr int kQB = 4;   // 4 lane-tasks (64 floats) in flight per thread: k = 16384 is one L2 round trip, not two
        const int tpc = nb * 2;   // lane-tasks per column; even, and a lane's partner (lane ^ 1) shares its block
#pragma unroll 1
        for (int c = 0; c < NCOLS; c++) {
            const float *xcol = reinterpret_cast<const float *>(reinterpret_cast<const char *>(p.x) + (size_t)c * p.nb11);
            unsigned char *col = act + (size_t)c * g.act_col;
#pragma unroll 1
            for (int base = 0; base < tpc; base += kConsumerThreads * kQB) {
                float4 v[kQB][4];
                if (ll_in) {
                    // every CTA needs the whole vector, so what matters is its LAST element to land: probe the end of
                    // this warp's span first (cheap, warp-uniform), then do the verified loads
                    const int tlast = min(base + (kQB - 1) * kConsumerThreads + warp * 32 + 31, tpc - 1);
                    ll_probe(reinterpret_cast<const char *>(p.x) + (size_t)tlast * 128 + 120, gd_src_tag, lwait);
                    if (threadIdx.x == 0 && base == 0) stamp(p.trace, 1);
                }
#pragma unroll
                for (int u = 0; u < kQB; u++) {
                    const int t = min(base + u * kConsumerThreads + (int)threadIdx.x, tpc - 1);
                    if (base + u * kConsumerThreads + warp * 32 >= tpc) continue;   // whole warp past the end: nothing to load
                    if (ll_in) {
                        // src1 is the LL vector the previous launch scattered to every rank: 8 bytes per element
                        const int tb = base + u * kConsumerThreads + warp * 32;   // first lane-task of this warp
                        const int nvalid = max(0, min(32, tpc - tb));
                        if (nvalid > 0)
                            ll_load16_warp(reinterpret_cast<const char *>(p.x) + (size_t)tb * 128, nvalid, gd_src_tag,
                                           reinterpret_cast<float *>(smem + g.llstage_off) + warp * 512, lane, v[u], lwait);
                    } else {
                        const float4 *src = reinterpret_cast<const float4 *>(xcol + (size_t)t * 16);
#pragma unroll
                        for (int j = 0; j < 4; j++) v[u][j] = src[j];
                    }
                }
                if (ll_in && threadIdx.x == 0 && base == 0) stamp(p.trace, 6);
#pragma unroll
                for (int u = 0; u < kQB; u++) {
                    if (base + u * kConsumerThreads + warp * 32 >= tpc) continue;   // (warp-uniform, so the shuffles below stay full)
                    const int tt = base + u * kConsumerThreads + (int)threadIdx.x;
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
                        // plane h holds elements 16h..16h+15 of every block -> conflict-free 16-byte reads later
                        *reinterpret_cast<uint4 *>(col + (size_t)h * (k >> 1) + (size_t)b * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        if (h == 0) {
                            reinterpret_cast<float *>(col + k)[b] = __half2float(__float2half_rn(__fdiv_rn(amax, 127.f)));
                            if (TYPE == B200_TYPE_Q4_0) reinterpret_cast<int *>(col + k + (size_t)nb * 4)[b] = 8 * sq;
                        }
                    }
                }
            }
        }
    }
    consumer_bar_sync();
    if (threadIdx.x == 0) stamp(p.trace, 3);


    const uint32_t ring_a = smem_u32(ring), act_a = smem_u32(act);
    const uint32_t full_a = smem_u32(full_bar), empty_a = smem_u32(empty_bar);

    // per-lane loop invariants: the lane's 4 blocks are b0 + lane + 32 i, i.e. fixed byte offsets inside a row plus
    // compile-time multiples of 32 blocks.  Lanes past the end of a short row ("dead", k < 4096 * segments) read
    // whatever follows in shared memory (the allocation is padded) and are predicated off at the accumulate.
    const uint32_t woff0 = (uint32_t)((b0 + lane) * QSB);
    const uint32_t soff0 = (uint32_t)(g.stage_qs + (b0 + lane) * 2);
    bool blive[4];
#pragma unroll
    for (int i = 0; i < 4; i++) blive[i] = b0 + lane + 32 * i < nb;

    // n == 1: the lane's activation blocks live in registers for the whole kernel
    uint4 alo[4], ahi[4];
    float da[4];
    int s8[4];
    if (NCOLS == 1) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t b = (uint32_t)min(b0 + lane + 32 * i, nb - 1);
            alo[i] = lds128(act_a + b * 16);
            ahi[i] = lds128(act_a + (uint32_t)(k >> 1) + b * 16);
            da[i] = lds_f32(act_a + (uint32_t)k + b * 4);
            s8[i] = TYPE == B200_TYPE_Q4_0 ? lds_s32(act_a + (uint32_t)k + (uint32_t)nb * 4 + b * 4) : 0;
        }
    }

    // The warp's rows form a sequence of "slots" (stage, pass).  kU slots are processed as ONE straight-line block:
    // wait for the stages they live in, kU independent row dots, one joint shuffle reduction, release the stages.
    // All stage/pass bookkeeping is incremental (no divisions) and all shared addresses are 32-bit.
    constexpr int kU = NCOLS == 1 ? 4 : (NCOLS <= 2 ? 2 : 1);
    const int ppst = g.ppst;                          // passes per stage
    const int total_slots = nstage_iters * ppst;
    int rows_in_chunk = 0, chunk_row0 = 0, cpar = 0;  // k-split bookkeeping (G > 1, where ppst == 1)
    int it = 0, ps = 0, st = 0;                       // stage use / pass / ring slot of the next slot
    uint32_t par = 0;                                 // parity of the full barrier for `it`
    for (int j0 = 0; j0 < total_slots; j0 += kU) {
        const int nslot = min(kU, total_slots - j0);
        float acc[kU][NCOLS];
        int grow[kU];          // row index relative to r_begin
        int prow[kU];          // row index relative to the k-split chunk
        bool rlive[kU];
        int chunk_rows_added = 0;
        // Slots are processed one after the other -- wait for the slot's stage, 4 independent block dots per lane, count
        // this warp out of the stage when it was its last pass -- so stages are released (and refilled) one at a time and
        // the ring never drains behind a whole group.  Only the shuffle reduction is shared by the kU rows of the group.
#pragma unroll
        for (int u = 0; u < kU; u++) {
            const bool valid = u < nslot;
#pragma unroll
            for (int c = 0; c < NCOLS; c++) acc[u][c] = 0.0f;
            rlive[u] = false;
            grow[u] = 0;
            prow[u] = 0;
            if (valid) {
                if (ps == 0) {
                    mbar_wait_a(full_a + 8u * (uint32_t)st, par);
                    if (j0 == 0 && u == 0 && threadIdx.x == 0) stamp(p.trace, 4);
                }
                const int rows = min(g.rs, nrows - it * g.rs);
                const int r = ps * rows_per_pass + row_in_pass;
                rlive[u] = r < rows;
                const int rr = min(r, rows - 1);
                const uint32_t stage_a = ring_a + (uint32_t)(st * g.stage_bytes);
                const uint32_t wbase = stage_a + (uint32_t)(rr * row_qs) + woff0;
                const uint32_t sbase = stage_a + (uint32_t)(rr * row_sc) + soff0;   // soff0 already includes stage_qs
                grow[u] = it * g.rs + rr;
                prow[u] = rows_in_chunk + chunk_rows_added + rr;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const uint4 w0 = lds128(wbase + (uint32_t)(i * 32 * QSB));
                    uint4 w1 = make_uint4(0, 0, 0, 0);
                    if (TYPE == B200_TYPE_Q8_0) w1 = lds128(wbase + (uint32_t)(i * 32 * QSB + 16));
                    const float dw = lds_h2f(sbase + (uint32_t)(i * 64));
                    if (NCOLS == 1) {
                        const int sumi = block_dot<TYPE>(w0, w1, alo[i], ahi[i], s8[i]);
                        if (DOTS) {
                            if (blive[i] && rlive[u]) p.dots[(r_begin + grow[u]) * nb + b0 + lane + 32 * i] = sumi;
                        } else if (blive[i]) {
                            acc[u][0] = fmaf((float)sumi, dw * da[i], acc[u][0]);
                        }
                    } else {
                        const uint32_t b = (uint32_t)min(b0 + lane + 32 * i, nb - 1);
#pragma unroll
                        for (int c = 0; c < NCOLS; c++) {
                            const uint32_t col = act_a + (uint32_t)(c * g.act_col);
                            const uint4 xlo = lds128(col + b * 16);
                            const uint4 xhi = lds128(col + (uint32_t)(k >> 1) + b * 16);
                            const float dx = lds_f32(col + (uint32_t)k + b * 4);
                            const int sx = TYPE == B200_TYPE_Q4_0 ? lds_s32(col + (uint32_t)k + (uint32_t)nb * 4 + b * 4) : 0;
                            const int sumi = block_dot<TYPE>(w0, w1, xlo, xhi, sx);
                            if (DOTS) {
                                if (blive[i] && rlive[u]) p.dots[((int64_t)c * m_rows + r_begin + grow[u]) * nb + b] = sumi;
                            } else if (blive[i]) {
                                acc[u][c] = fmaf((float)sumi, dw * dx, acc[u][c]);
                            }
                        }
                    }
                }
                if (ps + 1 == ppst) {
                    // last pass over this stage: count the warp out; whoever is last re-arms the slot for use it + stages.
                    // (shared-memory loads and the atomic go through the same in-order LSU path of the warp, so every
                    //  lane's reads of the stage have been performed when the atomic is.)
                    __syncwarp();
                    if (lane == 0) {
                        __threadfence_block();
                        if (atomicAdd(&stage_cnt[st], 1) == kConsumerWarps - 1) {
                            stage_cnt[st] = 0;
                            __threadfence_block();
                            if (it + g.stages < nstage_iters) issue_stage(it + g.stages, st);
                        }
                    }
                    if (G > 1) chunk_rows_added += rows;
                    ps = 0;
                    it++;
                    if (++st == g.stages) { st = 0; par ^= 1u; }
                } else {
                    ps++;
                }
            }
        }
        rows_in_chunk += chunk_rows_added;

        if (!DOTS) {
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1)
#pragma unroll
                for (int u = 0; u < kU; u++)
#pragma unroll
                    for (int c = 0; c < NCOLS; c++) acc[u][c] += __shfl_xor_sync(0xffffffffu, acc[u][c], off);
            // lane (u * NCOLS + c) publishes value (u, c)
            if (lane < kU * NCOLS) {
                float v = 0.0f;
                int gr = 0, pr = 0;
                bool lv = false;
#pragma unroll
                for (int u = 0; u < kU; u++)
#pragma unroll
                    for (int c = 0; c < NCOLS; c++)
                        if (lane == u * NCOLS + c) { v = acc[u][c]; gr = grow[u]; pr = prow[u]; lv = rlive[u]; }
                const int c = lane % NCOLS;
                if (lv) {
                    if (G == 1) {
                        if (gd.world > 0) {
                            // fused all-gather, producer side: the element goes to every rank's full vector (NVLink stores)
#pragma unroll
                            for (int r = 0; r < B200_MAX_RANKS; r++)
                                if (r < gd.world) ll_store(gd.peer_dst[r], m_ll0 + r_begin + gr, v, gd_tag);
                        } else {
                            m_dst[(int64_t)c * m_rows + r_begin + gr] = b200_gemv_epilogue(p, v, r_begin + gr, (int64_t)c * m_rows + r_begin + gr);
                        }
                    }
                    else part[((cpar * g.pr + pr) * kConsumerWarps + seg) * NCOLS + c] = v;   // k-split: park the partial
                }
            }
        }

        if (!DOTS && G > 1) {
            // k-split: partials of up to g.pr rows are parked in shared memory (double-buffered by chunk parity) and
            // combined in segment order after ONE barrier per chunk -- for decode shapes that is once per kernel.
            if (j0 + kU >= total_slots || rows_in_chunk + kU * g.rs > g.pr) {
                consumer_bar_sync();
                for (int t = threadIdx.x; t < rows_in_chunk * NCOLS; t += kConsumerThreads) {
                    const int r = t / NCOLS, c = t - r * NCOLS;
                    float v = 0.0f;
                    for (int sg = 0; sg < G; sg++) v += part[((cpar * g.pr + r) * kConsumerWarps + sg) * NCOLS + c];
                    if (gd.world > 0) {
#pragma unroll
                        for (int pr = 0; pr < B200_MAX_RANKS; pr++)
                            if (pr < gd.world) ll_store(gd.peer_dst[pr], m_ll0 + r_begin + chunk_row0 + r, v, gd_tag);
                    } else {
                        m_dst[(int64_t)c * m_rows + r_begin + chunk_row0 + r] =
                            b200_gemv_epilogue(p, v, r_begin + chunk_row0 + r, (int64_t)c * m_rows + r_begin + chunk_row0 + r);
                    }
                }
                chunk_row0 += rows_in_chunk;
                rows_in_chunk = 0;
                cpar ^= 1;
            }
        }
    }
    if (threadIdx.x == 0) stamp(p.trace, 5);
    if (gd.world > 0) {
        // bookkeeping only (device-scope): the last CTA of the grid bumps the slot's execution count
        consumer_bar_sync();
        if (threadIdx.x == 0) {
            uint32_t *st = gd.state + 2 * (size_t)m_slot;
            const uint32_t arrived = atomicAdd(st, 1u);
            if (arrived == (uint32_t)m_ctas - 1u) {
                st[0] = 0;
                st[1] = gd_tag >> 10;
            }
        }
    }
}

bool stream_geometry(const b200_gemv_params &p, StreamGeom *g) {
    const int k = (int)p.k, nb = k >> 5;
    const int qsb = p.type == B200_TYPE_Q4_0 ? 16 : 32;
    const int row_qs = nb * qsb, row_sc = nb * 2, row_bytes = row_qs + row_sc;
    int G = (nb + kSegBlocks - 1) / kSegBlocks;
    if (G > 8) return false;
    if (G == 3) G = 4;
    if (G > 4 && G < 8) G = 8;
    const int rows_per_pass = kConsumerWarps / G;
    int rs = (16 * 1024) / row_bytes;
    if (rs < rows_per_pass) rs = rows_per_pass;
    if (rs > 64) rs = 64;
    rs = rs / rows_per_pass * rows_per_pass;
    g->rs = rs;
    g->g = G;
    g->log2g = G == 1 ? 0 : (G == 2 ? 1 : (G == 4 ? 2 : 3));
    g->ppst = rs / rows_per_pass;
    g->stage_qs = rs * row_qs;
    g->stage_sc = rs * row_sc;
    g->stage_bytes = (int)b200_align_up((size_t)g->stage_qs + g->stage_sc, 128);
    g->act_col = (int)stream_act_col_bytes(p.type, k);
    const int act_bytes = (int)b200_align_up((size_t)g->act_col * p.n, 128);
    int pr = 64 / (int)p.n;
    if (pr < 4 * rs) pr = 4 * rs;   // at least one group of kU (<= 4) stages
    g->pr = pr;
    const int part_bytes = G > 1 ? (int)b200_align_up((size_t)2 * pr * kConsumerWarps * p.n * 4, 128) : 128;
    const int bar_bytes = 2 * kMaxStages * 8 + 16;   // + the broadcast slot of the gather tag
    // two of these kernels must be co-resident per SM (current + programmatic dependent): <= ~110 KB each
    const int ll_bytes = p.gather ? kConsumerWarps * 2048 : 0;   // per-warp staging of the fused all-gather's LL loads
    const int budget = 110 * 1024 - act_bytes - part_bytes - bar_bytes - ll_bytes - 8192 - 256;
    int stages = budget / g->stage_bytes;
    if (stages < 2) return false;
    if (stages > kMaxStages) stages = kMaxStages;
    g->stages = stages;
    g->ring_off = 0;
    g->act_off = stages * g->stage_bytes;
    g->part_off = g->act_off + act_bytes;
    g->bar_off = g->part_off + part_bytes;
    g->llstage_off = g->bar_off + bar_bytes;
    g->total = g->llstage_off + ll_bytes + 8192;   // slack for the unclamped reads of dead lanes
    return true;
}

template <int TYPE, int NCOLS>
int launch_stream_typed(b200_ctx *ctx, const b200_gemv_params &p, const StreamGeom &g, bool dots, const StreamBatch *batch = nullptr) {
    auto kern = dots ? gemv_stream_kernel<TYPE, NCOLS, true> : gemv_stream_kernel<TYPE, NCOLS, false>;
    if (dots) B200_SMEM_LIMIT_ONCE(ctx, (gemv_stream_kernel<TYPE, NCOLS, true>), 112 * 1024);
    else B200_SMEM_LIMIT_ONCE(ctx, (gemv_stream_kernel<TYPE, NCOLS, false>), 112 * 1024);
    StreamGeom gg = g;
    StreamBatch sb;
    memset(&sb, 0, sizeof(sb));
    int64_t ctas;
    if (batch) {
        sb = *batch;
        ctas = sb.total_ctas;
    } else {
        ctas = ctx->sm_count;
        if (ctas > p.m) ctas = p.m;
        sb.count = 1;
        sb.sub[0].qs = p.qs;
        sb.sub[0].d = p.d;
        sb.sub[0].dst = p.dst;
        sb.sub[0].m = (int)p.m;
        sb.sub[0].cta0 = 0;
        sb.sub[0].rows_q = (int)(p.m / ctas);
        sb.sub[0].rows_rem = (int)(p.m % ctas);
        sb.sub[0].ll_off = 0;
        sb.sub[0].row0 = p.gather ? p.gather->row0 : 0;
        sb.sub[0].slot = p.gather ? p.gather->slot : 0;
        sb.sub[0].ctas = (int)ctas;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)ctas, 1, 1);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = (size_t)g.total;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = ctx->opt_pdl ? 1 : 0;
    b200_gemv_params pp = p;
    pp.abort_dev = ctx->abort_dev;
    pp.abort_host = ctx->abort_host_dev;
    pp.wait_timeout_ns = (unsigned long long)(ctx->opt_plan_timeout_ms > 0 ? ctx->opt_plan_timeout_ms : 120000) * 1000000ull;
    pp.trace = NULL;
    if (ctx->trace && ctx->trace_next < ctx->trace_capacity)
        pp.trace = ctx->trace + (size_t)(ctx->trace_next++) * B200_TRACE_MAX_CTAS * B200_TRACE_STAMPS;
    b200_gather gd;
    memset(&gd, 0, sizeof(gd));
    if (p.gather) gd = *p.gather;
    pp.gather = NULL;
    B200_CUDA_TRY(ctx, cudaLaunchKernelEx(&cfg, kern, pp, gg, gd, sb));
    ctx->launches++;
    return B200_OK;
}

// divide the grid among the matrices of a batch in proportion to their rows (same k => same bytes per row)
void build_batch(const b200_ctx *ctx, const b200_gemv_params *ps, int count, StreamBatch *sb) {
    memset(sb, 0, sizeof(*sb));
    sb->count = count;
    int64_t rows = 0;
    for (int j = 0; j < count; j++) rows += ps[j].m;
    int64_t grid = ctx->sm_count;
    if (grid > rows) grid = rows;
    int64_t given = 0;
    int share[kMaxBatch];
    for (int j = 0; j < count; j++) {
        int64_t c = grid * ps[j].m / rows;
        if (c < 1) c = 1;
        if (c > ps[j].m) c = ps[j].m;
        share[j] = (int)c;
        given += c;
    }
    // hand out what integer division left over to the matrices with the most rows per CTA
    while (given < grid) {
        int best = -1;
        double worst = 0;
        for (int j = 0; j < count; j++) {
            const double load = (double)ps[j].m / share[j];
            if (share[j] < ps[j].m && load > worst) { worst = load; best = j; }
        }
        if (best < 0) break;
        share[best]++;
        given++;
    }
    int cta0 = 0;
    for (int j = 0; j < count; j++) {
        sb->sub[j].qs = ps[j].qs;
        sb->sub[j].d = ps[j].d;
        sb->sub[j].dst = ps[j].dst;
        sb->sub[j].m = (int)ps[j].m;
        sb->sub[j].cta0 = cta0;
        sb->sub[j].rows_q = (int)(ps[j].m / share[j]);
        sb->sub[j].rows_rem = (int)(ps[j].m % share[j]);
        sb->sub[j].ctas = share[j];
        if (ps[j].gather) {
            const b200_gather *g0 = ps[0].gather, *gj = ps[j].gather;
            sb->sub[j].ll_off = ((const char *)gj->peer_dst[gj->rank] - (const char *)g0->peer_dst[g0->rank]) / 8;
            sb->sub[j].row0 = gj->row0;
            sb->sub[j].slot = gj->slot;
        }
        cta0 += share[j];
    }
    sb->total_ctas = cta0;
}

template <int TYPE>
int launch_stream_cols(b200_ctx *ctx, const b200_gemv_params &p, const StreamGeom &g, bool dots) {
    switch (p.n) {
        case 1: return launch_stream_typed<TYPE, 1>(ctx, p, g, dots);
        case 2: return launch_stream_typed<TYPE, 2>(ctx, p, g, dots);
        case 3: return launch_stream_typed<TYPE, 3>(ctx, p, g, dots);
        case 4: return launch_stream_typed<TYPE, 4>(ctx, p, g, dots);
        case 5: return launch_stream_typed<TYPE, 5>(ctx, p, g, dots);
        case 6: return launch_stream_typed<TYPE, 6>(ctx, p, g, dots);
        case 7: return launch_stream_typed<TYPE, 7>(ctx, p, g, dots);
        case 8: return launch_stream_typed<TYPE, 8>(ctx, p, g, dots);
    }
    return B200_ERR_INVALID;
}

}  // namespace

int b200_launch_gather_finish(b200_ctx *ctx, const b200_gather &gd, const void *ll_src, float *out, int64_t count) {
    LLWait lw;
    lw.abort_dev = ctx->abort_dev;
    lw.abort_host = ctx->abort_host_dev;
    lw.timeout_ns = (unsigned long long)(ctx->opt_plan_timeout_ms > 0 ? ctx->opt_plan_timeout_ms : 120000) * 1000000ull;
    gather_finish_kernel<<<1, 1024, 0, ctx->stream>>>(gd, (const uint2 *)ll_src, out, count, lw);
    ctx->launches++;
    B200_CUDA_TRY(ctx, cudaGetLastError());
    return B200_OK;
}

// One launch for `count` (2..4) decode mul_mats that share type, k and src1 (n == 1).  Returns false when the batch does
// not qualify (the caller then launches them one by one).
bool b200_try_launch_gemv_stream_batch(b200_ctx *ctx, const b200_gemv_params *ps, int count, int *rc) {
    if (count < 2 || count > kMaxBatch) return false;
    for (int j = 0; j < count; j++) {
        const b200_gemv_params &p = ps[j];
        if (p.n != 1 || p.dst_n != 1 || p.dots || p.bias || p.residual || p.residual2 || p.act) return false;
        if ((p.gather != NULL) != (ps[0].gather != NULL)) return false;
        if (p.gather) {
            // one gather description must fit all: same group, same state, same producer to wait for, and the matrices' LL
            // vectors at the same relative offset on every rank (symmetric buffers)
            const b200_gather *g0 = ps[0].gather, *gj = p.gather;
            if (gj->world != g0->world || gj->rank != g0->rank || gj->state != g0->state || gj->wait_slot != g0->wait_slot) return false;
            const ptrdiff_t delta = (const char *)gj->peer_dst[gj->rank] - (const char *)g0->peer_dst[g0->rank];
            if (delta % 8 != 0) return false;
            for (int r = 0; r < g0->world; r++)
                if ((const char *)gj->peer_dst[r] - (const char *)g0->peer_dst[r] != delta) return false;
        }
        if (p.ne12 != 1 || p.ne13 != 1 || p.ne02 != 1 || p.ne03 != 1) return false;
        if (p.k % 256 != 0 || p.k > 32768 || p.m < 1 || p.m >= (1ll << 31)) return false;
        if (((uintptr_t)p.qs & 15) != 0 || ((uintptr_t)p.d & 15) != 0) return false;
        if (p.type != ps[0].type || p.k != ps[0].k || p.x != ps[0].x) return false;
    }
    StreamGeom g;
    if (!stream_geometry(ps[0], &g)) return false;
    StreamBatch sb;
    build_batch(ctx, ps, count, &sb);
    if (ps[0].type == B200_TYPE_Q4_0) *rc = launch_stream_typed<B200_TYPE_Q4_0, 1>(ctx, ps[0], g, false, &sb);
    else if (ps[0].type == B200_TYPE_Q8_0) *rc = launch_stream_typed<B200_TYPE_Q8_0, 1>(ctx, ps[0], g, false, &sb);
    else return false;
    return true;
}

// returns true when the streaming kernel takes this shape; *rc then holds the launch status
bool b200_try_launch_gemv_stream(b200_ctx *ctx, const b200_gemv_params &p, int *rc) {
    if (p.ne12 != 1 || p.ne13 != 1 || p.ne02 != 1 || p.ne03 != 1) return false;
    if (p.k % 256 != 0 || p.k > 32768 || p.n < 1 || p.n > 8) return false;
    if (((uintptr_t)p.qs & 15) != 0 || ((uintptr_t)p.d & 15) != 0) return false;
    if (p.dots == NULL && p.dst_n != p.n) return false;  // column-chunked dst keeps the generic addressing
    if (p.gather && (p.n != 1 || p.dots || p.bias || p.residual || p.residual2 || p.act)) return false;
    StreamGeom g;
    if (!stream_geometry(p, &g)) return false;
    const bool dots = p.dots != NULL;
    if (p.type == B200_TYPE_Q4_0) *rc = launch_stream_cols<B200_TYPE_Q4_0>(ctx, p, g, dots);
    else if (p.type == B200_TYPE_Q8_0) *rc = launch_stream_cols<B200_TYPE_Q8_0>(ctx, p, g, dots);
    else return false;
    return true;
}
