// b200_gemm_f16.cu -- the prefill path (every n the GEMV does not take): dst[n][m] = W[m,k] x X[n,k]^T as ONE dense contraction on the 5th-generation
// tensor cores, fp16 operands, fp32 accumulation in TMEM.
//
// Stands in for the COMPUTE phase of ggml_compute_forward_mul_mat (src/ggml.c:12056-12096) for prefill-sized batches, the way the
// reference's own CUDA backend does it for large batches (ggml_cuda_op_mul_mat_cublas, src/ggml-cuda.cu:1208-1306:
// to_fp16_cuda(src0), to_fp16_cuda(src1), fp16 GEMM):
//   W'[m][k] = fp16_rn(quant * d_w)      -- dequantize_row_q4_0 / _q8_0 (src/ggml-quants.c:980-998, :1074-1088), rounded once to fp16
//   X'[n][k] = fp16_rn(q * d_x)          -- the Q8_0-quantized activations (quantize_row_q8_0, src/ggml-quants.c:535-618): the
//                                           same values the exact path multiplies, rounded once to fp16
//   dst      = W' X'^T                   -- tcgen05.mma kind::f16, fp32 accumulators
// Relative error 2^-12 per operand element: NMSE ~1e-7 against the 5e-4 bound of test-backend-ops (measured in
// tests/test_gpu_gemm_f16.py against the oracle AND the exact kernel).  Bit-exact per-block int32 dots remain available through
// b200_block_dots / the "gemm_exact" option (b200_gemm_tc.cu), which is tensor-pipe-starved by design: every 32-wide block's
// partial must be scaled separately on CUDA cores (m*n*k/32 accumulator updates, profiles/r01_gemm_experiments.md).
//
// Why this shape (numbers: tools/mma_peak.cu on this pool's B200: 2230 TFLOP/s issue-only for kind::f16, 128 cycles per
// M128/N256/K16 instruction; the round-1 fp16 kernels ran at 100 us on C2 = 21 % of that):
//   * the round-1 kernels were L2-bandwidth-bound: a 128 x 256 tile re-reads X' for every 128 rows of W (360 MB for C2) and
//     materialised W' (2 B per weight) in HBM.  Here the weights are dequantized INSIDE the kernel from the raw repacked blocks
//     (0.56 B per weight from L2) and a CTA PAIR works on one 256 x 256 tile (tcgen05 cta_group::2: each CTA stages only its half
//     of X', the tensor cores of the pair share it), so X' traffic halves and the per-SM shared-memory traffic fits;
//   * persistent: one CTA pair per TPC walks a static list of work units; two 256-column accumulators in TMEM (all 512 columns)
//     so the epilogue of one unit overlaps the MMAs of the next;
//   * wave quantization (C2: 86 tiles on 74 pairs): the tiles of the last, partial round are split along k across the idle
//     pairs; the partial accumulators meet in an L2-resident workspace and are summed in a FIXED order (part 0, 1, 2, ...), each
//     pair reducing its own column slice -- deterministic, no atomics.
//
// Measured alternative, not shipped (profiles/r02_experiments.md): W' written by tcgen05.st into TENSOR memory and taken from there as the A operand
// (no shared-memory W' stages, no proxy fence) -- correct (all parity tests), +3-5 % on shapes with >= 3 tiles per pair, but it leaves room
// for only ONE accumulator, so the ~3 us accumulator drain per tile is exposed: -6 % on C2 and -5 % on the GPT-J prefill.  The same pass showed
// where the main loop stands: the MMA-issuing thread needs 520-580 cycles per k-step against the tensor pipe's 512.
//
// CTA = 16 warps: warp 0 TMA producer of the X' tile, warp 3 TMA producer of the raw weight rows (+ L2 prefetch), warp 1 MMA issuer
// (leader CTA only), warp 2 TMEM allocator, warps 4-7 epilogue (one per TMEM lane quadrant), warps 8-15 dequantization (thread = weight row x one block of 32).
#include "b200_tc_common.cuh"

using namespace b200tc;

namespace {

constexpr int TM = 128;                 // weight rows per CTA (tile M = 256 per pair)
constexpr int TN = 256;                 // activation rows (dst columns) per tile; each CTA stages 128 of them
constexpr int KSTEP = 64;               // k per pipeline stage = 128 bytes of fp16 = 2 quant blocks
constexpr int kStagesF = 5;
constexpr int kATile = TM * 128;        // 16 KB
constexpr int kBTile = (TN / 2) * 128;  // 16 KB
constexpr int kStageF = kATile + kBTile;               // 32 KB, a multiple of 1024
constexpr int kThreadsF = 16 * 32;
constexpr int kDeqWarps = 8;
constexpr int kDeqGroups = 2;             // groups of kDeqWarps / kDeqGroups warps that take alternate k-steps
constexpr int kEpiWarpsF = 4;
constexpr int kBarsF = (3 + kDeqGroups) * kStagesF + 4;   // raw_full (x groups), b_full, a_ready, empty (x stages), tmem_full[2], tmem_empty[2]
// raw weight tiles: per stage 128 rows x (32 | 64) bytes of the qs plane (both blocks of the k-step), TMA with the 32 / 64-byte
// swizzle so that thread = row reads its 16-byte pieces without bank conflicts
template <int TYPE> struct RawTile {
    static constexpr int kRow = TYPE == B200_TYPE_Q4_0 ? 32 : 64;
    static constexpr int kBytes = TM * kRow;                                 // 4 KB / 8 KB
    static constexpr int kSmem = kStagesF * (kStageF + kBytes) + 1024 + kBarsF * 8 + 64;
    static constexpr int kPfSteps = 256 / kRow;                              // k-steps per 256-byte L2 prefetch box row (8 / 4)
};
constexpr int kPfAhead = 24;             // k-steps the L2 prefetch cursor stays ahead of the copies (~6 us of MMAs)
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;            // shared::cluster address of the same location in the pair's leader CTA

// instruction descriptor: dense, D = F32 (1 << 4), A = B = F16 (0), both K-major, N = 256, M = 256 (the pair)
constexpr uint32_t kIdescF16x2 = (1u << 4) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)((2 * TM) >> 4) << 24);

__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on a barrier that may live in the peer CTA (shared::cluster address).  Default (CTA-scope) semantics on purpose: what
// crosses the pair is never generic-proxy data -- the W' tile is read by this SM's own tensor core (ordered by
// fence.proxy.async), the accumulator by tcgen05.fence -- and a cluster-scope release / acquire costs a MEMBAR.ALL.GPU plus an
// L1 invalidation (CCTL.IVALL) per k-step per warp: measured 4x the MMA time.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar_addr, uint32_t parity) {
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
// X' tile of one CTA of the pair; the bytes are counted on the LEADER's barrier (the MMA needs both halves)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t leader_bar) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
                 "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
                 : "memory");
}
// HBM -> L2 only: the box of a tensor map, no shared-memory destination, no barrier
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap *map, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_mma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives on the barrier at this offset in BOTH CTAs of the pair once every MMA issued so far has completed
__device__ __forceinline__ void tc_commit_pair(uint32_t bar_addr) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar_addr), "h"((unsigned short)3)
                 : "memory");
}
__device__ __forceinline__ uint4 lds128f(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128f(uint32_t a, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Work units of a launch.  Tile t = (tm, tn) with tn fastest (consecutive tiles share their weight rows in L2).  Pair p runs the
// full tiles p, p + P, ... of the complete rounds, then -- when the last round is partial -- one k-slice of one of its tiles.
struct Sched {
    int tiles_n, tiles, pairs;
    int full_rounds;      // rounds in which every pair has a whole tile
    int rem, split;       // tiles of the partial round, k-slices per such tile (rem * split <= pairs)
    int ksteps;           // k-steps of a whole tile
};

struct Unit {
    int tm, tn, ks0, ks1;     // tile, k-step range
    int part, nparts, rtile;  // k-slice index / count (1 = whole tile, stored directly) and index among the split tiles
};

__device__ __forceinline__ bool get_unit(const Sched &sc, int pair, int i, Unit &u) {
    int tile;
    if (i < sc.full_rounds) {
        tile = i * sc.pairs + pair;
        if (tile >= sc.tiles) return false;
        u.ks0 = 0; u.ks1 = sc.ksteps; u.part = 0; u.nparts = 1; u.rtile = 0;
    } else if (i == sc.full_rounds && sc.rem > 0 && pair < sc.rem * sc.split) {
        u.rtile = pair / sc.split;
        u.part = pair - u.rtile * sc.split;
        u.nparts = sc.split;
        tile = sc.full_rounds * sc.pairs + u.rtile;
        u.ks0 = (int)((long long)sc.ksteps * u.part / sc.split);
        u.ks1 = (int)((long long)sc.ksteps * (u.part + 1) / sc.split);
    } else {
        return false;
    }
    u.tm = tile / sc.tiles_n;
    u.tn = tile - u.tm * sc.tiles_n;
    return true;
}

struct GemmF16Args {
    const __half *dw;        // weight scales [m][nb]
    float *dst;              // [n][m]
    float *partial;          // split-k workspace: [rem][split][TN][2 * TM] fp32
    uint32_t *counters;      // [rem] arrivals of finished k-slices (zeroed before the launch)
    uint32_t *abort_flag;    // the context's abort word (device memory)
    uint32_t *abort_host;
    int m, n, k;
    int dw_pf;               // the scale plane can be prefetched through its tensor map (row pitch a multiple of 16 bytes)
    unsigned long long *trace;   // optional: this launch's [gridDim.x][8] %globaltimer stamps (b200_ctx_set_trace): 0 start, 1 first MMA, 2 MMAs of the
                                 // first unit issued, 3 first accumulator complete, 4 first unit stored, 5 last unit stored, 6 all k-slices arrived, 7 end
    Sched sc;
};
__device__ __forceinline__ void gstamp(unsigned long long *trace, int slot) {
    if (trace) trace[(size_t)blockIdx.x * 8 + slot] = gtimer();
}

// two small unsigned integers (< 1024) at bits 0.. and 16.. of `bits`, already OR-ed into the mantissa of 2^e ->
// half2((u0 - bias) * d, (u1 - bias) * d): the subtraction is exact, the product is rounded once
__device__ __forceinline__ uint32_t cvt2(uint32_t hbits, __half2 bias, __half2 d2) {
    __half2 h = *reinterpret_cast<const __half2 *>(&hbits);
    h = __hmul2(__hsub2(h, bias), d2);
    return *reinterpret_cast<const uint32_t *>(&h);
}

// k-order inside the operand tiles: within every aligned group of four k the order is (0, 2, 1, 3) -- what the masks below
// produce without byte permutes; quantize_to_f16_kernel writes X' in the same order, and a contraction does not care.
template <int TYPE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreadsF, 1)
gemm_f16_pair_kernel(const __grid_constant__ CUtensorMap map_raw, const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_raw_pf,
                     const __grid_constant__ CUtensorMap map_dw_pf, const GemmF16Args g) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    using RT = RawTile<TYPE>;
    unsigned char *rawtiles = smem + kStagesF * kStageF;
    uint64_t *bars = reinterpret_cast<uint64_t *>(rawtiles + kStagesF * RT::kBytes);
    // raw_full: [group][stages], local: this CTA's raw weight rows of a k-step have landed.  One barrier per (group, stage): a
    // group only waits for ITS k-steps (every other fill of a stage), and a parity wait that skips a phase can pass a fill early
    // (seen as one warp's 32 rows of a tile going wrong once in ~40 launches): every waiter must see consecutive phases.
    uint64_t *raw_full = bars;
    uint64_t *b_full = bars + kDeqGroups * kStagesF;  // [stages] leader: both halves of the X' tile have landed
    uint64_t *a_ready = b_full + kStagesF;            // [stages] leader: both CTAs' W' tiles are written (one group's warps of each CTA)
    uint64_t *empty = a_ready + kStagesF;             // [stages] local: the MMAs that read this stage have completed
    uint64_t *tmem_full = empty + kStagesF;           // [2] local: the unit's accumulator is complete
    uint64_t *tmem_empty = tmem_full + 2;             // [2] leader: both CTAs' epilogues have drained the accumulator (8 arrivals)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);

    constexpr int QSB = TYPE == B200_TYPE_Q4_0 ? 16 : 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank();
    const int pair = blockIdx.x >> 1;
    const int nb = g.k >> 5;
    const Sched sc = g.sc;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_raw) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_raw_pf) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_dw_pf) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kDeqGroups * kStagesF; s++) mbar_init(&raw_full[s], 1);
        for (int s = 0; s < kStagesF; s++) {
            mbar_init(&b_full[s], 1);
            mbar_init(&a_ready[s], 2 * kDeqWarps / kDeqGroups);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; b++) {
            mbar_init(&tmem_full[b], 1);
            mbar_init(&tmem_empty[b], 2 * kEpiWarpsF);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) gstamp(g.trace, 0);
    const uint32_t smem_a = smem_u32(smem);

    if (warp == 0) {
        // ===== TMA producer for X' (one thread per CTA): this CTA's 128 activation rows of the tile, stage by stage.  X' comes from
        // the quantize kernel right before this one in the stream: with programmatic dependent launch this grid is already
        // resident (weights are being staged and dequantized) when that kernel finishes. =====
        if (lane == 0) {
            asm volatile("griddepcontrol.wait;" ::: "memory");
            int it = 0;
            Unit u;
            for (int i = 0; get_unit(sc, pair, i, u); i++) {
                const int n0 = u.tn * TN + (int)rank * (TN / 2);
                for (int ks = u.ks0; ks < u.ks1; ks++, it++) {
                    const int s = it % kStagesF;
                    const uint32_t ph = (uint32_t)(it / kStagesF) & 1u;
                    mbar_wait(&empty[s], ph ^ 1u);
                    if (rank == 0) mbar_expect_tx(&b_full[s], 2 * kBTile);
                    tma_load_2d_pair(smem_a + (uint32_t)(s * kStageF + kATile), &map_b, ks * 128, n0, smem_u32(&b_full[s]) & kPeerMask);
                }
            }
        }
    } else if (warp == 3) {
        // ===== TMA producer for the raw weight rows (its own thread: a bulk-class instruction occupies the issuing thread for
        // ~150 cycles, and one k-step of MMAs is only 512).  The weights come from HBM (every byte exactly once per launch), ~2 us
        // away, while the five stages of the ring cover ~1.3 us of MMAs: a second cursor runs kPfAhead k-steps ahead of the copies,
        // across unit boundaries, and asks L2 for 256 bytes per row at a time (and for the block scales, 32 bytes per row). =====
        if (lane == 0) {
            int pf_i = 0, pf_ahead = 0;
            Unit pu;
            bool pf_live = get_unit(sc, pair, 0, pu);
            int pf_ks = pf_live ? pu.ks0 : 0;
            auto prefetch_more = [&]() {
                while (pf_live && pf_ahead < kPfAhead) {
                    const int pm0 = pu.tm * 2 * TM + (int)rank * TM;
                    const int blk0 = pf_ks & ~(RT::kPfSteps - 1);       // aligned boxes: the innermost coordinate stays a multiple of 16 bytes
                    tma_prefetch_2d(&map_raw_pf, blk0 * RT::kRow, pm0);
                    if (g.dw_pf && (pf_ks == pu.ks0 || (blk0 & 7) == 0)) tma_prefetch_2d(&map_dw_pf, (pf_ks >> 3) * 32, pm0);     // 16 scales = 8 k-steps
                    const int next = min(blk0 + RT::kPfSteps, pu.ks1);
                    pf_ahead += next - pf_ks;
                    pf_ks = next;
                    if (pf_ks >= pu.ks1) {
                        pf_live = get_unit(sc, pair, ++pf_i, pu);
                        pf_ks = pf_live ? pu.ks0 : 0;
                    }
                }
            };
            int it = 0;
            Unit u;
            for (int i = 0; get_unit(sc, pair, i, u); i++) {
                const int m0 = u.tm * 2 * TM + (int)rank * TM;
                for (int ks = u.ks0; ks < u.ks1; ks++, it++) {
                    prefetch_more();
                    pf_ahead--;
                    const int s = it % kStagesF;
                    const uint32_t ph = (uint32_t)(it / kStagesF) & 1u;
                    mbar_wait(&empty[s], ph ^ 1u);
                    uint64_t *rf = &raw_full[(it % kDeqGroups) * kStagesF + s];      // the barrier of the group that takes this k-step
                    mbar_expect_tx(rf, RT::kBytes);
                    tma_load_2d(rawtiles + s * RT::kBytes, &map_raw, ks * RT::kRow, m0, rf);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one thread of the LEADER CTA issues for the pair =====
        if (lane == 0 && rank == 0) {
            int it = 0, ui = 0;
            Unit u;
            for (int i = 0; get_unit(sc, pair, i, u); i++, ui++) {
                const int acc = ui & 1;
                const uint32_t aph = (uint32_t)(ui >> 1) & 1u;
                mbar_wait_cluster(smem_u32(&tmem_empty[acc]), aph ^ 1u);     // both epilogues have drained this accumulator
                tc_fence_after();
                const uint32_t td = tmem_base + (uint32_t)(acc * TN);
                for (int ks = u.ks0; ks < u.ks1; ks++, it++) {
                    const int s = it % kStagesF;
                    const uint32_t ph = (uint32_t)(it / kStagesF) & 1u;
                    mbar_wait_cluster(smem_u32(&b_full[s]), ph);
                    mbar_wait_cluster(smem_u32(&a_ready[s]), ph);
                    tc_fence_after();
                    if (it == 0) gstamp(g.trace, 1);
                    const uint32_t sa = smem_a + (uint32_t)(s * kStageF);
                    const uint64_t da = make_desc_sw128(sa), db = make_desc_sw128(sa + kATile);
#pragma unroll
                    for (int j = 0; j < 4; j++)      // K = 16 fp16 = 32 bytes per MMA: +2 in the (>>4) start-address field, inside the swizzle atom
                        tc_mma_f16_pair(td, da + (uint64_t)(j * 2), db + (uint64_t)(j * 2), kIdescF16x2, (ks > u.ks0 || j > 0) ? 1u : 0u);
                    tc_commit_pair(smem_u32(&empty[s]));          // both CTAs: stage s may be refilled
                }
                tc_commit_pair(smem_u32(&tmem_full[acc]));        // both CTAs: the accumulator is complete
                if (ui == 0) gstamp(g.trace, 2);
            }
        }
    } else if (warp >= 8) {
        // ===== dequantization: two groups of four warps take alternate k-steps; thread = one weight row of this CTA's 128, both
        // blocks of the k-step (the whole 128-byte row of the W' tile).  Why groups: the generic -> async proxy fence that must
        // precede the arrival costs a warp ~300 cycles (ncu: 40 % of these warps' time when every warp paid it every k-step);
        // two groups halve the fences per k-step and give each warp two k-step periods per pass. =====
        const int grp = (warp - 8) >> 2;
        const int r = (warp & 3) * 32 + lane;
        const __half2 bias_lo = TYPE == B200_TYPE_Q4_0 ? __floats2half2_rn(1032.f, 1032.f) : __floats2half2_rn(1152.f, 1152.f);
        const __half2 bias_hi = __floats2half2_rn(72.f, 72.f);
        const uint32_t leader_ready = smem_u32(a_ready) & kPeerMask;
        const uint32_t sw = (uint32_t)(r & 7);
        // where TMA's 32 / 64-byte swizzle puts 16-byte piece j of this row: j ^ (bits 7.. of the row's offset)
        const uint32_t rsw = TYPE == B200_TYPE_Q4_0 ? (uint32_t)((r >> 2) & 1) : (uint32_t)((r >> 1) & 3);
        const uint32_t raw_a = smem_u32(rawtiles) + (uint32_t)(r * RT::kRow);
        int it = grp, s = grp;                  // this group's k-steps of the pair's flattened sequence: grp, grp + 2, ...
        int visits = 0;                         // ... of which every kStagesF-th comes back to the same stage: the phase of ITS barrier
        uint32_t ph = 0;
        int base = 0;                           // flattened index of the unit's first k-step
        Unit u;
        for (int i = 0; get_unit(sc, pair, i, u); i++) {
            const int len = u.ks1 - u.ks0;
            const int row = min(u.tm * 2 * TM + (int)rank * TM + r, g.m - 1);      // (rows past m: TMA delivered zeros; any finite scale will do)
            const unsigned short *dwp = reinterpret_cast<const unsigned short *>(g.dw) + (int64_t)row * nb;
            auto ldd = [&](int ks, int b) -> unsigned short {      // scale of block b of k-step ks; past the unit or past a ragged k: 0
                const int blk = 2 * ks + b;
                return (ks < u.ks1 && blk < nb) ? __ldg(dwp + blk) : (unsigned short)0;
            };
            int ks = u.ks0 + (it - base);
            // the scales two of this thread's k-steps (four k-steps of the stream) ahead of their use; L2 has them (prefetch above)
            unsigned short a0 = ldd(ks, 0), a1 = ldd(ks, 1), b0 = ldd(ks + kDeqGroups, 0), b1 = ldd(ks + kDeqGroups, 1);
            for (; ks < u.ks1; ks += kDeqGroups, it += kDeqGroups) {
                const unsigned short c0 = a0, c1 = a1;
                a0 = b0; a1 = b1;
                b0 = ldd(ks + 2 * kDeqGroups, 0);
                b1 = ldd(ks + 2 * kDeqGroups, 1);
                mbar_wait(&raw_full[grp * kStagesF + s], ph);
                uint4 q[RT::kRow / 16];
#pragma unroll
                for (int j = 0; j < RT::kRow / 16; j++) q[j] = lds128f(raw_a + (uint32_t)(s * RT::kBytes) + ((((uint32_t)j) ^ rsw) << 4));
                const uint32_t arow = smem_a + (uint32_t)(s * kStageF + r * 128);
#pragma unroll
                for (int b = 0; b < 2; b++) {
                    const __half dh = __ushort_as_half(b == 0 ? c0 : c1);
                    const __half2 d2 = __halves2half2(dh, dh);
                    if (TYPE == B200_TYPE_Q4_0) {
                        // the block's 16 bytes: low nibbles = elements 0..15, high nibbles = elements 16..31
                        const uint32_t w[4] = {q[b].x, q[b].y, q[b].z, q[b].w};
                        uint32_t lo[8], hi[8];
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const uint32_t x = w[j], y = w[j] >> 8;
                            lo[2 * j + 0] = cvt2((x & 0x000F000Fu) | 0x64006400u, bias_lo, d2);     // bytes 0, 2 of the word: 1024 + nib
                            lo[2 * j + 1] = cvt2((y & 0x000F000Fu) | 0x64006400u, bias_lo, d2);     // bytes 1, 3
                            hi[2 * j + 0] = cvt2((x & 0x00F000F0u) | 0x54005400u, bias_hi, d2);     // 64 + nib (the nibble sits 4 bits up: ulp 1/16)
                            hi[2 * j + 1] = cvt2((y & 0x00F000F0u) | 0x54005400u, bias_hi, d2);
                        }
                        // 16-byte chunk c of the row goes where TMA with SWIZZLE_128B would put it: c ^ (row & 7)
                        sts128f(arow + (((uint32_t)(b * 4 + 0) ^ sw) << 4), make_uint4(lo[0], lo[1], lo[2], lo[3]));
                        sts128f(arow + (((uint32_t)(b * 4 + 1) ^ sw) << 4), make_uint4(lo[4], lo[5], lo[6], lo[7]));
                        sts128f(arow + (((uint32_t)(b * 4 + 2) ^ sw) << 4), make_uint4(hi[0], hi[1], hi[2], hi[3]));
                        sts128f(arow + (((uint32_t)(b * 4 + 3) ^ sw) << 4), make_uint4(hi[4], hi[5], hi[6], hi[7]));
                    } else {
#pragma unroll
                        for (int h16 = 0; h16 < 2; h16++) {
                            const uint4 qq = q[(b * 2 + h16) % (RT::kRow / 16)];
                            const uint32_t w[4] = {qq.x ^ 0x80808080u, qq.y ^ 0x80808080u, qq.z ^ 0x80808080u, qq.w ^ 0x80808080u};   // int8 + 128
                            uint32_t o[8];
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                o[2 * j + 0] = cvt2((w[j] & 0x00FF00FFu) | 0x64006400u, bias_lo, d2);
                                o[2 * j + 1] = cvt2(((w[j] >> 8) & 0x00FF00FFu) | 0x64006400u, bias_lo, d2);
                            }
                            sts128f(arow + (((uint32_t)(b * 4 + h16 * 2 + 0) ^ sw) << 4), make_uint4(o[0], o[1], o[2], o[3]));
                            sts128f(arow + (((uint32_t)(b * 4 + h16 * 2 + 1) ^ sw) << 4), make_uint4(o[4], o[5], o[6], o[7]));
                        }
                    }
                }
                // generic-proxy stores -> visible to the tensor core (async proxy), then one arrival per warp on the leader's barrier
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(leader_ready + (uint32_t)(s * 8));
                s += kDeqGroups;
                if (s >= kStagesF) s -= kStagesF;
                if (++visits == kStagesF) { visits = 0; ph ^= 1u; }
            }
            base += len;
        }
    } else if (warp >= 4) {
        // ===== epilogue: warp q of 4 owns TMEM lanes 32q.. = weight rows; whole tiles go straight to dst, k-slices through
        // the split-k workspace =====
        const int quad = warp & 3;
        const uint32_t leader_tempty = smem_u32(tmem_empty) & kPeerMask;
        int ui = 0;
        Unit u;
        for (int i = 0; get_unit(sc, pair, i, u); i++, ui++) {
            const int acc = ui & 1;
            const uint32_t aph = (uint32_t)(ui >> 1) & 1u;
            const int rloc = (int)rank * TM + quad * 32 + lane;      // row inside the 256-row tile
            const int row = u.tm * 2 * TM + rloc;
            const int n0 = u.tn * TN;
            mbar_wait(&tmem_full[acc], aph);
            tc_fence_after();
            if (ui == 0 && quad == 0 && lane == 0) gstamp(g.trace, 3);
            const uint32_t tcol = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * TN);
            float *pbase = g.partial + ((size_t)(u.rtile * sc.split + u.part) * TN) * (2 * TM) + rloc;
            // only the column groups that hold activation columns: a short prompt (n = 16) leaves 7 of the 8 groups of the tile unused,
            // and draining them was 3 of the 15 us of a 4096 x 4096 x 16 mul_mat
            const int ngrp = min(TN / 32, (min(TN, (int)g.n - n0) + 31) / 32);
#pragma unroll 1
            for (int c32 = 0; c32 < ngrp; c32++) {
                uint32_t v[32];
                tc_ld32(tcol + (uint32_t)(c32 * 32), v);
                tc_wait_ld();
                if (c32 == ngrp - 1) {
                    // the accumulator is in registers: hand it back to the MMA thread before the stores
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(leader_tempty + (uint32_t)(acc * 8));
                }
                if (u.nparts == 1) {
                    if (row < g.m) {
#pragma unroll
                        for (int j = 0; j < 32; j++) {
                            const int c = n0 + c32 * 32 + j;
                            if (c < g.n) g.dst[(int64_t)c * g.m + row] = __uint_as_float(v[j]);   // 32 lanes -> 128 contiguous bytes
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; j++) pbase[(size_t)(c32 * 32 + j) * (2 * TM)] = __uint_as_float(v[j]);
                }
            }
            if (u.nparts > 1) {
                // this warp's share of the k-slice is in the workspace: publish (the reduction follows below, by all twelve warps)
                __threadfence();
                __syncwarp();
                if (lane == 0) atomicAdd(g.counters + u.rtile, 1u);
            }
            if (quad == 0 && lane == 0) gstamp(g.trace, ui == 0 ? 4 : 5);
        }
    }

    // ===== split-k reduction: the pair's k-slice (always its last unit) is in the workspace; once every slice of the tile is,
    // this pair sums ITS column range of the tile over all slices in the fixed order 0, 1, 2, ... and writes dst.  Twelve warps
    // per CTA (epilogue + dequantization, idle by now): warp = 32 rows x a third of the columns, all loads of a batch in flight. =====
    if (warp >= 4 && sc.rem > 0 && pair < sc.rem * sc.split) {
        Unit u;
        get_unit(sc, pair, sc.full_rounds, u);
        const int wslot = warp - 4;                   // 0..11
        const int rgrp = wslot & 3, cthird = wslot >> 2;
        const int rloc = (int)rank * TM + rgrp * 32 + lane;
        const int row = u.tm * 2 * TM + rloc;
        const int n0 = u.tn * TN;
        if (lane == 0) {
            const uint32_t target = (uint32_t)(u.nparts * 2 * kEpiWarpsF);
            unsigned long long t0 = 0;
            unsigned spins = 0;
            for (;;) {
                uint32_t seen;
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(g.counters + u.rtile) : "memory");
                if (seen >= target) break;
                __nanosleep(100);
                if ((++spins & 255u) == 0u) {      // bounded: a pair that never delivers must not hang the GPU
                    if (*reinterpret_cast<volatile uint32_t *>(g.abort_flag) != 0u) break;
                    const unsigned long long now = gtimer();
                    if (t0 == 0ull) t0 = now;
                    else if (now - t0 > 5000000000ull) {
                        if (atomicCAS(g.abort_flag, 0u, 0x80000004u) == 0u) *reinterpret_cast<volatile uint32_t *>(g.abort_host) = 0x80000004u;
                        break;
                    }
                }
            }
        }
        if (wslot == 0 && lane == 0) gstamp(g.trace, 6);
        __syncwarp();
        const int ncols = min(TN, (int)g.n - n0);                                              // columns of the tile that exist
        const int p_lo = ncols * u.part / u.nparts, p_hi = ncols * (u.part + 1) / u.nparts;     // the pair's share of them
        const int span = p_hi - p_lo;
        const int c_lo = p_lo + (span > 0 ? span * cthird / 3 : 0), c_hi = p_lo + (span > 0 ? span * (cthird + 1) / 3 : 0);
        const float *tp = g.partial + ((size_t)(u.rtile * sc.split) * TN) * (2 * TM) + rloc;
        if (row < g.m) {
            constexpr int CB = 8, PB = 8;      // (all k-slices of eight columns in flight at once: one L2 round trip per batch for up to eight slices)
            for (int c = c_lo; c < c_hi; c += CB) {
                float acc[CB];
#pragma unroll
                for (int cc = 0; cc < CB; cc++) acc[cc] = 0.0f;
                for (int p0 = 0; p0 < u.nparts; p0 += PB) {
                    float v[PB][CB];
#pragma unroll
                    for (int pp = 0; pp < PB; pp++)
#pragma unroll
                        for (int cc = 0; cc < CB; cc++)
                            v[pp][cc] = (p0 + pp < u.nparts && c + cc < c_hi) ? __ldcg(tp + ((size_t)(p0 + pp) * TN + (c + cc)) * (2 * TM)) : 0.0f;
#pragma unroll
                    for (int pp = 0; pp < PB; pp++)
                        if (p0 + pp < u.nparts) {
#pragma unroll
                            for (int cc = 0; cc < CB; cc++) acc[cc] += v[pp][cc];
                        }
                }
#pragma unroll
                for (int cc = 0; cc < CB; cc++)
                    if (c + cc < c_hi) g.dst[(int64_t)(n0 + c + cc) * g.m + row] = acc[cc];
            }
        }
    }

    if (warp == 4 && lane == 0) gstamp(g.trace, 7);
    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
}

// F32 activations -> X'[n][k] fp16 = fp16_rn(q * d) of their Q8_0 quantization (quantize_row_q8_0, src/ggml-quants.c:535-618,
// same explicitly rounded arithmetic as b200_quantize.cu), written in the k-order of the operand tiles: within every aligned
// group of four the order is (0, 2, 1, 3).  8 lanes per block, one 128-bit load and one 64-bit store per lane.
__global__ void __launch_bounds__(256) quantize_to_f16_kernel(const float *__restrict__ x, int64_t k, int64_t nrows, size_t row_stride, __half *__restrict__ out,
                                                              uint32_t *__restrict__ counters, int n_counters) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");     // the GEMM may take the SMs as they free up (it waits for X' itself)
    asm volatile("griddepcontrol.wait;" ::: "memory");                  // this grid's own launch overlaps the kernel in front; x, X' and the counters are that kernel's until here
    // the split-k arrival counters of the GEMM behind: zeroed here rather than by a memset between the two kernels, which would make the
    // GEMM the programmatic dependent of a copy-engine node instead of this grid (the GEMM touches them only after its griddepcontrol.wait)
    if (blockIdx.x == 0)
        for (int i = threadIdx.x; i < n_counters; i += blockDim.x) counters[i] = 0u;
    const int64_t nb = k / 32;
    const int64_t total = nrows * nb * 8;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < (int64_t)b200_align_up((size_t)total, 32); t += (int64_t)gridDim.x * blockDim.x) {
        const bool live = t < total;
        const int64_t blk = (live ? t : total - 1) >> 3;
        const int sub = (int)(t & 7);
        const int64_t row = blk / nb, b = blk - row * nb;
        const float *src = reinterpret_cast<const float *>(reinterpret_cast<const char *>(x) + row * row_stride) + b * 32 + sub * 4;
        const float4 v = *reinterpret_cast<const float4 *>(src);
        float amax = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 4));
        const float id = amax != 0.0f ? __fdiv_rn(127.f, amax) : 0.0f;
        const float d = __half2float(__float2half_rn(__fdiv_rn(amax, 127.f)));
        const float q0 = (float)__float2int_rn(__fmul_rn(v.x, id)), q1 = (float)__float2int_rn(__fmul_rn(v.y, id));
        const float q2 = (float)__float2int_rn(__fmul_rn(v.z, id)), q3 = (float)__float2int_rn(__fmul_rn(v.w, id));
        if (!live) continue;
        const __half2 h01 = __halves2half2(__float2half_rn(__fmul_rn(q0, d)), __float2half_rn(__fmul_rn(q2, d)));     // positions 0, 1 <- elements 0, 2
        const __half2 h23 = __halves2half2(__float2half_rn(__fmul_rn(q1, d)), __float2half_rn(__fmul_rn(q3, d)));     // positions 2, 3 <- elements 1, 3
        uint2 o;
        o.x = *reinterpret_cast<const uint32_t *>(&h01);
        o.y = *reinterpret_cast<const uint32_t *>(&h23);
        *reinterpret_cast<uint2 *>(out + row * k + b * 32 + sub * 4) = o;
    }
}

// a plane as bytes: rows x row_bytes; box = 128 rows x box_bytes with the given swizzle; out-of-bounds (rows past m, bytes past the
// end of a ragged k) -> zeros
bool make_rows_map(CUtensorMap *map, const void *base, int64_t rows, int64_t row_bytes, int box_bytes, CUtensorMapSwizzle swz) {
    auto fn = get_encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {(cuuint64_t)row_bytes, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)row_bytes};
    cuuint32_t box[2] = {(cuuint32_t)box_bytes, (cuuint32_t)TM};
    cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// X' as bytes: n rows x (k * 2) bytes; box = 128 rows x 128 bytes, 128-byte swizzle, out-of-bounds -> zeros
bool make_xp_map(CUtensorMap *map, const void *base, int64_t rows, int64_t row_bytes) {
    auto fn = get_encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {(cuuint64_t)row_bytes, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)row_bytes};
    cuuint32_t box[2] = {128u, (cuuint32_t)(TN / 2)};
    cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

Sched make_sched(int64_t m, int64_t n, int64_t k, int pairs) {
    Sched sc;
    sc.tiles_n = (int)((n + TN - 1) / TN);
    sc.tiles = (int)((m + 2 * TM - 1) / (2 * TM)) * sc.tiles_n;
    sc.pairs = pairs;
    sc.ksteps = (int)((k + KSTEP - 1) / KSTEP);
    sc.full_rounds = sc.tiles / pairs;
    sc.rem = sc.tiles % pairs;
    sc.split = 1;
    if (sc.rem > 0) {
        int split = pairs / sc.rem;
        const int max_split = sc.ksteps / 8 > 0 ? sc.ksteps / 8 : 1;      // at least 8 k-steps per slice
        if (split > max_split) split = max_split;
        if (split > 16) split = 16;
        sc.split = split < 1 ? 1 : split;
    }
    if (sc.split == 1) {      // nothing to gain: the partial round runs as whole tiles
        sc.full_rounds = (sc.tiles + pairs - 1) / pairs;
        sc.rem = 0;
    }
    return sc;
}

}  // namespace

// scratch of the fp16 path beyond the activations: X' [n][k] fp16, split-k partials, counters
size_t b200_gemm_f16_scratch_bytes(int64_t k, int64_t m, int64_t n, int sm_count) {
    const Sched sc = make_sched(m, n, k, sm_count / 2);
    return b200_align_up((size_t)n * k * 2, 1024) + (size_t)sc.rem * sc.split * TN * 2 * TM * 4 + 1024;
}

int b200_launch_gemm_f16(b200_ctx *ctx, int type, const uint8_t *qs, const __half *d, int64_t k, int64_t m, const float *x, int64_t n,
                         size_t x_row_stride, float *dst, void *scratch) {
    B200_REQUIRE(ctx, k % 32 == 0 && k >= 32 && m >= 1 && n >= 1, B200_ERR_INVALID);
    B200_REQUIRE(ctx, scratch != NULL && dst != NULL, B200_ERR_INVALID);
    B200_REQUIRE(ctx, m < (1 << 30) && n < (1 << 30) && k < (1 << 30), B200_ERR_UNSUPPORTED);
    B200_REQUIRE(ctx, ((uintptr_t)x & 15) == 0 && (x_row_stride & 15) == 0, B200_ERR_UNSUPPORTED);
    const int pairs = ctx->sm_count / 2;
    B200_REQUIRE(ctx, pairs >= 1, B200_ERR_UNSUPPORTED);
    const int64_t nb = k / 32;
    const int qsb = b200_qs_bytes(type);
    const Sched sc = make_sched(m, n, k, pairs);
    __half *xp = (__half *)scratch;
    float *partial = (float *)((uint8_t *)scratch + b200_align_up((size_t)n * k * 2, 1024));
    uint32_t *counters = (uint32_t *)((uint8_t *)partial + (size_t)sc.rem * sc.split * TN * 2 * TM * 4);
    {
        const int64_t total = n * nb * 8;
        int64_t grid = (total + 255) / 256;
        const int64_t cap = (int64_t)ctx->sm_count * 16;
        if (grid > cap) grid = cap;
        cudaLaunchConfig_t qcfg;
        memset(&qcfg, 0, sizeof(qcfg));
        qcfg.gridDim = dim3((unsigned)grid, 1, 1);
        qcfg.blockDim = dim3(256, 1, 1);
        qcfg.stream = ctx->stream;
        cudaLaunchAttribute qattr[1];
        qattr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        qattr[0].val.programmaticStreamSerializationAllowed = 1;
        qcfg.attrs = qattr;
        qcfg.numAttrs = ctx->opt_pdl ? 1 : 0;
        B200_CUDA_TRY(ctx, cudaLaunchKernelEx(&qcfg, quantize_to_f16_kernel, x, k, n, x_row_stride, xp, counters, (int)sc.rem));
        ctx->launches++;
        B200_CUDA_TRY(ctx, cudaGetLastError());
    }
    CUtensorMap map_raw, map_b, map_raw_pf, map_dw_pf;
    B200_REQUIRE(ctx, ((uintptr_t)qs & 15) == 0 && ((uintptr_t)d & 1) == 0, B200_ERR_UNSUPPORTED);
    const bool dw_pf = (nb * 2) % 16 == 0 && nb >= 16 && ((uintptr_t)d & 15) == 0;
    const int64_t row_bytes = nb * qsb;
    const int pf_box = row_bytes >= 256 ? 256 : (int)row_bytes;
    if (!make_rows_map(&map_raw, qs, m, row_bytes, 2 * qsb, type == B200_TYPE_Q4_0 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_64B) ||
        !make_xp_map(&map_b, xp, n, k * 2) || !make_rows_map(&map_raw_pf, qs, m, row_bytes, pf_box, CU_TENSOR_MAP_SWIZZLE_NONE) ||
        !(dw_pf ? make_rows_map(&map_dw_pf, d, m, nb * 2, 32, CU_TENSOR_MAP_SWIZZLE_NONE) : make_rows_map(&map_dw_pf, qs, m, row_bytes, pf_box, CU_TENSOR_MAP_SWIZZLE_NONE))) {
        b200_set_error(ctx, "cuTensorMapEncodeTiled failed (fp16 prefill path, m=%lld n=%lld k=%lld)", (long long)m, (long long)n, (long long)k);
        return B200_ERR_CUDA;
    }
    GemmF16Args g;
    memset(&g, 0, sizeof(g));
    g.dw = d;
    g.dw_pf = dw_pf ? 1 : 0;
    g.dst = dst;
    g.partial = partial;
    g.counters = counters;
    g.abort_flag = ctx->abort_dev;
    g.abort_host = ctx->abort_host_dev;
    g.m = (int)m;
    g.n = (int)n;
    g.k = (int)k;
    g.sc = sc;
    if (ctx->trace && ctx->trace_next < ctx->trace_capacity) g.trace = ctx->trace + (size_t)(ctx->trace_next++) * B200_TRACE_MAX_CTAS * B200_TRACE_STAMPS;
    // one CTA pair per TPC, or fewer when the launch has fewer units than pairs (idle pairs would only spin up and exit)
    int use_pairs = pairs;
    const int units = sc.full_rounds > 0 ? (sc.tiles < pairs ? sc.tiles : pairs) : sc.rem * sc.split;
    if (units < use_pairs) use_pairs = units;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(2 * use_pairs), 1, 1);
    cfg.blockDim = dim3(kThreadsF, 1, 1);
    cfg.dynamicSmemBytes = type == B200_TYPE_Q4_0 ? RawTile<B200_TYPE_Q4_0>::kSmem : RawTile<B200_TYPE_Q8_0>::kSmem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;      // overlap the prologue and the first weight stages with the quantize kernel
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = ctx->opt_pdl ? 1 : 0;
    if (type == B200_TYPE_Q4_0) {
        B200_SMEM_LIMIT_ONCE(ctx, gemm_f16_pair_kernel<B200_TYPE_Q4_0>, RawTile<B200_TYPE_Q4_0>::kSmem);
        B200_CUDA_TRY(ctx, cudaLaunchKernelEx(&cfg, gemm_f16_pair_kernel<B200_TYPE_Q4_0>, map_raw, map_b, map_raw_pf, map_dw_pf, g));
    } else {
        B200_SMEM_LIMIT_ONCE(ctx, gemm_f16_pair_kernel<B200_TYPE_Q8_0>, RawTile<B200_TYPE_Q8_0>::kSmem);
        B200_CUDA_TRY(ctx, cudaLaunchKernelEx(&cfg, gemm_f16_pair_kernel<B200_TYPE_Q8_0>, map_raw, map_b, map_raw_pf, map_dw_pf, g));
    }
    ctx->launches++;
    B200_CUDA_TRY(ctx, cudaGetLastError());
    return B200_OK;
}
