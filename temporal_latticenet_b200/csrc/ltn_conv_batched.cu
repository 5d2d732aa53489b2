// Fused lattice convolution for SEVERAL lattices per launch: persistent CTAs over one tile queue.
//
//   out_b[v, f] = sum_{s<S} sum_{c<C} act_b( x_b[nbr_b[v,s], c] ) * W[s*C + c, f]  (+ bias[f]) (+ res_b[v, f])      b < nb
//
// Same contraction, operands and numerics as k_conv_tc<3, ATMEM, F16> of ltn_conv.cu (fp16 hi/lo split, three
// tcgen05 passes, A operand written from the gather registers straight into tensor memory, weights streamed by TMA):
// each output tile sees the same MMAs in the same k order, so the two kernels agree BIT FOR BIT on `out`.
// What changes is the shape of the launch (SURVEY.md 8b "Threading: B lattices per launch"):
//  * one launch serves the same layer of all `nb` independent windows in flight (identical weights, per-window
//    values / neighbour tables / GroupNorm statistics / device-side row counts in a small parameter table); the
//    window runner used to issue one 108-172-CTA launch per window and layer, each a partial wave of one CTA per SM;
//  * a fixed grid of at most one CTA per SM walks the tile list (problem, row tile, channel tile) round-robin, so
//    barrier set-up, tensor-memory allocation and pipeline fill are paid once per CTA instead of once per tile, and the
//    tail quantisation is that of ALL windows' tiles together;
//  * warp-specialised roles with register reallocation (setmaxnreg): 2 producer warpgroups (gather -> tensor memory),
//    1 epilogue warpgroup, and a fourth holding the MMA issuer, the TMA weight streamer and two "tile prologue" warps
//    that stage the next tile's neighbour slice and folded GroupNorm affine while the producers work on the current one;
//  * TWO accumulators in tensor memory: tile i's epilogue (TMEM -> registers -> swizzled transpose -> 128-byte lines,
//    residual, output statistics) runs under tile i+1's main loop.
#include "ltn_common.cuh"
#include "ltn_conv_common.cuh"
#include <cstdlib>
#include <cstring>

namespace {

constexpr int kBlockM = 128;
constexpr int kKB = 64;               // fp16 K elements per k-block = one 128-byte swizzle row
constexpr int kGroups = 2;            // producer warpgroups, alternate k-blocks
constexpr int kMaxBatch = 8;
constexpr int kWarpEpi0 = 8;          // warps 0..7: the two producer warpgroups
constexpr int kMetaThreads = 64;
constexpr int kMetaBufs = 3;          // neighbour-slice buffers: the gather's issue cursor runs up to two tiles ahead of its consume cursor
constexpr int kNbrPerMeta = (LTN_FEXT * kBlockM + kMetaThreads - 1) / kMetaThreads;
constexpr int kColMax = 192;          // output channels per tile at most (two accumulators + the A ring in 512 columns of tensor memory)
constexpr int kMaxSA = 4, kMaxSB = 9;   // weight ring: up to 9 stages so that a whole small weight tile (all k-blocks) can stay RESIDENT

// Role layout and register budgets for EPI = 4 or 8 epilogue warps (one or two warpgroups; with two, they take alternate
// 32-column chunks).  Budgets after setmaxnreg: the pool the roles share is what the CTA was given AT LAUNCH -- threads x
// the launch-time register count (65536 / threads rounded down to a multiple of 8 under __launch_bounds__(threads, 1)); a
// setmaxnreg.inc that asks for more than the decs have released never returns, so the budgets must sum to at most that.
template <int EPI>
struct Roles {
    static constexpr int kWarpMma = kWarpEpi0 + EPI, kWarpTma = kWarpMma + 1, kWarpMeta0 = kWarpMma + 2;
    static constexpr int kThreads = 32 * (kWarpMma + 4);   // producers (2 warpgroups) + epilogue (1 or 2) + auxiliary warpgroup
    static constexpr int kLaunchRegs = (65536 / kThreads) / 8 * 8;
    static constexpr int kRegsProducer = EPI == 8 ? 144 : 168, kRegsEpilogue = EPI == 8 ? 72 : 96, kRegsAux = EPI == 8 ? 48 : 56;
    static_assert(EPI == 4 || EPI == 8, "one or two epilogue warpgroups");
    static_assert(32 * (8 * kRegsProducer + EPI * kRegsEpilogue + 4 * kRegsAux) <= kThreads * kLaunchRegs,
                  "the per-role register budgets exceed the CTA's register pool: setmaxnreg.inc would spin forever");
};

struct BatchParams {
    const float* x[kMaxBatch];        // [Vx_b, C]
    const int* nbr[kMaxBatch];        // [Vq_b, S] or null
    const double* gn_sums[kMaxBatch]; // [G,2] or null
    const float* res[kMaxBatch];      // [Vq_b, F] or null
    float* out[kMaxBatch];            // [Vq_b, ldo]
    double* out_sums[kMaxBatch];      // [Gout,2] or null
    const int* vq_dev[kMaxBatch];     // device-side row counts (nullable)
    const int* vx_dev[kMaxBatch];
    int* flag[kMaxBatch];             // fp16 range flags (nullable)
    int Vq[kMaxBatch];
    int Vx[kMaxBatch];
    int nb;
    const float* gn_gamma;
    const float* gn_beta;
    const float* bias;
    float gn_eps;
    int gn_cpg, out_cpg;
    int C, S, F, ldo, relu, has_gn, has_sums;
    int n_tile, ny;
    int stages_a, stages_b, acc_bufs, acc_stride;
    int staged;                       // x[] hold PRE-STAGED operands (k_stage_a): per 4 channels one 16-byte quad {hi01, hi23, lo01, lo23}
    int res_async;                    // residual rows of the NEXT 32-column chunk prefetched with cp.async (4 epilogue warps, pre-staged operands)
    int resident;                     // stages_b == number of k-blocks: a channel tile's weights are loaded once per CTA and kept across its row tiles
    float a_mul, out_mul;
    int debug;                        // timing experiments only (LTN_CONVB_DEBUG): 1 no gather loads, 2 no staging, 4 no MMA issue, 8 no epilogue work
    unsigned long long* detail;       // nullable: [64 tiles][16] globaltimer stamps of CTA 0's roles (ltn_conv_batched_detail)
    unsigned long long* trace;        // nullable: this launch's trace record: [sum of live rows, tiles, then per CTA (entry, exit)] globaltimer ns
};

struct Tile {
    int b, row0, n0, N, Nmma;
};

constexpr int kTraceStride = 2 + 2 * 148;   // u64 per launch record
unsigned long long* g_trace_b = nullptr;
unsigned long long* g_detail_b = nullptr;
int g_trace_slots = 0, g_trace_next = 0;

__device__ __forceinline__ void detail_stamp(unsigned long long* buf, int it, int slot) {
    if (buf && blockIdx.x == 0 && it < 64) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        buf[it * 16 + slot] = t;
    }
}

template <int kEpiWarps>
__global__ void __launch_bounds__(Roles<kEpiWarps>::kThreads, 1)
k_conv_tc_batched(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
                  const __grid_constant__ BatchParams p) {
    using R = Roles<kEpiWarps>;
    constexpr int kThreadsB = R::kThreads, kWarpMma = R::kWarpMma, kWarpTma = R::kWarpTma, kWarpMeta0 = R::kWarpMeta0;
    constexpr int kRegsProducer = R::kRegsProducer, kRegsEpilogue = R::kRegsEpilogue, kRegsAux = R::kRegsAux;
    // The role layout fixes the operand form: 4 epilogue warps <=> pre-staged operands (the gathering layers), 8 <=> fp32 rows
    // converted in the gather loop (the dense layers).  Each instantiation carries ONE producer pipeline: the kernel's
    // instruction footprint is what its 20 warps in five roles share one instruction cache for (ncu: "no instruction" was
    // the top stall reason with both pipelines and the timing switches compiled in).
    constexpr bool kStaged = kEpiWarps == 4;
#ifdef LTN_CONVB_DEBUG_BUILD
    const int dbg = p.debug;
#else
    constexpr int dbg = 0;
#endif
    extern __shared__ uint8_t smem_raw[];
    // barriers: a_full[4] a_empty[4] b_full[4] b_empty[4] acc_full[2] acc_empty[2] meta_full[3] meta_empty[3]
    __shared__ __align__(8) uint64_t bars[2 * kMaxSA + 2 * kMaxSB + 4 + 2 * kMetaBufs];
    __shared__ uint32_t tmem_slot;
    __shared__ int s_prefix[kMaxBatch + 1];   // row tiles before problem b
    __shared__ int s_vq[kMaxBatch], s_vx[kMaxBatch];
    __shared__ __align__(16) float s_affine[kMaxBatch][512];  // per PROBLEM: folded GroupNorm scale [256] | shift [256], computed once per CTA
    __shared__ int s_nbr[kMetaBufs][LTN_FEXT * kBlockM];      // per meta buffer: [slot][tile row], -1 = absent
    // Output statistics, per tile parity and row quadrant: column sums [192] | sums of squares [192].  Every (quadrant, column)
    // has exactly ONE writer per tile (the epilogue warp of that quadrant and 32-column chunk), so the slots are plainly
    // stored, never cleared and never added to: shared-memory float atomics compile to a compare-and-swap loop, and eight of
    // them per lane and chunk were a third of the epilogue's instructions (ncu source page, round 2).
    __shared__ __align__(16) float s_colsum[2][4][2 * kColMax];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid < p.nb) {
        s_vq[tid] = p.vq_dev[tid] ? min(p.Vq[tid], __ldg(p.vq_dev[tid])) : p.Vq[tid];
        s_vx[tid] = p.vx_dev[tid] ? min(p.Vx[tid], __ldg(p.vx_dev[tid])) : p.Vx[tid];
    }
    __syncthreads();
    if (p.has_gn) {
        // GroupNorm folded to scale = rstd*gamma, shift = beta - mean*scale (statistics over all Vx rows), x the fp16 pre-scale
        for (int i = tid; i < p.nb * p.C; i += kThreadsB) {
            const int b = i / p.C, c = i - b * p.C;
            const int g = c / p.gn_cpg;
            const double* sums = p.gn_sums[b];
            const double n = (double)s_vx[b] * p.gn_cpg;
            const double mean = sums[2 * g] / n;
            double var = sums[2 * g + 1] / n - mean * mean;
            if (var < 0.0) var = 0.0;
            const float rstd = (float)(1.0 / sqrt(var + (double)p.gn_eps));
            const float sc = rstd * (p.gn_gamma ? __ldg(p.gn_gamma + c) : 1.0f);
            const float sh = (p.gn_beta ? __ldg(p.gn_beta + c) : 0.0f) - (float)mean * sc;
            s_affine[b][c] = sc * p.a_mul;
            s_affine[b][256 + c] = sh * p.a_mul;
        }
    }
    if (tid == 0) {
        int acc = 0;
        for (int b = 0; b < p.nb; ++b) {
            s_prefix[b] = acc;
            acc += (max(s_vq[b], 0) + kBlockM - 1) / kBlockM;
        }
        for (int b = p.nb; b <= kMaxBatch; ++b) s_prefix[b] = acc;
    }
    __syncthreads();
    const int total_tiles = s_prefix[p.nb] * p.ny;
    if ((int)blockIdx.x >= total_tiles) return;   // uniform per CTA, before any barrier / tensor-memory allocation

    const int C = p.C, S = p.S;
    const int num_kb = S * C / kKB;
    const int kb_per_slot = C / kKB;
    const int SA = p.stages_a, SB = p.stages_b;
    const uint32_t b_bytes = (uint32_t)p.n_tile * 128u, b_stage = 2u * b_bytes;
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* smem_epi = smem + (size_t)SB * b_stage;   // kEpiWarps x 4 KB: one transpose block per epilogue warp

    const uint32_t bar_afull = smem_u32(&bars[0]), bar_aempty = smem_u32(&bars[kMaxSA]);
    const uint32_t bar_bfull = smem_u32(&bars[2 * kMaxSA]), bar_bempty = smem_u32(&bars[2 * kMaxSA + kMaxSB]);
    const uint32_t bar_accfull = smem_u32(&bars[2 * kMaxSA + 2 * kMaxSB]), bar_accempty = bar_accfull + 16;
    const uint32_t bar_metafull = bar_accfull + 32, bar_metaempty = bar_metafull + 8 * kMetaBufs;
    const uint32_t a_col0 = (uint32_t)(p.acc_bufs * p.acc_stride);   // A ring behind the accumulators, 64 columns per stage
    constexpr uint32_t kTmemCols = 512;

    auto tile_at = [&](int it, Tile& t) -> bool {
        const int idx = (int)blockIdx.x + it * (int)gridDim.x;
        if (idx >= total_tiles) return false;
        const int row_tiles = s_prefix[p.nb];
        const int n = idx / row_tiles, r = idx - n * row_tiles;   // channel tile slowest: a CTA's consecutive tiles share their weights
        int b = 0;
        while (b + 1 < p.nb && r >= s_prefix[b + 1]) ++b;
        t.b = b;
        t.row0 = (r - s_prefix[b]) * kBlockM;
        t.n0 = n * p.n_tile;
        t.N = min(p.n_tile, p.F - t.n0);
        t.Nmma = (t.N + 15) & ~15;
        return true;
    };

    if (tid == 0) {
        for (int s = 0; s < SA; ++s) {
            mbar_init(bar_afull + 8 * s, 4);      // one arrive per gather warp of the owning group
            mbar_init(bar_aempty + 8 * s, 1);     // tcgen05.commit
        }
        for (int s = 0; s < SB; ++s) {
            mbar_init(bar_bfull + 8 * s, 1);      // the TMA thread's expect_tx arrive
            mbar_init(bar_bempty + 8 * s, 1);     // tcgen05.commit
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(bar_accfull + 8 * s, 1);            // tcgen05.commit
            mbar_init(bar_accempty + 8 * s, kEpiWarps);   // the epilogue warps
        }
        for (int s = 0; s < kMetaBufs; ++s) {
            mbar_init(bar_metafull + 8 * s, 2);   // the two tile-prologue warps
            mbar_init(bar_metaempty + 8 * s, 8);  // the eight gather warps
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kWarpMma) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    unsigned long long t_first = 0;
    if (p.trace && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_first));

    if (warp < kWarpEpi0) {
        // ===================== producers: gather A rows into tensor memory ==================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsProducer));
        const int group = warp >> 2;
        const int wrow0 = (warp & 3) * 32;
        const int ra = lane >> 2, cq = lane & 3;
        const uint32_t c4 = (uint32_t)C >> 2;
        const float a_mul = p.a_mul;
        const bool affine = p.has_gn != 0;
        // The gather is ONE software pipeline over all tiles of this CTA: the loads of the next unit (16 tile rows x 64
        // channels) are in flight while the current one is converted and written to tensor memory -- also across a tile
        // boundary, so the L2 latency of a tile's first rows hides under the previous tile's last k-block.  Two cursors walk
        // the k-blocks this group owns (those of its parity in the GLOBAL k-block sequence): `i_*` where loads are issued,
        // `c_*` (one k-block behind) where they are consumed.
        int c_stage = group % SA;
        uint32_t c_par = 1u;
        // issue cursor
        int i_it = -1, i_left = 0, i_slot = 0, i_c0 = 0, i_cur = -1;
        uint32_t i_base = 0;
        bool i_valid = false;
        const float4* i_x4 = nullptr;
        const int* i_nbr = nullptr;
        uint32_t rowidx[4];
        uint32_t rmask = 0;
        // consume cursor
        int c_it = -1, c_left = 0, c_c0 = 0;
        uint32_t c_base = 0;
        bool c_valid = false;
        const float* c_aff = s_affine[0];
        int* c_flag = nullptr;
        float amax = 0.f;
        auto issue_next_tile = [&]() {
            for (;;) {
                if (i_it >= 0) i_base += (uint32_t)num_kb;
                ++i_it;
                Tile t;
                if (!tile_at(i_it, t)) { i_valid = false; return; }
                const int q = i_it % kMetaBufs;
                if (lane == 0) mbar_wait(bar_metafull + 8 * q, (uint32_t)(i_it / kMetaBufs) & 1u);   // every tile, owned or not
                __syncwarp();
                const int first = (group + (int)(i_base & 1u)) & 1;
                if (first >= num_kb) continue;
                i_left = (num_kb - first + 1) >> 1;
                i_x4 = reinterpret_cast<const float4*>(p.x[t.b]);
                i_nbr = s_nbr[q];
                i_slot = first / kb_per_slot;
                i_c0 = (first - i_slot * kb_per_slot) * kKB;
                i_cur = -1;
                i_valid = true;
                if (tid == 0) detail_stamp(p.detail, i_it, 0);
                return;
            }
        };
        auto consume_next_tile = [&]() {
            for (;;) {
                if (c_it >= 0) {   // leaving tile c_it: its neighbour slice is no longer read by this warp (the issue cursor is ahead)
                    if (tid == 0) detail_stamp(p.detail, c_it, 2);
                    if (c_flag && !(amax < 65504.f)) atomicOr(c_flag, 1);
                    amax = 0.f;
                    c_flag = nullptr;
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_metaempty + 8 * (c_it % kMetaBufs));
                    c_base += (uint32_t)num_kb;
                }
                ++c_it;
                Tile t;
                if (!tile_at(c_it, t)) { c_valid = false; return; }
                const int first = (group + (int)(c_base & 1u)) & 1;
                if (first >= num_kb) continue;
                c_left = (num_kb - first + 1) >> 1;
                c_aff = s_affine[t.b];
                c_flag = p.flag[t.b];
                const int sl = first / kb_per_slot;
                c_c0 = (first - sl * kb_per_slot) * kKB;
                c_valid = true;
                if (tid == 0) detail_stamp(p.detail, c_it, 1);
                return;
            }
        };
        float4 buf[2][8];
        uint32_t bmask[2];
        auto issue = [&](float4* dst, uint32_t& dmask, const int half) {
            if (i_slot != i_cur) {
                i_cur = i_slot;
                const int* tap = i_nbr + i_slot * kBlockM + wrow0 + ra;
                rmask = 0;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int sv = tap[8 * r];
                    rowidx[r] = (uint32_t)(sv >= 0 ? sv : 0) * c4 + (uint32_t)cq;   // row 0 stands in for an absent neighbour
                    rmask |= (sv >= 0 ? 1u : 0u) << r;
                }
            }
            const uint32_t o = (uint32_t)i_c0 >> 2;
            dmask = half ? (rmask >> 2) : rmask;
            if (!(dbg & 1)) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t ri = (half ? rowidx[2 + h] : rowidx[h]) + o;
#pragma unroll
                    for (int g = 0; g < 4; ++g) dst[4 * h + g] = __ldg(i_x4 + ri + 4 * g);
                }
            }
            if (!half) return;
            if (--i_left == 0) { issue_next_tile(); return; }
            i_c0 += kGroups * kKB;
            while (i_c0 >= C) { i_c0 -= C; ++i_slot; }
        };
        auto consume = [&](const float4* cur, uint32_t cmask, const int half) {
            if (half == 0) {
                if (lane == 0) mbar_wait(bar_aempty + 8 * c_stage, c_par);
                __syncwarp();
                tc_fence_after();
            }
            uint32_t hi[16], lo[16];
            if (!(dbg & 2)) {
            {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                float4 sc = make_float4(a_mul, a_mul, a_mul, a_mul), sh = make_float4(0.f, 0.f, 0.f, 0.f);
                if (affine) {
                    sc = *reinterpret_cast<const float4*>(c_aff + c_c0 + 16 * g + 4 * cq);
                    sh = *reinterpret_cast<const float4*>(c_aff + 256 + c_c0 + 16 * g + 4 * cq);
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float4 v = cur[4 * h + g];
                    v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y);
                    v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
                    if (p.relu) { v.x = relu_nan(v.x); v.y = relu_nan(v.y); v.z = relu_nan(v.z); v.w = relu_nan(v.w); }
                    if (!((cmask >> h) & 1u)) v = make_float4(0.f, 0.f, 0.f, 0.f);   // absent neighbour / row beyond the tile
                    amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
                    const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
                    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                    const __half2 l01 = __floats2half2_rn(v.x - f01.x, v.y - f01.y), l23 = __floats2half2_rn(v.z - f23.x, v.w - f23.y);
                    hi[4 * g + 2 * h] = *reinterpret_cast<const uint32_t*>(&h01);
                    hi[4 * g + 2 * h + 1] = *reinterpret_cast<const uint32_t*>(&h23);
                    lo[4 * g + 2 * h] = *reinterpret_cast<const uint32_t*>(&l01);
                    lo[4 * g + 2 * h + 1] = *reinterpret_cast<const uint32_t*>(&l23);
                }
            }
            }
            const uint32_t ta = tmem_base + ((uint32_t)(wrow0 + 16 * half) << 16) + a_col0 + (uint32_t)c_stage * 64u;
            tmem_st_16x256b_x4(ta, hi);
            tmem_st_16x256b_x4(ta + 32u, lo);
            }
            if (half == 0) return;
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_afull + 8 * c_stage);
            c_stage += kGroups;
            while (c_stage >= SA) { c_stage -= SA; c_par ^= 1u; }
            if (--c_left == 0) { consume_next_tile(); return; }
            c_c0 += kGroups * kKB;
            while (c_c0 >= C) c_c0 -= C;
        };
        if constexpr (kStaged) {
            // ---- pre-staged operands (k_stage_a): the gather only moves 16-byte quads, so what bounds it is the number of
            // loads in flight.  THREE register buffers: the loads run two units (one k-block) ahead of the tensor-memory
            // stores.  The issue cursor changes tile LAZILY (at its next first-half issue), so that it never waits for a
            // neighbour slice whose buffer this very warp has not released yet.
            float4 sb3[3][8];
            uint32_t sm3[3];
            bool sv3[3] = {false, false, false};
            bool i_pending = false;
            auto issue_s = [&](float4* dst, uint32_t& dmask, bool& dvalid, const int half) {
                if (!half) {
                    if (i_pending) { issue_next_tile(); i_pending = false; }
                    if (!i_valid) { dvalid = false; return; }
                } else if (!i_valid) { dvalid = false; return; }
                dvalid = true;
                if (i_slot != i_cur) {
                    i_cur = i_slot;
                    const int* tap = i_nbr + i_slot * kBlockM + wrow0 + ra;
                    rmask = 0;
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const int sv = tap[8 * r];
                        rowidx[r] = (uint32_t)(sv >= 0 ? sv : 0) * c4 + (uint32_t)cq;
                        rmask |= (sv >= 0 ? 1u : 0u) << r;
                    }
                }
                const uint32_t o = (uint32_t)i_c0 >> 2;
                dmask = half ? (rmask >> 2) : rmask;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t ri = (half ? rowidx[2 + h] : rowidx[h]) + o;
#pragma unroll
                    for (int g = 0; g < 4; ++g) dst[4 * h + g] = __ldg(i_x4 + ri + 4 * g);
                }
                if (!half) return;
                if (--i_left == 0) { i_pending = true; return; }
                i_c0 += kGroups * kKB;
                while (i_c0 >= C) { i_c0 -= C; ++i_slot; }
            };
            auto consume_s = [&](const float4* cur, uint32_t cmask, bool cvalid, const int half) -> bool {
                if (!cvalid || !c_valid) return false;
                if (half == 0) {
                    if (lane == 0) mbar_wait(bar_aempty + 8 * c_stage, c_par);
                    __syncwarp();
                    tc_fence_after();
                }
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const bool ok = (cmask >> h) & 1u;   // absent neighbour / row beyond the tile -> zeros
                        const float4 v = cur[4 * h + g];
                        hi[4 * g + 2 * h] = ok ? __float_as_uint(v.x) : 0u;
                        hi[4 * g + 2 * h + 1] = ok ? __float_as_uint(v.y) : 0u;
                        lo[4 * g + 2 * h] = ok ? __float_as_uint(v.z) : 0u;
                        lo[4 * g + 2 * h + 1] = ok ? __float_as_uint(v.w) : 0u;
                    }
                }
                const uint32_t ta = tmem_base + ((uint32_t)(wrow0 + 16 * half) << 16) + a_col0 + (uint32_t)c_stage * 64u;
                tmem_st_16x256b_x4(ta, hi);
                tmem_st_16x256b_x4(ta + 32u, lo);
                if (half == 0) return true;
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_afull + 8 * c_stage);
                c_stage += kGroups;
                while (c_stage >= SA) { c_stage -= SA; c_par ^= 1u; }
                if (--c_left == 0) consume_next_tile();
                return true;
            };
            issue_next_tile();
            consume_next_tile();
            issue_s(sb3[0], sm3[0], sv3[0], 0);
            issue_s(sb3[1], sm3[1], sv3[1], 1);
            for (;;) {   // unit n is consumed while units n+1, n+2 are in flight; buffers and halves repeat every six units
                issue_s(sb3[2], sm3[2], sv3[2], 0); if (!consume_s(sb3[0], sm3[0], sv3[0], 0)) break;
                issue_s(sb3[0], sm3[0], sv3[0], 1); if (!consume_s(sb3[1], sm3[1], sv3[1], 1)) break;
                issue_s(sb3[1], sm3[1], sv3[1], 0); if (!consume_s(sb3[2], sm3[2], sv3[2], 0)) break;
                issue_s(sb3[2], sm3[2], sv3[2], 1); if (!consume_s(sb3[0], sm3[0], sv3[0], 1)) break;
                issue_s(sb3[0], sm3[0], sv3[0], 0); if (!consume_s(sb3[1], sm3[1], sv3[1], 0)) break;
                issue_s(sb3[1], sm3[1], sv3[1], 1); if (!consume_s(sb3[2], sm3[2], sv3[2], 1)) break;
            }
        } else {
        issue_next_tile();
        consume_next_tile();
        if (i_valid) issue(buf[0], bmask[0], 0);
        while (c_valid) {
            issue(buf[1], bmask[1], 1);     // second half of the k-block the issue cursor is on; moves the cursor on (maybe to the next tile)
            consume(buf[0], bmask[0], 0);
            if (i_valid) issue(buf[0], bmask[0], 0);
            consume(buf[1], bmask[1], 1);   // completes the k-block; moves the consume cursor on
        }
        }
    } else if (warp < kWarpMma) {
        // ===================== epilogue warpgroup: TMEM -> registers -> global ==============================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsEpilogue));
        const int qd = warp & 3;
        const int eh = (warp - kWarpEpi0) >> 2;   // which share of the 32-column chunks (0 when there is one epilogue warpgroup)
        const int et = tid - kWarpEpi0 * 32;      // 0 .. 32 * kEpiWarps - 1
        const int chunk = lane & 7, sub = lane >> 3;
        float* stage = reinterpret_cast<float*>(smem_epi + (warp - kWarpEpi0) * 4096);
        const uint32_t t_lane = tmem_base + ((uint32_t)(qd * 32) << 16);
        const float out_mul = p.out_mul;
        // Residual prefetch (res_async): a synchronous __ldg per row in the store loop made the residual layers 1.5x slower
        // than their twins without one (64 -> 64 on 56k rows: 39 against 26 us; 192 -> 192 on 93k rows: 215 against 144 us).
        // The residual quad a lane adds to row r = 4 * i8 + sub, columns 4 * chunk .. + 3 is fetched by THAT lane one chunk
        // ahead with cp.async (also across a tile boundary, i.e. under the wait for the next accumulator) into one of two
        // 4 KB buffers per warp: the first is the folded-GroupNorm table, unused on pre-staged operands, the second follows
        // the transpose blocks in dynamic shared memory.
        const bool ra = p.res_async != 0;
        const uint8_t* rbuf0 = reinterpret_cast<const uint8_t*>(&s_affine[0][0]) + ((warp - kWarpEpi0) & 3) * 4096 + lane * 16;
        const uint8_t* rbuf1 = smem_epi + kEpiWarps * 4096 + ((warp - kWarpEpi0) & 3) * 4096 + lane * 16;
        int rq = 0;
        auto res_issue = [&](const Tile& tt, const int cb, const int q) {
            const float* res = p.res[tt.b];
            if (res && 4 * chunk < tt.N - cb) {
                const int vq = s_vq[tt.b];
                const float* src = res + (size_t)(tt.row0 + qd * 32 + sub) * p.F + tt.n0 + cb + 4 * chunk;
                const uint32_t dst = smem_u32(q ? rbuf1 : rbuf0);
#pragma unroll
                for (int i8 = 0; i8 < 8; ++i8)
                    if (tt.row0 + qd * 32 + i8 * 4 + sub < vq)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)i8 * 512u), "l"(src + (size_t)i8 * 4 * p.F) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        if (ra) {
            Tile t0;
            if (tile_at(0, t0) && eh * 32 < t0.N) res_issue(t0, eh * 32, 0);
            else asm volatile("cp.async.commit_group;" ::: "memory");
        }
        for (int it = 0;; ++it) {
            Tile t;
            if (!tile_at(it, t)) break;
            const int a = it % p.acc_bufs, use = it / p.acc_bufs;
            if (lane == 0) mbar_wait(bar_accfull + 8 * a, (uint32_t)use & 1u);
            __syncwarp();
            tc_fence_after();
            if (warp == kWarpEpi0 && lane == 0) detail_stamp(p.detail, it, 6);
            const int Vq = s_vq[t.b];
            float* out = p.out[t.b];
            const float* res = p.res[t.b];
            float* colsum = s_colsum[it & 1][qd];
            for (int cb = eh * 32; cb < t.N && !(dbg & 8); cb += 32 * (kEpiWarps / 4)) {
                const bool col_ok = 4 * chunk < t.N - cb;
                const int col = t.n0 + cb + 4 * chunk;
                float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.bias && col_ok) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
                if (ra) {   // the next chunk's residual rows (this tile's, or the first chunk of the next tile) go on their way now
                    int ncb = cb + 32 * (kEpiWarps / 4);
                    Tile tn = t;
                    bool more = true;
                    if (ncb >= t.N) {
                        ncb = eh * 32;
                        more = tile_at(it + 1, tn) && ncb < tn.N;
                    }
                    if (more) res_issue(tn, ncb, rq ^ 1);
                    else asm volatile("cp.async.commit_group;" ::: "memory");
                }
                float acc[32];
                tmem_ld32(t_lane + (uint32_t)(a * p.acc_stride + cb), acc);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4*>(stage + lane * 32 + ((j ^ (lane & 7)) << 2)) =
                        make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
                if (ra) asm volatile("cp.async.wait_group 1;" ::: "memory");   // all but the group just committed: this chunk's rows are in
                __syncwarp();
                const uint8_t* rcur = rq ? rbuf1 : rbuf0;
                rq ^= ra ? 1 : 0;
                float4 cs = make_float4(0.f, 0.f, 0.f, 0.f), cq2 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int i8 = 0; i8 < 8; ++i8) {
                    const int r = i8 * 4 + sub;
                    const int v = t.row0 + qd * 32 + r;
                    float4 o = *reinterpret_cast<const float4*>(stage + r * 32 + ((chunk ^ (r & 7)) << 2));
                    o.x *= out_mul; o.y *= out_mul; o.z *= out_mul; o.w *= out_mul;   // exact: power of two
                    o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
                    if (v < Vq && col_ok) {
                        if (res) {
                            const float4 rr = ra ? *reinterpret_cast<const float4*>(rcur + i8 * 512)
                                                 : __ldg(reinterpret_cast<const float4*>(res + (size_t)v * p.F + col));
                            o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
                        }
                        *reinterpret_cast<float4*>(out + (size_t)v * p.ldo + col) = o;
                        cs.x += o.x; cs.y += o.y; cs.z += o.z; cs.w += o.w;
                        cq2.x = fmaf(o.x, o.x, cq2.x); cq2.y = fmaf(o.y, o.y, cq2.y); cq2.z = fmaf(o.z, o.z, cq2.z); cq2.w = fmaf(o.w, o.w, cq2.w);
                    }
                }
                __syncwarp();   // the block is consumed before the next chunk overwrites it
                if (p.has_sums) {
#pragma unroll
                    for (int o = 8; o <= 16; o <<= 1) {
                        cs.x += __shfl_xor_sync(0xffffffffu, cs.x, o); cs.y += __shfl_xor_sync(0xffffffffu, cs.y, o);
                        cs.z += __shfl_xor_sync(0xffffffffu, cs.z, o); cs.w += __shfl_xor_sync(0xffffffffu, cs.w, o);
                        cq2.x += __shfl_xor_sync(0xffffffffu, cq2.x, o); cq2.y += __shfl_xor_sync(0xffffffffu, cq2.y, o);
                        cq2.z += __shfl_xor_sync(0xffffffffu, cq2.z, o); cq2.w += __shfl_xor_sync(0xffffffffu, cq2.w, o);
                    }
                    if (sub == 0 && col_ok) {
                        *reinterpret_cast<float4*>(colsum + cb + 4 * chunk) = cs;
                        *reinterpret_cast<float4*>(colsum + kColMax + cb + 4 * chunk) = cq2;
                    }
                }
            }
            // the accumulator has been read: hand it back to the MMA issuer before the statistics are flushed
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_accempty + 8 * a);
            if (warp == kWarpEpi0 && lane == 0) detail_stamp(p.detail, it, 7);
            if (p.has_sums) {
                // per-tile column sums of the four row quadrants -> group sums -> one double atomic per group; the slots of this
                // tile parity are next written two tiles later, i.e. after the next tile's barrier
                asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
                double* osums = p.out_sums[t.b];
                const int g0 = t.n0 / p.out_cpg, g1 = (t.n0 + t.N) / p.out_cpg;   // n_tile is a multiple of out_cpg
                const float* all = s_colsum[it & 1][0];
                for (int g = g0 + et; g < g1; g += 32 * kEpiWarps) {
                    float sa = 0.f, sb = 0.f;
                    for (int c = g * p.out_cpg - t.n0; c < (g + 1) * p.out_cpg - t.n0; ++c) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            sa += all[q * 2 * kColMax + c]; sb += all[q * 2 * kColMax + kColMax + c];
                        }
                    }
                    if (osums) {
                        atomicAdd(osums + 2 * g, (double)sa);
                        atomicAdd(osums + 2 * g + 1, (double)sb);
                    }
                }
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsAux));
        if (warp == kWarpMma) {
            // ===================== MMA issuer: one thread ================================================
            if (lane == 0) {
                int sa = 0, sb = 0;
                uint32_t pa = 0, pb = 0;
                for (int it = 0;; ++it) {
                    Tile t;
                    if (!tile_at(it, t)) break;
                    const int a = it % p.acc_bufs, use = it / p.acc_bufs;
                    detail_stamp(p.detail, it, 3);
                    mbar_wait(bar_accempty + 8 * a, ((uint32_t)use & 1u) ^ 1u);   // the epilogue has drained this accumulator
                    tc_fence_after();
                    detail_stamp(p.detail, it, 4);
                    // instruction descriptor: D = F32 [4,6), A / B = F16 (0) [7,10) [10,13), K-major both, N>>3 [17,23), M>>4 [24,29)
                    const uint32_t idesc = (1u << 4) | ((uint32_t)(t.Nmma >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);
                    const uint32_t td = tmem_base + (uint32_t)(a * p.acc_stride);
                    for (int kb = 0; kb < num_kb; ++kb) {
                        mbar_wait(bar_bfull + 8 * sb, pb);
                        if (kb == 0) detail_stamp(p.detail, it, 8);
                        mbar_wait(bar_afull + 8 * sa, pa);
                        tc_fence_after();
                        if (kb == 0) detail_stamp(p.detail, it, 9);
                        const uint32_t bb = smem_u32(smem + (size_t)sb * b_stage);
                        const uint64_t b_hi = make_desc(bb), b_lo = make_desc(bb + b_bytes);
                        const uint32_t ta = tmem_base + a_col0 + (uint32_t)sa * 64u;
                        if (!(dbg & 4)) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint64_t adv = (uint64_t)((k * 32) >> 4);
                                umma_ts<true>(td, ta + k * 8, b_hi + adv, idesc, (kb | k) != 0);
                                umma_ts<true>(td, ta + 32u + k * 8, b_hi + adv, idesc, 1);
                                umma_ts<true>(td, ta + k * 8, b_lo + adv, idesc, 1);
                            }
                        }
                        umma_commit(bar_aempty + 8 * sa);
                        umma_commit(bar_bempty + 8 * sb);
                        if (++sa == SA) { sa = 0; pa ^= 1u; }
                        if (++sb == SB) { sb = 0; pb ^= 1u; }
                    }
                    umma_commit(bar_accfull + 8 * a);
                    detail_stamp(p.detail, it, 5);
                }
            }
            __syncwarp();
        } else if (warp == kWarpTma) {
            // ===================== TMA: stream the pre-split weight tiles, continuously across tiles ===========
            if (lane == 0) {
                int sb = 0;
                uint32_t pb = 1u;
                int loaded_n0 = -1;   // resident mode: the channel tile whose k-blocks the stages hold (stage = k-block)
                for (int it = 0;; ++it) {
                    Tile t;
                    if (!tile_at(it, t)) break;
                    const bool keep = p.resident && t.n0 == loaded_n0;
                    for (int kb = 0; kb < num_kb; ++kb) {
                        mbar_wait(bar_bempty + 8 * sb, pb);
                        if (keep) {
                            mbar_arrive(bar_bfull + 8 * sb);   // the stage already holds this k-block: hand it over again, no L2 read
                        } else {
                            const uint32_t bb = smem_u32(smem + (size_t)sb * b_stage);
                            mbar_arrive_expect_tx(bar_bfull + 8 * sb, b_stage);
                            tma_load_2d(bb, &map_hi, bar_bfull + 8 * sb, kb * kKB, t.n0);
                            tma_load_2d(bb + b_bytes, &map_lo, bar_bfull + 8 * sb, kb * kKB, t.n0);
                        }
                        if (++sb == SB) { sb = 0; pb ^= 1u; }
                    }
                    loaded_n0 = t.n0;
                    detail_stamp(p.detail, it, 10);
                }
            }
            __syncwarp();
        } else {
            // ===================== tile prologue: the neighbour slices of the NEXT tiles ======================
            const int mt = tid - kWarpMeta0 * 32;   // 0..63
            for (int it = 0;; ++it) {
                Tile t;
                if (!tile_at(it, t)) break;
                const int q = it % kMetaBufs;
                if (lane == 0) mbar_wait(bar_metaempty + 8 * q, ((uint32_t)(it / kMetaBufs) & 1u) ^ 1u);
                __syncwarp();
                const int Vq = s_vq[t.b], Vx = s_vx[t.b];
                const int* nbr = p.nbr[t.b];
                int* dst = s_nbr[q];
                constexpr int kChunk = 6;   // loads in flight per thread (the auxiliary warps run on a small register budget)
                for (int u0 = 0; u0 < kNbrPerMeta; u0 += kChunk) {
                    int pref[kChunk];
#pragma unroll
                    for (int u = 0; u < kChunk; ++u) {
                        const int i = mt + (u0 + u) * kMetaThreads;
                        const int r = i / S;
                        int sv = -1;
                        if (i < kBlockM * S && t.row0 + r < Vq) sv = nbr ? __ldg(nbr + (size_t)t.row0 * S + i) : t.row0 + r;
                        pref[u] = sv;
                    }
#pragma unroll
                    for (int u = 0; u < kChunk; ++u) {
                        const int i = mt + (u0 + u) * kMetaThreads;
                        if (i < kBlockM * S) {
                            const int r = i / S, sl = i - r * S;
                            dst[sl * kBlockM + r] = pref[u] >= Vx ? -1 : pref[u];
                        }
                    }
                    if ((u0 + kChunk) * kMetaThreads >= kBlockM * S) break;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_metafull + 8 * q);
                if (mt == 0) detail_stamp(p.detail, it, 11);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kWarpMma) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
    }
    if (p.trace && tid == 0) {
        unsigned long long t_last;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_last));
        p.trace[2 + (size_t)blockIdx.x * 2] = t_first;
        p.trace[3 + (size_t)blockIdx.x * 2] = t_last;
        if (blockIdx.x == 0) {
            unsigned long long rows = 0;
            for (int b = 0; b < p.nb; ++b) rows += (unsigned long long)max(s_vq[b], 0);
            p.trace[0] = rows;
            p.trace[1] = (unsigned long long)total_tiles;
        }
    }
}

// A-operand pre-staging for the gathering (S = 9) layers.  The gather loop of the convolution touches every vertex row
// nine times (once per neighbour slot, from different tiles); converting the row each time -- GroupNorm affine, ReLU,
// fp16 range check, hi / lo split: ~16 instructions per value -- made the producers, not the tensor pipe or the memory
// system, the bound of the batched kernel (measured: 1.1 us of staging per k-block against 0.6 us of MMA at N = 192).
// This pass does that work ONCE per row and layer and leaves, per 4 channels, the 16 bytes {hi01, hi23, lo01, lo23} that
// tcgen05.st takes for them -- same size and indexing as the fp32 row, so the gather loop only moves data.  The
// arithmetic is the inline path's, operation for operation: results are bit-identical.
struct StageParams {
    const float* x[kMaxBatch];
    const double* gn_sums[kMaxBatch];
    uint4* out[kMaxBatch];
    const int* vx_dev[kMaxBatch];
    int* flag[kMaxBatch];
    int Vx[kMaxBatch];
    const float* gn_gamma;
    const float* gn_beta;
    float gn_eps, a_mul;
    int gn_cpg, C, relu, has_gn;
};

__global__ void __launch_bounds__(256)
k_stage_a(const __grid_constant__ StageParams p) {
    __shared__ __align__(16) float s_aff[512];
    const int b = blockIdx.y;
    const int Vx = p.vx_dev[b] ? min(p.Vx[b], __ldg(p.vx_dev[b])) : p.Vx[b];
    const int C = p.C, C4 = C >> 2;
    if (p.has_gn) {
        const double* sums = p.gn_sums[b];
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            const int g = c / p.gn_cpg;
            const double n = (double)Vx * p.gn_cpg;
            const double mean = sums[2 * g] / n;
            double var = sums[2 * g + 1] / n - mean * mean;
            if (var < 0.0) var = 0.0;
            const float rstd = (float)(1.0 / sqrt(var + (double)p.gn_eps));
            const float sc = rstd * (p.gn_gamma ? __ldg(p.gn_gamma + c) : 1.0f);
            const float sh = (p.gn_beta ? __ldg(p.gn_beta + c) : 0.0f) - (float)mean * sc;
            s_aff[c] = sc * p.a_mul;
            s_aff[256 + c] = sh * p.a_mul;
        }
    }
    __syncthreads();
    const float4* x4 = reinterpret_cast<const float4*>(p.x[b]);
    uint4* out = p.out[b];
    const long long total = (long long)Vx * C4;
    float amax = 0.f;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int q = (int)(t % C4);
        float4 sc = make_float4(p.a_mul, p.a_mul, p.a_mul, p.a_mul), sh = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.has_gn) {
            sc = *reinterpret_cast<const float4*>(s_aff + 4 * q);
            sh = *reinterpret_cast<const float4*>(s_aff + 256 + 4 * q);
        }
        float4 v = __ldg(x4 + t);
        v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y);
        v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
        if (p.relu) { v.x = relu_nan(v.x); v.y = relu_nan(v.y); v.z = relu_nan(v.z); v.w = relu_nan(v.w); }
        amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
        const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn(v.x - f01.x, v.y - f01.y), l23 = __floats2half2_rn(v.z - f23.x, v.w - f23.y);
        uint4 o;
        o.x = *reinterpret_cast<const uint32_t*>(&h01); o.y = *reinterpret_cast<const uint32_t*>(&h23);
        o.z = *reinterpret_cast<const uint32_t*>(&l01); o.w = *reinterpret_cast<const uint32_t*>(&l23);
        out[t] = o;
    }
    if (p.flag[b] && !(amax < 65504.f)) atomicOr(p.flag[b], 1);
}

}  // namespace

extern "C" {

// Tracing of the following batched launches: launch i (i < nr_records) writes record i of buf ([nr_records, 2 + 2*148] u64,
// zeroed by the caller): [0] = live rows summed over its problems, [1] = tiles, then (entry, exit) %globaltimer stamps per
// CTA.  The record address is baked into the launch, so launches captured into a CUDA graph keep writing their record on
// every replay.  buf = NULL switches tracing off.  Returns the number of records handed out since the last call.
int ltn_conv_batched_trace(unsigned long long* buf, int nr_records) {
    const int used = g_trace_next;
    g_trace_b = buf;
    g_trace_slots = buf ? nr_records : 0;
    g_trace_next = 0;
    return used;
}

// Role timeline of CTA 0 of the following batched launches: buf [64 tiles][16] u64 (zeroed by the caller) receives %globaltimer
// stamps -- 0 gather issue cursor enters the tile, 1 / 2 gather consume cursor enters / leaves, 3 / 4 MMA issuer before / after
// the accumulator hand-back wait, 8 / 9 first weight stage / first A stage ready, 5 accumulator committed, 6 / 7 epilogue
// sees the accumulator / hands it back, 10 last weight k-block issued, 11 neighbour slice staged.  NULL switches it off.
int ltn_conv_batched_detail(unsigned long long* buf) {
    g_detail_b = buf;
    return 0;
}

// A-operand pre-staging (k_stage_a) for nb problems: out[b] [Vx_b, C] (same bytes as x[b]) receives, per 4 channels, the
// fp16 quads {hi01, hi23, lo01, lo23} of relu?(GroupNorm(x)) * 2^a_log2; flag[b] is OR-ed with 1 when a value leaves the
// fp16 range.  ltn_conv_tc_f16_batched(..., staged = 1) then takes out[] in place of x[].
int ltn_stage_a_batched(int nb, const float* const* x, const int* Vx, const int* const* vx_dev, int C, const double* const* gn_sums,
                        const float* gn_gamma, const float* gn_beta, float gn_eps, int gn_groups, int relu, int a_log2,
                        void* const* out, int* const* flag, void* stream) {
    if (nb <= 0) return 0;
    if (nb > kMaxBatch || C <= 0 || C % 4 || C > 256) return -2;
    const bool has_gn = gn_sums && gn_sums[0];
    if (has_gn && (gn_groups <= 0 || C % gn_groups)) return -2;
    static StageParams p;
    memset(&p, 0, sizeof(p));
    int live = 0, vmax = 0;
    for (int b = 0; b < nb; ++b) {
        if (Vx[b] <= 0) continue;
        p.x[live] = x[b]; p.gn_sums[live] = has_gn ? gn_sums[b] : nullptr; p.out[live] = reinterpret_cast<uint4*>(out[b]);
        p.vx_dev[live] = vx_dev ? vx_dev[b] : nullptr; p.flag[live] = flag ? flag[b] : nullptr; p.Vx[live] = Vx[b];
        if (Vx[b] > vmax) vmax = Vx[b];
        ++live;
    }
    if (live == 0) return 0;
    p.gn_gamma = gn_gamma; p.gn_beta = gn_beta; p.gn_eps = gn_eps; p.a_mul = ldexpf(1.0f, a_log2);
    p.gn_cpg = has_gn ? C / gn_groups : 1; p.C = C; p.relu = relu; p.has_gn = has_gn ? 1 : 0;
    long long items = (long long)vmax * (C / 4);
    int bx = (int)((items + 256 * 4 - 1) / (256 * 4));   // ~4 quads per thread
    if (bx < 1) bx = 1;
    if (bx > 148 * 8) bx = 148 * 8;
    k_stage_a<<<dim3(bx, live), 256, 0, (cudaStream_t)stream>>>(p);
    LTN_CHECK_LAUNCH();
    return 0;
}

// The fp16-operand fused convolution of ltn_conv_tc_f16 for nb <= 8 independent problems that share the weights (the same
// layer of several windows in flight): every per-problem argument is a HOST array of nb entries.  Requirements as
// ltn_conv_tc_f16 (C % 64 == 0, C <= 256 with a folded GroupNorm, F % 8 == 0, ldo % 4 == 0); an explicit a_scale/a_shift
// affine is not offered here.
int ltn_conv_tc_f16_batched(int nb, const float* const* x, const int* Vx, const int* const* vx_dev, const int* const* nbr,
                            const int* Vq, const int* const* vq_dev, int C, int S, const void* wt_hi, const void* wt_lo, int w_log2,
                            int a_log2, int F, const double* const* gn_sums, const float* gn_gamma, const float* gn_beta,
                            float gn_eps, int gn_groups, int relu, const float* bias, const float* const* res, float* const* out,
                            int ldo, double* const* out_sums, int out_groups, int* const* flag, int staged, void* stream) {
    if (nb <= 0) return 0;
    if (nb > kMaxBatch) return -2;
    const bool has_nbr = nbr && nbr[0];
    const bool has_gn = gn_sums && gn_sums[0];
    const bool has_sums = out_sums && out_sums[0];
    if (C <= 0 || C % kKB || (has_gn && C > 256) || F <= 0 || F % 8 || S < 1 || ldo % 4) return -2;
    if (has_gn && (gn_groups <= 0 || C % gn_groups)) return -2;
    if (has_sums && (out_groups <= 0 || F % out_groups)) return -2;
    if (!wt_hi || !wt_lo) return -2;
    static BatchParams p;   // ~700 bytes, filled per launch (single host thread per process: see SURVEY.md 8b "Threading")
    memset(&p, 0, sizeof(p));
    long long row_tiles = 0;
    int live = 0;
    for (int b = 0; b < nb; ++b) {
        if (Vq[b] <= 0) continue;
        if (Vx[b] <= 0) return -2;   // row 0 of x must be readable (stand-in address of absent neighbours)
        if ((nbr && nbr[b] != nullptr) != has_nbr || (gn_sums && gn_sums[b] != nullptr) != has_gn ||
            (out_sums && out_sums[b] != nullptr) != has_sums)
            return -2;   // the problems of one launch share their structure
        p.x[live] = x[b]; p.nbr[live] = has_nbr ? nbr[b] : nullptr; p.gn_sums[live] = has_gn ? gn_sums[b] : nullptr;
        p.res[live] = res ? res[b] : nullptr; p.out[live] = out[b]; p.out_sums[live] = has_sums ? out_sums[b] : nullptr;
        p.vq_dev[live] = vq_dev ? vq_dev[b] : nullptr; p.vx_dev[live] = vx_dev ? vx_dev[b] : nullptr;
        p.flag[live] = flag ? flag[b] : nullptr;
        p.Vq[live] = Vq[b]; p.Vx[live] = Vx[b];
        row_tiles += (Vq[b] + kBlockM - 1) / kBlockM;
        ++live;
    }
    if (live == 0) return 0;
    p.nb = live;
    p.gn_gamma = gn_gamma; p.gn_beta = gn_beta; p.bias = bias; p.gn_eps = gn_eps;
    p.gn_cpg = has_gn ? C / gn_groups : 1;
    p.out_cpg = has_sums ? F / out_groups : 1;
    p.C = C; p.S = has_nbr ? S : 1; p.F = F; p.ldo = ldo; p.relu = relu; p.has_gn = has_gn ? 1 : 0; p.has_sums = has_sums ? 1 : 0;
    p.a_mul = ldexpf(1.0f, a_log2);
    p.out_mul = ldexpf(1.0f, -(a_log2 + w_log2));
    p.staged = staged ? 1 : 0;       // x[] are ltn_stage_a_batched outputs: affine / ReLU / split already applied
    if (p.staged) p.has_gn = 0;
    static const int dbg = []() { const char* e = getenv("LTN_CONVB_DEBUG"); return e ? atoi(e) : 0; }();
    p.debug = dbg;
    p.detail = g_detail_b;
    p.trace = (g_trace_b && g_trace_next < g_trace_slots) ? g_trace_b + (size_t)kTraceStride * g_trace_next++ : nullptr;
    // Output channels per tile: at most 192, so that TWO accumulators and an A ring of >= 2 stages fit the 512 columns of
    // tensor memory; fewer when the tile list would leave SMs idle (small levels).
    static const int n_cap = []() { const char* e = getenv("LTN_CONVB_NCAP"); return e && atoi(e) >= 16 && atoi(e) <= kColMax ? atoi(e) : kColMax; }();
    const int unit = has_sums ? lcm16(p.out_cpg) : 16;
    int ny = (F + n_cap - 1) / n_cap;
    while (row_tiles * (ny + 1) <= 148 && (F / (ny + 1)) >= 32) ++ny;
    int n_tile = ((F + ny - 1) / ny + unit - 1) / unit * unit;
    if (n_tile > n_cap) n_tile = n_cap / unit * unit;
    if (n_tile <= 0) return -2;
    ny = (F + n_tile - 1) / n_tile;
    p.n_tile = n_tile;
    p.ny = ny;
    const int nmma = (n_tile + 15) & ~15;
    p.acc_stride = (nmma + 31) & ~31;
    static const int want_bufs = []() { const char* e = getenv("LTN_CONVB_ACC"); return e ? atoi(e) : 2; }();
    p.acc_bufs = (want_bufs >= 2 && 2 * p.acc_stride + 2 * 64 <= 512) ? 2 : 1;
    int sa = (512 - p.acc_bufs * p.acc_stride) / 64;
    if (sa > kMaxSA) sa = kMaxSA;
    sa &= ~1;   // the two producer groups own alternate stages
    if (sa < 2) return -3;
    p.stages_a = sa;
    const size_t b_stage = 2 * (size_t)n_tile * 128;
    static const int sb_cap = []() { const char* e = getenv("LTN_CONVB_SB"); return e && atoi(e) >= 2 ? atoi(e) : 4; }();
    // Epilogue warps: 8 for the dense layers (their tiles are epilogue-bound), 4 for the gathering layers on pre-staged
    // operands: that role layout leaves the producers 168 registers, enough for the three-deep gather pipeline.
    static const int epi_env = []() { const char* e = getenv("LTN_CONVB_EPI"); return e ? atoi(e) : 0; }();
    (void)epi_env;
    const int epi = p.staged ? 4 : 8;   // the role layout IS the operand form (see the kernel)
    static const int want_resident = []() { const char* e = getenv("LTN_CONVB_RESIDENT"); return e ? atoi(e) : 1; }();
    bool any_res = false;
    for (int b = 0; b < live; ++b) any_res = any_res || p.res[b] != nullptr;
    static const int want_res_async = []() { const char* e = getenv("LTN_CONVB_RESASYNC"); return e ? atoi(e) : 1; }();
    p.res_async = (want_res_async && any_res && p.staged && epi == 4) ? 1 : 0;
    const int res_kb = p.res_async ? 16 : 0;   // the second residual buffer of each epilogue warp
    const size_t b_budget = (size_t)(226 - 44 - 4 * epi - res_kb - 2) * 1024;   // 227 KB - static (~43 KB) - epilogue staging
    const int num_kb = p.S * C / kKB;
    int sb = (int)(b_budget / b_stage);
    // Weights RESIDENT: when all k-blocks of a channel tile fit the ring (every dense layer, the 64 -> 64 convolutions), the
    // ring gets one stage per k-block and the TMA lane loads each stage once per CTA and channel tile; the following row
    // tiles re-use it.  Otherwise every row tile streams the whole K x N tile again from L2 -- which is what bounds the
    // batched kernel (measured: 5-7 TB/s of L2 -> SM traffic, 40-60 % of it weights).
    p.resident = (want_resident && num_kb <= kMaxSB && num_kb <= sb) ? 1 : 0;
    if (p.resident) {
        sb = num_kb;                 // stage = k-block (a single-k-block layer runs on a ring of one stage: nothing is ever re-loaded)
    } else {
        if (sb > 4) sb = 4;
        if (sb > sb_cap) sb = sb_cap;
        if (sb < 2) return -3;
    }
    p.stages_b = sb;
    const size_t smem = (size_t)sb * b_stage + (size_t)epi * 4096 + (size_t)res_kb * 1024 + 1024;
    alignas(64) CUtensorMap map_hi, map_lo;
    int rc = make_weight_map(&map_hi, wt_hi, F, p.S * C, n_tile, true);
    if (rc) return rc;
    rc = make_weight_map(&map_lo, wt_lo, F, p.S * C, n_tile, true);
    if (rc) return rc;
    const void* fn = epi == 8 ? (const void*)k_conv_tc_batched<8> : (const void*)k_conv_tc_batched<4>;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    long long tiles = row_tiles * ny;
    static const int grid_cap = []() { const char* e = getenv("LTN_CONVB_GRID"); return e && atoi(e) >= 1 ? atoi(e) : 148; }();
    const int grid = (int)(tiles < grid_cap ? tiles : grid_cap);
    if (epi == 8) k_conv_tc_batched<8><<<grid, Roles<8>::kThreads, smem, (cudaStream_t)stream>>>(map_hi, map_lo, p);
    else k_conv_tc_batched<4><<<grid, Roles<4>::kThreads, smem, (cudaStream_t)stream>>>(map_hi, map_lo, p);
    LTN_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
