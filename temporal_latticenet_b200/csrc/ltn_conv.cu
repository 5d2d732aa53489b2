// Fused lattice convolution on the Blackwell tensor cores: neighbour gather (+ folded GroupNorm/ReLU)
// + GEMM + bias + residual in ONE kernel, no [V,9C] im2row buffer.
//
//   out[v, f] = sum_{s<S} sum_{c<C} act( x[nbr[v,s], c] ) * W[s*C + c, f]  (+ bias[f]) (+ res[v, f])
//   act(t)    = relu?( t * a_scale[c] + a_shift[c] )     -- GroupNorm folded to a per-channel affine,
//                                                           applied to PRESENT rows only (absent -> 0,
//                                                           exactly what im2row of the normalised values gives)
//
// serves ConvLatticeModule / CoarsenLattice / FinefyLattice (S = 9, seq_lattice/lattice_modules.py:440,573;
// models.py:353,398), GnRelu1x1 / Conv1x1 / nn.Linear and the GRU/LSTM gate GEMMs (S = 1, nbr = NULL).
//
// Blackwell mapping
//  * tcgen05.mma.cta_group::1, M = 128 vertices per CTA, N = up to 256 output channels so the A rows are gathered
//    exactly once; fp32 accumulator lives in TMEM (N columns x 128 lanes).
//  * fp32 parity: every fp32 operand is split x = hi + lo and the product is issued as hi*hi + lo*hi + hi*lo -- three
//    tensor-core passes, ~2^-21 relative error per product, i.e. fp32-class results (the reference ran these GEMMs
//    in fp32 / TF32 cuBLAS).  Two operand types carry the split:
//      - fp16 (ltn_conv_tc_f16, what the window runners use): 11 + 11 significant bits in 2 bytes per element, kind::f16
//        at twice the tf32 rate; exact power-of-two scaling of activations and weights plus a device-side range flag
//        cover fp16's narrow exponent (the caller redoes flagged work with tf32).  The A operand never touches shared
//        memory: the gather registers are converted and written straight into TENSOR memory with tcgen05.st.16x256b
//        (the gather is laid out like that fragment) and the MMA reads A from there (TS mode).
//      - tf32 (ltn_conv_tc; hi = round-to-tf32, lo = x - hi exactly): A staged K-major in 128-byte-swizzled shared
//        memory by the producers (16-byte chunks at chunk ^ (row & 7), fence.proxy.async), kind::tf32.  PASSES = 1 is
//        the single-pass TF32 variant (reported separately, never the default).
//  * A (gathered, data dependent): two producer warpgroups take alternate k-blocks; the tile's slice of the neighbour
//    table sits in shared memory, row segments are fetched with 128-bit loads one unit ahead, the folded
//    GroupNorm+ReLU and the hi/lo split happen in registers.
//  * B (weights, dense): pre-split hi/lo K-major copies streamed by TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B)
//    from a dedicated lane that owns the ring's barriers and starts streaming before the block-wide set-up barrier.
//    mbarrier rings decouple both from the single MMA-issuing thread; tcgen05.commit releases stages and finally
//    hands the accumulator to the epilogue.
//  * epilogue (8 warps): tcgen05.ld 32x32b (thread = row, 32 columns at a time) -> per-warp swizzled transpose in
//    the idle operand ring -> quarter-warp per row: + bias, + residual, whole 128-byte lines stored; optionally
//    the GroupNorm statistics of the OUTPUT (sum, sum of squares per group) are reduced here, so the next
//    layer's normalisation costs no extra pass over the data.
//  * sizes that only the device knows (vertex counts after hash insertion) are read from device
//    memory (vq_dev / vx_dev), so the launch needs no host synchronisation.
#include "ltn_common.cuh"
#include "ltn_tc.cuh"
#include "ltn_conv_common.cuh"
#include <cstdlib>

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 32;        // fp32 elements per k-block = one 128-byte swizzle row
constexpr int kGroups = 2;         // producer warpgroups (round-robin k-blocks); 3 measured equal (fp16, A in tensor memory) or slower (tf32: register spills)
constexpr int kGroupThreads = 128;
constexpr int kMmaWarp = 4 * kGroups, kTmaWarp = 4 * kGroups + 1;
constexpr int kThreads = 32 * (4 * kGroups + 2);   // producer/epilogue warps + MMA warp + TMA warp
constexpr int kMaxStages = 4;    // A ring (tensor-memory columns / shared memory)
constexpr int kMaxStagesB = 8;   // weight ring: with A in tensor memory the shared memory is all its own

struct ConvParams {
    const float* x;        // [Vx, C]
    const int* nbr;        // [Vq, S] or null (identity, S = 1)
    const float* a_scale;  // [C] or null   explicit per-channel affine on A
    const float* a_shift;
    const double* gn_sums; // [G,2] or null GroupNorm of x folded into A: sums over the Vx rows
    const float* gn_gamma;
    const float* gn_beta;
    const float* bias;     // [F] or null
    const float* res;      // [Vq, F] or null
    float* out;            // [Vq, ldo]
    double* out_sums;      // [Gout,2] or null: accumulate GroupNorm statistics of the output
    const int* vq_dev;     // device-side row counts (nullable)
    const int* vx_dev;
    float gn_eps;
    int gn_cpg, out_cpg;
    int Vq, Vx, C, S, F, ldo, relu;
    int n_tile;            // output channels per CTA (<= 256, multiple of 16)
    int stages_a, stages_b; // depth of the A (gathered) and B (TMA) shared-memory rings
    int cluster;            // CTAs per cluster sharing the weight tiles by TMA multicast (1, 2 or 4)
    float a_mul, out_mul;   // fp16 operands: power-of-two scale applied to A, and its (and the weights') inverse for the output
    int* flag;              // fp16 operands: set to 1 when a staged magnitude leaves the half range (nullable)
    unsigned long long* trace;  // nullable: [CTA][8] globaltimer stamps of the phases (ltn_conv_trace)
};

__device__ __forceinline__ void trace_stamp(const ConvParams& p, int phase) {
    if (p.trace) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.trace[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 8 + phase] = t;
    }
}

unsigned long long* g_trace = nullptr;

// OPERANDS: F16 = false -> tf32 operands (4 bytes / element, 32 K-elements per 128-byte swizzle row, MMA K = 8)
//           F16 = true  -> fp16 operands (2 bytes / element, 64 K-elements per row, MMA K = 16): the hi/lo split
//                          carries the same 11 + 11 significant bits as the tf32 split in HALF the shared-memory
//                          bytes and at twice the tensor rate.  fp16's narrow exponent is handled by exact
//                          power-of-two scaling (a_mul on the activations, a per-tensor factor baked into the
//                          weight copies, out_mul undoes both in the epilogue) and an overflow flag the caller
//                          checks (ops.py falls back to the tf32 operands when it is raised).
template <int PASSES, bool ATMEM, bool F16>
__global__ void __launch_bounds__(kThreads, 1)
k_conv_tc(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo, ConvParams p) {
    constexpr int KB = F16 ? 64 : kBlockK;          // K elements per k-block
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * kMaxStages + 2 * kMaxStagesB + 1];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float s_affine[2 * 256];
    __shared__ float s_colsum[2 * 256];
    __shared__ int s_nbr[LTN_FEXT * kBlockM];   // [slot][tile row]: source vertex of every tap of the tile, -1 = absent

    if (threadIdx.x == 0) trace_stamp(p, 0);   // entry
    const int row0 = blockIdx.x * kBlockM;
    // The tile's slice of the neighbour table ([128, S] ints, contiguous in global memory) is fetched ONCE, coalesced, and
    // kept slot-major in shared memory: a slot change in the gather loop then costs a shared-memory read instead of a
    // dependent global load + warp shuffles in front of every row fetch.  The loads are issued FIRST, bounded by the host
    // capacity only, so their latency overlaps the device-side row counts read next (the masks are applied at the store).
    constexpr int kNbrPerThread = (LTN_FEXT * kBlockM + kThreads - 1) / kThreads;
    int nbr_pref[kNbrPerThread];
#pragma unroll
    for (int q = 0; q < kNbrPerThread; ++q) {
        const int i = (int)threadIdx.x + q * kThreads;
        const int r = i / p.S;
        nbr_pref[q] = -1;
        if (i < kBlockM * p.S && row0 + r < p.Vq) nbr_pref[q] = p.nbr ? __ldg(p.nbr + (size_t)row0 * p.S + i) : row0 + r;
    }
    const int Vq = p.vq_dev ? min(p.Vq, __ldg(p.vq_dev)) : p.Vq;
    const int Vx = p.vx_dev ? min(p.Vx, __ldg(p.vx_dev)) : p.Vx;
    const int CL = p.cluster;
    // uniform per CLUSTER, before any barrier / TMEM allocation: a CTA without rows still takes part in its
    // cluster's weight multicast when a sibling has rows
    if ((int)(blockIdx.x / CL * CL) * kBlockM >= Vq) return;
    const uint32_t cta_rank = CL > 1 ? cluster_ctarank() : 0u;
    const uint16_t cta_mask = (uint16_t)((1u << CL) - 1u);
    const int n0 = blockIdx.y * p.n_tile;
    const int N = min(p.n_tile, p.F - n0);     // valid output channels of this CTA (multiple of 8)
    const int Nmma = (N + 15) & ~15;           // MMA N (multiple of 16 at M = 128); the extra weight rows are TMA zero fill

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K = p.S * p.C;
    const int num_kb = K / KB;
    const int kb_per_slot = p.C / KB;
    const int SA = p.stages_a, SB = p.stages_b;

    // shared-memory carve-up: two INDEPENDENT rings, A stages [A_hi | A_lo] and B stages [B_hi | B_lo], every
    // block 1024-byte aligned, so each ring can have the depth the shared-memory budget allows.
    const uint32_t a_bytes = kBlockM * 128;
    const uint32_t b_bytes = (uint32_t)p.n_tile * 128;
    const uint32_t a_stage = (PASSES == 3 ? 2 : 1) * a_bytes, b_stage = (PASSES == 3 ? 2 : 1) * b_bytes;
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* smem_b = smem + (ATMEM ? 0 : (size_t)SA * a_stage);   // ATMEM: the A ring lives in tensor memory

    const uint32_t bar_full = smem_u32(&bars[0]), bar_empty = smem_u32(&bars[kMaxStages]);
    const uint32_t bar_fullb = smem_u32(&bars[2 * kMaxStages]), bar_emptyb = smem_u32(&bars[2 * kMaxStages + kMaxStagesB]);
    const uint32_t bar_acc = smem_u32(&bars[2 * kMaxStages + 2 * kMaxStagesB]);
    // tensor memory: accumulator columns [0, Nmma); ATMEM: A ring behind it, 64 columns per stage (hi | lo)
    const uint32_t a_col0 = ((uint32_t)Nmma + 31u) & ~31u;
    uint32_t tmem_cols = 32;
    while (tmem_cols < (ATMEM ? a_col0 + (uint32_t)SA * 64u : (uint32_t)Nmma)) tmem_cols <<= 1;

    if (tid == 0) {
        for (int s = 0; s < SA; ++s) {
            mbar_init(bar_full + 8 * s, kGroupThreads / 32);   // one arrive per gather warp of the group that owns the k-block
            mbar_init(bar_empty + 8 * s, 1);
        }
        mbar_init(bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // The TMA lane owns the weight ring's barriers: it initialises them and, when no cluster sibling has to be waited
    // for, streams the first SB weight stages right away -- the TMA round trip (~1 us) then runs under the rest of the
    // set-up instead of after the block barrier.
    int tma_kb0 = 0;
    if (warp == kTmaWarp && lane == 0) {
        for (int s = 0; s < SB; ++s) {
            mbar_init(bar_fullb + 8 * s, 1);                   // the TMA thread's expect_tx arrive (+ the bytes of all CL slices)
            mbar_init(bar_emptyb + 8 * s, CL);                 // every CTA of the cluster has finished reading the stage
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (CL == 1) {
            const uint32_t tx = (PASSES == 3 ? 2u : 1u) * b_bytes;
            tma_kb0 = min(SB, num_kb);
            for (int kb = 0; kb < tma_kb0; ++kb) {
                const uint32_t bb = smem_u32(smem_b + (size_t)kb * b_stage);
                mbar_arrive_expect_tx(bar_fullb + 8 * kb, tx);
                tma_load_2d(bb, &map_hi, bar_fullb + 8 * kb, kb * KB, n0);
                if (PASSES == 3) tma_load_2d(bb + b_bytes, &map_lo, bar_fullb + 8 * kb, kb * KB, n0);
            }
        }
    }
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    const bool affine = (p.a_scale != nullptr) || (p.gn_sums != nullptr);
    const float a_mul = F16 ? p.a_mul : 1.0f;   // exact power of two, folded into the affine
    if (p.gn_sums) {
        // GroupNorm folded to scale = rstd*gamma, shift = beta - mean*scale (statistics over all Vx rows)
        for (int c = tid; c < p.C; c += kThreads) {
            const int g = c / p.gn_cpg;
            const double n = (double)Vx * p.gn_cpg;
            const double mean = p.gn_sums[2 * g] / n;
            double var = p.gn_sums[2 * g + 1] / n - mean * mean;
            if (var < 0.0) var = 0.0;
            const float rstd = (float)(1.0 / sqrt(var + (double)p.gn_eps));
            const float a = rstd * (p.gn_gamma ? __ldg(p.gn_gamma + c) : 1.0f);
            const float sh = (p.gn_beta ? __ldg(p.gn_beta + c) : 0.0f) - (float)mean * a;
            s_affine[c] = a * a_mul;
            s_affine[256 + c] = sh * a_mul;
        }
    } else if (p.a_scale) {
        for (int c = tid; c < p.C; c += kThreads) {
            s_affine[c] = __ldg(p.a_scale + c) * a_mul;
            s_affine[256 + c] = __ldg(p.a_shift + c) * a_mul;
        }
    }
    if (p.out_sums)
        for (int c = tid; c < 2 * 256; c += kThreads) s_colsum[c] = 0.f;
#pragma unroll
    for (int q = 0; q < kNbrPerThread; ++q) {
        const int i = tid + q * kThreads;
        if (i < kBlockM * p.S) {
            const int r = i / p.S, sl = i - r * p.S;
            int sv = nbr_pref[q];
            if (row0 + r >= Vq || sv >= Vx) sv = -1;
            s_nbr[sl * kBlockM + r] = sv;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (CL > 1) cluster_sync_all();   // sibling barriers are initialised before any remote arrive / multicast lands
    const uint32_t tmem_base = tmem_slot;
    if (tid == 0) trace_stamp(p, 1);   // set-up done

    if (warp < 4 * kGroups) {
        // ===================== producers: gather A rows (two warpgroups, alternate k-blocks) ==========
        // Quarter-warp per row: 8 lanes fetch the 8 16-byte chunks of one 128-byte row segment, so one
        // warp-wide LDG.128 touches 4 whole cache lines (not 32 partial ones) and the swizzled stores of a
        // quarter-warp cover one shared-memory row conflict-free.  A warp owns 32 tile rows and walks them
        // 4 at a time; a gather UNIT is 8 float4 per thread (tf32: 8 rows x 32 channels; fp16: 4 rows x 64
        // channels, two units per k-block); the loads of the NEXT unit are issued before the current one is
        // converted, so the L2 latency of the gather overlaps the staging work.
        const int group = warp >> 2;
        const int wrow0 = (warp & 3) * 32;         // first tile row of this warp
        const int chunk = lane & 7, sub = lane >> 3;
        // All ring / k-block bookkeeping is incremental: no integer division in the loop.
        const float4* x4 = reinterpret_cast<const float4*>(p.x);
        const uint32_t c4 = (uint32_t)p.C >> 2;
        if constexpr (F16 && ATMEM) {
            // ---- fp16 operands staged in TENSOR memory: no shared-memory A ring at all -------------------------------
            // tcgen05.st.16x256b hands thread T the 32-bit columns 8g + 2(T%4) + {0,1} of lanes T/4 and T/4 + 8 -- four
            // consecutive fp16 channels of two rows per column group g.  The gather is laid out the SAME way (thread T
            // fetches the float4 at channel 16g + 4(T%4) of its rows; 4 lanes cover 64 contiguous bytes of a row), so the
            // converted values go from registers straight to the operand's place: no transpose, no shuffle, no st.shared,
            // and the tensor core no longer reads A through the shared-memory pipe (80 of the 120-200 KB per k-block).
            // A unit is 16 tile rows x 64 channels (8 float4 per thread), two units per k-block and warp.
            const int ra = lane >> 2, cq = lane & 3;
            float4 buf[2][8];
            uint32_t bmask[2];
            uint32_t rowidx[4];    // float4 index of (source vertex of tile row wrow0 + ra + 8r, this lane's channel quad)
            uint32_t rmask = 0;
            float amax = 0.f;
            int i_slot = group / kb_per_slot, i_c0 = (group - i_slot * kb_per_slot) * KB, i_cur = -1;   // issue position
            int c_c0 = i_c0, c_stage = group % SA;                                                    // consume position
            uint32_t c_par = ((group / SA) & 1) ^ 1;
            auto issue = [&](float4* dst, uint32_t& dmask, const int half) {
                if (i_slot != i_cur) {
                    i_cur = i_slot;
                    const int* tap = s_nbr + i_slot * kBlockM + wrow0 + ra;
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
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t ri = (half ? rowidx[2 + h] : rowidx[h]) + o;
#pragma unroll
                    for (int g = 0; g < 4; ++g) dst[4 * h + g] = __ldg(x4 + ri + 4 * g);
                }
                if (!half) return;
                i_c0 += kGroups * KB;
                while (i_c0 >= p.C) { i_c0 -= p.C; ++i_slot; }
            };
            auto consume = [&](const float4* cur, uint32_t cmask, const int half) {
                if (half == 0) {
                    if (lane == 0) mbar_wait(bar_empty + 8 * c_stage, c_par);   // one waiter per warp
                    __syncwarp();
                    tc_fence_after();
                }
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    float4 sc = make_float4(a_mul, a_mul, a_mul, a_mul), sh = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (affine) {
                        sc = *reinterpret_cast<const float4*>(s_affine + c_c0 + 16 * g + 4 * cq);
                        sh = *reinterpret_cast<const float4*>(s_affine + 256 + c_c0 + 16 * g + 4 * cq);
                    }
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        float4 t = cur[4 * h + g];
                        t.x = fmaf(t.x, sc.x, sh.x); t.y = fmaf(t.y, sc.y, sh.y);
                        t.z = fmaf(t.z, sc.z, sh.z); t.w = fmaf(t.w, sc.w, sh.w);
                        if (p.relu) { t.x = relu_nan(t.x); t.y = relu_nan(t.y); t.z = relu_nan(t.z); t.w = relu_nan(t.w); }
                        if (!((cmask >> h) & 1u)) t = make_float4(0.f, 0.f, 0.f, 0.f);   // absent neighbour / row beyond the tile
                        amax = fmaxf(amax, fmaxf(fmaxf(fabsf(t.x), fabsf(t.y)), fmaxf(fabsf(t.z), fabsf(t.w))));
                        const __half2 h01 = __floats2half2_rn(t.x, t.y), h23 = __floats2half2_rn(t.z, t.w);
                        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                        const __half2 l01 = __floats2half2_rn(t.x - f01.x, t.y - f01.y), l23 = __floats2half2_rn(t.z - f23.x, t.w - f23.y);
                        hi[4 * g + 2 * h] = *reinterpret_cast<const uint32_t*>(&h01);
                        hi[4 * g + 2 * h + 1] = *reinterpret_cast<const uint32_t*>(&h23);
                        lo[4 * g + 2 * h] = *reinterpret_cast<const uint32_t*>(&l01);
                        lo[4 * g + 2 * h + 1] = *reinterpret_cast<const uint32_t*>(&l23);
                    }
                }
                const uint32_t ta = tmem_base + ((uint32_t)(wrow0 + 16 * half) << 16) + a_col0 + (uint32_t)c_stage * 64u;
                tmem_st_16x256b_x4(ta, hi);
                if (PASSES == 3) tmem_st_16x256b_x4(ta + 32u, lo);
                if (half == 0) return;
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_full + 8 * c_stage);   // one arrival per warp
                c_c0 += kGroups * KB;
                while (c_c0 >= p.C) c_c0 -= p.C;
                c_stage += kGroups;
                while (c_stage >= SA) { c_stage -= SA; c_par ^= 1u; }
            };
            if (group < num_kb) issue(buf[0], bmask[0], 0);
            for (int kb = group; kb < num_kb; kb += kGroups) {
                issue(buf[1], bmask[1], 1);
                consume(buf[0], bmask[0], 0);
                if (kb + kGroups < num_kb) issue(buf[0], bmask[0], 0);
                consume(buf[1], bmask[1], 1);
            }
            if (p.flag && !(amax < 65504.f)) atomicOr(p.flag, 1);
        } else {
        float4 buf[2][8];      // two register buffers, used with compile-time indices
        uint32_t bmask[2];     // bit j: tile row j of the warp's walk is a present neighbour
        uint32_t rowidx[8];    // float4 index of (row j's source vertex, this lane's chunk), refreshed per slot
        uint32_t rmask = 0;
        float amax = 0.f;      // F16: largest scaled magnitude staged by this thread (range check)
        int i_slot = group / kb_per_slot, i_c0 = (group - i_slot * kb_per_slot) * KB, i_cur = -1;   // issue position
        int c_c0 = i_c0, c_stage = group % SA;                                                    // consume position
        uint32_t c_par = ((group / SA) & 1) ^ 1;
        // part: which half of the warp's rows (fp16 operands only; compile-time at every call site)
        auto issue = [&](float4* dst, uint32_t& dmask, const int part) {
            if (i_slot != i_cur) {
                i_cur = i_slot;
                const int* tap = s_nbr + i_slot * kBlockM + wrow0;
                rmask = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int sv = tap[ATMEM ? 8 * sub + j : sub + 4 * j];
                    // row 0 stands in for an absent neighbour (loaded unconditionally, discarded at staging time): a
                    // predicated load would be followed by a predicated register move that waits for it on the spot
                    rowidx[j] = (uint32_t)(sv >= 0 ? sv : 0) * c4 + (uint32_t)chunk;
                    rmask |= (sv >= 0 ? 1u : 0u) << j;
                }
            }
            const uint32_t o = (uint32_t)i_c0 >> 2;
            if (F16) {
                dmask = part ? (rmask >> 4) : rmask;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t ri = (part ? rowidx[4 + j] : rowidx[j]) + o;
                    dst[2 * j] = __ldg(x4 + ri);          // channels c0 + 4*chunk .. (first 128-byte line of the segment)
                    dst[2 * j + 1] = __ldg(x4 + ri + 8);  // channels c0 + 32 + 4*chunk .. (second line)
                }
                if (!part) return;
            } else {
                dmask = rmask;
#pragma unroll
                for (int j = 0; j < 8; ++j) dst[j] = __ldg(x4 + rowidx[j] + o);
            }
            i_c0 += kGroups * KB;
            while (i_c0 >= p.C) { i_c0 -= p.C; ++i_slot; }
        };
        auto consume = [&](const float4* cur, uint32_t cmask, const int part) {
            if (!F16 || part == 0) {
                if (lane == 0) mbar_wait(bar_empty + 8 * c_stage, c_par);   // one waiter per warp
                __syncwarp();
            }
            uint8_t* a_hi0 = smem + (size_t)c_stage * a_stage;
            if (F16) {
                // 4 rows x 2 float4: each float4 becomes 4 fp16 hi + 4 fp16 lo = one 8-byte store each; a quarter-warp
                // covers channels [0,32) of its row with the even float4 and [32,64) with the odd one.
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    float4 sc = make_float4(a_mul, a_mul, a_mul, a_mul), sh = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (affine) {
                        sc = *reinterpret_cast<const float4*>(s_affine + c_c0 + 32 * q + 4 * chunk);
                        sh = *reinterpret_cast<const float4*>(s_affine + 256 + c_c0 + 32 * q + 4 * chunk);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float4 t = cur[2 * j + q];
                        t.x = fmaf(t.x, sc.x, sh.x); t.y = fmaf(t.y, sc.y, sh.y);
                        t.z = fmaf(t.z, sc.z, sh.z); t.w = fmaf(t.w, sc.w, sh.w);
                        if (p.relu) { t.x = relu_nan(t.x); t.y = relu_nan(t.y); t.z = relu_nan(t.z); t.w = relu_nan(t.w); }
                        if (!((cmask >> j) & 1u)) t = make_float4(0.f, 0.f, 0.f, 0.f);   // absent neighbour / row beyond the tile
                        amax = fmaxf(amax, fmaxf(fmaxf(fabsf(t.x), fabsf(t.y)), fmaxf(fabsf(t.z), fabsf(t.w))));
                        const __half2 h01 = __floats2half2_rn(t.x, t.y), h23 = __floats2half2_rn(t.z, t.w);
                        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                        const int r = wrow0 + sub + 4 * (4 * part + j);
                        uint8_t* dst = a_hi0 + r * 128 + (((4 * q + (chunk >> 1)) ^ (r & 7)) << 4) + ((chunk & 1) << 3);
                        uint2 hv, lv;
                        hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
                        *reinterpret_cast<uint2*>(dst) = hv;
                        if (PASSES == 3) {
                            const __half2 l01 = __floats2half2_rn(t.x - f01.x, t.y - f01.y), l23 = __floats2half2_rn(t.z - f23.x, t.w - f23.y);
                            lv.x = *reinterpret_cast<const uint32_t*>(&l01); lv.y = *reinterpret_cast<const uint32_t*>(&l23);
                            *reinterpret_cast<uint2*>(dst + a_bytes) = lv;
                        }
                    }
                }
                if (part == 0) return;
                fence_proxy_async();   // generic-proxy stores -> visible to the tensor core (async proxy)
            } else {
                float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
                if (affine) {
                    sc = *reinterpret_cast<const float4*>(s_affine + c_c0 + 4 * chunk);
                    sh = *reinterpret_cast<const float4*>(s_affine + 256 + c_c0 + 4 * chunk);
                }
                float4 tv[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 t = cur[j];
                    if (affine) {
                        t.x = fmaf(t.x, sc.x, sh.x); t.y = fmaf(t.y, sc.y, sh.y);
                        t.z = fmaf(t.z, sc.z, sh.z); t.w = fmaf(t.w, sc.w, sh.w);
                    }
                    if (p.relu) { t.x = relu_nan(t.x); t.y = relu_nan(t.y); t.z = relu_nan(t.z); t.w = relu_nan(t.w); }
                    if (!((cmask >> j) & 1u)) t = make_float4(0.f, 0.f, 0.f, 0.f);   // absent neighbour / row beyond the tile
                    tv[j] = t;
                }
                if (ATMEM) {
                    // The quarter-warp (lanes 8*sub .. 8*sub+7) holds an 8x8 block: register j = row 8*sub+j, lane c = chunk c.
                    // Three butterfly rounds transpose it in registers so that lane i of the warp owns ALL 32 floats of tile
                    // row wrow0 + i -- the layout tcgen05.st wants (thread = TMEM lane) -- without touching shared memory.
#pragma unroll
                    for (int m = 4; m >= 1; m >>= 1) {
                        const bool upper = (chunk & m) != 0;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if (j & m) continue;
                            const float4 a = tv[j], b = tv[j | m];
                            const float4 snd = upper ? a : b;
                            float4 rcv;
                            rcv.x = __shfl_xor_sync(0xffffffffu, snd.x, m); rcv.y = __shfl_xor_sync(0xffffffffu, snd.y, m);
                            rcv.z = __shfl_xor_sync(0xffffffffu, snd.z, m); rcv.w = __shfl_xor_sync(0xffffffffu, snd.w, m);
                            if (upper) tv[j] = rcv; else tv[j | m] = rcv;
                        }
                    }
                    float hi[32], lo[32];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        hi[4 * j] = tf32_hi(tv[j].x); hi[4 * j + 1] = tf32_hi(tv[j].y); hi[4 * j + 2] = tf32_hi(tv[j].z); hi[4 * j + 3] = tf32_hi(tv[j].w);
                        lo[4 * j] = tv[j].x - hi[4 * j]; lo[4 * j + 1] = tv[j].y - hi[4 * j + 1];
                        lo[4 * j + 2] = tv[j].z - hi[4 * j + 2]; lo[4 * j + 3] = tv[j].w - hi[4 * j + 3];
                    }
                    const uint32_t ta = tmem_base + ((uint32_t)wrow0 << 16) + a_col0 + (uint32_t)c_stage * 64u;
                    tmem_st32(ta, hi);
                    if (PASSES == 3) tmem_st32(ta + 32u, lo);
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    tc_fence_before();
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int r = wrow0 + sub + 4 * j;
                        const float4 t = tv[j];
                        const float4 h = make_float4(tf32_hi(t.x), tf32_hi(t.y), tf32_hi(t.z), tf32_hi(t.w));
                        uint8_t* dst = a_hi0 + r * 128 + ((chunk ^ (r & 7)) << 4);
                        *reinterpret_cast<float4*>(dst) = h;
                        if (PASSES == 3)
                            *reinterpret_cast<float4*>(dst + a_bytes) = make_float4(t.x - h.x, t.y - h.y, t.z - h.z, t.w - h.w);
                    }
                    fence_proxy_async();   // generic-proxy stores -> visible to the tensor core (async proxy)
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_full + 8 * c_stage);   // one arrival per warp, not 32 serialised ones
            c_c0 += kGroups * KB;
            while (c_c0 >= p.C) c_c0 -= p.C;
            c_stage += kGroups;
            while (c_stage >= SA) { c_stage -= SA; c_par ^= 1u; }
        };
        // software pipeline: the loads of the next unit are in flight while the current one is staged
        if (F16) {
            if (group < num_kb) issue(buf[0], bmask[0], 0);
            for (int kb = group; kb < num_kb; kb += kGroups) {
                issue(buf[1], bmask[1], 1);
                consume(buf[0], bmask[0], 0);
                if (kb + kGroups < num_kb) issue(buf[0], bmask[0], 0);
                consume(buf[1], bmask[1], 1);
            }
            // range check of the fp16 operands: anything at or beyond the largest finite half raises the caller's flag
            if (p.flag && !(amax < 65504.f)) atomicOr(p.flag, 1);
        } else {
            if (group < num_kb) issue(buf[0], bmask[0], 0);
            for (int kb = group; kb < num_kb; kb += 2 * kGroups) {
                const int kb1 = kb + kGroups, kb2 = kb + 2 * kGroups;
                if (kb1 < num_kb) issue(buf[1], bmask[1], 0);
                consume(buf[0], bmask[0], 0);
                if (kb1 < num_kb) {
                    if (kb2 < num_kb) issue(buf[0], bmask[0], 0);
                    consume(buf[1], bmask[1], 0);
                }
            }
        }

        }

        // ===================== epilogue: TMEM -> registers -> global (8 warps, alternate 32-column chunks) ====
        if (tid == 0) trace_stamp(p, 3);   // producers done
        mbar_wait(bar_acc, 0);
        tc_fence_after();
        if (tid == 0) trace_stamp(p, 4);   // accumulator complete
        // The accumulator leaves TMEM with thread = row (tcgen05.ld 32x32b), but a row-per-thread global store touches
        // 32 different cache lines per instruction, 16 bytes each.  Every warp therefore turns its 32 x 32 block around
        // in a private 4 KB of the (now idle) operand ring -- swizzled so that both directions are conflict-free -- and
        // finishes quarter-warp per row: whole 128-byte lines for the output stores and the residual loads, the
        // GroupNorm statistics of the output from 2 shuffle rounds per column quad.
        const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        float* stage = reinterpret_cast<float*>(smem + warp * 4096);
        const int wrow = (warp & 3) * 32;
        const float out_mul = F16 ? p.out_mul : 1.0f;
        for (int cb = group * 32; cb < N; cb += 32 * kGroups) {
            float acc[32];
            tmem_ld32(t_lane + (uint32_t)cb, acc);   // warp-collective: executed by every lane
#pragma unroll
            for (int j = 0; j < 8; ++j)
                *reinterpret_cast<float4*>(stage + lane * 32 + ((j ^ (lane & 7)) << 2)) =
                    make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
            __syncwarp();
            const bool col_ok = 4 * chunk < N - cb;   // N is a multiple of 8, so whole float4 groups
            const int col = n0 + cb + 4 * chunk;
            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bias && col_ok) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
            float4 cs = make_float4(0.f, 0.f, 0.f, 0.f), cq = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int r = it * 4 + sub;
                const int v = row0 + wrow + r;
                float4 t = *reinterpret_cast<const float4*>(stage + r * 32 + ((chunk ^ (r & 7)) << 2));
                if (F16) { t.x *= out_mul; t.y *= out_mul; t.z *= out_mul; t.w *= out_mul; }   // exact: power of two
                t.x += b4.x; t.y += b4.y; t.z += b4.z; t.w += b4.w;
                if (v < Vq && col_ok) {
                    if (p.res) {
                        const float4 q = __ldg(reinterpret_cast<const float4*>(p.res + (size_t)v * p.F + col));
                        t.x += q.x; t.y += q.y; t.z += q.z; t.w += q.w;
                    }
                    *reinterpret_cast<float4*>(p.out + (size_t)v * p.ldo + col) = t;
                    cs.x += t.x; cs.y += t.y; cs.z += t.z; cs.w += t.w;
                    cq.x = fmaf(t.x, t.x, cq.x); cq.y = fmaf(t.y, t.y, cq.y); cq.z = fmaf(t.z, t.z, cq.z); cq.w = fmaf(t.w, t.w, cq.w);
                }
            }
            __syncwarp();   // the block is consumed before the next chunk overwrites it
            if (p.out_sums) {
#pragma unroll
                for (int o = 8; o <= 16; o <<= 1) {   // the four row groups of a column quad
                    cs.x += __shfl_xor_sync(0xffffffffu, cs.x, o); cs.y += __shfl_xor_sync(0xffffffffu, cs.y, o);
                    cs.z += __shfl_xor_sync(0xffffffffu, cs.z, o); cs.w += __shfl_xor_sync(0xffffffffu, cs.w, o);
                    cq.x += __shfl_xor_sync(0xffffffffu, cq.x, o); cq.y += __shfl_xor_sync(0xffffffffu, cq.y, o);
                    cq.z += __shfl_xor_sync(0xffffffffu, cq.z, o); cq.w += __shfl_xor_sync(0xffffffffu, cq.w, o);
                }
                if (sub == 0 && col_ok) {
                    float* sc = s_colsum + cb + 4 * chunk;
                    atomicAdd(sc, cs.x); atomicAdd(sc + 1, cs.y); atomicAdd(sc + 2, cs.z); atomicAdd(sc + 3, cs.w);
                    atomicAdd(sc + 256, cq.x); atomicAdd(sc + 257, cq.y); atomicAdd(sc + 258, cq.z); atomicAdd(sc + 259, cq.w);
                }
            }
        }
        tc_fence_before();
        if (tid == 0) trace_stamp(p, 5);   // epilogue stores issued
    } else if (warp == kMmaWarp) {
        // ===================== MMA issuer: one thread =====================
        if (lane == 0) {
            // instruction descriptor: D = F32 [4,6), A / B format [7,10) [10,13) (2 = TF32, 0 = F16), K-major both,
            // N>>3 [17,23), M>>4 [24,29)
            const uint32_t fmt = F16 ? 0u : 2u;
            const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(Nmma >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);
            int sa = 0, sb = 0;
            uint32_t pa = 0, pb = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(bar_fullb + 8 * sb, pb);
                mbar_wait(bar_full + 8 * sa, pa);
                tc_fence_after();
                if (kb == 0) trace_stamp(p, 2);   // first operands staged
                const uint32_t st = smem_u32(smem + (size_t)sa * a_stage);
                const uint64_t a_hi = make_desc(st), a_lo = make_desc(st + a_bytes);
                const uint32_t bb = smem_u32(smem_b + (size_t)sb * b_stage);
                const uint64_t b_hi = make_desc(bb), b_lo = make_desc(bb + b_bytes);
                const uint32_t ta = tmem_base + a_col0 + (uint32_t)sa * 64u;   // ATMEM: this stage's A columns (hi at +0, lo at +32)
#pragma unroll
                for (int k = 0; k < 4; ++k) {   // 4 MMAs of 32 bytes of K per 128-byte swizzle row (8 tf32 / 16 fp16 elements)
                    const uint64_t adv = (uint64_t)((k * 32) >> 4);
                    if (ATMEM) {   // 8 tensor-memory columns per MMA: 8 tf32 or 16 fp16 elements of K
                        umma_ts<F16>(tmem_base, ta + k * 8, b_hi + adv, idesc, (kb | k) != 0);
                        if (PASSES == 3) {
                            umma_ts<F16>(tmem_base, ta + 32u + k * 8, b_hi + adv, idesc, 1);
                            umma_ts<F16>(tmem_base, ta + k * 8, b_lo + adv, idesc, 1);
                        }
                    } else {
                        umma_ss<F16>(tmem_base, a_hi + adv, b_hi + adv, idesc, (kb | k) != 0);
                        if (PASSES == 3) {
                            umma_ss<F16>(tmem_base, a_lo + adv, b_hi + adv, idesc, 1);
                            umma_ss<F16>(tmem_base, a_hi + adv, b_lo + adv, idesc, 1);
                        }
                    }
                }
                umma_commit(bar_empty + 8 * sa);      // both stages reusable once these MMAs have read them
                if (CL > 1) umma_commit_mcast(bar_emptyb + 8 * sb, cta_mask);   // ... the weight stage in every sibling
                else umma_commit(bar_emptyb + 8 * sb);
                if (++sa == SA) { sa = 0; pa ^= 1u; }
                if (++sb == SB) { sb = 0; pb ^= 1u; }
            }
            umma_commit(bar_acc);                      // accumulator complete -> epilogue
        }
        __syncwarp();
    } else {
        // ===================== TMA warp: stream the pre-split weight tiles =====================
        if (lane == 0) {
            const uint32_t tx = (PASSES == 3 ? 2u : 1u) * b_bytes;
            // stages [0, tma_kb0) were issued during the set-up (first lap of the ring: nothing to wait for)
            int sb = tma_kb0 == SB ? 0 : tma_kb0;
            uint32_t pb = tma_kb0 == SB ? 0u : 1u;
            for (int kb = tma_kb0; kb < num_kb; ++kb) {
                mbar_wait(bar_emptyb + 8 * sb, pb);
                const uint32_t bb = smem_u32(smem_b + (size_t)sb * b_stage);
                mbar_arrive_expect_tx(bar_fullb + 8 * sb, tx);   // the whole stage: own slice + the siblings' multicasts
                if (CL > 1) {
                    // this CTA fetches rows [rank * n_tile/CL, +n_tile/CL) of the tile and delivers them to all siblings
                    const int rows = p.n_tile / CL;
                    const uint32_t off = cta_rank * (uint32_t)rows * 128u;
                    tma_load_2d_mcast(bb + off, &map_hi, bar_fullb + 8 * sb, kb * KB, n0 + (int)cta_rank * rows, cta_mask);
                    if (PASSES == 3)
                        tma_load_2d_mcast(bb + b_bytes + off, &map_lo, bar_fullb + 8 * sb, kb * KB, n0 + (int)cta_rank * rows, cta_mask);
                } else {
                    tma_load_2d(bb, &map_hi, bar_fullb + 8 * sb, kb * KB, n0);
                    if (PASSES == 3) tma_load_2d(bb + b_bytes, &map_lo, bar_fullb + 8 * sb, kb * KB, n0);
                }
                if (++sb == SB) { sb = 0; pb ^= 1u; }
            }
        }
        __syncwarp();
    }
    __syncthreads();
    if (CL > 1) cluster_sync_all();   // nobody leaves while a sibling may still signal its barriers / fill its stages
    if (warp == kMmaWarp) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
    }
    if (tid == 0) trace_stamp(p, 6);   // teardown
    if (p.out_sums) {
        // per-CTA column sums -> group sums -> one double atomic per group
        const int g0 = n0 / p.out_cpg, g1 = (n0 + N) / p.out_cpg;   // n_tile is a multiple of out_cpg (checked on the host)
        for (int g = g0 + tid; g < g1; g += kThreads) {
            float a = 0.f, b = 0.f;
            for (int c = g * p.out_cpg - n0; c < (g + 1) * p.out_cpg - n0; ++c) { a += s_colsum[c]; b += s_colsum[256 + c]; }
            atomicAdd(p.out_sums + 2 * g, (double)a);
            atomicAdd(p.out_sums + 2 * g + 1, (double)b);
        }
    }
}

// weight -> K-major ([F,K]) copies split for the fp32-parity passes: hi = round-to-tf32(w), lo = w - hi.
// transposed_in = 0: w is [K,F] (reference conv layout);  1: w is [F,K] already (nn.Linear layout)
// HALF: fp16 copies of w * mul (mul a power of two chosen by the caller so the largest weight sits high in the
// half range): hi = half(w*mul), lo = half(w*mul - hi).
template <bool HALF>
__global__ void k_split(const float* __restrict__ in, int K, int F, int transposed_in, float mul, void* __restrict__ hi_v,
                        void* __restrict__ lo_v) {
    __shared__ float tile[32][33];
    int k0 = blockIdx.x * 32, f0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        if (transposed_in) {
            int f = f0 + i, k = k0 + threadIdx.x;
            tile[threadIdx.x][i] = (k < K && f < F) ? in[(size_t)f * K + k] : 0.f;
        } else {
            int k = k0 + i, f = f0 + threadIdx.x;
            tile[i][threadIdx.x] = (k < K && f < F) ? in[(size_t)k * F + f] : 0.f;
        }
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int f = f0 + i, k = k0 + threadIdx.x;
        if (f < F && k < K) {
            float w = tile[threadIdx.x][i];
            if (HALF) {
                __half* hi = reinterpret_cast<__half*>(hi_v);
                __half* lo = reinterpret_cast<__half*>(lo_v);
                w *= mul;
                const __half h = __float2half_rn(w);
                hi[(size_t)f * K + k] = h;
                if (lo) lo[(size_t)f * K + k] = __float2half_rn(w - __half2float(h));
            } else {
                float* hi = reinterpret_cast<float*>(hi_v);
                float* lo = reinterpret_cast<float*>(lo_v);
                float h = tf32_hi(w);
                hi[(size_t)f * K + k] = h;
                if (lo) lo[(size_t)f * K + k] = w - h;
            }
        }
    }
}

// common launcher.  half: fp16 operands (wt_* are __half copies scaled by 2^w_log2, activations are scaled by 2^a_log2).
int conv_launch(const float* x, int Vx, const int* vx_dev, const int* nbr, int Vq, const int* vq_dev, int C, int S,
                const void* wt_hi, const void* wt_lo, int F, const float* a_scale, const float* a_shift, const double* gn_sums,
                const float* gn_gamma, const float* gn_beta, float gn_eps, int gn_groups, int relu, const float* bias,
                const float* res, float* out, int ldo, double* out_sums, int out_groups, int passes, bool half, int a_log2,
                int w_log2, int* flag, void* stream) {
    if (Vq <= 0) return 0;
    if (Vx <= 0) return -2;   // row 0 of x must be readable (stand-in address of absent neighbours)
    const bool affine = a_scale || gn_sums;
    const int kb_elems = half ? 64 : kBlockK;
    if (C <= 0 || C % kb_elems || (affine && C > 256) || F <= 0 || F % 8 || S < 1 || (passes != 1 && passes != 3) || ldo % 4) return -2;
    if (half && passes != 3) return -2;
    if ((a_scale == nullptr) != (a_shift == nullptr) || (a_scale && gn_sums)) return -2;
    if (gn_sums && (gn_groups <= 0 || C % gn_groups)) return -2;
    if (out_sums && (out_groups <= 0 || F % out_groups)) return -2;
    if (passes == 3 && !wt_lo) return -2;
    ConvParams p;
    p.x = x; p.nbr = nbr; p.a_scale = a_scale; p.a_shift = a_shift; p.gn_sums = gn_sums; p.gn_gamma = gn_gamma; p.gn_beta = gn_beta;
    p.gn_eps = gn_eps; p.gn_cpg = gn_sums ? C / gn_groups : 1; p.bias = bias; p.res = res; p.out = out; p.out_sums = out_sums;
    p.out_cpg = out_sums ? F / out_groups : 1;
    p.vq_dev = vq_dev; p.vx_dev = vx_dev; p.Vq = Vq; p.Vx = Vx; p.C = C; p.S = nbr ? S : 1; p.F = F; p.ldo = ldo; p.relu = relu;
    p.a_mul = half ? ldexpf(1.0f, a_log2) : 1.0f;
    p.out_mul = half ? ldexpf(1.0f, -(a_log2 + w_log2)) : 1.0f;
    p.flag = half ? flag : nullptr;
    p.trace = g_trace;
    // Output channels per CTA.  All of them when they fit one accumulator (A rows gathered once); when that
    // leaves SMs idle (few row tiles) the channels are split further, but never beyond ONE wave of 148 CTAs
    // (one CTA per SM: a second, partly filled wave would cost a whole extra tile time).
    const int row_tiles = (Vq + kBlockM - 1) / kBlockM;
    const int unit = out_sums ? lcm16(p.out_cpg) : 16;   // tiles start on GroupNorm group boundaries
    int ny = (F + 255) / 256;
    while (row_tiles * (ny + 1) <= 148 && (F / (ny + 1)) >= 32) ++ny;
    int n_tile = ((F + ny - 1) / ny + unit - 1) / unit * unit;
    if (n_tile > 256) n_tile = 256 / unit * unit;
    if (n_tile <= 0) return -2;
    ny = (F + n_tile - 1) / n_tile;
    p.n_tile = n_tile;
    const size_t a_stage = (size_t)(passes == 3 ? 2 : 1) * kBlockM * 128, b_stage = (size_t)(passes == 3 ? 2 : 1) * n_tile * 128;
    const size_t budget = 208 * 1024;
    // the gather ring gets the depth first (measured: 3 A + 2 B stages beat 2 A + 3 B by 1.4x at N = 192), the
    // TMA ring takes what is left, at least two stages each
    // Measured (tools/bench_conv.py, 192->192 on 13.7k vertices): A staged in shared memory 60.6 us, A staged in tensor
    // memory 111.7 us (1-pass: 40.8 vs 79.1 us) -- correct, but the register transpose + tcgen05.st/wait::st chain in
    // the producers costs more than the tensor core's shared-memory reads it removes.  Off unless LTN_CONV_ATMEM=1.
    // fp16 operands: A staged in tensor memory is the default (the gather is laid out like the tcgen05.st fragment,
    // so nothing has to be transposed); LTN_CONV_ATMEM=0 keeps the shared-memory ring.
    static const int want_atmem = []() { const char* e = getenv("LTN_CONV_ATMEM"); return e ? atoi(e) : -1; }();
    const bool atmem = half ? (want_atmem != 0) : (want_atmem > 0);
    int sa, sb;
    size_t smem;
    if (atmem) {
        // A operand staged in TENSOR memory by the producers (tcgen05.st): shared memory only holds the weight ring,
        // and the tensor core's A reads (40 % of its shared-memory wavefronts at N = 192) disappear
        const int acc_cols = (((n_tile + 15) & ~15) + 31) & ~31;
        sa = (512 - acc_cols) / 64;
        if (sa > kMaxStages) sa = kMaxStages;
        if (sa < 2) return -3;
        static const int sb_cap = []() { const char* e = getenv("LTN_CONV_SB"); return e && atoi(e) >= 2 ? atoi(e) : 4; }();
        sb = (int)(budget / b_stage);
        if (sb > kMaxStagesB) sb = kMaxStagesB;
        if (sb > sb_cap) sb = sb_cap;   // measured: 2 / 3 / 4 / 8 stages -> see DESIGN.md (a shallower ring leaves more L1 to the gather)
        if (sb < 2) return -3;
        smem = sb * b_stage;
        if (smem < 12 * 4096) smem = 12 * 4096;   // the epilogue turns each warp's 32 x 32 block around in 4 KB of this region
        smem += 1024;
    } else {
        sa = (int)((budget - 2 * b_stage) / a_stage);
        if (sa > kMaxStages) sa = kMaxStages;
        if (sa < 2) return -3;
        sb = (int)((budget - sa * a_stage) / b_stage);
        if (sb > kMaxStages) sb = kMaxStages;
        if (sb < 2) return -3;
        smem = sa * a_stage + sb * b_stage + 1024;
    }
    p.stages_a = sa;
    p.stages_b = sb;
    // Row tiles that share a weight tile are grouped into thread-block clusters: each CTA fetches 1/CL of the
    // tile and multicasts it, so the L2 -> SM weight traffic (3x the gather traffic at N = 192) drops by CL.
    // Measured on B200 (bench_conv3, V = 13.7k): 192->192 72 us unclustered vs 78 us with clusters of 4, 128->128 on
    // 4.6k vertices 39 us vs 75 us -- the siblings' stages free in lock-step, which costs more than the L2 traffic
    // saves at these sizes.  Clusters therefore stay off unless LTN_CONV_CLUSTER asks for them.
    int cl = 1;
    static const int want_cl = []() { const char* e = getenv("LTN_CONV_CLUSTER"); return e ? atoi(e) : 1; }();
    if (want_cl >= 4 && row_tiles >= 4 && n_tile % 32 == 0) cl = 4;
    else if (want_cl >= 2 && row_tiles >= 2) cl = 2;
    p.cluster = cl;
    alignas(64) CUtensorMap map_hi, map_lo;
    int rc = make_weight_map(&map_hi, wt_hi, F, p.S * C, n_tile / cl, half);
    if (rc) return rc;
    rc = make_weight_map(&map_lo, passes == 3 ? wt_lo : wt_hi, F, p.S * C, n_tile / cl, half);
    if (rc) return rc;
    const void* fn = half ? (atmem ? (const void*)k_conv_tc<3, true, true> : (const void*)k_conv_tc<3, false, true>)
                   : passes == 3 ? (atmem ? (const void*)k_conv_tc<3, true, false> : (const void*)k_conv_tc<3, false, false>)
                                 : (atmem ? (const void*)k_conv_tc<1, true, false> : (const void*)k_conv_tc<1, false, false>);
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((row_tiles + cl - 1) / cl * cl, ny);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    void* args[3] = {(void*)&map_hi, (void*)&map_lo, (void*)&p};
    e = cudaLaunchKernelExC(&cfg, fn, args);
    if (e != cudaSuccess) return (int)e;
    LTN_CHECK_LAUNCH();
    return 0;
}

}  // namespace

extern "C" {

int ltn_split_tf32(const float* w, int K, int F, int transposed_in, float* wt_hi, float* wt_lo, void* stream) {
    if (K <= 0 || F <= 0) return 0;
    dim3 grid((K + 31) / 32, (F + 31) / 32), block(32, 8);
    k_split<false><<<grid, block, 0, (cudaStream_t)stream>>>(w, K, F, transposed_in, 1.0f, wt_hi, wt_lo);
    LTN_CHECK_LAUNCH();
    return 0;
}

int ltn_split_f16(const float* w, int K, int F, int transposed_in, int w_log2, void* wt_hi, void* wt_lo, void* stream) {
    if (K <= 0 || F <= 0) return 0;
    dim3 grid((K + 31) / 32, (F + 31) / 32), block(32, 8);
    k_split<true><<<grid, block, 0, (cudaStream_t)stream>>>(w, K, F, transposed_in, ldexpf(1.0f, w_log2), wt_hi, wt_lo);
    LTN_CHECK_LAUNCH();
    return 0;
}

// Phase tracing of the NEXT ltn_conv_tc* launches: buf (device, nullable to switch off) receives 8 globaltimer stamps
// per CTA: entry, set-up done, first operands staged, producers done, accumulator complete, epilogue stores issued, teardown.
int ltn_conv_trace(unsigned long long* buf) {
    g_trace = buf;
    return 0;
}

// Fused gather + GEMM on tcgen05.  See the header of this file; the C ABI is documented in
// include/latticenet_b200.h.  passes: 3 = fp32-parity split (default), 1 = single-pass TF32.
int ltn_conv_tc(const float* x, int Vx, const int* vx_dev, const int* nbr, int Vq, const int* vq_dev, int C, int S,
                const float* wt_hi, const float* wt_lo, int F, const float* a_scale, const float* a_shift,
                const double* gn_sums, const float* gn_gamma, const float* gn_beta, float gn_eps, int gn_groups, int relu,
                const float* bias, const float* res, float* out, int ldo, double* out_sums, int out_groups, int passes,
                void* stream) {
    return conv_launch(x, Vx, vx_dev, nbr, Vq, vq_dev, C, S, wt_hi, wt_lo, F, a_scale, a_shift, gn_sums, gn_gamma, gn_beta, gn_eps,
                       gn_groups, relu, bias, res, out, ldo, out_sums, out_groups, passes, false, 0, 0, nullptr, stream);
}

// The same operation with fp16 hi/lo operands (fp32-parity, three passes): wt_hi / wt_lo from ltn_split_f16 with
// the same w_log2; activations are scaled by 2^a_log2 before the split; *flag (nullable) is OR-ed with 1 when a
// scaled activation reaches the end of the half range (the result is then unusable: redo with ltn_conv_tc).
int ltn_conv_tc_f16(const float* x, int Vx, const int* vx_dev, const int* nbr, int Vq, const int* vq_dev, int C, int S,
                    const void* wt_hi, const void* wt_lo, int w_log2, int a_log2, int F, const float* a_scale,
                    const float* a_shift, const double* gn_sums, const float* gn_gamma, const float* gn_beta, float gn_eps,
                    int gn_groups, int relu, const float* bias, const float* res, float* out, int ldo, double* out_sums,
                    int out_groups, int* flag, void* stream) {
    return conv_launch(x, Vx, vx_dev, nbr, Vq, vq_dev, C, S, wt_hi, wt_lo, F, a_scale, a_shift, gn_sums, gn_gamma, gn_beta, gn_eps,
                       gn_groups, relu, bias, res, out, ldo, out_sums, out_groups, 3, true, a_log2, w_log2, flag, stream);
}

}  // extern "C"
