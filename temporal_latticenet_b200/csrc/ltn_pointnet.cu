// PointNetSeqModule front end (seq_lattice/lattice_modules.py:448-530) as two kernels instead of three
// materialised [4N, 16/32/64] activations + scatter_max + index_select + cat + masked_fill:
//
//   k_pointnet_mlp_max : per distributed row  MLP 4 -> 16 -> 32 -> 64 (ReLU between, none after the last,
//                        lattice_modules.py:460-473) entirely in registers, weights broadcast from shared
//                        memory, then the segmented max onto the row's vertex: first inside the block in shared
//                        memory, then one 64-bit atomicMax per (distinct vertex of the block, channel) (value in
//                        the high word, ~row in the low word: largest value wins, smallest row breaks ties --
//                        torch_scatter's result up to its own race on ties).
//   k_pointnet_decode  : per (vertex, channel) unpack max / arg-max, apply quirk Q3 literally
//                        (`argmax_clone[argmax > argmax.shape[0]] = 0`, lattice_modules.py:513-514), gather the
//                        barycentric weight of the winning row, concatenate [max(64) | bary(64)] and zero the
//                        vertices with fewer than 4 contributing rows (lattice_modules.py:519-530).
//
// HBM-bound by construction: reads 4N x 24 B of rows + ids once, writes V x 512 B; the 4N x 64 atomics
// resolve in L2 (V x 64 x 8 B = 7 MB at V = 14k).
#include "ltn_common.cuh"
#include "ltn_tc.cuh"
#include <cuda_fp16.h>
#include <cstdlib>

namespace {

constexpr int kThreads = 256;
constexpr int D0 = 4, D1 = 16, D2 = 32, D3 = 64;

__device__ __forceinline__ uint32_t ord_enc(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord_dec(uint32_t e) {
    return __uint_as_float((e & 0x80000000u) ? (e & 0x7FFFFFFFu) : ~e);
}

struct MlpWeights {
    const float *w1, *b1, *w2, *b2, *w3, *b3;  // nn.Linear layout [out, in]
};

// Layers 1-2 BY VALUE (2.5 KB of kernel parameters = constant bank): every FFMA of the fully unrolled loops takes its
// weight as a constant-bank operand, so the 576 multiply-adds per row cost no shared-memory reads at all (they were a
// third of the shared-memory wavefronts that bound k_pointnet_tc).  [in][out] order, as the inner loops walk it.
struct Mlp12 {
    float w1[D0 * D1], b1[D1], w2[D1 * D2], b2[D2];
};

long long* g_pn_trace = nullptr;
constexpr int kPointnetTcThreads = 128;   // rows per block of k_pointnet_tc (measured); LTN_PN_THREADS=256 for the other form
constexpr int kHash = 512;   // block-local vertex table (256 rows -> at most 256 distinct vertices)
constexpr int kSlots = 128;  // distinct vertices per block resolved in shared memory (typical: 30-60); the rest go straight to L2
constexpr int kPitch = kSlots + 1;

// One block = 256 consecutive distributed rows = 64 consecutive points.  Scan neighbours share their simplex
// vertices (~90 rows per vertex overall), so the segmented max is first resolved INSIDE the block in shared
// memory -- value with a 32-bit atomicMax, arg-max row with a 32-bit atomicMin among the rows that reach it --
// and only one 64-bit atomicMax per (distinct vertex of the block, channel) goes to L2, 128 contiguous bytes per
// vertex quarter.  (Measured before: one atomic per (row, channel), 32 different vertices per warp instruction,
// the L2 83 % busy and the kernel 183 us per 500k rows.)
__global__ void __launch_bounds__(kThreads, 2)
k_pointnet_mlp_max(const float* __restrict__ rows, int width, const int* __restrict__ idx, int R,
                   const int* __restrict__ r_dev, MlpWeights w, int V, const int* __restrict__ v_dev,
                   unsigned long long* packed) {
    // weights transposed to [in][out] so a thread's inner loop reads consecutive (broadcast) words
    __shared__ __align__(16) float s1[D0 * D1], s2[D1 * D2], s3[D2 * D3], sb1[D1], sb2[D2], sb3[D3];
    __shared__ int h_key[kHash], h_slot[kHash], slot_id[kThreads], nslots;
    // per-quarter tables [channel][slot], padded so that both the per-channel atomics of a warp's group leaders
    // (different slots) and the flush (16 channels of one slot) spread over the banks
    __shared__ uint32_t t_val[16 * kPitch], t_row[16 * kPitch];
    for (int i = threadIdx.x; i < D0 * D1; i += kThreads) s1[(i % D0) * D1 + i / D0] = __ldg(w.w1 + i);
    for (int i = threadIdx.x; i < D1 * D2; i += kThreads) s2[(i % D1) * D2 + i / D1] = __ldg(w.w2 + i);
    for (int i = threadIdx.x; i < D2 * D3; i += kThreads) s3[(i % D2) * D3 + i / D2] = __ldg(w.w3 + i);
    for (int i = threadIdx.x; i < D1; i += kThreads) sb1[i] = __ldg(w.b1 + i);
    for (int i = threadIdx.x; i < D2; i += kThreads) sb2[i] = __ldg(w.b2 + i);
    for (int i = threadIdx.x; i < D3; i += kThreads) sb3[i] = __ldg(w.b3 + i);
    if (r_dev) R = min(R, *r_dev);
    if (v_dev) V = min(V, *v_dev);
    for (int base = blockIdx.x * kThreads; base < R; base += gridDim.x * kThreads) {   // block-uniform trip count
        const int row = base + threadIdx.x;
        int id = -1;
        if (row < R) {
            id = __ldg(idx + row);
            id = id < 0 ? 0 : id;                   // lattice_modules.py:479-480
            if (id >= V) id = -1;
        }
        // ---- block-local numbering of the distinct vertices ------------------------------------------------
        for (int i = threadIdx.x; i < kHash; i += kThreads) h_key[i] = -1;
        if (threadIdx.x == 0) nslots = 0;
        __syncthreads();   // also orders the weight staging before its first use
        int hpos = 0;
        if (id >= 0) {
            hpos = (int)(((uint32_t)id * 2654435761u) >> 23) & (kHash - 1);
            while (true) {
                const int old = atomicCAS(&h_key[hpos], -1, id);
                if (old == -1 || old == id) break;
                hpos = (hpos + 1) & (kHash - 1);
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < kHash; i += kThreads)
            if (h_key[i] >= 0) {
                const int sl = atomicAdd(&nslots, 1);
                h_slot[i] = sl;
                slot_id[sl] = h_key[i];
            }
        __syncthreads();
        const int lslot = id >= 0 ? h_slot[hpos] : -1;
        const int ns = min(nslots, kSlots);
        const bool local = lslot >= 0 && lslot < kSlots;

        // ---- MLP 4 -> 16 -> 32 in registers ----------------------------------------------------------------
        float h2[D2];
        {
            float x[D0] = {0.f, 0.f, 0.f, 0.f};
            if (id >= 0) {
                const float* in = rows + (size_t)row * width;
#pragma unroll
                for (int i = 0; i < D0; ++i) x[i] = __ldg(in + i);
            }
            float h1[D1];
#pragma unroll
            for (int o = 0; o < D1; ++o) h1[o] = sb1[o];
#pragma unroll
            for (int i = 0; i < D0; ++i)
#pragma unroll
                for (int o = 0; o < D1; ++o) h1[o] = fmaf(x[i], s1[i * D1 + o], h1[o]);
#pragma unroll
            for (int o = 0; o < D2; ++o) h2[o] = sb2[o];
#pragma unroll
            for (int i = 0; i < D1; ++i) {
                const float a = fmaxf(h1[i], 0.f);
#pragma unroll
                for (int o = 0; o < D2; o += 4) {
                    const float4 ww = *reinterpret_cast<const float4*>(s2 + i * D2 + o);
                    h2[o] = fmaf(a, ww.x, h2[o]); h2[o + 1] = fmaf(a, ww.y, h2[o + 1]);
                    h2[o + 2] = fmaf(a, ww.z, h2[o + 2]); h2[o + 3] = fmaf(a, ww.w, h2[o + 3]);
                }
            }
#pragma unroll
            for (int i = 0; i < D2; ++i) h2[i] = fmaxf(h2[i], 0.f);
        }
        // ---- last layer in four quarters of 16 outputs (bounds the live registers: two blocks per SM) ---------
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
            for (int e = threadIdx.x; e < ns * 16; e += kThreads) {
                const int at = (e & 15) * kPitch + (e >> 4);
                t_val[at] = 0u;
                t_row[at] = 0xFFFFFFFFu;
            }
            float y[16];
#pragma unroll
            for (int o = 0; o < 16; ++o) y[o] = sb3[q * 16 + o];
#pragma unroll
            for (int i = 0; i < D2; ++i) {
                const float a = h2[i];
#pragma unroll
                for (int o = 0; o < 16; o += 4) {
                    const float4 ww = *reinterpret_cast<const float4*>(s3 + i * D3 + q * 16 + o);
                    y[o] = fmaf(a, ww.x, y[o]); y[o + 1] = fmaf(a, ww.y, y[o + 1]);
                    y[o + 2] = fmaf(a, ww.z, y[o + 2]); y[o + 3] = fmaf(a, ww.w, y[o + 3]);
                }
            }
            uint32_t enc[16];
#pragma unroll
            for (int o = 0; o < 16; ++o) enc[o] = ord_enc(y[o]);
            __syncthreads();
            // (a warp reduction per vertex group first -- __reduce_max_sync over __match_any_sync masks -- was measured
            // 2.5x SLOWER: divergent masks make the warp run the whole loop once per group)
            if (local) {
#pragma unroll
                for (int o = 0; o < 16; ++o) atomicMax(&t_val[o * kPitch + lslot], enc[o]);
            }
            __syncthreads();
            if (local) {
                unsigned win = 0u;   // loads first, then atomics for the winners only (see k_pointnet_tc)
#pragma unroll
                for (int o = 0; o < 16; ++o) win |= (unsigned)(t_val[o * kPitch + lslot] == enc[o]) << o;
#pragma unroll
                for (int o = 0; o < 16; ++o)
                    if ((win >> o) & 1u) atomicMin(&t_row[o * kPitch + lslot], (uint32_t)row);   // smallest row wins ties
            }
            __syncthreads();
            // one 64-bit atomic per (distinct vertex, channel): value in the high word, ~row in the low word
            for (int e = threadIdx.x; e < ns * 16; e += kThreads) {
                const int sl = e >> 4, c = e & 15;
                const unsigned long long key = ((unsigned long long)t_val[c * kPitch + sl] << 32) |
                                               (unsigned long long)(0xFFFFFFFFu - t_row[c * kPitch + sl]);
                // fire-and-forget reduction: no read-back, so nothing waits for an L2 round trip (a `__ldcg` pre-check
                // here made every thread sit through ~3 dependent L2 latencies per quarter: 157 us -> see profiles/)
                atomicMax(packed + (size_t)slot_id[sl] * D3 + q * 16 + c, key);
            }
            if (lslot >= kSlots) {   // more distinct vertices than the shared tables hold: this row goes to L2 directly
                unsigned long long* dst = packed + (size_t)id * D3 + q * 16;
                const unsigned long long low = (unsigned long long)(0xFFFFFFFFu - (uint32_t)row);
#pragma unroll
                for (int o = 0; o < 16; ++o) {
                    const unsigned long long key = ((unsigned long long)enc[o] << 32) | low;
                    if (__ldcg(dst + o) < key) atomicMax(dst + o, key);
                }
            }
            __syncthreads();
        }
    }
}

// The same front end with the LAST MLP layer (32 -> 64: 2048 of the 2624 multiply-adds per row, and most of the
// shared-memory weight reads that bound the kernel above) on the tensor cores.
//   * a block owns 256 consecutive rows = two M = 128 accumulator tiles; thread t computes layers 1-2 for ITS row in
//     registers, splits relu(h2) * 2^a_log2 into fp16 hi + lo (the 11 + 11 significant bits of the fp32-parity
//     convolution, csrc/ltn_conv.cu) and writes both straight into tensor memory with tcgen05.st.32x32b -- thread =
//     row = TMEM lane is exactly the A-operand layout of a TS-mode MMA, so the activations never touch shared memory;
//   * W3 is split the same way by every block into a 128-byte-swizzled K-major tile (K padded to one 128-byte row),
//     scaled by a power of two taken from max|W3|;
//   * 2 k-steps x 3 passes (hi*hi + lo*hi + hi*lo) per tile by one thread, fp32 accumulation in TMEM;
//   * each thread reads its row of the accumulator back 16 columns at a time and feeds the same block-local
//     segmented max as above.
// *flag is OR-ed with 1 when a staged activation leaves the fp16 range (the caller then uses k_pointnet_mlp_max).
template <int NT>   // threads = rows per block: 256 (two M = 128 tiles, 2 blocks / SM) or 128 (one tile, 4 blocks / SM)
__global__ void __launch_bounds__(NT, NT == 128 ? 4 : 2)
k_pointnet_tc(const float* __restrict__ rows, int width, const int* __restrict__ idx, int R, const int* __restrict__ r_dev,
              MlpWeights w, const __grid_constant__ Mlp12 M, int V, const int* __restrict__ v_dev, unsigned long long* packed,
              float a_mul, int* flag, long long* trace) {
    __shared__ __align__(1024) uint8_t s_w3[2 * 64 * 128];   // [hi | lo] 64 output rows x 128 bytes, SWIZZLE_128B
    __shared__ __align__(16) float sb3[D3];
    constexpr int kThreads = NT, kTiles = NT / 128, kHash = 2 * NT, kSlots = NT / 2, kPitch = kSlots + 1;
    constexpr int kTmemCols = NT;          // kTiles x (64 accumulator + 32 operand columns) rounded up to a power of two
    __shared__ int h_key[kHash], h_slot[kHash], slot_id[kThreads], nslots;
    extern __shared__ uint32_t t_dyn[];   // all 64 channels at once: 7 block barriers per tile instead of 20
    uint32_t* const t_val = t_dyn;
    uint32_t* const t_row = t_dyn + D3 * kPitch;
    __shared__ __align__(8) uint64_t bar_mma;
    __shared__ uint32_t tmem_slot;
    __shared__ float s_wmax[kThreads / 32];
    __shared__ float s_wmul;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < D3; i += kThreads) sb3[i] = __ldg(w.b3 + i);
    // ---- W3 [64, 32] -> fp16 hi / lo, K-major, 128-byte swizzle; item t: output row t / 4, 8 inputs (one 16-byte chunk)
    {
        float m = 0.f;
        for (int t = tid; t < 256; t += kThreads) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(w.w3) + t * 2), b = __ldg(reinterpret_cast<const float4*>(w.w3) + t * 2 + 1);
            m = fmaxf(m, fmaxf(fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))),
                               fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w)))));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) s_wmax[warp] = m;
    }
    if (tid == 0) {
        mbar_init(smem_u32(&bar_mma), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    __syncthreads();
    if (tid == 0) {
        float m = 0.f;
        for (int i = 0; i < kThreads / 32; ++i) m = fmaxf(m, s_wmax[i]);
        // largest |w| -> [2^13, 2^14): exact power-of-two scale, undone in the epilogue
        int e = 0;
        if (m > 0.f && m < 3.0e38f) e = 13 - (int)((__float_as_uint(m) >> 23) & 0xFFu) + 127;
        e = max(-60, min(60, e));
        s_wmul = __uint_as_float((uint32_t)(127 + e) << 23);
    }
    __syncthreads();
    const float w_mul = s_wmul;
    for (int t = tid; t < 256; t += kThreads) {
        const int n = t >> 2, chunk = t & 3;
        const float4 a = __ldg(reinterpret_cast<const float4*>(w.w3) + t * 2), b = __ldg(reinterpret_cast<const float4*>(w.w3) + t * 2 + 1);
        const float wv[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float x0 = wv[2 * i] * w_mul, x1 = wv[2 * i + 1] * w_mul;
            const __half2 h = __floats2half2_rn(x0, x1);
            const float2 f = __half22float2(h);
            const __half2 l = __floats2half2_rn(x0 - f.x, x1 - f.y);
            hi[i] = *reinterpret_cast<const uint32_t*>(&h);
            lo[i] = *reinterpret_cast<const uint32_t*>(&l);
        }
        uint8_t* dst = s_w3 + n * 128 + ((chunk ^ (n & 7)) << 4);
        *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(dst + 64 * 128) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        // the upper half of every 128-byte row (K = 32..63) is never read: the MMAs stop after two 32-byte k-steps
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t acc_col = (uint32_t)(warp >> 2) * 64u;                     // accumulator of this thread's tile
    const uint32_t a_col = (uint32_t)kTiles * 64u + (uint32_t)(warp >> 2) * 32u;   // its A operand: 16 columns hi, 16 columns lo
    const float out_mul = 1.0f / (a_mul * w_mul);                  // both powers of two: exact
    if (r_dev) R = min(R, *r_dev);
    if (v_dev) V = min(V, *v_dev);
    float amax = 0.f;
    uint32_t phase = 0;
    for (int base = blockIdx.x * kThreads; base < R; base += gridDim.x * kThreads) {   // block-uniform trip count
        long long tk[8];
        tk[0] = clock64();
        const int row = base + tid;
        int id = -1;
        if (row < R) {
            id = __ldg(idx + row);
            id = id < 0 ? 0 : id;                   // lattice_modules.py:479-480
            if (id >= V) id = -1;
        }
        for (int i = tid; i < kHash; i += kThreads) h_key[i] = -1;
        if (tid == 0) nslots = 0;
        // ---- MLP 4 -> 16 -> 32 in registers, relu, fp16 hi / lo straight into tensor memory ---------------------
        {
            float x[D0] = {0.f, 0.f, 0.f, 0.f};
            if (id >= 0) {
                const float* in = rows + (size_t)row * width;
#pragma unroll
                for (int i = 0; i < D0; ++i) x[i] = __ldg(in + i);
            }
            float h1[D1], h2[D2];
#pragma unroll
            for (int o = 0; o < D1; ++o) h1[o] = M.b1[o];
#pragma unroll
            for (int i = 0; i < D0; ++i)
#pragma unroll
                for (int o = 0; o < D1; ++o) h1[o] = fmaf(x[i], M.w1[i * D1 + o], h1[o]);
#pragma unroll
            for (int o = 0; o < D2; ++o) h2[o] = M.b2[o];
#pragma unroll
            for (int i = 0; i < D1; ++i) {
                const float a = fmaxf(h1[i], 0.f);
#pragma unroll
                for (int o = 0; o < D2; ++o) h2[o] = fmaf(a, M.w2[i * D2 + o], h2[o]);   // constant-bank operand: no LDS
            }
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float x0 = fmaxf(h2[2 * i], 0.f) * a_mul, x1 = fmaxf(h2[2 * i + 1], 0.f) * a_mul;
                amax = fmaxf(amax, fmaxf(x0, x1));
                const __half2 h = __floats2half2_rn(x0, x1);
                const float2 f = __half22float2(h);
                const __half2 l = __floats2half2_rn(x0 - f.x, x1 - f.y);
                hi[i] = *reinterpret_cast<const uint32_t*>(&h);
                lo[i] = *reinterpret_cast<const uint32_t*>(&l);
            }
            tmem_st16(t_lane + a_col, hi);
            tmem_st16(t_lane + a_col + 16u, lo);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        tk[1] = clock64();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            // D = F32, A = B = F16, K-major, N = 64, M = 128
            const uint32_t idesc = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint64_t b_hi = make_desc(smem_u32(s_w3)), b_lo = make_desc(smem_u32(s_w3 + 64 * 128));
#pragma unroll
            for (int m = 0; m < kTiles; ++m) {
                const uint32_t d = tmem_base + (uint32_t)m * 64u, a = tmem_base + (uint32_t)kTiles * 64u + (uint32_t)m * 32u;
#pragma unroll
                for (int k = 0; k < 2; ++k) {   // 16 fp16 of K = 8 columns of A = 32 bytes of the B row
                    const uint64_t adv = (uint64_t)((k * 32) >> 4);
                    umma_f16_ts(d, a + k * 8, b_hi + adv, idesc, k);
                    umma_f16_ts(d, a + 16u + k * 8, b_hi + adv, idesc, 1);
                    umma_f16_ts(d, a + k * 8, b_lo + adv, idesc, 1);
                }
            }
            umma_commit(smem_u32(&bar_mma));
        }
        // ---- block-local numbering of the distinct vertices (overlaps the MMAs) -------------------------------
        int hpos = 0;
        if (id >= 0) {
            hpos = (int)(((uint32_t)id * 2654435761u) >> 22) & (kHash - 1);
            while (true) {
                const int old = atomicCAS(&h_key[hpos], -1, id);
                if (old == -1 || old == id) break;
                hpos = (hpos + 1) & (kHash - 1);
            }
        }
        __syncthreads();
        for (int i = tid; i < kHash; i += kThreads)
            if (h_key[i] >= 0) {
                const int sl = atomicAdd(&nslots, 1);
                h_slot[i] = sl;
                slot_id[sl] = h_key[i];
            }
        __syncthreads();
        const int lslot = id >= 0 ? h_slot[hpos] : -1;
        const int ns = min(nslots, kSlots);
        const bool local = lslot >= 0 && lslot < kSlots;
        tk[2] = clock64();
        for (int e = tid; e < ns * D3; e += kThreads) {
            const int at = (e & (D3 - 1)) * kPitch + (e >> 6);
            t_val[at] = 0u;
            t_row[at] = 0xFFFFFFFFu;
        }
        mbar_wait(smem_u32(&bar_mma), phase);
        phase ^= 1u;
        tc_fence_after();
        uint32_t enc[D3];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float y[16];
            tmem_ld16(t_lane + acc_col + (uint32_t)(q * 16), y);   // warp-collective
#pragma unroll
            for (int o = 0; o < 16; ++o) enc[q * 16 + o] = ord_enc(fmaf(y[o], out_mul, sb3[q * 16 + o]));
        }
        tk[3] = clock64();
        __syncthreads();
        if (local) {
#pragma unroll
            for (int o = 0; o < D3; ++o) atomicMax(&t_val[o * kPitch + lslot], enc[o]);
        }
        tk[4] = clock64();
        __syncthreads();
        if (local) {
            // which channels did this row win?  Pure loads first (they pipeline; interleaved with the atomics every load
            // waited for the previous atomic: 9.8k of the 18k cycles of a tile), then predicated atomics for the few winners
            unsigned long long win = 0ull;
#pragma unroll
            for (int o = 0; o < D3; ++o) win |= (unsigned long long)(t_val[o * kPitch + lslot] == enc[o]) << o;
            uint32_t my_row = (uint32_t)row;
            asm volatile("" : "+r"(my_row));   // keep it in a register: rematerialising it from %tid.x costs an S2R per channel
            const uint32_t tr = smem_u32(t_row + lslot);
            const uint32_t w_lo = (uint32_t)win, w_hi = (uint32_t)(win >> 32);
#pragma unroll
            for (int o = 0; o < D3; ++o) {   // predicated reductions, no branch per channel; smallest row wins ties
                const uint32_t bit = ((o < 32 ? w_lo : w_hi) >> (o & 31)) & 1u;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p red.shared.min.u32 [%0], %1;\n\t}"
                             ::"r"(tr + (uint32_t)(o * kPitch * 4)), "r"(my_row), "r"(bit) : "memory");
            }
        }
        tk[5] = clock64();
        __syncthreads();
        // one fire-and-forget 64-bit reduction per (distinct vertex, channel): value in the high word, ~row in the low word;
        // 64 consecutive channels of a vertex = 512 contiguous bytes per two warps
        for (int e = tid; e < ns * D3; e += kThreads) {
            const int sl = e >> 6, c = e & (D3 - 1);
            const unsigned long long key = ((unsigned long long)t_val[c * kPitch + sl] << 32) |
                                           (unsigned long long)(0xFFFFFFFFu - t_row[c * kPitch + sl]);
            atomicMax(packed + (size_t)slot_id[sl] * D3 + c, key);
        }
        if (lslot >= kSlots) {   // more distinct vertices than the shared tables hold: this row goes to L2 directly
            unsigned long long* dst = packed + (size_t)id * D3;
            const unsigned long long low = (unsigned long long)(0xFFFFFFFFu - (uint32_t)row);
#pragma unroll
            for (int o = 0; o < D3; ++o) atomicMax(dst + o, ((unsigned long long)enc[o] << 32) | low);
        }
        tk[6] = clock64();
        __syncthreads();
        tk[7] = clock64();
        if (trace && blockIdx.x == 0 && tid == 0 && base == (int)(gridDim.x * kThreads)) {   // second tile of block 0
            for (int i = 0; i < 8; ++i) trace[i] = tk[i];
            trace[8] = ns;
        }
        tc_fence_before();   // the next tile's tcgen05.st / MMAs follow this tile's tcgen05.ld across the block barrier above
    }
    if (flag && !(amax < 65504.f)) atomicOr(flag, 1);
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
    }
}

// out [V, 128] = [ max | bary of the arg-max row (Q3) ], rows with fewer than min_rows contributors zeroed
__global__ void __launch_bounds__(kThreads)
k_pointnet_decode(const unsigned long long* __restrict__ packed, int V, const int* __restrict__ v_dev, int R,
                  const int* __restrict__ r_dev, const float* __restrict__ rows, int width, const double* __restrict__ vert_acc,
                  int min_rows, float* __restrict__ out) {
    if (r_dev) R = min(R, *r_dev);
    if (v_dev) V = min(V, *v_dev);
    long long total = (long long)V * D3;
    for (long long t = (long long)blockIdx.x * kThreads + threadIdx.x; t < total; t += (long long)gridDim.x * kThreads) {
        const int v = (int)(t / D3), c = (int)(t - (long long)v * D3);
        const unsigned long long k = packed[t];
        float val = 0.f;
        long long arg = R;                                  // torch_scatter's empty-segment sentinel
        if (k != 0ull) {
            val = ord_dec((uint32_t)(k >> 32));
            arg = (long long)(0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFull));
        }
        if (arg > V) arg = 0;                               // quirk Q3, literally (row index compared with V)
        if (arg > R - 1) arg = R - 1;
        float bary = (R > 0) ? __ldg(rows + (size_t)arg * width + (width - 1)) : 0.f;
        if (min_rows > 0 && vert_acc[(size_t)v * 4 + 3] < (double)min_rows) { val = 0.f; bary = 0.f; }
        out[(size_t)v * (2 * D3) + c] = val;
        out[(size_t)v * (2 * D3) + D3 + c] = bary;
    }
}

}  // namespace

extern "C" {

// rows [R, width] (first 4 columns feed the MLP, last column = barycentric weight), idx [R];
// packed [V,64] u64 scratch (zeroed here); vert_acc: the accumulator ltn_distribute left ([cap,4] double,
// [:,3] = rows per vertex); min_rows = 4 (0 disables the mask: early max-pool fusion); out [V,128].
int ltn_pointnet(const float* rows, int width, const int* idx, int R, const int* r_dev, const float* w1, const float* b1,
                 const float* w2, const float* b2, const float* w3, const float* b3, int V, const int* v_dev,
                 unsigned long long* packed, const double* vert_acc, int min_rows, float* out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (V <= 0) return 0;
    if (width != D0 + 1) return -2;
    cudaError_t e = cudaMemsetAsync(packed, 0, sizeof(unsigned long long) * (size_t)V * D3, st);
    if (e != cudaSuccess) return (int)e;
    if (R > 0) {
        MlpWeights w{w1, b1, w2, b2, w3, b3};
        int blocks = (R + kThreads - 1) / kThreads;
        if (blocks > 148 * 8) blocks = 148 * 8;
        k_pointnet_mlp_max<<<blocks, kThreads, 0, st>>>(rows, width, idx, R, r_dev, w, V, v_dev, packed);
        LTN_CHECK_LAUNCH();
    }
    long long total = (long long)V * D3;
    int blocks = (int)((total + kThreads - 1) / kThreads);
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_pointnet_decode<<<blocks, kThreads, 0, st>>>(packed, V, v_dev, R, r_dev, rows, width, vert_acc, min_rows, out);
    LTN_CHECK_LAUNCH();
    return 0;
}

// debugging aid: clock64 stamps of the phases of one tile of k_pointnet_tc (block 0, second tile) -> buf[0..7], buf[8] = distinct vertices
int ltn_pointnet_trace(long long* buf) {
    g_pn_trace = buf;
    return 0;
}

// ltn_pointnet with the 32 -> 64 layer on the tensor cores (fp16 hi/lo operands, fp32-parity three passes; activations
// staged as relu(h2) * 2^a_log2).  *flag (int32) is OR-ed with 1 when an activation leaves the fp16 range: the result is
// then unusable and the caller redoes the frame with ltn_pointnet.
int ltn_pointnet_tc(const float* rows, int width, const int* idx, int R, const int* r_dev, const float* w12_host,
                    const float* w3, const float* b3, int V, const int* v_dev, unsigned long long* packed,
                    const double* vert_acc, int min_rows, float* out, int a_log2, int* flag, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (V <= 0) return 0;
    if (width != D0 + 1 || !flag || !w12_host) return -2;
    Mlp12 M;
    // host copy of layers 1-2 in nn.Linear layout: w1 [16,4], b1 [16], w2 [32,16], b2 [32] -> [in][out]
    for (int o = 0; o < D1; ++o)
        for (int i = 0; i < D0; ++i) M.w1[i * D1 + o] = w12_host[o * D0 + i];
    for (int o = 0; o < D1; ++o) M.b1[o] = w12_host[D0 * D1 + o];
    for (int o = 0; o < D2; ++o)
        for (int i = 0; i < D1; ++i) M.w2[i * D2 + o] = w12_host[D0 * D1 + D1 + o * D1 + i];
    for (int o = 0; o < D2; ++o) M.b2[o] = w12_host[D0 * D1 + D1 + D1 * D2 + o];
    cudaError_t e = cudaMemsetAsync(packed, 0, sizeof(unsigned long long) * (size_t)V * D3, st);
    if (e != cudaSuccess) return (int)e;
    if (R > 0) {
        MlpWeights w{nullptr, nullptr, nullptr, nullptr, w3, b3};
        int blocks = (R + kThreads - 1) / kThreads;
        if (blocks > 148 * 2) blocks = 148 * 2;   // persistent: two blocks per SM (256 tensor-memory columns each)
        static const int nt = []() { const char* e = getenv("LTN_PN_THREADS"); return e && atoi(e) == 256 ? 256 : kPointnetTcThreads; }();
        blocks = (R + nt - 1) / nt;
        const int per_sm = nt == 128 ? 4 : 2;
        if (blocks > 148 * per_sm) blocks = 148 * per_sm;   // persistent
        const size_t dyn = sizeof(uint32_t) * 2 * D3 * (nt / 2 + 1);
        const void* fn = nt == 128 ? (const void*)k_pointnet_tc<128> : (const void*)k_pointnet_tc<256>;
        e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        if (e != cudaSuccess) return (int)e;
        if (nt == 128) k_pointnet_tc<128><<<blocks, 128, dyn, st>>>(rows, width, idx, R, r_dev, w, M, V, v_dev, packed, ldexpf(1.0f, a_log2), flag, g_pn_trace);
        else k_pointnet_tc<256><<<blocks, 256, dyn, st>>>(rows, width, idx, R, r_dev, w, M, V, v_dev, packed, ldexpf(1.0f, a_log2), flag, g_pn_trace);
        LTN_CHECK_LAUNCH();
    }
    long long total = (long long)V * D3;
    int blocks = (int)((total + kThreads - 1) / kThreads);
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_pointnet_decode<<<blocks, kThreads, 0, st>>>(packed, V, v_dev, R, r_dev, rows, width, vert_acc, min_rows, out);
    LTN_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
