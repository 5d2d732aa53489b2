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
#pragma unroll
                for (int o = 0; o < 16; ++o)
                    if (t_val[o * kPitch + lslot] == enc[o]) atomicMin(&t_row[o * kPitch + lslot], (uint32_t)row);   // smallest row wins ties
            }
            __syncthreads();
            // one 64-bit atomic per (distinct vertex, channel): value in the high word, ~row in the low word
            for (int e = threadIdx.x; e < ns * 16; e += kThreads) {
                const int sl = e >> 4, c = e & 15;
                const unsigned long long key = ((unsigned long long)t_val[c * kPitch + sl] << 32) |
                                               (unsigned long long)(0xFFFFFFFFu - t_row[c * kPitch + sl]);
                unsigned long long* dst = packed + (size_t)slot_id[sl] * D3 + q * 16 + c;
                if (__ldcg(dst) < key) atomicMax(dst, key);
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

}  // extern "C"
