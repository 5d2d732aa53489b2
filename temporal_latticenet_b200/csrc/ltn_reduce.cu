// Segmented reductions and normalisation over lattice vertices: scatter_max (+argmax), scatter_add,
// GroupNorm statistics / apply (+ReLU).  sm_100a.
#include "ltn_common.cuh"

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ uint32_t ord_enc(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord_dec(uint32_t e) {
    return __uint_as_float((e & 0x80000000u) ? (e & 0x7FFFFFFFu) : ~e);
}

// packed[v,c] = max over rows of (ordered(value) << 32 | ~row): the largest value wins and, between
// equal values, the SMALLEST source row (deterministic; torch_scatter's CUDA path is a race there).
__global__ void __launch_bounds__(kThreads)
k_scatter_max(const float* __restrict__ src, const int* __restrict__ idx, int R, int C,
              unsigned long long* packed, int V) {
    long long total = (long long)R * C;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        int row = (int)(t / C);
        int c = (int)(t - (long long)row * C);
        int id = __ldg(idx + row);
        id = id < 0 ? 0 : id;  // lattice_modules.py:479-480
        if (id >= V) continue;
        unsigned long long key = ((unsigned long long)ord_enc(__ldg(src + t)) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)row);
        atomicMax(packed + (size_t)id * C + c, key);
    }
}

// empty segment: value 0, argmax = R (torch_scatter 2.0.4 sentinel, quirk Q3)
__global__ void __launch_bounds__(kThreads)
k_scatter_max_decode(const unsigned long long* __restrict__ packed, long long n, int R, float* __restrict__ out,
                     long long* __restrict__ arg) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        unsigned long long k = packed[t];
        if (k == 0ull) { out[t] = 0.f; arg[t] = R; }
        else { out[t] = ord_dec((uint32_t)(k >> 32)); arg[t] = (long long)(0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFull)); }
    }
}

__global__ void __launch_bounds__(kThreads)
k_scatter_add(const float* __restrict__ src, const int* __restrict__ idx, int R, int C, float* out, int V) {
    long long total = (long long)R * C;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        int row = (int)(t / C);
        int c = (int)(t - (long long)row * C);
        int id = __ldg(idx + row);
        id = id < 0 ? 0 : id;
        if (id < V) atomicAdd(out + (size_t)id * C + c, __ldg(src + t));
    }
}

// GroupNorm statistics over x [V,C]: per group sum and sum of squares in double (var = E[x^2]-mean^2
// must not cancel in fp32).  One thread per float4 column and row phase: a warp reads whole rows with
// 128-bit loads; per-thread double partials -> shared per-channel sums -> one global atomic per group
// and block.  Grid = at most one block per SM.
__global__ void __launch_bounds__(kThreads)
k_gn_stats(const float* __restrict__ x, int V, const int* __restrict__ v_dev, int C, int cpg, double* sums /*[G,2]*/) {
    extern __shared__ double sh[];  // [C][2]
    if (v_dev) V = min(V, *v_dev);
    const int C4 = C >> 2;
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sh[i] = 0.0;
    __syncthreads();
    const int rpb = blockDim.x / C4;                      // rows in flight per block
    const int col = threadIdx.x % C4, rph = threadIdx.x / C4;
    const int rows_per_block = (V + gridDim.x - 1) / gridDim.x;
    const int r0 = blockIdx.x * rows_per_block, r1 = min(V, r0 + rows_per_block);
    if (rph < rpb) {
        double s1[4] = {0.0, 0.0, 0.0, 0.0}, s2[4] = {0.0, 0.0, 0.0, 0.0};
        // fp32 partials over short runs, folded into double every 8 rows
        for (int r = r0 + rph; r < r1; r += rpb * 8) {
            float f1[4] = {0.f, 0.f, 0.f, 0.f}, f2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                int rr = r + u * rpb;
                if (rr < r1) {
                    float4 a = __ldg(reinterpret_cast<const float4*>(x + (size_t)rr * C) + col);
                    f1[0] += a.x; f1[1] += a.y; f1[2] += a.z; f1[3] += a.w;
                    f2[0] = fmaf(a.x, a.x, f2[0]); f2[1] = fmaf(a.y, a.y, f2[1]);
                    f2[2] = fmaf(a.z, a.z, f2[2]); f2[3] = fmaf(a.w, a.w, f2[3]);
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) { s1[q] += (double)f1[q]; s2[q] += (double)f2[q]; }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            atomicAdd(&sh[2 * (col * 4 + q)], s1[q]);
            atomicAdd(&sh[2 * (col * 4 + q) + 1], s2[q]);
        }
    }
    __syncthreads();
    const int G = C / cpg;
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        double a = 0.0, b = 0.0;
        for (int c = g * cpg; c < (g + 1) * cpg; ++c) { a += sh[2 * c]; b += sh[2 * c + 1]; }
        atomicAdd(sums + 2 * g, a);
        atomicAdd(sums + 2 * g + 1, b);
    }
}

// y = relu?( (x - mean_g) * rstd_g * gamma_c + beta_c ), folded into per-channel a_c, b_c in smem
__global__ void __launch_bounds__(kThreads)
k_gn_apply(const float* __restrict__ x, int V, const int* __restrict__ v_dev, int C, int cpg,
           const double* __restrict__ sums, const float* __restrict__ gamma, const float* __restrict__ beta,
           float eps, int relu, float* __restrict__ y) {
    extern __shared__ float ab[];  // a[C], b[C]
    if (v_dev) V = min(V, *v_dev);
    double n = (double)V * cpg;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        int g = c / cpg;
        double mean = sums[2 * g] / n;
        double var = sums[2 * g + 1] / n - mean * mean;
        if (var < 0.0) var = 0.0;
        float rstd = (float)(1.0 / sqrt(var + (double)eps));
        float a = rstd * (gamma ? __ldg(gamma + c) : 1.0f);
        ab[c] = a;
        ab[C + c] = (beta ? __ldg(beta + c) : 0.0f) - (float)mean * a;
    }
    __syncthreads();
    long long total = (long long)V * C;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        int c = (int)(t % C);
        float r = fmaf(__ldg(x + t), ab[c], ab[C + c]);
        y[t] = (relu && r < 0.f) ? 0.f : r;
    }
}

// ---- GroupNorm(+ReLU) backward (torch.nn.GroupNorm over [1,C,V] as the lattice modules use it) -------------------------
// With xh = (x - mean_g) * rstd_g and gy' = gy * [y > 0] (ReLU fused in the forward):
//   per channel:  A_c = sum_v gy'[v,c]  (= grad beta),   B_c = sum_v gy'[v,c] * xh[v,c]  (= grad gamma)
//   per group:    m1_g = sum_{c in g} gamma_c A_c / n,   m2_g = sum_{c in g} gamma_c B_c / n,   n = V * C/G
//   grad x[v,c] = (gamma_c gy'[v,c] - m1_g - xh[v,c] m2_g) * rstd_g
// Pass 1 (k_gn_bwd_stats) reduces A, B in double (same shape as k_gn_stats); pass 2 (k_gn_bwd_apply) is one
// read-modify-write pass.  Replaces ~10 eager torch ops with [V,C] temporaries.
__device__ __forceinline__ void gn_group_stats(const double* __restrict__ sums, int g, double n, float eps, float& mean, float& rstd) {
    const double m = sums[2 * g] / n;
    double var = sums[2 * g + 1] / n - m * m;
    if (var < 0.0) var = 0.0;
    mean = (float)m;
    rstd = (float)(1.0 / sqrt(var + (double)eps));
}

__global__ void __launch_bounds__(kThreads)
k_gn_bwd_stats(const float* __restrict__ x, const float* __restrict__ gy, const float* __restrict__ y, int V, int C, int cpg,
               const double* __restrict__ sums, float eps, double* chan /*[C,2]*/) {
    extern __shared__ double sh[];  // [C][2]
    const int C4 = C >> 2;
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sh[i] = 0.0;
    __syncthreads();
    const int rpb = blockDim.x / C4;
    const int col = threadIdx.x % C4, rph = threadIdx.x / C4;
    const int rows_per_block = (V + gridDim.x - 1) / gridDim.x;
    const int r0 = blockIdx.x * rows_per_block, r1 = min(V, r0 + rows_per_block);
    if (rph < rpb) {
        float mean[4], rstd[4];
        const double n = (double)V * cpg;
#pragma unroll
        for (int q = 0; q < 4; ++q) gn_group_stats(sums, (col * 4 + q) / cpg, n, eps, mean[q], rstd[q]);
        double s1[4] = {0.0, 0.0, 0.0, 0.0}, s2[4] = {0.0, 0.0, 0.0, 0.0};
        for (int r = r0 + rph; r < r1; r += rpb * 8) {
            float f1[4] = {0.f, 0.f, 0.f, 0.f}, f2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int rr = r + u * rpb;
                if (rr < r1) {
                    const float4 a = __ldg(reinterpret_cast<const float4*>(x + (size_t)rr * C) + col);
                    float4 g = __ldg(reinterpret_cast<const float4*>(gy + (size_t)rr * C) + col);
                    if (y) {
                        const float4 o = __ldg(reinterpret_cast<const float4*>(y + (size_t)rr * C) + col);
                        if (!(o.x > 0.f)) g.x = 0.f;
                        if (!(o.y > 0.f)) g.y = 0.f;
                        if (!(o.z > 0.f)) g.z = 0.f;
                        if (!(o.w > 0.f)) g.w = 0.f;
                    }
                    f1[0] += g.x; f1[1] += g.y; f1[2] += g.z; f1[3] += g.w;
                    f2[0] = fmaf(g.x, (a.x - mean[0]) * rstd[0], f2[0]); f2[1] = fmaf(g.y, (a.y - mean[1]) * rstd[1], f2[1]);
                    f2[2] = fmaf(g.z, (a.z - mean[2]) * rstd[2], f2[2]); f2[3] = fmaf(g.w, (a.w - mean[3]) * rstd[3], f2[3]);
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) { s1[q] += (double)f1[q]; s2[q] += (double)f2[q]; }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            atomicAdd(&sh[2 * (col * 4 + q)], s1[q]);
            atomicAdd(&sh[2 * (col * 4 + q) + 1], s2[q]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(chan + i, sh[i]);
}

__global__ void __launch_bounds__(kThreads)
k_gn_bwd_apply(const float* __restrict__ x, const float* __restrict__ gy, const float* __restrict__ y, int V, int C, int cpg,
               const double* __restrict__ sums, const double* __restrict__ chan, const float* __restrict__ gamma, float eps,
               float* __restrict__ gx) {
    extern __shared__ float k[];  // per channel: mean, rstd, gamma, m1, m2
    const double n = (double)V * cpg;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int g = c / cpg;
        float mean, rstd;
        gn_group_stats(sums, g, n, eps, mean, rstd);
        double m1 = 0.0, m2 = 0.0;
        for (int j = g * cpg; j < (g + 1) * cpg; ++j) {
            const double gj = gamma ? (double)__ldg(gamma + j) : 1.0;
            m1 += gj * chan[2 * j];
            m2 += gj * chan[2 * j + 1];
        }
        k[c] = mean; k[C + c] = rstd; k[2 * C + c] = gamma ? __ldg(gamma + c) : 1.0f;
        k[3 * C + c] = (float)(m1 / n); k[4 * C + c] = (float)(m2 / n);
    }
    __syncthreads();
    const long long total = (long long)V * C;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(t % C);
        float g = __ldg(gy + t);
        if (y && !(__ldg(y + t) > 0.f)) g = 0.f;
        const float xh = (__ldg(x + t) - k[c]) * k[C + c];
        gx[t] = (k[2 * C + c] * g - k[3 * C + c] - xh * k[4 * C + c]) * k[C + c];
    }
}

inline int grid_for(long long work_items, int threads) {
    long long b = (work_items + threads - 1) / threads;
    const long long cap = 148LL * 32;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// IoU accumulation of callbacks/scores.py:13-30 on the device.  One thread per point: arg-max of its K class scores
// (first maximum wins, like torch.argmax), then the (gt, pred) pair is counted in a block-local K x K histogram that
// is flushed with one 64-bit atomic per non-empty cell.
__global__ void __launch_bounds__(kThreads)
k_confusion(const float* __restrict__ scores, const long long* __restrict__ gt, int N, const int* __restrict__ n_dev, int K,
            unsigned long long* conf) {
    extern __shared__ unsigned int s_hist[];
    if (n_dev) N = min(N, __ldg(n_dev));
    for (int i = threadIdx.x; i < K * K; i += blockDim.x) s_hist[i] = 0u;
    __syncthreads();
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < N; p += gridDim.x * blockDim.x) {
        const float* row = scores + (size_t)p * K;
        int best = 0;
        float bv = __ldg(row);
        for (int k = 1; k < K; ++k) {
            const float v = __ldg(row + k);
            if (v > bv || (v != v && bv == bv)) { bv = v; best = k; }   // NaN counts as the maximum (torch.argmax)
        }
        const long long g = __ldg(gt + p);
        if (g >= 0 && g < K) atomicAdd(&s_hist[(int)g * K + best], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K * K; i += blockDim.x)
        if (s_hist[i]) atomicAdd(conf + i, (unsigned long long)s_hist[i]);
}

// folds one cloud's confusion matrix into the running per-class intersection / union (scores.py:24-30: only classes
// PRESENT in this cloud's ground truth, the unlabeled class skipped) and clears the matrix for the next cloud
__global__ void k_scores_fold(unsigned long long* conf, int K, int unlabeled, long long* inter, long long* uni) {
    const int l = threadIdx.x;
    unsigned long long g = 0, pr = 0, d = 0;
    if (l < K) {
        for (int j = 0; j < K; ++j) { g += conf[l * K + j]; pr += conf[j * K + l]; }
        d = conf[l * K + l];
    }
    __syncthreads();
    if (l < K && g > 0 && l != unlabeled) {
        inter[l] += (long long)d;
        uni[l] += (long long)(g + pr - d);
    }
    for (int i = threadIdx.x; i < K * K; i += blockDim.x) conf[i] = 0ull;
}

}  // namespace

extern "C" {

// packed: [V,C] u64 scratch (zeroed here); out [V,C] f32; arg [V,C] i64
int ltn_scatter_max(const float* src, const int* idx, int R, int C, int V, unsigned long long* packed, float* out,
                    long long* arg, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (V <= 0 || C <= 0) return 0;
    cudaError_t e = cudaMemsetAsync(packed, 0, sizeof(unsigned long long) * (size_t)V * C, st);
    if (e != cudaSuccess) return (int)e;
    if (R > 0) {
        k_scatter_max<<<grid_for((long long)R * C, kThreads), kThreads, 0, st>>>(src, idx, R, C, packed, V);
        LTN_CHECK_LAUNCH();
    }
    k_scatter_max_decode<<<grid_for((long long)V * C, kThreads), kThreads, 0, st>>>(packed, (long long)V * C, R, out, arg);
    LTN_CHECK_LAUNCH();
    return 0;
}

// out [V,C] must be zeroed (or hold the accumulator) on entry
int ltn_scatter_add(const float* src, const int* idx, int R, int C, float* out, int V, void* stream) {
    if (R <= 0 || C <= 0) return 0;
    k_scatter_add<<<grid_for((long long)R * C, kThreads), kThreads, 0, (cudaStream_t)stream>>>(src, idx, R, C, out, V);
    LTN_CHECK_LAUNCH();
    return 0;
}

// sums: [G,2] double, zeroed here.  C % 4 == 0, C/4 <= 256
int ltn_gn_stats(const float* x, int V, const int* v_dev, int C, int G, double* sums, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (G <= 0 || C % G || C % 4 || C / 4 > kThreads) return -2;
    cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(double) * 2 * (size_t)G, st);
    if (e != cudaSuccess) return (int)e;
    if (V <= 0) return 0;
    int rpb = kThreads / (C / 4);
    int blocks = (V + rpb * 8 - 1) / (rpb * 8);
    if (blocks > 148) blocks = 148;
    if (blocks < 1) blocks = 1;
    k_gn_stats<<<blocks, kThreads, sizeof(double) * 2 * C, st>>>(x, V, v_dev, C, C / G, sums);
    LTN_CHECK_LAUNCH();
    return 0;
}

int ltn_gn_apply(const float* x, int V, const int* v_dev, int C, int G, const double* sums, const float* gamma,
                 const float* beta, float eps, int relu, float* y, void* stream) {
    if (V <= 0) return 0;
    if (G <= 0 || C % G) return -2;
    k_gn_apply<<<grid_for((long long)V * C, kThreads), kThreads, sizeof(float) * 2 * C, (cudaStream_t)stream>>>(
        x, V, v_dev, C, C / G, sums, gamma, beta, eps, relu, y);
    LTN_CHECK_LAUNCH();
    return 0;
}

// scores [N,K] (any monotone transform of the class probabilities), gt [N] int64; conf [K,K] u64 scratch (zero on first
// use, left zero), inter / uni [K] int64 accumulators.  K <= 64.
int ltn_scores_accumulate(const float* scores, const long long* gt, int N, const int* n_dev, int K, int unlabeled,
                          unsigned long long* conf, long long* inter, long long* uni, void* stream) {
    if (K <= 0 || K > 64) return -2;
    cudaStream_t st = (cudaStream_t)stream;
    if (N > 0) {
        int blocks = (N + kThreads - 1) / kThreads;
        if (blocks > 4 * 148) blocks = 4 * 148;
        k_confusion<<<blocks, kThreads, sizeof(unsigned int) * K * K, st>>>(scores, gt, N, n_dev, K, conf);
        LTN_CHECK_LAUNCH();
    }
    k_scores_fold<<<1, 64, 0, st>>>(conf, K, unlabeled, inter, uni);
    LTN_CHECK_LAUNCH();
    return 0;
}

// GroupNorm(+ReLU) backward.  x, gy [V,C]; y (nullable) = the forward's output when ReLU was fused (its sign is the mask);
// sums [G,2] = the forward's statistics; chan [C,2] double scratch (zeroed here) returns (grad beta, grad gamma) per
// channel; gx [V,C].  C % 4 == 0, C/4 <= 256.
int ltn_gn_bwd(const float* x, const float* gy, const float* y, int V, int C, int G, const double* sums, const float* gamma,
               float eps, double* chan, float* gx, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (G <= 0 || C % G || C % 4 || C / 4 > kThreads) return -2;
    cudaError_t e = cudaMemsetAsync(chan, 0, sizeof(double) * 2 * (size_t)C, st);
    if (e != cudaSuccess) return (int)e;
    if (V <= 0) return 0;
    int rpb = kThreads / (C / 4);
    int blocks = (V + rpb * 8 - 1) / (rpb * 8);
    if (blocks > 148) blocks = 148;
    if (blocks < 1) blocks = 1;
    k_gn_bwd_stats<<<blocks, kThreads, sizeof(double) * 2 * C, st>>>(x, gy, y, V, C, C / G, sums, eps, chan);
    LTN_CHECK_LAUNCH();
    k_gn_bwd_apply<<<grid_for((long long)V * C, kThreads), kThreads, sizeof(float) * 5 * C, st>>>(x, gy, y, V, C, C / G, sums, chan, gamma,
                                                                                                 eps, gx);
    LTN_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
