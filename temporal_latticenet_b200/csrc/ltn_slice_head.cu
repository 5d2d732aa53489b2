// The per-POINT tail of the network in two kernels (inference): SliceFastCUDALatticeModule after its 1x1 chain
// (/root/reference/seq_lattice/models.py:232,465; recipe SURVEY.md E12):
//
//   g[p, r*9 + c] = w[p,r] * bott[idx[p,r], c]  (c < 8),  g[p, r*9 + 8] = w[p,r]          gather      [N, 4*9]
//   g            -= gamma[c] * max_r g[p, r*9 + c] + beta[c]                               max over the simplex
//   h             = W1 g                                                                    Linear 36 -> 36, no bias
//   y             = relu(GroupNorm_18(h))        statistics over ALL points                 GroupNorm + ReLU
//   dw            = W2 y + b2                                                               Linear 36 -> 4
//   logit[p, k]   = b[k] + sum_r (w[p,r] + dw[p,r]) * scores[idx[p,r], k]                   slice of the vertices' class scores
//   logsm[p, :]   = log_softmax(logit[p, :])
//
// Op by op this was 14 launches per frame and window (gather, max, three elementwise, two SIMT GEMMs, GroupNorm statistics
// and apply, add, slice, bias, log-softmax) moving [N, 36] tensors back and forth: 535 us of a lock-step group's last frame
// with nothing for the tensor cores to do (profiles/r2_timeline_group.txt).  The GroupNorm in the middle needs statistics
// over all points, so there are two passes: k_head_stats computes h and only its group sums, k_head_apply RECOMPUTES h
// (36 x 36 multiply-adds per point from 4 x 32 bytes of input: cheaper than writing and re-reading [N, 36]) and finishes.
#include "ltn_common.cuh"
#include <cstdlib>

namespace {

constexpr int kHeadThreads = 128;
constexpr int kG = 36;        // 4 simplex vertices x (8 bottleneck values + the barycentric weight)
constexpr int kGroups = 18;   // gn_groups(36) = 36 / 2

struct HeadIn {
    const float* bott;   // [V, 8]
    const int* idx;      // [4N]
    const float* w;      // [4N]
    const float* gamma;  // [9]
    const float* beta;   // [9]
};

// h = W1 (g - (gamma * max_r g + beta)) of one point; sW1 [36][36] in shared memory (uniform reads: broadcasts).  Every h[j] is
// handed to `fold(j, h_j)` as soon as it exists: neither kernel keeps the 36 values (the statistics kernel adds them to its
// group sums, the apply kernel normalises them and adds them to the four delta weights), which is the difference between
// 168 and ~100 registers per thread, i.e. between 3 and 5 blocks per SM for a kernel that waits on shared-memory broadcasts.
template <class Fold>
__device__ __forceinline__ void head_h(const HeadIn& in, int V, long long p, const float* sW1, const float* sGB, int4& id4, float4& w4,
                                       Fold&& fold) {
    id4 = __ldg(reinterpret_cast<const int4*>(in.idx) + p);
    w4 = __ldg(reinterpret_cast<const float4*>(in.w) + p);
    const int ids[4] = {id4.x, id4.y, id4.z, id4.w};
    const float ws[4] = {w4.x, w4.y, w4.z, w4.w};
    float g[kG];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const bool ok = ids[r] >= 0 && ids[r] < V;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (ok) {
            const float4* row = reinterpret_cast<const float4*>(in.bott + (size_t)ids[r] * 8);
            a = __ldg(row);
            b = __ldg(row + 1);
        }
        const float wv = ok ? ws[r] : 0.f;
        g[r * 9 + 0] = __fmul_rn(wv, a.x); g[r * 9 + 1] = __fmul_rn(wv, a.y); g[r * 9 + 2] = __fmul_rn(wv, a.z); g[r * 9 + 3] = __fmul_rn(wv, a.w);
        g[r * 9 + 4] = __fmul_rn(wv, b.x); g[r * 9 + 5] = __fmul_rn(wv, b.y); g[r * 9 + 6] = __fmul_rn(wv, b.z); g[r * 9 + 7] = __fmul_rn(wv, b.w);
        g[r * 9 + 8] = wv;
    }
#pragma unroll
    for (int c = 0; c < 9; ++c) {
        const float mx = fmaxf(fmaxf(g[c], g[9 + c]), fmaxf(g[18 + c], g[27 + c]));
        const float sub = __fadd_rn(__fmul_rn(sGB[c], mx), sGB[9 + c]);   // gamma * max + beta, rounded as the reference's three ops
#pragma unroll
        for (int r = 0; r < 4; ++r) g[r * 9 + c] = __fsub_rn(g[r * 9 + c], sub);
    }
#pragma unroll
    for (int j = 0; j < kG; ++j) {
        const float4* wr = reinterpret_cast<const float4*>(sW1 + j * kG);
        float acc = 0.f;
#pragma unroll
        for (int q = 0; q < kG / 4; ++q) {
            const float4 ww = wr[q];
            acc = fmaf(ww.x, g[4 * q], acc); acc = fmaf(ww.y, g[4 * q + 1], acc);
            acc = fmaf(ww.z, g[4 * q + 2], acc); acc = fmaf(ww.w, g[4 * q + 3], acc);
        }
        fold(j, acc);
    }
}

__global__ void __launch_bounds__(kHeadThreads)
k_head_stats(HeadIn in, int V, const int* __restrict__ v_dev, int N, const int* __restrict__ n_dev, const float* __restrict__ W1,
             double* __restrict__ sums) {
    __shared__ __align__(16) float sW1[kG * kG];
    __shared__ float sGB[18];
    __shared__ float sRed[kHeadThreads / 32][2 * kGroups];
    if (v_dev) V = min(V, __ldg(v_dev));
    if (n_dev) N = min(N, __ldg(n_dev));
    for (int i = threadIdx.x; i < kG * kG; i += kHeadThreads) sW1[i] = __ldg(W1 + i);
    if (threadIdx.x < 9) { sGB[threadIdx.x] = __ldg(in.gamma + threadIdx.x); sGB[9 + threadIdx.x] = __ldg(in.beta + threadIdx.x); }
    __syncthreads();
    float s[kGroups], q[kGroups];
#pragma unroll
    for (int g = 0; g < kGroups; ++g) { s[g] = 0.f; q[g] = 0.f; }
    for (long long p = (long long)blockIdx.x * kHeadThreads + threadIdx.x; p < N; p += (long long)gridDim.x * kHeadThreads) {
        int4 id4; float4 w4;
        head_h(in, V, p, sW1, sGB, id4, w4, [&](const int j, const float hj) {
            s[j >> 1] += hj;                          // j is a compile-time constant after unrolling: plain register adds
            q[j >> 1] = fmaf(hj, hj, q[j >> 1]);
        });
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int g = 0; g < kGroups; ++g) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s[g] += __shfl_xor_sync(0xffffffffu, s[g], o);
            q[g] += __shfl_xor_sync(0xffffffffu, q[g], o);
        }
        if (lane == 0) { sRed[warp][2 * g] = s[g]; sRed[warp][2 * g + 1] = q[g]; }
    }
    __syncthreads();
    if (threadIdx.x < 2 * kGroups) {
        double t = 0.0;
#pragma unroll
        for (int wv = 0; wv < kHeadThreads / 32; ++wv) t += (double)sRed[wv][threadIdx.x];
        atomicAdd(sums + threadIdx.x, t);
    }
}

__global__ void __launch_bounds__(kHeadThreads)
k_head_apply(HeadIn in, int V, const int* __restrict__ v_dev, int N, const int* __restrict__ n_dev, const float* __restrict__ W1,
             const double* __restrict__ sums, const float* __restrict__ gn_w, const float* __restrict__ gn_b, float gn_eps,
             const float* __restrict__ W2, const float* __restrict__ b2, int no_deform, const float* __restrict__ scores, int ld_scores,
             const float* __restrict__ cls_bias, int K, float* __restrict__ logits, float* __restrict__ logsm) {
    __shared__ __align__(16) float sW1[kG * kG];
    __shared__ __align__(16) float sW2[4 * kG];
    __shared__ float sGB[18], sAff[2 * kG], sB2[4], sCb[32];
    __shared__ float sOut[2][kHeadThreads * 32];   // logits | log-softmax of the block's points, written out coalesced
    if (v_dev) V = min(V, __ldg(v_dev));
    if (n_dev) N = min(N, __ldg(n_dev));
    for (int i = threadIdx.x; i < kG * kG; i += kHeadThreads) sW1[i] = __ldg(W1 + i);
    for (int i = threadIdx.x; i < 4 * kG; i += kHeadThreads) sW2[i] = __ldg(W2 + i);
    if (threadIdx.x < 9) { sGB[threadIdx.x] = __ldg(in.gamma + threadIdx.x); sGB[9 + threadIdx.x] = __ldg(in.beta + threadIdx.x); }
    if (threadIdx.x < 4) sB2[threadIdx.x] = b2 ? __ldg(b2 + threadIdx.x) : 0.f;
    if (threadIdx.x < 32) sCb[threadIdx.x] = (threadIdx.x < K && cls_bias) ? __ldg(cls_bias + threadIdx.x) : 0.f;
    if (threadIdx.x < kG) {
        const int g = threadIdx.x >> 1;
        const double n = 2.0 * (double)N;
        const double mean = sums[2 * g] / n;
        double var = sums[2 * g + 1] / n - mean * mean;
        if (var < 0.0) var = 0.0;
        const float rstd = (float)(1.0 / sqrt(var + (double)gn_eps));
        const float sc = rstd * (gn_w ? __ldg(gn_w + threadIdx.x) : 1.f);
        sAff[threadIdx.x] = sc;
        sAff[kG + threadIdx.x] = (gn_b ? __ldg(gn_b + threadIdx.x) : 0.f) - (float)mean * sc;
    }
    __syncthreads();
    const long long nblk = ((long long)N + kHeadThreads - 1) / kHeadThreads;
    for (long long blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        const long long p0 = blk * kHeadThreads, p = p0 + threadIdx.x;
        if (p < N) {
            int4 id4; float4 w4;
            float dw[4] = {sB2[0], sB2[1], sB2[2], sB2[3]};
            head_h(in, V, p, sW1, sGB, id4, w4, [&](const int j, const float hj) {
                float y = fmaf(hj, sAff[j], sAff[kG + j]);
                y = y > 0.f ? y : (y != y ? y : 0.f);   // ReLU that keeps NaN
#pragma unroll
                for (int r = 0; r < 4; ++r) dw[r] = fmaf(sW2[r * kG + j], y, dw[r]);
            });
            const int ids[4] = {id4.x, id4.y, id4.z, id4.w};
            const float ws[4] = {w4.x, w4.y, w4.z, w4.w};
            float acc[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) acc[k] = 0.f;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if (ids[r] < 0 || ids[r] >= V) continue;
                const float ww = no_deform ? ws[r] : __fadd_rn(ws[r], dw[r]);
                const float4* row = reinterpret_cast<const float4*>(scores + (size_t)ids[r] * ld_scores);
#pragma unroll
                for (int k4 = 0; k4 < 8; ++k4) {
                    if (4 * k4 < K) {
                        const float4 v = __ldg(row + k4);
                        acc[4 * k4] = fmaf(ww, v.x, acc[4 * k4]); acc[4 * k4 + 1] = fmaf(ww, v.y, acc[4 * k4 + 1]);
                        acc[4 * k4 + 2] = fmaf(ww, v.z, acc[4 * k4 + 2]); acc[4 * k4 + 3] = fmaf(ww, v.w, acc[4 * k4 + 3]);
                    }
                }
            }
            float mx = -INFINITY;
#pragma unroll
            for (int k = 0; k < 32; ++k)
                if (k < K) { acc[k] += sCb[k]; mx = fmaxf(mx, acc[k]); }
            float se = 0.f;
#pragma unroll
            for (int k = 0; k < 32; ++k)
                if (k < K) se += __expf(acc[k] - mx);
            const float lse = mx + __logf(se);
#pragma unroll
            for (int k = 0; k < 32; ++k)
                if (k < K) {
                    sOut[0][threadIdx.x * K + k] = acc[k];
                    sOut[1][threadIdx.x * K + k] = acc[k] - lse;
                }
        }
        __syncthreads();
        const long long live = min((long long)kHeadThreads, (long long)N - p0) * K;
        float* o0 = logits ? logits + p0 * K : nullptr;
        float* o1 = logsm + p0 * K;
        for (long long i = threadIdx.x; i < live; i += kHeadThreads) {
            if (o0) o0[i] = sOut[0][i];
            o1[i] = sOut[1][i];
        }
        __syncthreads();
    }
}

}  // namespace

extern "C" {

// The per-point tail of SliceFastCUDALatticeModule + the model's log-softmax in two kernels (see the top of this file).
// bott [V, 8]: output of the 1x1 chain; scores [V, ld_scores]: the vertices' class scores (lv @ Wc^T, K <= 32 live columns);
// idx / w [4N]: the splatting simplex of every point; sums [18, 2] double: scratch for the GroupNorm statistics (zeroed here);
// logits (nullable) / logsm [N, K].  v_dev / n_dev: device-side live counts (static-capacity graphs), nullable.
int ltn_slice_head(const float* bott, int V, const int* v_dev, const float* scores, int ld_scores, const int* idx, const float* w,
                   int N, const int* n_dev, const float* gamma, const float* beta, const float* W1, const float* gn_w,
                   const float* gn_b, float gn_eps, const float* W2, const float* b2, const float* cls_bias, int K, int no_deform,
                   double* sums, float* logits, float* logsm, void* stream) {
    if (N <= 0) return 0;
    if (K < 1 || K > 32 || ld_scores % 4 || ld_scores < ((K + 3) & ~3) || !sums || !logsm) return -2;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(double) * 2 * kGroups, st);
    if (e != cudaSuccess) return (int)e;
    HeadIn in{bott, idx, w, gamma, beta};
    long long blocks = ((long long)N + kHeadThreads - 1) / kHeadThreads;
    if (blocks > 148 * 8) blocks = 148 * 8;
    // the statistics kernel ends every block with 36 double atomics on the same 36 addresses: fewer, longer blocks
    // (grid-stride loop) trade latency hiding for fewer serialised reductions; LTN_HEAD_STATS_BPS = blocks per SM.  Measured on one
    // scan (tools/bench_slice_head.py, both kernels): 1 -> 57.6, 2 -> 55.6, 3 -> 55.5, 4 -> 57.7, 8 -> 60.5 us per call
    static const int stats_bps = []() { const char* e = getenv("LTN_HEAD_STATS_BPS"); return e && atoi(e) >= 1 ? atoi(e) : 3; }();
    long long sblocks = blocks < 148LL * stats_bps ? blocks : 148LL * stats_bps;
    k_head_stats<<<(int)sblocks, kHeadThreads, 0, st>>>(in, V, v_dev, N, n_dev, W1, sums);
    LTN_CHECK_LAUNCH();
    k_head_apply<<<(int)blocks, kHeadThreads, 0, st>>>(in, V, v_dev, N, n_dev, W1, sums, gn_w, gn_b, gn_eps, W2, b2, no_deform, scores,
                                                       ld_scores, cls_bias, K, logits, logsm);
    LTN_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
