// Bandwidth-bound gather / scatter kernels over the lattice: im2row (+ its transpose), splat, slice,
// gather, slice_classify and their backward passes.  sm_100a; 128-bit vectorised row accesses.
//
// Neighbour tables ([V,9] int32, -1 = absent) are built once per lattice state by ltn_neighbours and
// shared by every convolution on that level, so none of these kernels touches the hash table.
#include "ltn_common.cuh"
#include <cstdlib>

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// out[v, s*C + c] = vals[nbr[v,s], c] ; one thread per float4.  C % 4 == 0.
__global__ void __launch_bounds__(kThreads)
k_im2row(const float* __restrict__ vals, int Vvals, const int* __restrict__ vvals_dev, const int* __restrict__ nbr,
         int Vq, const int* __restrict__ vq_dev, int C4, float4* __restrict__ out) {
    if (vq_dev) Vq = min(Vq, *vq_dev);
    if (vvals_dev) Vvals = min(Vvals, *vvals_dev);
    long long total = (long long)Vq * LTN_FEXT * C4;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        long long vs = t / C4;
        int c4 = (int)(t - vs * C4);
        int id = __ldg(nbr + vs);
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (id >= 0 && id < Vvals) x = ld4(vals + ((size_t)id * C4 + c4) * 4);
        out[t] = x;
    }
}

// transpose of im2row as a GATHER (no atomics): grad_vals[u, c] = sum_s grad_rows[nbrT[u, s^1], s*C + c]
// (+ centre slot 8 -> nbrT[u, 8]).  nbrT is the table of the opposite direction: the same table for a
// same-level convolution (neighbour relation is symmetric with the slot pair swapped), the finefy
// table for coarsen and the coarsen table for finefy.  Vrows = number of rows of grad_rows.
__global__ void __launch_bounds__(kThreads)
k_row2im(const float* __restrict__ grad_rows, int Vrows, const int* __restrict__ nbrT, int Vu, int C4,
         float4* __restrict__ grad_vals) {
    long long total = (long long)Vu * C4;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        int u = (int)(t / C4);
        int c4 = (int)(t - (long long)u * C4);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int s = 0; s < LTN_FEXT; ++s) {
            int st = (s < 2 * LTN_D1) ? (s ^ 1) : s;
            int v = __ldg(nbrT + (size_t)u * LTN_FEXT + st);
            if (v >= 0 && v < Vrows) {
                float4 g = ld4(grad_rows + (((size_t)v * LTN_FEXT + s) * C4 + c4) * 4);
                acc.x += g.x; acc.y += g.y; acc.z += g.z; acc.w += g.w;
            }
        }
        grad_vals[t] = acc;
    }
}

// splat: out[idx, 0:C] += w * val[p] ; out[idx, C] += w.   One thread per point; the four simplex rows of the
// 32 points of a warp are aggregated per vertex inside the warp (scan neighbours share vertices), so one lane
// per distinct vertex issues the atomics instead of every row hammering the same address.
__global__ void __launch_bounds__(kThreads)
k_splat(const float* __restrict__ val, int N, int C, const int* __restrict__ idx, const float* __restrict__ w,
        float* out, int V) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = p < N;
    int ids[4] = {-1, -1, -1, -1};
    float ws[4] = {0.f, 0.f, 0.f, 0.f};
    if (valid) {
        int4 i4 = __ldg(reinterpret_cast<const int4*>(idx) + p);
        float4 w4 = __ldg(reinterpret_cast<const float4*>(w) + p);
        ids[0] = i4.x; ids[1] = i4.y; ids[2] = i4.z; ids[3] = i4.w;
        ws[0] = w4.x; ws[1] = w4.y; ws[2] = w4.z; ws[3] = w4.w;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int id = (valid && ids[r] >= 0 && ids[r] < V) ? ids[r] : -1;
        int cnt;
        {   // homogeneous coordinate
            float s[1] = {ws[r]};
            if (ltn_warp_group_sum<1>(id, s, cnt)) atomicAdd(out + (size_t)id * (C + 1) + C, s[0]);
        }
        for (int c = 0; c < C; c += 4) {   // values, four channels per aggregation round
            float s[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) s[q] = (valid && c + q < C) ? ws[r] * __ldg(val + (size_t)p * C + c + q) : 0.f;
            if (ltn_warp_group_sum<4>(id, s, cnt)) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (c + q < C) atomicAdd(out + (size_t)id * (C + 1) + c + q, s[q]);
            }
        }
    }
}

// splat for few channels (C <= 3, the reflectance-like inputs of the model and of BASELINE config 1), run-length
// form: a thread walks kSplatRun CONSECUTIVE points and keeps, per simplex rank r, the vertex id it is currently
// accumulating into and the partial sums in registers.  Consecutive scan points fall into the same simplex for
// a dozen points on average (3 cm point spacing against a 0.6 m lattice), and then the rank-r vertex of point p and
// of point p+1 is the same vertex -- so the sums are carried in registers and only a CHANGE of vertex issues the
// atomics: ~10x fewer of them, no warp votes, no shuffles.  Unordered input degrades to one atomic per row.
constexpr int kSplatRun = 16;   // measured on the accumulated 4-scan cloud: 2 -> 49 us, 4 -> 29, 8 -> 17, 16 -> 13.4, 32 -> 23 (too few threads)

template <int C>
__global__ void __launch_bounds__(kThreads)
k_splat_runs(const float* __restrict__ val, int N, const int* __restrict__ idx, const float* __restrict__ w, float* out, int V, int run) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int p0 = t * run;
    if (p0 >= N) return;
    const int p1 = min(N, p0 + run);
    int cur[4] = {-1, -1, -1, -1};
    float acc[4][C + 1];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c <= C; ++c) acc[r][c] = 0.f;
    auto flush = [&](int r) {
        if (cur[r] >= 0) {
            float* o = out + (size_t)cur[r] * (C + 1);
#pragma unroll
            for (int c = 0; c <= C; ++c) atomicAdd(o + c, acc[r][c]);
        }
    };
    for (int p = p0; p < p1; ++p) {
        const int4 i4 = __ldg(reinterpret_cast<const int4*>(idx) + p);
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(w) + p);
        float v[C];
#pragma unroll
        for (int c = 0; c < C; ++c) v[c] = __ldg(val + (size_t)p * C + c);
        const int ids[4] = {i4.x, i4.y, i4.z, i4.w};
        const float ws[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int id = ((unsigned)ids[r] < (unsigned)V) ? ids[r] : -1;
            if (id != cur[r]) {
                flush(r);
                cur[r] = id;
#pragma unroll
                for (int c = 0; c <= C; ++c) acc[r][c] = 0.f;
            }
#pragma unroll
            for (int c = 0; c < C; ++c) acc[r][c] = fmaf(ws[r], v[c], acc[r][c]);
            acc[r][C] += ws[r];
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) flush(r);
}

// splat for few channels, warp-scan form: one thread per point (coalesced 128-bit index / weight loads, full occupancy),
// then per simplex rank a SEGMENTED inclusive scan over the warp: consecutive scan points share their rank-r vertex for
// a dozen points on average, so the lanes of a run add up in five shuffle rounds and only the LAST lane of each run
// issues the atomics.  Same number of atomics as the run-length walk (~0.13 per row) at 16x its parallelism.
template <int C>
__global__ void __launch_bounds__(kThreads)
k_splat_seg(const float* __restrict__ val, int N, const int* __restrict__ idx, const float* __restrict__ w, float* out, int V) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool valid = p < N;
    int4 i4 = make_int4(-1, -1, -1, -1);
    float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float v[C];
#pragma unroll
    for (int c = 0; c < C; ++c) v[c] = 0.f;
    if (valid) {
        i4 = __ldg(reinterpret_cast<const int4*>(idx) + p);
        w4 = __ldg(reinterpret_cast<const float4*>(w) + p);
#pragma unroll
        for (int c = 0; c < C; ++c) v[c] = __ldg(val + (size_t)p * C + c);
    }
    const int ids[4] = {i4.x, i4.y, i4.z, i4.w};
    const float ws[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int id = ((unsigned)ids[r] < (unsigned)V) ? ids[r] : -1;
        float a[C + 1];
#pragma unroll
        for (int c = 0; c < C; ++c) a[c] = id >= 0 ? ws[r] * v[c] : 0.f;
        a[C] = id >= 0 ? ws[r] : 0.f;
        const int prev = __shfl_up_sync(0xffffffffu, id, 1);
        const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || prev != id);
        const int start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));   // first lane of my run
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
            for (int c = 0; c <= C; ++c) {
                const float t = __shfl_up_sync(0xffffffffu, a[c], o);
                if (lane - o >= start) a[c] += t;
            }
        }
        const bool tail = lane == 31 || ((heads >> (lane + 1)) & 1u);
        if (tail && id >= 0) {
            float* o = out + (size_t)id * (C + 1);
#pragma unroll
            for (int c = 0; c <= C; ++c) atomicAdd(o + c, a[c]);
        }
    }
}

// slice: out[p, :] = sum_r w[p,r] * vals[idx[p,r], :]   (r in order 0..3).  C % 4 == 0.
// Every block walks ONE CONTIGUOUS run of points: neighbouring scan points share their simplex vertices (~90 rows per
// vertex), so the rows a block gathers stay in its SM's L1 and the L2 -> SM traffic falls from 4 rows per point towards
// the distinct rows.  (A grid-stride walk scatters consecutive points over all SMs: measured 16 % L1 hits, 0.45-0.53 of
// the copy bandwidth.  Staging the distinct rows in shared memory through a block-local hash was also built and
// measured: 57 us against 24 us -- five block barriers per 128 points cost more than the L1 does for free.)
__global__ void __launch_bounds__(kThreads)
k_slice(const float* __restrict__ vals, int V, int C4, const int* __restrict__ idx, const float* __restrict__ w, int N,
        float4* __restrict__ out) {
    const long long total = (long long)N * C4;
    const long long per = ((total + gridDim.x - 1) / gridDim.x + kThreads - 1) / kThreads * kThreads;
    const long long t_end = min(total, per * (blockIdx.x + 1));
    long long t = per * blockIdx.x + threadIdx.x;
    if (t >= t_end) return;
    // software pipeline: the streamed index / weight quads of the NEXT item are in flight (HBM latency) while the
    // rows of the current one are gathered (L1 / L2 latency) -- the two latencies no longer add up per item
    int p = (int)(t / C4);
    int4 i4 = __ldg(reinterpret_cast<const int4*>(idx) + p);
    float4 w4 = __ldg(reinterpret_cast<const float4*>(w) + p);
    while (true) {
        const int c4 = (int)(t - (long long)p * C4);
        const long long tn = t + kThreads;
        const bool more = tn < t_end;
        const int pn = more ? (int)(tn / C4) : p;
        const int4 i4n = __ldg(reinterpret_cast<const int4*>(idx) + pn);
        const float4 w4n = __ldg(reinterpret_cast<const float4*>(w) + pn);
        const int ids[4] = {i4.x, i4.y, i4.z, i4.w};
        const float ws[4] = {w4.x, w4.y, w4.z, w4.w};
        float4 x[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {   // row 0 stands in for an absent vertex: the loads stay unconditional and independent
            const bool ok = ids[r] >= 0 && ids[r] < V;
            x[r] = ld4(vals + ((size_t)(ok ? ids[r] : 0) * C4 + c4) * 4);
        }
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < 4; ++r)
            if (ids[r] >= 0 && ids[r] < V) {
                acc.x = fmaf(ws[r], x[r].x, acc.x); acc.y = fmaf(ws[r], x[r].y, acc.y);
                acc.z = fmaf(ws[r], x[r].z, acc.z); acc.w = fmaf(ws[r], x[r].w, acc.w);
            }
        __stcs(out + t, acc);   // written once, read by a later kernel: stream past L1
        if (!more) break;
        t = tn; p = pn; i4 = i4n; w4 = w4n;
    }
}

// The same for rows whose half (H = C4/2 float4) divides the block: a thread owns the float4 pair (j, j + H) of a
// point, so the index / weight quads and the presence tests are shared by two outputs, there is no division in the
// loop (the thread's column never changes, its point advances by kThreads / H) and all indexing is 32-bit.
__global__ void __launch_bounds__(kThreads)
k_slice_pair(const float* __restrict__ vals, int V, int H, const int* __restrict__ idx, const float* __restrict__ w, int N,
             float4* __restrict__ out) {
    const int ppb = kThreads / H;                              // points per block step
    const int per = ((N + gridDim.x - 1) / gridDim.x + ppb - 1) / ppb * ppb;   // contiguous run of points per block
    const int p_end = min(N, per * ((int)blockIdx.x + 1));
    const int j = threadIdx.x % H;
    int p = per * (int)blockIdx.x + threadIdx.x / H;
    if (p >= p_end) return;
    const float4* vals4 = reinterpret_cast<const float4*>(vals);
    const int C4 = 2 * H;
    int4 i4 = __ldg(reinterpret_cast<const int4*>(idx) + p);
    float4 w4 = __ldg(reinterpret_cast<const float4*>(w) + p);
    while (true) {
        const int pn = p + ppb;
        const bool more = pn < p_end;
        const int pl = more ? pn : p;
        const int4 i4n = __ldg(reinterpret_cast<const int4*>(idx) + pl);
        const float4 w4n = __ldg(reinterpret_cast<const float4*>(w) + pl);
        const int ids[4] = {i4.x, i4.y, i4.z, i4.w};
        const float ws[4] = {w4.x, w4.y, w4.z, w4.w};
        float4 xa[4], xb[4];
        bool ok[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {   // row 0 stands in for an absent vertex: the loads stay unconditional and independent
            ok[r] = (unsigned)ids[r] < (unsigned)V;
            const float4* row = vals4 + (unsigned)(ok[r] ? ids[r] : 0) * (unsigned)C4 + j;
            xa[r] = __ldg(row);
            xb[r] = __ldg(row + H);
        }
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < 4; ++r)
            if (ok[r]) {
                a.x = fmaf(ws[r], xa[r].x, a.x); a.y = fmaf(ws[r], xa[r].y, a.y); a.z = fmaf(ws[r], xa[r].z, a.z); a.w = fmaf(ws[r], xa[r].w, a.w);
                b.x = fmaf(ws[r], xb[r].x, b.x); b.y = fmaf(ws[r], xb[r].y, b.y); b.z = fmaf(ws[r], xb[r].z, b.z); b.w = fmaf(ws[r], xb[r].w, b.w);
            }
        float4* o = out + (size_t)p * C4 + j;
        __stcs(o, a);
        __stcs(o + H, b);
        if (!more) break;
        p = pn; i4 = i4n; w4 = w4n;
    }
}

// slice backward wrt vals: grad_vals[idx[p,r], :] += w[p,r] * grad_out[p, :]  (atomics)
__global__ void __launch_bounds__(kThreads)
k_slice_bwd(const float* __restrict__ grad_out, int N, int C, const int* __restrict__ idx,
            const float* __restrict__ w, float* grad_vals, int V) {
    long long total = (long long)N * C;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        int p = (int)(t / C);
        int c = (int)(t - (long long)p * C);
        float g = __ldg(grad_out + t);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int id = __ldg(idx + (size_t)p * 4 + r);
            if (id >= 0 && id < V) atomicAdd(grad_vals + (size_t)id * C + c, __ldg(w + (size_t)p * 4 + r) * g);
        }
    }
}

// gather: out[p, r*(C+1) + c] = w*vals[id, c] ; out[p, r*(C+1) + C] = w ; absent -> zeros (conv. U6)
__global__ void __launch_bounds__(kThreads)
k_gather(const float* __restrict__ vals, int V, const int* __restrict__ v_dev, int C, const int* __restrict__ idx,
         const float* __restrict__ w, int N, const int* __restrict__ n_dev, float* __restrict__ out) {
    if (v_dev) V = min(V, *v_dev);
    if (n_dev) N = min(N, *n_dev);
    const int W1 = C + 1;
    long long total = (long long)N * 4 * W1;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        long long row = t / W1;
        int c = (int)(t - row * W1);
        int id = __ldg(idx + row);
        float r = 0.f;
        if (id >= 0 && id < V) {
            float ww = __ldg(w + row);
            r = (c < C) ? __fmul_rn(ww, __ldg(vals + (size_t)id * C + c)) : ww;
        }
        out[t] = r;
    }
}

// gather backward wrt vals: grad_vals[id, c] += w * grad_out[p, r*(C+1)+c]
__global__ void __launch_bounds__(kThreads)
k_gather_bwd(const float* __restrict__ grad_out, int N, int C, const int* __restrict__ idx,
             const float* __restrict__ w, float* grad_vals, int V) {
    const int W1 = C + 1;
    long long total = (long long)N * 4 * C;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        long long row = t / C;
        int c = (int)(t - row * C);
        int id = __ldg(idx + row);
        if (id >= 0 && id < V) atomicAdd(grad_vals + (size_t)id * C + c, __ldg(w + row) * __ldg(grad_out + row * W1 + c));
    }
}

// slice_classify forward: one warp per point.
//   s[c]      = sum_r (w+dw)[p,r] * vals[id_r, c]
//   logit[k]  = b[k] + sum_c Wc[k,c] * s[c]
// Wc (K x C) is staged in shared memory once per block.
__global__ void __launch_bounds__(kThreads)
k_slice_classify(const float* __restrict__ vals, int V, const int* __restrict__ v_dev, int C, const int* __restrict__ idx,
                 const float* __restrict__ w, const float* __restrict__ dw, int N, const int* __restrict__ n_dev,
                 const float* __restrict__ Wc, const float* __restrict__ bias, int K, float* __restrict__ out,
                 float* __restrict__ sliced) {
    extern __shared__ float sW[];  // K*C
    if (v_dev) V = min(V, *v_dev);
    if (n_dev) N = min(N, *n_dev);
    for (int i = threadIdx.x; i < K * C; i += blockDim.x) sW[i] = __ldg(Wc + i);
    __syncthreads();
    int lane = threadIdx.x & 31;
    int warps_per_block = blockDim.x >> 5;
    for (int p = blockIdx.x * warps_per_block + (threadIdx.x >> 5); p < N; p += gridDim.x * warps_per_block) {
        int4 i4 = __ldg(reinterpret_cast<const int4*>(idx) + p);
        float4 w4 = __ldg(reinterpret_cast<const float4*>(w) + p);
        float4 d4 = __ldg(reinterpret_cast<const float4*>(dw) + p);
        int ids[4] = {i4.x, i4.y, i4.z, i4.w};
        float ws[4] = {w4.x + d4.x, w4.y + d4.y, w4.z + d4.z, w4.w + d4.w};
        // channel sums for this lane's channels (C <= 256 -> at most 8 per lane), gathered once
        float sc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int c = lane + 32 * j;
            float s = 0.f;
            if (c < C) {
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    if (ids[r] >= 0 && ids[r] < V) s = fmaf(ws[r], __ldg(vals + (size_t)ids[r] * C + c), s);
                if (sliced) sliced[(size_t)p * C + c] = s;
            }
            sc[j] = s;
        }
        for (int k = 0; k < K; ++k) {
            float v = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                int c = lane + 32 * j;
                if (c < C) v = fmaf(sW[k * C + c], sc[j], v);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) out[(size_t)p * K + k] = v + __ldg(bias + k);
        }
    }
}

// slice_classify backward.  Given grad_logit [N,K]:
//   gs[c]          = sum_k grad_logit[p,k] * Wc[k,c]
//   grad_dw[p,r]   = sum_c gs[c] * vals[id_r, c]
//   grad_vals[id_r, c] += (w+dw)[p,r] * gs[c]                       (atomics)
// grad_Wc = grad_logit^T @ sliced and grad_b = column sums are done with a GEMM / reduction on the
// host side from the `sliced` tensor saved by the forward.
__global__ void __launch_bounds__(kThreads)
k_slice_classify_bwd(const float* __restrict__ grad_logit, const float* __restrict__ vals, int V, int C,
                     const int* __restrict__ idx, const float* __restrict__ w, const float* __restrict__ dw, int N,
                     const float* __restrict__ Wc, int K, float* grad_vals, float* __restrict__ grad_dw) {
    extern __shared__ float sW[];
    for (int i = threadIdx.x; i < K * C; i += blockDim.x) sW[i] = __ldg(Wc + i);
    __syncthreads();
    int lane = threadIdx.x & 31;
    int warps_per_block = blockDim.x >> 5;
    for (int p = blockIdx.x * warps_per_block + (threadIdx.x >> 5); p < N; p += gridDim.x * warps_per_block) {
        int4 i4 = __ldg(reinterpret_cast<const int4*>(idx) + p);
        float4 w4 = __ldg(reinterpret_cast<const float4*>(w) + p);
        float4 d4 = __ldg(reinterpret_cast<const float4*>(dw) + p);
        int ids[4] = {i4.x, i4.y, i4.z, i4.w};
        float ws[4] = {w4.x + d4.x, w4.y + d4.y, w4.z + d4.z, w4.w + d4.w};
        float gd[4] = {0.f, 0.f, 0.f, 0.f};
        for (int c = lane; c < C; c += 32) {
            float gs = 0.f;
            for (int k = 0; k < K; ++k) gs = fmaf(__ldg(grad_logit + (size_t)p * K + k), sW[k * C + c], gs);
#pragma unroll
            for (int r = 0; r < 4; ++r)
                if (ids[r] >= 0 && ids[r] < V) {
                    gd[r] = fmaf(gs, __ldg(vals + (size_t)ids[r] * C + c), gd[r]);
                    atomicAdd(grad_vals + (size_t)ids[r] * C + c, ws[r] * gs);
                }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            float v = gd[r];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) grad_dw[(size_t)p * 4 + r] = v;
        }
    }
}

inline int grid_for(long long work_items, int threads) {
    long long b = (work_items + threads - 1) / threads;
    const long long cap = 148LL * 32;  // persistent-ish: a few waves of 148 SMs, grid-stride inside
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

extern "C" {

int ltn_im2row(const float* vals, int Vvals, const int* vvals_dev, const int* nbr, int Vq, const int* vq_dev, int C,
               float* out, void* stream) {
    if (Vq <= 0 || C <= 0) return 0;
    if (C % 4) return -2;
    long long total = (long long)Vq * LTN_FEXT * (C / 4);
    k_im2row<<<grid_for(total, kThreads), kThreads, 0, (cudaStream_t)stream>>>(vals, Vvals, vvals_dev, nbr, Vq, vq_dev,
                                                                              C / 4, reinterpret_cast<float4*>(out));
    LTN_CHECK_LAUNCH();
    return 0;
}

int ltn_row2im(const float* grad_rows, int Vrows, const int* nbrT, int Vu, int C, float* grad_vals, void* stream) {
    if (Vu <= 0 || C <= 0) return 0;
    if (C % 4) return -2;
    long long total = (long long)Vu * (C / 4);
    k_row2im<<<grid_for(total, kThreads), kThreads, 0, (cudaStream_t)stream>>>(grad_rows, Vrows, nbrT, Vu, C / 4,
                                                                              reinterpret_cast<float4*>(grad_vals));
    LTN_CHECK_LAUNCH();
    return 0;
}

// out [V, C+1] must be zeroed by the caller
int ltn_splat(const float* val, int N, int C, const int* idx, const float* w, float* out, int V, void* stream) {
    if (N <= 0) return 0;
    if (C >= 1 && C <= 3) {
        static const int run = []() { const char* e = getenv("LTN_SPLAT_RUN"); return e && atoi(e) > 0 ? atoi(e) : 0; }();
        if (run > 0) {   // the run-length walk of round 1, kept for comparison (LTN_SPLAT_RUN=16)
            const int threads = (N + run - 1) / run;
            const int blocks = ltn_blocks(threads, kThreads);
            if (C == 1) k_splat_runs<1><<<blocks, kThreads, 0, (cudaStream_t)stream>>>(val, N, idx, w, out, V, run);
            else if (C == 2) k_splat_runs<2><<<blocks, kThreads, 0, (cudaStream_t)stream>>>(val, N, idx, w, out, V, run);
            else k_splat_runs<3><<<blocks, kThreads, 0, (cudaStream_t)stream>>>(val, N, idx, w, out, V, run);
        } else {
            const int blocks = ltn_blocks(N, kThreads);
            if (C == 1) k_splat_seg<1><<<blocks, kThreads, 0, (cudaStream_t)stream>>>(val, N, idx, w, out, V);
            else if (C == 2) k_splat_seg<2><<<blocks, kThreads, 0, (cudaStream_t)stream>>>(val, N, idx, w, out, V);
            else k_splat_seg<3><<<blocks, kThreads, 0, (cudaStream_t)stream>>>(val, N, idx, w, out, V);
        }
        LTN_CHECK_LAUNCH();
        return 0;
    }
    k_splat<<<ltn_blocks(N, kThreads), kThreads, 0, (cudaStream_t)stream>>>(val, N, C, idx, w, out, V);
    LTN_CHECK_LAUNCH();
    return 0;
}

int ltn_slice(const float* vals, int V, int C, const int* idx, const float* w, int N, float* out, void* stream) {
    if (N <= 0 || C <= 0) return 0;
    if (C % 4) return -2;
    const int C4 = C / 4;
    long long total = (long long)N * C4;
    long long blocks = (total + kThreads - 1) / kThreads;
    if (blocks > 148 * 8) blocks = 148 * 8;   // one contiguous run of points per block (see k_slice)
    if (C4 % 2 == 0 && kThreads % (C4 / 2) == 0 && (long long)V * C4 < (1ll << 31)) {
        const int H = C4 / 2;
        long long b2 = ((long long)N * H + kThreads - 1) / kThreads;
        if (b2 > 148 * 8) b2 = 148 * 8;
        k_slice_pair<<<(int)b2, kThreads, 0, (cudaStream_t)stream>>>(vals, V, H, idx, w, N, reinterpret_cast<float4*>(out));
        LTN_CHECK_LAUNCH();
        return 0;
    }
    k_slice<<<(int)blocks, kThreads, 0, (cudaStream_t)stream>>>(vals, V, C4, idx, w, N,
                                                                             reinterpret_cast<float4*>(out));
    LTN_CHECK_LAUNCH();
    return 0;
}

// grad_vals [V,C] must be zeroed by the caller
int ltn_slice_bwd(const float* grad_out, int N, int C, const int* idx, const float* w, float* grad_vals, int V,
                  void* stream) {
    if (N <= 0 || C <= 0) return 0;
    k_slice_bwd<<<grid_for((long long)N * C, kThreads), kThreads, 0, (cudaStream_t)stream>>>(grad_out, N, C, idx, w,
                                                                                            grad_vals, V);
    LTN_CHECK_LAUNCH();
    return 0;
}

int ltn_gather(const float* vals, int V, const int* v_dev, int C, const int* idx, const float* w, int N, const int* n_dev,
               float* out, void* stream) {
    if (N <= 0) return 0;
    long long total = (long long)N * 4 * (C + 1);
    k_gather<<<grid_for(total, kThreads), kThreads, 0, (cudaStream_t)stream>>>(vals, V, v_dev, C, idx, w, N, n_dev, out);
    LTN_CHECK_LAUNCH();
    return 0;
}

// grad_vals [V,C] must be zeroed by the caller
int ltn_gather_bwd(const float* grad_out, int N, int C, const int* idx, const float* w, float* grad_vals, int V,
                   void* stream) {
    if (N <= 0) return 0;
    k_gather_bwd<<<grid_for((long long)N * 4 * C, kThreads), kThreads, 0, (cudaStream_t)stream>>>(grad_out, N, C, idx,
                                                                                                 w, grad_vals, V);
    LTN_CHECK_LAUNCH();
    return 0;
}

// sliced (nullable): [N,C] = sum_r (w+dw) vals[id_r] saved for the weight gradient
int ltn_slice_classify(const float* vals, int V, const int* v_dev, int C, const int* idx, const float* w, const float* dw,
                       int N, const int* n_dev, const float* Wc, const float* bias, int K, float* out, float* sliced,
                       void* stream) {
    if (N <= 0) return 0;
    size_t smem = sizeof(float) * (size_t)K * C;
    if (smem > 200 * 1024 || C > 256) return -3;
    cudaError_t e = cudaFuncSetAttribute(k_slice_classify, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int blocks = (int)((N + 7) / 8);
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_slice_classify<<<blocks, kThreads, smem, (cudaStream_t)stream>>>(vals, V, v_dev, C, idx, w, dw, N, n_dev, Wc, bias, K, out,
                                                                       sliced);
    LTN_CHECK_LAUNCH();
    return 0;
}

// grad_vals [V,C] must be zeroed by the caller; grad_dw [N,4] is fully written
int ltn_slice_classify_bwd(const float* grad_logit, const float* vals, int V, int C, const int* idx, const float* w,
                           const float* dw, int N, const float* Wc, int K, float* grad_vals, float* grad_dw,
                           void* stream) {
    if (N <= 0) return 0;
    size_t smem = sizeof(float) * (size_t)K * C;
    if (smem > 200 * 1024) return -3;
    cudaError_t e = cudaFuncSetAttribute(k_slice_classify_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int blocks = (int)((N + 7) / 8);
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_slice_classify_bwd<<<blocks, kThreads, smem, (cudaStream_t)stream>>>(grad_logit, vals, V, C, idx, w, dw, N, Wc, K,
                                                                           grad_vals, grad_dw);
    LTN_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
