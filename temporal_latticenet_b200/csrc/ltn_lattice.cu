// Lattice structure kernels: hash-table build with deterministic vertex numbering, distribute,
// coarse-vertex creation, neighbour tables.  sm_100a.
//
// Design (B200-first, not the reference's lock+counter insert):
//  * the table maps a 63-bit packed key to a slot with ONE 64-bit CAS -- no lock word, no spin;
//  * a warp first dedupes its 32 keys with __match_any_sync, so one lane per distinct key probes;
//  * vertex ids are NOT handed out by a racing atomic counter.  Each new slot records the smallest
//    row (point*4+r) that touched it; a flag-scan over the rows then numbers new vertices in order
//    of first appearance.  That is exactly the order a sequential insert loop produces, so ids are
//    reproducible run to run and equal to the scalar oracle's, and they stay append-only across the
//    frames of a window (seq_lattice/models.py:287-289 relies on that for hidden-state alignment).
#include "ltn_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kScanBlock = 1024;

__global__ void k_hash_clear(uint64_t* slot_keys, int* slot_ids, int* slot_first, int nslots, int* counters) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nslots) {
        slot_keys[i] = LTN_EMPTY;
        slot_ids[i] = -1;
        slot_first[i] = LTN_INT_MAX;
    }
    if (i < 8) counters[i] = 0;
}

// one thread per point; 4 dedupe rounds per warp
__global__ void __launch_bounds__(kThreads)
k_insert_points(const float* __restrict__ pos, int N, const int* __restrict__ n_dev, float sx, float sy, float sz,
                uint64_t* slot_keys, const int* __restrict__ slot_ids, int* slot_first, int nslots, int* counters,
                int* __restrict__ row_slot, float* __restrict__ row_w) {
    if (n_dev) N = min(N, *n_dev);
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = p < N;
    int key[LTN_D1][LTN_D];
    float bary[LTN_D1];
    if (valid) {
        ltn_simplex(pos[(size_t)p * 3], pos[(size_t)p * 3 + 1], pos[(size_t)p * 3 + 2], sx, sy, sz, key, bary);
    }
    const unsigned lane = threadIdx.x & 31u;
    int slots[LTN_D1];
#pragma unroll
    for (int r = 0; r < LTN_D1; ++r) {
        bool ok = valid && ltn_in_range(key[r][0], key[r][1], key[r][2]);
        // lanes without a key take a private dummy so they never match a real one
        uint64_t k = ok ? ltn_pack(key[r][0], key[r][1], key[r][2]) : (LTN_EMPTY - 1ull - lane);
        unsigned peers = __match_any_sync(0xffffffffu, k);
        int leader = __ffs(peers) - 1;
        int slot = -1;
        if ((int)lane == leader && ok) {
            slot = ltn_claim_slot(k, slot_keys, nslots);
            // rows grow with the lane index, so the leader holds the group's smallest row
            if (slot >= 0 && __ldcg(slot_ids + slot) < 0) atomicMin(slot_first + slot, p * LTN_D1 + r);
        }
        slot = __shfl_sync(0xffffffffu, slot, leader);
        if (valid && !ok) atomicAdd(counters + LTN_CNT_RANGE, 1);
        slots[r] = ok ? slot : -1;
    }
    if (valid) {
        *reinterpret_cast<int4*>(row_slot + (size_t)p * LTN_D1) = make_int4(slots[0], slots[1], slots[2], slots[3]);
        if (row_w) *reinterpret_cast<float4*>(row_w + (size_t)p * LTN_D1) = make_float4(bary[0], bary[1], bary[2], bary[3]);
    }
}

__device__ __forceinline__ int is_first_row(int row, int R, const int* __restrict__ row_slot,
                                            const int* __restrict__ slot_ids, const int* __restrict__ slot_first) {
    if (row >= R) return 0;
    int s = row_slot[row];
    if (s < 0) return 0;
    return (__ldcg(slot_ids + s) < 0 && __ldcg(slot_first + s) == row) ? 1 : 0;
}

__global__ void __launch_bounds__(kScanBlock)
k_number_count(const int* __restrict__ row_slot, int R, const int* __restrict__ n_dev, const int* __restrict__ slot_ids,
               const int* __restrict__ slot_first, int* __restrict__ block_sums) {
    __shared__ int warp_sums[kScanBlock / 32];
    if (n_dev) R = min(R, *n_dev * LTN_D1);
    int row = blockIdx.x * kScanBlock + threadIdx.x;
    int f = is_first_row(row, R, row_slot, slot_ids, slot_first);
    unsigned b = __ballot_sync(0xffffffffu, f);
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = __popc(b);
    __syncthreads();
    if (threadIdx.x < 32) {
        int v = warp_sums[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) block_sums[blockIdx.x] = v;
    }
}

// single block: exclusive scan of block_sums in place, then publish the new vertex count
__global__ void __launch_bounds__(1024)
k_number_scan(int* block_sums, int nblocks, int* counters, int cap) {
    __shared__ int sh[1024];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += 1024) {
        int i = base + threadIdx.x;
        int v = (i < nblocks) ? block_sums[i] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            int t = (threadIdx.x >= o) ? sh[threadIdx.x - o] : 0;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        int incl = sh[threadIdx.x];
        if (i < nblocks) block_sums[i] = carry + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        int before = counters[LTN_CNT_FILLED];
        int after = before + carry;
        counters[LTN_CNT_PREV] = before;
        if (after > cap) { counters[LTN_CNT_OVERFLOW] += after - cap; after = cap; }
        counters[LTN_CNT_FILLED] = after;
    }
}

__global__ void __launch_bounds__(kScanBlock)
k_number_assign(const int* __restrict__ row_slot, int R, const int* __restrict__ n_dev, const uint64_t* __restrict__ slot_keys,
                int* slot_ids, int* slot_first, const int* __restrict__ block_sums, const int* __restrict__ counters, int cap,
                int4* __restrict__ keys) {
    __shared__ int warp_off[kScanBlock / 32];
    if (n_dev) R = min(R, *n_dev * LTN_D1);
    int row = blockIdx.x * kScanBlock + threadIdx.x;
    int f = is_first_row(row, R, row_slot, slot_ids, slot_first);
    unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    unsigned b = __ballot_sync(0xffffffffu, f);
    int in_warp = __popc(b & ((1u << lane) - 1u));
    if (lane == 0) warp_off[warp] = __popc(b);
    __syncthreads();
    if (threadIdx.x < 32) {
        int v = warp_off[threadIdx.x], incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)lane >= o) incl += t;
        }
        warp_off[threadIdx.x] = incl - v;
    }
    __syncthreads();
    if (f) {
        int id = counters[LTN_CNT_PREV] + block_sums[blockIdx.x] + warp_off[warp] + in_warp;
        int s = row_slot[row];
        slot_first[s] = LTN_INT_MAX;
        if (id < cap) {
            int x, y, z;
            ltn_unpack(slot_keys[s], x, y, z);
            keys[id] = make_int4(x, y, z, -(x + y + z));
            slot_ids[s] = id;
        }
    }
}

// distribute, second half: slot -> id, rows [4N, 3+vd+1], per-vertex position sums (double) + count.
// The sums are warp-aggregated: the 32 points of a warp are scan neighbours and mostly share their simplex
// vertices, so lanes with the same vertex id add up inside the warp and one lane issues the double atomics.
__global__ void __launch_bounds__(kThreads)
k_distribute_rows(const float* __restrict__ pos, const float* __restrict__ val, int N, const int* __restrict__ n_dev, int vd,
                  const int* __restrict__ row_slot, const float* __restrict__ row_w,
                  const int* __restrict__ slot_ids, float* __restrict__ rows, int* __restrict__ idx,
                  double* vert_acc) {
    if (n_dev) N = min(N, *n_dev);
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = p < N;   // no early return: the warp-wide aggregation below needs every lane
    const int width = LTN_D + vd + 1;
    float px = 0.f, py = 0.f, pz = 0.f;
    int s[4] = {-1, -1, -1, -1};
    float w[4] = {0.f, 0.f, 0.f, 0.f};
    if (valid) {
        px = pos[(size_t)p * 3]; py = pos[(size_t)p * 3 + 1]; pz = pos[(size_t)p * 3 + 2];
        int4 s4 = *reinterpret_cast<const int4*>(row_slot + (size_t)p * LTN_D1);
        float4 w4 = *reinterpret_cast<const float4*>(row_w + (size_t)p * LTN_D1);
        s[0] = s4.x; s[1] = s4.y; s[2] = s4.z; s[3] = s4.w;
        w[0] = w4.x; w[1] = w4.y; w[2] = w4.z; w[3] = w4.w;
    }
    int ids[4];
#pragma unroll
    for (int r = 0; r < LTN_D1; ++r) {
        int id = (s[r] >= 0) ? __ldg(slot_ids + s[r]) : -1;
        ids[r] = id;
        if (valid) {
            float* o = rows + ((size_t)p * LTN_D1 + r) * width;
            o[0] = px; o[1] = py; o[2] = pz;
            for (int i = 0; i < vd; ++i) o[LTN_D + i] = val[(size_t)p * vd + i];
            o[LTN_D + vd] = w[r];
        }
        const int a = valid ? (id < 0 ? 0 : id) : -1;  // ids < 0 fold onto vertex 0 (lattice_modules.py:479-480)
        float sum[3] = {px, py, pz};
        int cnt;
        if (ltn_warp_group_sum<3>(a, sum, cnt)) {
            atomicAdd(vert_acc + (size_t)a * 4 + 0, (double)sum[0]);
            atomicAdd(vert_acc + (size_t)a * 4 + 1, (double)sum[1]);
            atomicAdd(vert_acc + (size_t)a * 4 + 2, (double)sum[2]);
            atomicAdd(vert_acc + (size_t)a * 4 + 3, (double)cnt);
        }
    }
    if (valid) *reinterpret_cast<int4*>(idx + (size_t)p * LTN_D1) = make_int4(ids[0], ids[1], ids[2], ids[3]);
}

__global__ void __launch_bounds__(kThreads)
k_local_mean_sub(float* __restrict__ rows, const int* __restrict__ idx, int R, const int* __restrict__ n_dev, int width,
                 const double* __restrict__ vert_acc) {
    if (n_dev) R = min(R, *n_dev * LTN_D1);
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    int id = idx[r];
    id = id < 0 ? 0 : id;
    double cnt = vert_acc[(size_t)id * 4 + 3];
    float* o = rows + (size_t)r * width;
#pragma unroll
    for (int i = 0; i < LTN_D; ++i) {
        float mean = (float)(vert_acc[(size_t)id * 4 + i] / cnt);
        o[i] = __fsub_rn(o[i], mean);
    }
}

__global__ void k_vertex_counts(const double* __restrict__ vert_acc, int V, float* __restrict__ counts) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < V) counts[v] = (float)vert_acc[(size_t)v * 4 + 3];
}

// neighbour table: one thread per (vertex, slot).  SURVEY.md appendix B.6.
__global__ void __launch_bounds__(kThreads)
k_neighbours(const int4* __restrict__ keys_q, int Vq, const int* __restrict__ vq_dev,
             const uint64_t* __restrict__ slot_keys, const int* __restrict__ slot_ids, int nslots, int mode,
             int dil, int same_table, int* __restrict__ out) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int v = t / LTN_FEXT, s = t - v * LTN_FEXT;
    if (vq_dev) Vq = min(Vq, *vq_dev);
    if (v >= Vq) return;
    int4 k = keys_q[v];
    int c[4] = {k.x, k.y, k.z, k.w};
    int tap[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int o = 0;
        if (s < 2 * LTN_D1) {
            o = (i == (s >> 1)) ? -LTN_D : 1;
            if (s & 1) o = -o;
            o *= dil;
        }
        tap[i] = (mode == 1) ? 2 * c[i] + o : c[i] + o;
    }
    int id = -1;
    bool ok = true;
    if (mode == 2) {
        ok = (((tap[0] | tap[1] | tap[2] | tap[3]) & 1) == 0);
        tap[0] /= 2; tap[1] /= 2; tap[2] /= 2;
    }
    if (s == LTN_FEXT - 1 && mode == 0 && same_table) {
        id = v;  // centre of a same-lattice table is the vertex itself (convention U5)
    } else if (ok && ltn_in_range(tap[0], tap[1], tap[2])) {
        int slot = ltn_find_slot(ltn_pack(tap[0], tap[1], tap[2]), slot_keys, nslots);
        if (slot >= 0) id = __ldg(slot_ids + slot);
    }
    out[t] = id;
}

}  // namespace

extern "C" {

int ltn_hash_clear(uint64_t* slot_keys, int* slot_ids, int* slot_first, int nslots, int* counters, void* stream) {
    k_hash_clear<<<ltn_blocks(nslots > 8 ? nslots : 8, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        slot_keys, slot_ids, slot_first, nslots, counters);
    LTN_CHECK_LAUNCH();
    return 0;
}

// Inserts the 4 simplex vertices of N points and numbers the new vertices deterministically.
// row_slot [4N] int32 (scratch, returned), row_w [4N] float (nullable), block_sums scratch of
// ceil(4N/1024)+1 ints.  keys is [cap,4] int32.
int ltn_insert_points(const float* pos, int N, const int* n_dev, float sx, float sy, float sz, uint64_t* slot_keys,
                      int* slot_ids, int* slot_first, int nslots, int* counters, int* keys, int cap, int* row_slot,
                      float* row_w, int* block_sums, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (N <= 0) return 0;
    k_insert_points<<<ltn_blocks(N, kThreads), kThreads, 0, st>>>(pos, N, n_dev, sx, sy, sz, slot_keys, slot_ids, slot_first,
                                                                  nslots, counters, row_slot, row_w);
    LTN_CHECK_LAUNCH();
    int R = N * LTN_D1;
    int nb = ltn_blocks(R, kScanBlock);
    k_number_count<<<nb, kScanBlock, 0, st>>>(row_slot, R, n_dev, slot_ids, slot_first, block_sums);
    LTN_CHECK_LAUNCH();
    k_number_scan<<<1, 1024, 0, st>>>(block_sums, nb, counters, cap);
    LTN_CHECK_LAUNCH();
    k_number_assign<<<nb, kScanBlock, 0, st>>>(row_slot, R, n_dev, slot_keys, slot_ids, slot_first, block_sums, counters, cap,
                                              reinterpret_cast<int4*>(keys));
    LTN_CHECK_LAUNCH();
    return 0;
}

// distribute = insert + rows/idx/w outputs (+ per-vertex position sums for the local mean).
// vert_acc: [cap,4] double scratch, zeroed here.  subtract_mean: 0 keeps raw positions.
int ltn_distribute(const float* pos, const float* val, int N, const int* n_dev, int val_dim, float sx, float sy, float sz,
                   uint64_t* slot_keys, int* slot_ids, int* slot_first, int nslots, int* counters, int* keys, int cap,
                   int* row_slot, int* block_sums, double* vert_acc, float* rows, int* idx, float* w,
                   int subtract_mean, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (N <= 0) return 0;
    int rc = ltn_insert_points(pos, N, n_dev, sx, sy, sz, slot_keys, slot_ids, slot_first, nslots, counters, keys, cap,
                               row_slot, w, block_sums, stream);
    if (rc) return rc;
    cudaError_t e = cudaMemsetAsync(vert_acc, 0, sizeof(double) * 4 * (size_t)cap, st);
    if (e != cudaSuccess) return (int)e;
    k_distribute_rows<<<ltn_blocks(N, kThreads), kThreads, 0, st>>>(pos, val, N, n_dev, val_dim, row_slot, w, slot_ids, rows,
                                                                    idx, vert_acc);
    LTN_CHECK_LAUNCH();
    if (subtract_mean) {
        int R = N * LTN_D1;
        k_local_mean_sub<<<ltn_blocks(R, kThreads), kThreads, 0, st>>>(rows, idx, R, n_dev, LTN_D + val_dim + 1, vert_acc);
        LTN_CHECK_LAUNCH();
    }
    return 0;
}

int ltn_vertex_counts(const double* vert_acc, int V, float* counts, void* stream) {
    if (V <= 0) return 0;
    k_vertex_counts<<<ltn_blocks(V, kThreads), kThreads, 0, (cudaStream_t)stream>>>(vert_acc, V, counts);
    LTN_CHECK_LAUNCH();
    return 0;
}

// mode 0 same level, 1 query coarse / table fine (coarsen), 2 query fine / table coarse (finefy)
int ltn_neighbours(const int* keys_q, int Vq, const int* vq_dev, const uint64_t* slot_keys, const int* slot_ids,
                   int nslots, int mode, int dilation, int same_table, int* out, void* stream) {
    if (Vq <= 0) return 0;
    long long n = (long long)Vq * LTN_FEXT;
    k_neighbours<<<ltn_blocks(n, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const int4*>(keys_q), Vq, vq_dev, slot_keys, slot_ids, nslots, mode, dilation, same_table, out);
    LTN_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
