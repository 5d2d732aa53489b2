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

// one thread per point; 4 dedupe rounds per warp.  The four ranks of a point are INDEPENDENT table operations, so their
// dependent L2 round trips (first probe -> id of the slot -> atomicMin) are issued rank-interleaved: 3-4 latencies per
// warp instead of 12.
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
    const uint32_t m = (uint32_t)nslots - 1u;
    uint64_t k[LTN_D1], first[LTN_D1];
    uint32_t h[LTN_D1];
    int leader[LTN_D1];
    bool ok[LTN_D1], mine[LTN_D1];
    int nbad = 0;
#pragma unroll
    for (int r = 0; r < LTN_D1; ++r) {
        ok[r] = valid && ltn_in_range(key[r][0], key[r][1], key[r][2]);
        nbad += (valid && !ok[r]) ? 1 : 0;
        // lanes without a key take a private dummy so they never match a real one
        k[r] = ok[r] ? ltn_pack(key[r][0], key[r][1], key[r][2]) : (LTN_EMPTY - 1ull - lane);
        const unsigned peers = __match_any_sync(0xffffffffu, k[r]);
        leader[r] = __ffs(peers) - 1;
        mine[r] = ((int)lane == leader[r]) && ok[r];
        h[r] = ltn_hash(k[r]) & m;
    }
#pragma unroll
    for (int r = 0; r < LTN_D1; ++r) first[r] = mine[r] ? __ldcg(slot_keys + h[r]) : 0ull;   // four probes in flight
    int slots[LTN_D1], ids[LTN_D1];
#pragma unroll
    for (int r = 0; r < LTN_D1; ++r) {
        int slot = -1;
        if (mine[r]) slot = (first[r] == k[r]) ? (int)h[r] : ltn_claim_slot(k[r], slot_keys, nslots);
        slots[r] = slot;
    }
#pragma unroll
    for (int r = 0; r < LTN_D1; ++r) ids[r] = (mine[r] && slots[r] >= 0) ? __ldcg(slot_ids + slots[r]) : 0;
#pragma unroll
    for (int r = 0; r < LTN_D1; ++r) {
        // rows grow with the lane index, so the leader holds the group's smallest row
        if (mine[r] && slots[r] >= 0 && ids[r] < 0) atomicMin(slot_first + slots[r], p * LTN_D1 + r);
        const int sl = __shfl_sync(0xffffffffu, slots[r], leader[r]);
        slots[r] = ok[r] ? sl : -1;
    }
    if (nbad) atomicAdd(counters + LTN_CNT_RANGE, nbad);
    if (valid) {
        *reinterpret_cast<int4*>(row_slot + (size_t)p * LTN_D1) = make_int4(slots[0], slots[1], slots[2], slots[3]);
        if (row_w) *reinterpret_cast<float4*>(row_w + (size_t)p * LTN_D1) = make_float4(bary[0], bary[1], bary[2], bary[3]);
    }
}

// Is `row` the first appearance of a vertex that is new in this batch?  slot_first[s] is only ever lowered (atomicMin in
// k_insert_points) for slots that have no id yet and is put back to INT_MAX when the id is assigned, so
// slot_first[s] == row says both "new" and "first": ONE random table read per row.
__device__ __forceinline__ int is_first_row(int row, int R, const int* __restrict__ row_slot,
                                            const int* __restrict__ slot_ids, const int* __restrict__ slot_first) {
    if (row >= R) return 0;
    int s = row_slot[row];
    if (s < 0) return 0;
    return __ldcg(slot_first + s) == row ? 1 : 0;
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

// distribute, second half (a): slot -> id, per-vertex position sums (double) + count.  No row is written yet: the
// local mean has to be known first, and every [4N, 3+vd+1] row is then written exactly once (k_distribute_write).
// The sums are warp-aggregated: the 32 points of a warp are scan neighbours and mostly share their simplex
// vertices, so lanes with the same vertex id add up inside the warp and one lane issues the double atomics.
// (Sums of <= 2^29 float32 values in double are exact, so the result does not depend on the order of the atomics.)
__global__ void __launch_bounds__(kThreads)
k_distribute_sums(const float* __restrict__ pos, int N, const int* __restrict__ n_dev, const int* __restrict__ row_slot,
                  const int* __restrict__ slot_ids, int* __restrict__ idx, double* vert_acc) {
    if (n_dev) N = min(N, *n_dev);
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = p < N;   // no early return: the warp-wide aggregation below needs every lane
    float px = 0.f, py = 0.f, pz = 0.f;
    int s[4] = {-1, -1, -1, -1};
    if (valid) {
        px = pos[(size_t)p * 3]; py = pos[(size_t)p * 3 + 1]; pz = pos[(size_t)p * 3 + 2];
        int4 s4 = *reinterpret_cast<const int4*>(row_slot + (size_t)p * LTN_D1);
        s[0] = s4.x; s[1] = s4.y; s[2] = s4.z; s[3] = s4.w;
    }
    int ids[4];
#pragma unroll
    for (int r = 0; r < LTN_D1; ++r) ids[r] = (s[r] >= 0) ? __ldg(slot_ids + s[r]) : -1;   // four lookups in flight
    if (valid) *reinterpret_cast<int4*>(idx + (size_t)p * LTN_D1) = make_int4(ids[0], ids[1], ids[2], ids[3]);
#pragma unroll
    for (int r = 0; r < LTN_D1; ++r) {
        const int a = valid ? (ids[r] < 0 ? 0 : ids[r]) : -1;  // ids < 0 fold onto vertex 0 (lattice_modules.py:479-480)
        float sum[3] = {px, py, pz};
        int cnt;
        if (ltn_warp_group_sum<3>(a, sum, cnt)) {
            atomicAdd(vert_acc + (size_t)a * 4 + 0, (double)sum[0]);
            atomicAdd(vert_acc + (size_t)a * 4 + 1, (double)sum[1]);
            atomicAdd(vert_acc + (size_t)a * 4 + 2, (double)sum[2]);
            atomicAdd(vert_acc + (size_t)a * 4 + 3, (double)cnt);
        }
    }
}

// distribute, second half (b): rows[4p + r] = [pos - mean(vertex), val..., w_r], written ONCE.  A block stages the rows
// of its 256 points in shared memory (thread = point) and streams them out as one contiguous run of 128-bit stores.
__global__ void __launch_bounds__(kThreads)
k_distribute_write(const float* __restrict__ pos, const float* __restrict__ val, int N, const int* __restrict__ n_dev, int vd,
                   const int* __restrict__ idx, const float* __restrict__ row_w, const double* __restrict__ vert_acc,
                   int subtract_mean, float* __restrict__ rows) {
    extern __shared__ __align__(16) float s_rows[];
    if (n_dev) N = min(N, *n_dev);
    const int width = LTN_D + vd + 1;
    const int p0 = blockIdx.x * kThreads;
    const int p = p0 + threadIdx.x;
    if (p < N) {
        const float px = pos[(size_t)p * 3], py = pos[(size_t)p * 3 + 1], pz = pos[(size_t)p * 3 + 2];
        const int4 i4 = *reinterpret_cast<const int4*>(idx + (size_t)p * LTN_D1);
        const float4 w4 = *reinterpret_cast<const float4*>(row_w + (size_t)p * LTN_D1);
        const int ids[4] = {i4.x, i4.y, i4.z, i4.w};
        const float w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
        for (int r = 0; r < LTN_D1; ++r) {
            float* o = s_rows + ((size_t)threadIdx.x * LTN_D1 + r) * width;
            float mx = 0.f, my = 0.f, mz = 0.f;
            if (subtract_mean) {
                const int id = ids[r] < 0 ? 0 : ids[r];
                const double* a = vert_acc + (size_t)id * 4;
                const double cnt = a[3];
                mx = (float)(a[0] / cnt); my = (float)(a[1] / cnt); mz = (float)(a[2] / cnt);
            }
            o[0] = __fsub_rn(px, mx); o[1] = __fsub_rn(py, my); o[2] = __fsub_rn(pz, mz);
            for (int i = 0; i < vd; ++i) o[LTN_D + i] = val[(size_t)p * vd + i];
            o[LTN_D + vd] = w[r];
        }
    }
    __syncthreads();
    const int npts = min(kThreads, N - p0);
    if (npts <= 0) return;
    const int nfl = npts * LTN_D1 * width;                    // floats of this block's contiguous run
    float* dst = rows + (size_t)p0 * LTN_D1 * width;          // 256 * 4 * width floats per block: 16-byte aligned
    const int n4 = nfl >> 2;
    for (int i = threadIdx.x; i < n4; i += kThreads)
        reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(s_rows)[i];
    for (int i = (n4 << 2) + threadIdx.x; i < nfl; i += kThreads) dst[i] = s_rows[i];
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
    k_distribute_sums<<<ltn_blocks(N, kThreads), kThreads, 0, st>>>(pos, N, n_dev, row_slot, slot_ids, idx, vert_acc);
    LTN_CHECK_LAUNCH();
    const int width = LTN_D + val_dim + 1;
    const size_t smem = sizeof(float) * kThreads * LTN_D1 * (size_t)width;
    if (smem > 200 * 1024) return -2;
    if (smem > 48 * 1024) {
        e = cudaFuncSetAttribute(k_distribute_write, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    k_distribute_write<<<ltn_blocks(N, kThreads), kThreads, smem, st>>>(pos, val, N, n_dev, val_dim, idx, w, vert_acc, subtract_mean, rows);
    LTN_CHECK_LAUNCH();
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
