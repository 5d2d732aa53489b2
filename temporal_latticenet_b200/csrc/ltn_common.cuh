// Shared device helpers for the sm_100a lattice kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define LTN_D 3
#define LTN_D1 4
#define LTN_FEXT 9               // 2(d+1)+1, seq_lattice/lattice_modules.py:299
#define LTN_EMPTY 0xFFFFFFFFFFFFFFFFull
#define LTN_KEY_BIAS (1 << 20)   // 21 bits per packed coordinate
#define LTN_INT_MAX 0x7FFFFFFF

// counters[] layout (int32[8]) shared with the Python side
#define LTN_CNT_FILLED 0      // number of vertices == next id
#define LTN_CNT_PREV 1        // value of FILLED before the last insertion batch
#define LTN_CNT_OVERFLOW 2    // vertices dropped because the table was full (convention U4)
#define LTN_CNT_RANGE 3       // keys outside the packable +-2^20 range

// host-side count of kernel launches issued by this library (ltn_launch_count in the C ABI)
extern unsigned long long g_ltn_launches;

#define LTN_CHECK_LAUNCH()                       \
    do {                                         \
        cudaError_t e__ = cudaGetLastError();    \
        if (e__ != cudaSuccess) return (int)e__; \
        ++g_ltn_launches;                        \
    } while (0)

static inline int ltn_blocks(long long n, int threads) { return (int)((n + threads - 1) / threads); }

__device__ __forceinline__ uint64_t ltn_pack(int x, int y, int z) {
    return ((uint64_t)(uint32_t)(x + LTN_KEY_BIAS) << 42) | ((uint64_t)(uint32_t)(y + LTN_KEY_BIAS) << 21) |
           (uint64_t)(uint32_t)(z + LTN_KEY_BIAS);
}
__device__ __forceinline__ bool ltn_in_range(int x, int y, int z) {
    return (unsigned)(x + LTN_KEY_BIAS) < (2u << 20) && (unsigned)(y + LTN_KEY_BIAS) < (2u << 20) &&
           (unsigned)(z + LTN_KEY_BIAS) < (2u << 20);
}
__device__ __forceinline__ void ltn_unpack(uint64_t k, int& x, int& y, int& z) {
    x = (int)((k >> 42) & 0x1FFFFF) - LTN_KEY_BIAS;
    y = (int)((k >> 21) & 0x1FFFFF) - LTN_KEY_BIAS;
    z = (int)(k & 0x1FFFFF) - LTN_KEY_BIAS;
}
__device__ __forceinline__ uint32_t ltn_hash(uint64_t k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return (uint32_t)k;
}

// lookup only: slot index or -1
__device__ __forceinline__ int ltn_find_slot(uint64_t key, const uint64_t* __restrict__ slot_keys, int nslots) {
    uint32_t m = (uint32_t)nslots - 1u, s = ltn_hash(key) & m;
    for (int probe = 0; probe < nslots; ++probe) {
        uint64_t cur = __ldg(slot_keys + s);
        if (cur == key) return (int)s;
        if (cur == LTN_EMPTY) return -1;
        s = (s + 1u) & m;
    }
    return -1;
}

// find or claim with one 64-bit CAS on the packed key (no lock word, no spin)
__device__ __forceinline__ int ltn_claim_slot(uint64_t key, uint64_t* slot_keys, int nslots) {
    uint32_t m = (uint32_t)nslots - 1u, s = ltn_hash(key) & m;
    for (int probe = 0; probe < nslots; ++probe) {
        uint64_t cur = __ldcg(slot_keys + s);
        if (cur == key) return (int)s;
        if (cur == LTN_EMPTY) {
            unsigned long long old = atomicCAS((unsigned long long*)(slot_keys + s), (unsigned long long)LTN_EMPTY,
                                               (unsigned long long)key);
            if (old == LTN_EMPTY || old == key) return (int)s;
        }
        s = (s + 1u) & m;
    }
    return -1;
}

// Enclosing simplex of one point.  Every float op is an explicit
// round-to-nearest intrinsic so the result is bit-identical to oracle/lattice_oracle.c::orc_simplex
// regardless of -fmad.
__device__ __forceinline__ void ltn_simplex(float px, float py, float pz, float sx, float sy, float sz,
                                            int key[LTN_D1][LTN_D], float bary[LTN_D1]) {
    float cf[LTN_D] = {__fmul_rn(px, sx), __fmul_rn(py, sy), __fmul_rn(pz, sz)};
    float e[LTN_D1];
    float sm = 0.0f;
#pragma unroll
    for (int i = LTN_D; i > 0; --i) {
        e[i] = __fsub_rn(sm, __fmul_rn((float)i, cf[i - 1]));
        sm = __fadd_rn(sm, cf[i - 1]);
    }
    e[0] = sm;
    int rem0[LTN_D1], rank[LTN_D1] = {0, 0, 0, 0}, sum = 0;
    float res[LTN_D1];
#pragma unroll
    for (int i = 0; i < LTN_D1; ++i) {
        float v = __fmul_rn(e[i], 0.25f);
        float up = __fmul_rn(ceilf(v), 4.0f);
        float dn = __fmul_rn(floorf(v), 4.0f);
        rem0[i] = (__fsub_rn(up, e[i]) < __fsub_rn(e[i], dn)) ? (int)up : (int)dn;
        sum += rem0[i];
    }
    sum /= LTN_D1;
#pragma unroll
    for (int i = 0; i < LTN_D1; ++i) res[i] = __fsub_rn(e[i], (float)rem0[i]);
#pragma unroll
    for (int i = 0; i < LTN_D; ++i)
#pragma unroll
        for (int j = i + 1; j < LTN_D1; ++j) {
            if (res[i] < res[j]) rank[i]++; else rank[j]++;
        }
#pragma unroll
    for (int i = 0; i < LTN_D1; ++i) {
        rank[i] += sum;
        if (rank[i] < 0) { rank[i] += LTN_D1; rem0[i] += LTN_D1; }
        else if (rank[i] > LTN_D) { rank[i] -= LTN_D1; rem0[i] -= LTN_D1; }
    }
    float b[LTN_D1 + 1] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < LTN_D1; ++i) {
        float delta = __fmul_rn(__fsub_rn(e[i], (float)rem0[i]), 0.25f);
        // b[D - rank] += delta ; b[D+1 - rank] -= delta   (static indexing keeps b[] in registers)
#pragma unroll
        for (int q = 0; q <= LTN_D1; ++q) {
            if (q == LTN_D - rank[i]) b[q] = __fadd_rn(b[q], delta);
            if (q == LTN_D1 - rank[i]) b[q] = __fsub_rn(b[q], delta);
        }
    }
    b[0] = __fadd_rn(b[0], __fadd_rn(1.0f, b[LTN_D1]));
#pragma unroll
    for (int r = 0; r < LTN_D1; ++r) {
#pragma unroll
        for (int i = 0; i < LTN_D; ++i) key[r][i] = rem0[i] + r - ((rank[i] > LTN_D - r) ? LTN_D1 : 0);
        bary[r] = b[r];
    }
}

// Warp-aggregated accumulation: lanes that hold the same key (vertex id) first add their K floats together
// inside the warp, then ONE lane per distinct key issues the global atomics.  Neighbouring LiDAR points land
// on the same lattice vertices (40-400 rows per vertex), so this removes most of the same-address contention
// a per-row atomicAdd has.  Returns true on the leader lane of each group, with the group's sums in v[] and
// the group size in count.  All 32 lanes must call it (key < 0 = this lane has nothing to add).
template <int K>
__device__ __forceinline__ bool ltn_warp_group_sum(int key, float (&v)[K], int& count) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned peers = __match_any_sync(0xffffffffu, key < 0 ? -1 - (int)lane : key);
    const bool leader = (int)lane == __ffs(peers) - 1;
    unsigned rest = peers & ~(1u << lane);   // the other lanes of my group, visited one per round
    count = __popc(peers);
    float mine[K];
#pragma unroll
    for (int k = 0; k < K; ++k) mine[k] = v[k];   // peers read the ORIGINAL values
    while (__any_sync(0xffffffffu, rest != 0u)) {
        const int src = rest ? __ffs(rest) - 1 : (int)lane;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float o = __shfl_sync(0xffffffffu, mine[k], src);
            if (rest) v[k] += o;
        }
        rest &= rest - 1u;
    }
    return leader && key >= 0;
}
