/*
 * lattice_oracle.c -- scalar CPU restatement of the permutohedral-lattice core that sits under
 * Temporal LatticeNet.  TEST INFRASTRUCTURE ONLY: it is the checker for the CUDA path (tests/,
 * __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs).  Nothing under
 * temporal_latticenet_b200/ may import, link or call it.
 *
 * PARITY UNPINNED at the `latticenet` boundary: the reference obtains this arithmetic from an
 * un-vendored, un-pinned third-party module (github.com/peerschuett/lattice_net, HEAD, cloned at
 * docker run time -- /root/reference/README.md:47-48) and ships no tests, golden vectors or
 * fixtures for it.  What is restated here is
 *   (1) the published algorithm of Adams, Baek, Davis, "Fast High-Dimensional Filtering Using the
 *       Permutohedral Lattice" (Eurographics 2010): elevate -> closest remainder-0 point -> rank ->
 *       barycentric weights -> the d+1 simplex vertex keys, and
 *   (2) the contracts visible at the reference's own call sites:
 *       seq_lattice/models.py:297-298   distribute(ls, positions, values, reset_hashmap)
 *                                        -> distributed [4N, 3+val_dim+1], indices [4N], weights [4N]
 *       seq_lattice/models.py:452       point-major rows, d+1 = 4 consecutive rows per point
 *       seq_lattice/lattice_modules.py:299,310-311,316-320
 *                                        filter_extent = 9, slot 8 = centre, -1 = neighbour absent
 *       seq_lattice/lattice_modules.py:448-452,477-480
 *                                        last column = barycentric weight, indices may be -1
 *       seq_lattice/models.py:287-289   append-only vertex ids across the frames of a window
 *       seq_lattice/models.py:465       slice_classify(lv, ls, positions, indices, weights)
 * Conventions that cannot be verified are fixed here and documented in
 * DESIGN.md; the CUDA path follows the same conventions.
 *
 * pos_dim d = 3 throughout (the reference hard-codes 2(d+1)+1 = 9, lattice_modules.py:310-311).
 * All float arithmetic is plain IEEE fp32, one rounding per operation (build with
 * -ffp-contract=off); the CUDA path uses the matching __f*_rn intrinsics so keys AND barycentric
 * weights are bit-identical.
 *
 * Vertex ids are assigned in insertion order: row order p*4+r, first occurrence wins.  That is the
 * order a single-threaded run of the reference's `insert` would produce; the CUDA path reproduces
 * it deterministically (first-row ranking) so ids compare without canonicalisation as well.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define D 3
#define D1 4
#define FEXT 9 /* 2*(D+1)+1, seq_lattice/lattice_modules.py:299 */

typedef struct OrcTable {
    int cap;     /* max number of vertices (cfg lattice_gpu.hash_table_capacity) */
    int size;    /* vertices stored so far == next id */
    int nslots;  /* open-addressing slots, power of two >= 2*cap */
    int *slots;  /* slot -> vertex id, -1 empty */
    int *keys;   /* [cap,3] first d coordinates; coordinate d is minus their sum */
} OrcTable;

static unsigned orc_hash(const int *k) {
    /* hash recalled from the upstream project: k = (k + key[i]) * 2531011 ; any hash would do,
       vertex ids do not depend on it */
    unsigned h = 0;
    for (int i = 0; i < D; i++) { h += (unsigned)k[i]; h *= 2531011u; }
    return h;
}

OrcTable *orc_table_create(int cap) {
    OrcTable *t = (OrcTable *)calloc(1, sizeof(OrcTable));
    t->cap = cap;
    int n = 8;
    while (n < 2 * cap) n <<= 1;
    t->nslots = n;
    t->slots = (int *)malloc(sizeof(int) * (size_t)n);
    t->keys = (int *)calloc((size_t)(cap > 0 ? cap : 1) * D, sizeof(int));
    memset(t->slots, 0xff, sizeof(int) * (size_t)n);
    return t;
}

void orc_table_free(OrcTable *t) {
    if (!t) return;
    free(t->slots);
    free(t->keys);
    free(t);
}

void orc_table_clear(OrcTable *t) {
    memset(t->slots, 0xff, sizeof(int) * (size_t)t->nslots);
    t->size = 0;
}

int orc_table_size(const OrcTable *t) { return t->size; }
int orc_table_capacity(const OrcTable *t) { return t->cap; }
const int *orc_table_keys(const OrcTable *t) { return t->keys; }

int orc_table_find(const OrcTable *t, const int *key) {
    unsigned m = (unsigned)t->nslots - 1u;
    unsigned s = orc_hash(key) & m;
    for (;;) {
        int id = t->slots[s];
        if (id < 0) return -1;
        const int *k = t->keys + (size_t)id * D;
        if (k[0] == key[0] && k[1] == key[1] && k[2] == key[2]) return id;
        s = (s + 1u) & m;
    }
}

/* returns the vertex id, inserting if new; -1 when the table already holds `cap` vertices
   (convention U4: overflow yields index -1, lattice_modules.py:479-480 tolerates it) */
int orc_table_insert(OrcTable *t, const int *key) {
    unsigned m = (unsigned)t->nslots - 1u;
    unsigned s = orc_hash(key) & m;
    for (;;) {
        int id = t->slots[s];
        if (id < 0) {
            if (t->size >= t->cap) return -1;
            id = t->size++;
            t->slots[s] = id;
            memcpy(t->keys + (size_t)id * D, key, sizeof(int) * D);
            return id;
        }
        const int *k = t->keys + (size_t)id * D;
        if (k[0] == key[0] && k[1] == key[1] && k[2] == key[2]) return id;
        s = (s + 1u) & m;
    }
}

/* SURVEY.md appendix B.1-B.3.  scale[i] = inv_std_dev / (sigma_i * sqrt((i+1)(i+2))) is computed by
   the caller in double and rounded once to fp32 so that oracle and CUDA consume the same bits. */
void orc_simplex(const float *p, const float *scale, int *keys /*[4][3]*/, float *bary /*[4]*/) {
    float cf[D], e[D1];
    for (int i = 0; i < D; i++) cf[i] = p[i] * scale[i];
    float sm = 0.0f;
    for (int i = D; i > 0; i--) {
        float t = (float)i * cf[i - 1];
        e[i] = sm - t;
        sm = sm + cf[i - 1];
    }
    e[0] = sm;

    int rem0[D1], rank[D1] = {0, 0, 0, 0}, sum = 0;
    for (int i = 0; i < D1; i++) {
        float v = e[i] * 0.25f;
        float up = ceilf(v) * 4.0f;
        float dn = floorf(v) * 4.0f;
        rem0[i] = ((up - e[i]) < (e[i] - dn)) ? (int)up : (int)dn;
        sum += rem0[i];
    }
    sum /= D1; /* exact: every rem0 is a multiple of d+1 */
    for (int i = 0; i < D; i++)
        for (int j = i + 1; j < D1; j++) {
            float di = e[i] - (float)rem0[i];
            float dj = e[j] - (float)rem0[j];
            if (di < dj) rank[i]++; else rank[j]++;
        }
    for (int i = 0; i < D1; i++) {
        rank[i] += sum;
        if (rank[i] < 0) { rank[i] += D1; rem0[i] += D1; }
        else if (rank[i] > D) { rank[i] -= D1; rem0[i] -= D1; }
    }
    float b[D1 + 1] = {0, 0, 0, 0, 0};
    for (int i = 0; i < D1; i++) {
        float delta = (e[i] - (float)rem0[i]) * 0.25f;
        b[D - rank[i]] = b[D - rank[i]] + delta;
        b[D1 - rank[i]] = b[D1 - rank[i]] - delta;
    }
    b[0] = b[0] + (1.0f + b[D1]);
    for (int r = 0; r < D1; r++) {
        for (int i = 0; i < D; i++)
            keys[r * D + i] = rem0[i] + r - ((rank[i] > D - r) ? D1 : 0);
        bary[r] = b[r];
    }
}

/* models.py:297-298 / SURVEY B.5.  rows [4N, 3+val_dim+1] = [p_xyz, val.., bary]; idx [4N]; w [4N].
   The per-vertex mean subtraction is a separate step (orc_local_mean_sub). */
void orc_distribute(OrcTable *t, const float *pos, const float *val, int N, int val_dim,
                    const float *scale, float *rows, int *idx, float *w) {
    const int width = D + val_dim + 1;
    for (int p = 0; p < N; p++) {
        int keys[D1 * D];
        float bary[D1];
        orc_simplex(pos + (size_t)p * D, scale, keys, bary);
        for (int r = 0; r < D1; r++) {
            size_t row = (size_t)p * D1 + r;
            int id = orc_table_insert(t, keys + r * D);
            float *o = rows + row * width;
            for (int i = 0; i < D; i++) o[i] = pos[(size_t)p * D + i];
            for (int i = 0; i < val_dim; i++) o[D + i] = val[(size_t)p * val_dim + i];
            o[D + val_dim] = bary[r];
            idx[row] = id;
            w[row] = bary[r];
        }
    }
}

/* coarse-vertex creation (convention U3): the frame's points are inserted at the coarser sigma
   into a persistent table; same key math as distribute, no per-point outputs. */
void orc_insert_points(OrcTable *t, const float *pos, int N, const float *scale) {
    for (int p = 0; p < N; p++) {
        int keys[D1 * D];
        float bary[D1];
        orc_simplex(pos + (size_t)p * D, scale, keys, bary);
        for (int r = 0; r < D1; r++) (void)orc_table_insert(t, keys + r * D);
    }
}

/* rows[:, 0:3] -= mean over all rows with the same vertex id (ids < 0 count as 0,
   lattice_modules.py:479-480 convention).  Accumulates in double, rounds once: the CUDA path
   accumulates fp32 partials and is compared within tolerance. */
void orc_local_mean_sub(float *rows, const int *idx, int R, int width, int V) {
    double *acc = (double *)calloc((size_t)(V > 0 ? V : 1) * 4, sizeof(double));
    for (int r = 0; r < R; r++) {
        int id = idx[r] < 0 ? 0 : idx[r];
        for (int i = 0; i < D; i++) acc[(size_t)id * 4 + i] += rows[(size_t)r * width + i];
        acc[(size_t)id * 4 + 3] += 1.0;
    }
    for (int r = 0; r < R; r++) {
        int id = idx[r] < 0 ? 0 : idx[r];
        for (int i = 0; i < D; i++) {
            float mean = (float)(acc[(size_t)id * 4 + i] / acc[(size_t)id * 4 + 3]);
            rows[(size_t)r * width + i] = rows[(size_t)r * width + i] - mean;
        }
    }
    free(acc);
}

/* Neighbour table [Vq, 9] of ids in `nbr` for the first Vq vertices of `query`.
   mode 0: same resolution               tap = k + dil*o
   mode 1: query coarse, nbr fine        tap = 2k + dil*o          (coarsen, models.py:353)
   mode 2: query fine,  nbr coarse       tap = (k + dil*o)/2, valid only if all 4 coords even
                                                                    (finefy, models.py:398)
   slot 2a = +o_a, slot 2a+1 = -o_a, o_a = (1,1,1,1) with -d at axis a; slot 8 = centre.
   -1 = absent (lattice_modules.py:318). */
void orc_neighbours(const OrcTable *query, int Vq, const OrcTable *nbr, int mode, int dil, int *out) {
    for (int v = 0; v < Vq; v++) {
        const int *k3 = query->keys + (size_t)v * D;
        int k[D1] = {k3[0], k3[1], k3[2], -(k3[0] + k3[1] + k3[2])};
        for (int s = 0; s < FEXT; s++) {
            int t[D1];
            for (int i = 0; i < D1; i++) {
                int o = 0;
                if (s < 2 * D1) {
                    int a = s >> 1;
                    o = (i == a) ? -D : 1;
                    if (s & 1) o = -o;
                    o *= dil;
                }
                t[i] = (mode == 1) ? 2 * k[i] + o : k[i] + o;
            }
            int id = -1;
            if (mode == 2) {
                if (((t[0] | t[1] | t[2] | t[3]) & 1) == 0) {
                    int h[D] = {t[0] / 2, t[1] / 2, t[2] / 2};
                    id = orc_table_find(nbr, h);
                }
            } else {
                id = orc_table_find(nbr, t);
            }
            out[(size_t)v * FEXT + s] = id;
        }
    }
}

/* im2row from a neighbour table: out[v, s*C + c] = vals[nbr[v,s], c], zeros when the neighbour
   is absent or has no value row yet (id >= Vvals) -- lattice_modules.py:301,316. */
void orc_im2row(const int *nbr, int Vq, const float *vals, int Vvals, int C, float *out) {
    for (int v = 0; v < Vq; v++)
        for (int s = 0; s < FEXT; s++) {
            int id = nbr[(size_t)v * FEXT + s];
            float *o = out + ((size_t)v * FEXT + s) * C;
            if (id < 0 || id >= Vvals) memset(o, 0, sizeof(float) * (size_t)C);
            else memcpy(o, vals + (size_t)id * C, sizeof(float) * (size_t)C);
        }
}

/* splat: out[idx, 0:C] += w * val_p ; out[idx, C] += w  (homogeneous coordinate).
   Sequential row order. */
void orc_splat(const float *val, int N, int C, const int *idx, const float *w, float *out, int V) {
    for (int p = 0; p < N; p++)
        for (int r = 0; r < D1; r++) {
            int id = idx[(size_t)p * D1 + r];
            if (id < 0 || id >= V) continue;
            float ww = w[(size_t)p * D1 + r];
            float *o = out + (size_t)id * (C + 1);
            for (int c = 0; c < C; c++) o[c] = o[c] + ww * val[(size_t)p * C + c];
            o[C] = o[C] + ww;
        }
}

/* slice: out[p,:] = sum_r w[p,r] * vals[idx[p,r],:], r = 0..3 in order */
void orc_slice(const float *vals, int V, int C, const int *idx, const float *w, int N, float *out) {
    for (int p = 0; p < N; p++) {
        float *o = out + (size_t)p * C;
        for (int c = 0; c < C; c++) o[c] = 0.0f;
        for (int r = 0; r < D1; r++) {
            int id = idx[(size_t)p * D1 + r];
            if (id < 0 || id >= V) continue;
            float ww = w[(size_t)p * D1 + r];
            for (int c = 0; c < C; c++) o[c] = o[c] + ww * vals[(size_t)id * C + c];
        }
    }
}

/* gather: g[p, r*(C+1)+c] = w*vals[id,c]; g[p, r*(C+1)+C] = w; absent -> zeros */
void orc_gather(const float *vals, int V, int C, const int *idx, const float *w, int N, float *out) {
    for (int p = 0; p < N; p++)
        for (int r = 0; r < D1; r++) {
            int id = idx[(size_t)p * D1 + r];
            float *o = out + ((size_t)p * D1 + r) * (C + 1);
            if (id < 0 || id >= V) { memset(o, 0, sizeof(float) * (size_t)(C + 1)); continue; }
            float ww = w[(size_t)p * D1 + r];
            for (int c = 0; c < C; c++) o[c] = ww * vals[(size_t)id * C + c];
            o[C] = ww;
        }
}

/* slice_classify: logit[p,k] = b[k] + sum_c W[k,c] * sum_r (w+dw)[p,r]*vals[id,c] */
void orc_slice_classify(const float *vals, int V, int C, const int *idx, const float *w,
                        const float *dw, int N, const float *W, const float *b, int K, float *out) {
    float *s = (float *)malloc(sizeof(float) * (size_t)C);
    for (int p = 0; p < N; p++) {
        for (int c = 0; c < C; c++) s[c] = 0.0f;
        for (int r = 0; r < D1; r++) {
            int id = idx[(size_t)p * D1 + r];
            if (id < 0 || id >= V) continue;
            float ww = w[(size_t)p * D1 + r] + dw[(size_t)p * D1 + r];
            for (int c = 0; c < C; c++) s[c] = s[c] + ww * vals[(size_t)id * C + c];
        }
        for (int k = 0; k < K; k++) {
            float acc = 0.0f;
            for (int c = 0; c < C; c++) acc = acc + W[(size_t)k * C + c] * s[c];
            out[(size_t)p * K + k] = acc + b[k];
        }
    }
    free(s);
}
