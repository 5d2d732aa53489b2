/*
 * latticenet_b200.h -- C ABI of libltn_b200.so, the sm_100a implementation of the permutohedral
 * lattice hot path under Temporal LatticeNet.
 *
 * What it replaces.  The reference has no FFI of its own for this path: it imports a pybind11
 * module, `latticenet` (C++/CUDA, github.com/peerschuett/lattice_net, un-vendored -- reference
 * README.md:47-48), and the Python package `latticenet_py` built on it.  Each entry point below
 * cites the reference call site whose work it performs; the Python binding that presents these
 * entry points under the reference's names lives in temporal_latticenet_b200/ (see INTEGRATION.md).
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller (torch tensors); the library allocates
 *    nothing persistent and frees nothing;  `stream` is a cudaStream_t passed as void*;
 *  - return value: 0 on success, a cudaError_t (>0) from the launch, or <0 for an argument the
 *    kernels do not support (-2 shape, -3 shared memory);
 *  - no exceptions, no host synchronisation inside any call;
 *  - sizes only the device knows: wherever a count comes with a nullable `*_dev` companion (n_dev, v_dev,
 *    vq_dev, vx_dev, vh_dev, r_dev), the kernel uses min(host bound, *dev) -- launches are sized by the host
 *    bound (a capacity), so whole frames replay from a CUDA graph without any host synchronisation;
 *  - pos_dim is 3 (4 simplex vertices per point, filter extent 9: lattice_modules.py:299,310-311);
 *  - rows are point-major: row = point*4 + r (models.py:452); indices are int32, -1 = absent.
 *
 * Hash-table storage (all allocated by the caller, `nslots` a power of two >= 2*capacity; the
 * Python host uses 4*capacity):
 *    slot_keys  uint64[nslots]  packed key (3 x 21 bit), all-ones = empty
 *    slot_ids   int32 [nslots]  vertex id of the slot, -1 until numbered
 *    slot_first int32 [nslots]  scratch: smallest row that touched a new slot in this batch
 *    keys       int32 [capacity,4]  key of vertex id (x,y,z,-(x+y+z))
 *    counters   int32 [8]       [0] vertices, [1] vertices before the last batch,
 *                               [2] dropped on overflow, [3] keys out of the packable range
 */
#ifndef LATTICENET_B200_H
#define LATTICENET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- lattice structure (latticenet.Lattice / HashTable: train_ln.py:106,239;
 *      lattice_modules.py:7-8) ------------------------------------------------------------- */

/* HashTable::clear -- begin a new sequence (models.py:287-289, reset_hashmap=True) */
int ltn_hash_clear(uint64_t* slot_keys, int* slot_ids, int* slot_first, int nslots, int* counters, void* stream);

/* Insert the 4 enclosing-simplex vertices of N points (scale s = inv_std_dev/(sigma*sqrt((i+1)(i+2))))
 * and number NEW vertices in order of first appearance (append-only ids).  Used for the coarse
 * lattices of GnReluCoarsen (models.py:182,353).  row_slot [4N] and block_sums [ceil(4N/1024)+1]
 * are scratch; row_w [4N] (nullable) receives the barycentric weights. */
int ltn_insert_points(const float* pos, int N, const int* n_dev, float sx, float sy, float sz, uint64_t* slot_keys,
                      int* slot_ids, int* slot_first, int nslots, int* counters, int* keys, int capacity, int* row_slot,
                      float* row_w, int* block_sums, void* stream);

/* DistributeLatticeModule.forward (models.py:62,297-298): rows [4N, 3+val_dim+1] = [xyz, val, bary],
 * idx [4N], w [4N]; subtract_mean != 0 also subtracts the per-vertex mean position from rows[:,0:3].
 * vert_acc [capacity,4] double scratch; afterwards vert_acc[v,3] = number of rows on vertex v. */
int ltn_distribute(const float* pos, const float* val, int N, const int* n_dev, int val_dim, float sx, float sy, float sz,
                   uint64_t* slot_keys, int* slot_ids, int* slot_first, int nslots, int* counters, int* keys,
                   int capacity, int* row_slot, int* block_sums, double* vert_acc, float* rows, int* idx, float* w,
                   int subtract_mean, void* stream);

/* rows-per-vertex as float [V] (lattice_modules.py:519-521, scatter_add of ones) */
int ltn_vertex_counts(const double* vert_acc, int V, float* counts, void* stream);

/* Neighbour table [Vq,9] (slot 2a = +o_a, 2a+1 = -o_a, 8 = centre; -1 absent).  mode 0: same level
 * (ConvLatticeModule / Im2RowLattice / Im2RowIndicesLattice, lattice_modules.py:301,304,440,573);
 * mode 1: query coarse, table fine (CoarsenLattice, models.py:353); mode 2: query fine, table
 * coarse (FinefyLattice, models.py:398).  vq_dev (nullable): device-side vertex count. */
int ltn_neighbours(const int* keys_q, int Vq, const int* vq_dev, const uint64_t* slot_keys, const int* slot_ids,
                   int nslots, int mode, int dilation, int same_table, int* out, void* stream);

/* ---- gathers / scatters ---------------------------------------------------------------------- */

/* Im2RowLattice.apply (lattice_modules.py:301): out [Vq, 9*C], zeros where absent. C % 4 == 0 */
int ltn_im2row(const float* vals, int Vvals, const int* vvals_dev, const int* nbr, int Vq, const int* vq_dev, int C,
               float* out, void* stream);
/* backward of im2row as a gather through the opposite-direction table nbrT [Vu,9] */
int ltn_row2im(const float* grad_rows, int Vrows, const int* nbrT, int Vu, int C, float* grad_vals, void* stream);

/* SplatLatticeModule (models.py:234): out [V, C+1] += w*[val,1]; caller zeroes out */
int ltn_splat(const float* val, int N, int C, const int* idx, const float* w, float* out, int V, void* stream);
/* SliceLatticeModule (models.py:233): out [N,C] = sum_r w*vals[idx]. C % 4 == 0 */
int ltn_slice(const float* vals, int V, int C, const int* idx, const float* w, int N, float* out, void* stream);
int ltn_slice_bwd(const float* grad_out, int N, int C, const int* idx, const float* w, float* grad_vals, int V,
                  void* stream);
/* GatherLattice inside SliceFastCUDALatticeModule (models.py:232,465): out [N, 4*(C+1)] */
int ltn_gather(const float* vals, int V, const int* v_dev, int C, const int* idx, const float* w, int N, const int* n_dev,
               float* out, void* stream);
int ltn_gather_bwd(const float* grad_out, int N, int C, const int* idx, const float* w, float* grad_vals, int V,
                   void* stream);
/* SliceClassifyLattice (models.py:465): out [N,K]; sliced [N,C] (nullable) saved for the weight grad */
int ltn_slice_classify(const float* vals, int V, const int* v_dev, int C, const int* idx, const float* w, const float* dw,
                       int N, const int* n_dev, const float* Wc, const float* bias, int K, float* out, float* sliced,
                       void* stream);
int ltn_slice_classify_bwd(const float* grad_logit, const float* vals, int V, int C, const int* idx, const float* w,
                           const float* dw, int N, const float* Wc, int K, float* grad_vals, float* grad_dw,
                           void* stream);
/* The per-point tail of SliceFastCUDALatticeModule (models.py:232,465) and the model's LogSoftmax (models.py:251,470) in two
 * kernels, inference: gather [N,4*9] of the bottleneck values bott [V,8] -> minus gamma*max-over-simplex + beta -> Linear
 * 36->36 (W1) -> GroupNorm(18 groups, statistics over all points; sums [18,2] double scratch) + ReLU -> Linear 36->4 (W2, b2)
 * = delta weights -> logits [N,K] = cls_bias + sum_r (w + dw) * scores[idx] (scores [V, ld_scores]: the vertices' class scores,
 * K <= 32) -> logsm [N,K] = log_softmax.  no_deform = experiment "slice_no_deform" (delta weights ignored).  logits nullable. */
int ltn_slice_head(const float* bott, int V, const int* v_dev, const float* scores, int ld_scores, const int* idx, const float* w,
                   int N, const int* n_dev, const float* gamma, const float* beta, const float* W1, const float* gn_w,
                   const float* gn_b, float gn_eps, const float* W2, const float* b2, const float* cls_bias, int K, int no_deform,
                   double* sums, float* logits, float* logsm, void* stream);

/* ---- segmented reductions, normalisation ------------------------------------------------------ */

/* torch_scatter.scatter_max(src[R,C], idx[R], dim=0) (lattice_modules.py:512): packed [V,C] u64 scratch */
int ltn_scatter_max(const float* src, const int* idx, int R, int C, int V, unsigned long long* packed, float* out,
                    long long* arg, void* stream);
/* torch_scatter.scatter_add (lattice_modules.py:497,502,520); out pre-zeroed */
int ltn_scatter_add(const float* src, const int* idx, int R, int C, float* out, int V, void* stream);
/* GroupNorm over [1,C,V] (Gn / GnRelu1x1 / GnReluConv, lattice_modules.py:75,100,436-437): sums [G,2] double */
int ltn_gn_stats(const float* x, int V, const int* v_dev, int C, int G, double* sums, void* stream);
int ltn_gn_apply(const float* x, int V, const int* v_dev, int C, int G, const double* sums, const float* gamma,
                 const float* beta, float eps, int relu, float* y, void* stream);

/* ---- fused convolution on the tensor cores (csrc/ltn_conv.cu) ------------------------------------- */

/* GroupNorm(+ReLU) backward of the modules above (train_ln.py:229-231 loss.backward()): y (nullable) = forward output when
 * ReLU was fused; sums [G,2] = forward statistics; chan [C,2] double returns (grad beta, grad gamma) per channel; gx [V,C] */
int ltn_gn_bwd(const float* x, const float* gy, const float* y, int V, int C, int G, const double* sums, const float* gamma,
               float eps, double* chan, float* gx, void* stream);

/* Weight gradient of ConvLatticeModule / CoarsenLattice / FinefyLattice / the 1x1 layers on the tensor cores (train_ln.py:229-231
 * loss.backward(); SURVEY.md 8b `ltn_conv_bwd_weight`): dW [S*C, F] += sum_v act[nbr[v,s], :]^T dy[v, :], dW zeroed by the caller.
 * act [Vx, C]: the layer's input activations (after GroupNorm / ReLU); dy [Vq, F]; nbr [Vq, S] or NULL (S = 1, row v).
 * C % 4 == 0, F % 16 == 0, 16 <= F <= 256.  fp32 parity: 3-pass tf32 hi/lo split, both operands MN-major in shared memory. */
int ltn_conv_bwd_weight(const float* act, int Vx, const int* nbr, int Vq, int C, int S, const float* dy, int F, float* dW,
                        void* stream);

/* ConvLatticeModule / CoarsenLattice / FinefyLattice (lattice_modules.py:440,573; models.py:353,398) and
 * the dense layers around them (GnRelu1x1, Conv1x1, nn.Linear, GRU/LSTM gate GEMMs) WITHOUT the [V,9C]
 * im2row buffer:  out[v,f] = sum_{s<S} sum_c act(x[nbr[v,s],c]) * W[s*C+c, f] (+bias[f]) (+res[v,f]),
 * act(t) = relu?(t*scale[c] + shift[c]) on present rows, 0 for absent ones (nbr < 0 or >= Vx).
 * The per-channel affine is either explicit (a_scale/a_shift) or the GroupNorm of x folded from its
 * statistics gn_sums [G,2] (ltn_gn_stats or a previous call's out_sums), gn_gamma/gn_beta [C].
 * out_sums (nullable, caller-zeroed) [out_groups,2]: GroupNorm statistics of the OUTPUT, accumulated in
 * the epilogue for the next layer.  wt_hi / wt_lo: ltn_split_tf32 copies of the weight, [F, S*C]
 * (wt_lo may be NULL when passes == 1).  nbr == NULL means S = 1, identity rows.  vx_dev / vq_dev
 * (nullable): device-side row counts.  C % 32 == 0 (C <= 256 with an affine), F % 8 == 0, ldo % 4 == 0.
 * passes = 3: fp32-parity (hi/lo split, three tf32 tensor-core passes); passes = 1: plain TF32. */
int ltn_conv_tc(const float* x, int Vx, const int* vx_dev, const int* nbr, int Vq, const int* vq_dev, int C, int S,
                const float* wt_hi, const float* wt_lo, int F, const float* a_scale, const float* a_shift,
                const double* gn_sums, const float* gn_gamma, const float* gn_beta, float gn_eps, int gn_groups, int relu,
                const float* bias, const float* res, float* out, int ldo, double* out_sums, int out_groups, int passes,
                void* stream);
/* weight -> K-major [F,K] tf32 split: hi = round-to-tf32(w), lo = w - hi.  transposed_in = 0: w is the
 * reference's conv layout [K,F] (lattice_modules.py:291); 1: w is [F,K] (nn.Linear) */
int ltn_split_tf32(const float* w, int K, int F, int transposed_in, float* wt_hi, float* wt_lo, void* stream);
/* The same operation with fp16 hi/lo operands: the split keeps the same 11 + 11 significant bits per operand as
 * the tf32 one in half the bytes (half the shared-memory traffic, twice the tensor rate).  fp16's exponent range
 * is covered by exact power-of-two scaling: the weight copies hold w * 2^w_log2 (ltn_split_f16, w_log2 chosen by
 * the caller from max|w|), activations are staged as act(x) * 2^a_log2, the epilogue multiplies by
 * 2^-(w_log2 + a_log2).  *flag (int32, nullable) is OR-ed with 1 when a staged magnitude reaches 65504: the
 * result is then unusable and the caller redoes the work with ltn_conv_tc.  C % 64 == 0; wt_hi / wt_lo are
 * [F, S*C] fp16. */
int ltn_conv_tc_f16(const float* x, int Vx, const int* vx_dev, const int* nbr, int Vq, const int* vq_dev, int C, int S,
                    const void* wt_hi, const void* wt_lo, int w_log2, int a_log2, int F, const float* a_scale,
                    const float* a_shift, const double* gn_sums, const float* gn_gamma, const float* gn_beta, float gn_eps,
                    int gn_groups, int relu, const float* bias, const float* res, float* out, int ldo, double* out_sums,
                    int out_groups, int* flag, void* stream);
int ltn_split_f16(const float* w, int K, int F, int transposed_in, int w_log2, void* wt_hi, void* wt_lo, void* stream);
/* ltn_conv_tc_f16 for nb <= 8 independent problems that share the weights -- the same layer of several windows in
 * flight (SURVEY.md 8b "Threading: kernels must tolerate B lattices per launch").  One persistent launch walks the
 * tiles of all problems; x / Vx / vx_dev / nbr / Vq / vq_dev / gn_sums / res / out / out_sums / flag are HOST arrays of nb
 * entries (device pointers resp. ints), everything else is shared.  Results are bit-identical to nb calls of
 * ltn_conv_tc_f16 (out; the statistics differ in summation order only). */
int ltn_conv_tc_f16_batched(int nb, const float* const* x, const int* Vx, const int* const* vx_dev, const int* const* nbr,
                            const int* Vq, const int* const* vq_dev, int C, int S, const void* wt_hi, const void* wt_lo, int w_log2,
                            int a_log2, int F, const double* const* gn_sums, const float* gn_gamma, const float* gn_beta,
                            float gn_eps, int gn_groups, int relu, const float* bias, const float* const* res, float* const* out,
                            int ldo, double* const* out_sums, int out_groups, int* const* flag, int staged, void* stream);
/* A-operand pre-staging for the gathering layers (ConvLatticeModule / coarsen / finefy, S = 9): one pass per layer input does
 * what the gather loop would otherwise repeat for each of a row's nine uses -- folded GroupNorm, ReLU, x 2^a_log2, fp16 range
 * check, hi / lo split -- and writes, per 4 channels, the 16-byte quad {hi01, hi23, lo01, lo23} into out[b] ([Vx_b, C], the
 * size of x[b]).  ltn_conv_tc_f16_batched(..., staged = 1) takes out[] in place of x[] (gn_sums then unused).  Bit-identical
 * results. */
int ltn_stage_a_batched(int nb, const float* const* x, const int* Vx, const int* const* vx_dev, int C, const double* const* gn_sums,
                        const float* gn_gamma, const float* gn_beta, float gn_eps, int gn_groups, int relu, int a_log2,
                        void* const* out, int* const* flag, void* stream);
/* tracing of the following batched launches: launch i < nr_records writes record i of buf ([nr_records, 2 + 2*148] u64,
 * zeroed by the caller): live rows, tiles, then (entry, exit) globaltimer ns per CTA; the record address is baked into the
 * launch (captured launches keep writing on every replay).  NULL switches it off.  Returns the records used since the
 * previous call. */
int ltn_conv_batched_trace(unsigned long long* buf, int nr_records);
/* role timeline of CTA 0 of the following batched launches: buf [64 tiles][16] u64 globaltimer stamps (see the source for
 * the slots); NULL switches it off */
int ltn_conv_batched_detail(unsigned long long* buf);
/* phase tracing of the following ltn_conv_tc* launches (NULL switches it off): buf receives 8 globaltimer stamps (ns)
 * per CTA in launch-grid order -- entry, set-up done, first operands staged, producers done, accumulator complete,
 * epilogue stores issued, teardown, (unused) */
int ltn_conv_trace(unsigned long long* buf);

/* PointNetSeqModule front end (lattice_modules.py:448-530) fused: MLP 4->16->32->64 per distributed row,
 * segmented max per vertex (+arg-max), barycentric weight of the winning row (quirk Q3), concatenation
 * and the min-4-rows mask.  rows [R, 5] = ltn_distribute's output, idx [R]; w*/b*: nn.Linear weights
 * [out,in] / biases; packed [V,64] u64 scratch; vert_acc: ltn_distribute's accumulator; min_rows = 4
 * (0 = no mask, the early max-pool variant); out [V,128].  r_dev / v_dev (nullable): device-side sizes. */
int ltn_pointnet(const float* rows, int width, const int* idx, int R, const int* r_dev, const float* w1, const float* b1,
                 const float* w2, const float* b2, const float* w3, const float* b3, int V, const int* v_dev,
                 unsigned long long* packed, const double* vert_acc, int min_rows, float* out, void* stream);

/* The same with the 32 -> 64 layer on the tensor cores: fp16 hi/lo operands (three passes, fp32-class results as in
 * ltn_conv_tc_f16), activations staged as relu(h2) * 2^a_log2 straight into tensor memory.  w12_host: HOST pointer to
 * layers 1-2 in nn.Linear layout, concatenated (w1 [16,4], b1 [16], w2 [32,16], b2 [32] = 624 floats): they travel as
 * kernel parameters and are read as constant-bank operands.  *flag (int32) is OR-ed with 1 when an activation leaves
 * the fp16 range: the caller then redoes the work with ltn_pointnet. */
int ltn_pointnet_tc(const float* rows, int width, const int* idx, int R, const int* r_dev, const float* w12_host,
                    const float* w3, const float* b3, int V, const int* v_dev, unsigned long long* packed,
                    const double* vert_acc, int min_rows, float* out, int a_log2, int* flag, void* stream);

/* phase tracing of ltn_pointnet_tc (NULL switches it off): 8 clock64 stamps of one tile of block 0 (loop top, operands in
 * tensor memory, vertices numbered, accumulator read, block max done, arg-max done, flushed, next tile) + distinct vertices */
int ltn_pointnet_trace(long long* buf);

/* ---- temporal fusion (seq_lattice/lattice_modules.py:17-339) ------------------------------------ */

/* GRUModule.forward pointwise stage (lattice_modules.py:58-63); rows >= Vh are the zero padding */
int ltn_gru_pointwise(const float* gi, const float* gh, const float* h, const float* b_hh, int V, int Vh, const int* v_dev,
                      const int* vh_dev, int C, float* out, void* stream);
/* the same stage, leaving the GroupNorm statistics of its output behind for the layer that normalises it next (sums
 * [groups,2] double, zeroed by the caller): GnReluConv / GnReluCoarsen after a GRU fusion point, models.py:340-353,435 */
int ltn_gru_pointwise_stats(const float* gi, const float* gh, const float* h, const float* b_hh, int V, int Vh, const int* v_dev,
                            const int* vh_dev, int C, float* out, double* sums, int groups, void* stream);
/* LSTMModule.forward pointwise stage with c_prev = 0 (lattice_modules.py:32-37) */
int ltn_lstm_pointwise(const float* gi, const float* gh, const float* b_hh, int V, int Vh, const int* v_dev,
                       const int* vh_dev, int C, float* out, void* stream);
/* CustomKernelConvLatticeIm2RowModule.forward (lattice_modules.py:282-339) fused:
 * out [V,C], weights_out [V,9] (nullable) */
int ltn_aflow(const float* lv, const float* h, int V, int Vh, const int* v_dev, const int* vh_dev, int C, const int* nbr,
              const float* alpha, const float* beta, const float* bias, float pad_value, int use_center, float* out,
              float* weights_out, void* stream);

/* ---- window assembly (dataloader/kitti_dataloader.py:129-171) ------------------------------------- */

/* One SemanticKITTI scan from its wire format to the model's inputs on the device: raw [N,4] float32 = the .bin payload
 * (x, y, z, reflectance); mats = HOST pointer to nr_mats (1..3) row-major 4x4 float64 matrices applied in order to
 * [x,y,z,1] in float64 (velo -> world, world -> first scan of the window, -90 degrees about x), then divided by w and
 * rounded to float32 (kitti_dataloader.py:160-171, DataTransformer.py:88-91); pos [N,3], val [N,1] = reflectance. */
int ltn_assemble_scan(const float* raw, int N, const double* mats, int nr_mats, float* pos, float* val, void* stream);

/* Scores.accumulate_scores (callbacks/scores.py:13-30; train_ln.py:219 via StateCallback): arg-max of scores [N,K],
 * K x K confusion counts of one cloud folded into per-class intersection / union (classes present in gt only,
 * `unlabeled` skipped).  conf [K,K] u64 scratch (zeroed by the caller once), inter / uni [K] int64 accumulators. */
int ltn_scores_accumulate(const float* scores, const long long* gt, int N, const int* n_dev, int K, int unlabeled,
                          unsigned long long* conf, long long* inter, long long* uni, void* stream);

/* library version / build info */
int ltn_version(void);
/* dst[0] = a, dst[1] = b on the stream (per-frame point counts of the graph engine; values travel in the launch) */
int ltn_set_int2(int* dst, int a, int b, void* stream);
/* kernels launched by this library since it was loaded, modulo 2^31 (bench.py: gpu_launches) */
int ltn_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
