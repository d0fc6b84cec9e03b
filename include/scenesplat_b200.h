/* scenesplat_b200 -- C-ABI of the B200 (sm_100a) kernels behind SceneSplat's PTv3 3DGS encoder hot path.
 *
 * The reference (zenghjian/SceneSplat) has no C/FFI plugin interface for this path: its boundary is
 * Python (registries + the `Point` dict) over three third-party operator
 * libraries (spconv, torch_scatter, flash_attn) and ATen built-ins.  The native convention of the
 * reference's OWN extensions is free functions taking raw device pointers + ints, caller-allocated
 * outputs (libs/pointops/src/pointops_api.cpp:15-32, libs/pointops/functions/query.py:7-25).  This
 * header follows that convention without torch types in the signatures:
 *
 *   - every pointer is DEVICE memory owned by the caller (torch's caching allocator in practice),
 *     unless a parameter name ends in `_host`;
 *   - every function is asynchronous on `stream` (a cudaStream_t passed as void*), never allocates,
 *     and takes scratch from a caller-provided workspace sized by the matching *_workspace_bytes();
 *   - return value: 0 = ok, -1 = bad arguments, otherwise a cudaError_t;
 *   - no global state: thread-safe per stream.
 *
 * dtype flags: *_is_bf16 = 1 -> __nv_bfloat16 elements, 0 -> float.
 * order ids: 0 = "z", 1 = "z-trans", 2 = "hilbert", 3 = "hilbert-trans".
 */
#ifndef SCENESPLAT_B200_H_
#define SCENESPLAT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------
 * Serialization.  Replaces Point.serialization (pointcept/models/utils/structure.py:47-102):
 * encode x4 (serialization/default.py:8-24) + torch.argsort + scatter_ inverse, and offset2batch
 * (pointcept/models/utils/misc.py:19-24). */

/* max over all grid coordinates -> *out_max_dev (structure.py:66: depth = bit_length(max)). */
int ss_coord_max(const void* grid_coord, int coord_is_int32, int64_t n, int64_t* out_max_dev, void* stream);

size_t ss_serialize_workspace_bytes(int64_t n, int rows, int depth, int n_batch);

/* grid_coord [n,3] int64 (or int32), offset [n_batch] cumulative counts.  rows <= 4 output rows,
 * row r encoded with order_ids_host[r] (the shuffle_orders permutation is applied by the caller by
 * permuting the ids).  Outputs: batch_out [n] (nullable), code/order/inverse [rows, n] int64.
 * Ties (duplicate voxels inside one batch item) are ordered by original index (stable). */
int ss_serialize(const void* grid_coord, int coord_is_int32, const int64_t* offset, int n_batch, int64_t n, int depth,
                 int rows, const int* order_ids_host, int64_t* batch_out, int64_t* code, int64_t* order,
                 int64_t* inverse, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * GridSample.  Replaces GridSample.__call__ (pointcept/datasets/transform.py:1211-1330) and
 * fnv_hash_vec / ravel_hash_vec (:1384-1416).  hash_type: 0 = fnv, 1 = ravel. */
size_t ss_gridsample_workspace_bytes(int64_t n);

/* coord [n,3] f32.  Outputs: idx_sort [n] (points sorted by voxel hash, stable), inverse [n] (voxel rank
 * of every raw point), start [n+1] (first sorted position of every voxel, start[M] = n), *m_dev = M,
 * min_coord_dev [3] (voxel-index minimum that was subtracted). */
int ss_gridsample_index(const float* coord, int64_t n, double grid_size, int hash_type, int64_t* idx_sort,
                        int64_t* inverse, int64_t* start, int64_t* m_dev, int64_t* min_coord_dev, void* workspace,
                        size_t workspace_bytes, void* stream);

/* Representative of every voxel: member (rnd[v] % count[v]) when rnd != NULL (train mode,
 * transform.py:1264-1268) else member (frag % count[v]) (test-mode fragment `frag`, :1304-1306).
 * Outputs sized for the upper bound n: idx_unique [M], grid_coord_out [M,3] (nullable), count_out [M] (nullable). */
int ss_gridsample_select(const float* coord, int64_t n, double grid_size, const int64_t* min_coord_dev,
                         const int64_t* idx_sort, const int64_t* start, const int64_t* m_dev, const int64_t* rnd,
                         int64_t frag, int64_t* idx_unique, int64_t* grid_coord_out, int64_t* count_out, void* stream);

/* dst[v,:] = src[idx[v],:], rows of row_bytes; row count = *count_dev when non-NULL else count. */
int ss_gather_rows(const void* src, int64_t row_bytes, const int64_t* idx, const int64_t* count_dev, int64_t count,
                   void* dst, void* stream);

/* ------------------------------------------------------------------------------------------------
 * SerializedPooling / SerializedUnpooling (point_transformer_v3m1_base.py:371-444, 471-482). */
size_t ss_pool_workspace_bytes(int64_t n);

/* Parent code/order [k,n].  Clusters = runs of (code[0] >> 3*pooling_depth) along order[0].
 * Child row r' is derived from parent row src_row_host[r'] (the child's shuffle permutation).
 * Child arrays have row stride m_cap (>= n).  Outputs: cluster [n] (= pooling_inverse),
 * seg_start [n+1] (CSR pointer into order[0]), head [m_cap] (one member per cluster), *m_dev = M,
 * child code/order/inverse [k,m_cap], child grid_coord [m_cap,3] and batch [m_cap] (nullable, need the
 * parent's grid_coord / batch). */
int ss_pool_index(const int64_t* code, const int64_t* order, const int64_t* grid_coord, const int64_t* batch, int64_t n,
                  int k, int pooling_depth, const int* src_row_host, int64_t m_cap, int64_t* cluster, int64_t* seg_start,
                  int64_t* head, int64_t* m_dev, int64_t* child_code, int64_t* child_order, int64_t* child_inverse,
                  int64_t* child_grid_coord, int64_t* child_batch, void* workspace, size_t workspace_bytes,
                  void* stream);

/* torch_scatter.segment_csr(src[order], seg_start, reduce) (+ optional folded-BN affine and GELU).
 * reduce: 0 sum, 1 mean, 2 max, 3 min.  act: 0 none, 1 GELU(erf).  Row count = *m_dev if non-NULL else m.
 * order == NULL: rows in place (plain segment_csr(src, seg_start)); empty segments give 0 like torch_scatter. */
int ss_segment_reduce(const void* src, int src_is_bf16, const int64_t* order, const int64_t* seg_start,
                      const int64_t* m_dev, int64_t m, int channels, int reduce, const float* scale, const float* shift,
                      int act, void* out, int out_is_bf16, void* stream);

/* One launch for what SerializedPooling does after the index build (point_transformer_v3m1_base.py:400-404):
 * out = segment_csr(src[order], seg_start, reduce) (+ folded-BN affine, GELU) AND coord_out =
 * segment_csr(coord[order], seg_start, "mean") for the fp32 [n,3] coordinates, over the same segments. */
int ss_pool_reduce(const void* src, int src_is_bf16, const float* coord, const int64_t* order, const int64_t* seg_start,
                   int64_t m, int channels, int reduce, const float* scale, const float* shift, int act, void* out,
                   int out_is_bf16, float* coord_out, void* stream);

/* out[i,:] = f_a(a[i,:]) + f_b(b[cluster[i],:]); f = optional affine + activation.  out_a (nullable)
 * receives f_a(a) alone (what the reference's stale sparse_conv_feat holds after unpooling).
 * out_flags: bit 0 = out is bf16 (else fp32), bit 1 = out_a is bf16 even when out is fp32. */
int ss_unpool_gather_add(const void* a, const void* b, int in_is_bf16, const int64_t* cluster, int64_t n, int channels,
                         const float* scale_a, const float* shift_a, const float* scale_b, const float* shift_b, int act,
                         void* out, void* out_a, int out_flags, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Submanifold convolution (replaces spconv.SubMConv3d, call sites
 * point_transformer_v3m1_base.py:277-284 and :499-506). */
size_t ss_kmap_workspace_bytes(int64_t n, int k);

/* Neighbour table nbr [k^3, n] int32 (tap-major, -1 = inactive) from one serialization row
 * (code_row [n], encoded with order_id at `depth`; order_row [n] = its sorting permutation, part of the contract -- the
 * pair lists walk it -- but not read by the build since it keeps the codes in a transient open-addressing table inside
 * `workspace`: 16 B x 2^ceil(log2(2 n)) slots).  tap t = (i*k+j)*k+l <-> (i-r, j-r, l-r).  Equal codes (duplicated
 * voxels) resolve to the smallest voxel index.  tap_count_dev [k^3] int64 receives the number of active pairs per tap. */
int ss_kmap_build(const void* grid_coord, int coord_is_int32, const int64_t* batch, const int64_t* code_row,
                  const int64_t* order_row, int64_t n, int depth, int order_id, int k, int32_t* nbr,
                  int64_t* tap_count_dev, void* workspace, size_t workspace_bytes, void* stream);

/* nbr_to [k_to^3, n] / count_to [k_to^3] = the rows of the k_from^3 map of the SAME voxel set that belong to the centred
 * k_to^3 window (k_from - k_to even and positive): the level-0 3^3 map of the xCPE convs from the stem's 5^3 map, without a
 * second search.  Identical to ss_kmap_build(..., k_to). */
int ss_kmap_subset(const int32_t* nbr_from, const int64_t* count_from, int64_t n, int k_from, int k_to, int32_t* nbr_to,
                   int64_t* count_to, void* stream);

/* Compacted pair lists for the gather-GEMM path: tap t owns rows [tap_base[t], tap_base[t] + count[t])
 * of the product buffer; pair_in [p_pad] = input row of every product row, ypos [k^3, n] = product row
 * of (tap, output voxel) or -1; ypos_rank (nullable, k = 3) [n, 32] = the same positions indexed by the output's RANK
 * along order_row (columns >= 27 are -1): what the fused conv's reducer warps read, one 128-byte line per output;
 * tile_first_rank (nullable; tap bases multiples of 256) [p_pad / 256] = rank of the first output of every 256-row tile
 * of the product buffer; tile_order / tile_pos (nullable, both or none, need tile_first_rank) [p_pad / 256] = the tiles
 * sorted by (first rank, index) and the inverse permutation: the order in which the fused conv produces them. */
int ss_kmap_pairs(const int32_t* nbr, const int64_t* order_row, int64_t n, int k, const int64_t* tap_base_dev,
                  int64_t p_pad, int32_t* pair_in, int32_t* ypos, int32_t* ypos_rank, int32_t* tile_first_rank,
                  int32_t* tile_order, int32_t* tile_pos, void* workspace, size_t workspace_bytes, void* stream);

/* SIMT fp32-accumulate conv: out[p,co] = bias[co] + sum_t sum_ci wt[t][ci][co] * in[nbr[t][p]][ci],
 * then optional affine (folded BN) + activation.  wt is the [k^3, cin, cout] fp32 re-layout of the
 * reference weight [cout, k, k, k, cin]. */
int ss_subm_conv_simt(const void* in, int in_is_bf16, const int32_t* nbr, const float* wt, const float* bias,
                      const float* scale, const float* shift, int act, int64_t n, int k3, int cin, int cout, void* out,
                      int out_is_bf16, void* stream);

/* Training: weight gradient of a stem-like conv (fp32, Cout == 32, k^3 * Cin <= 1536):
 * dw[t][ci][co] = sum_p in[nbr[t][p]][ci] * dy[p][co]  (the [k^3, cin, cout] layout of ss_subm_conv_simt's wt).
 * Autograd of spconv.SubMConv3d(11 -> 32, k = 5) w.r.t. the weight (point_transformer_v3m1_base.py:499-506). */
size_t ss_stem_conv_wgrad_workspace_bytes(int k3, int cin);
int ss_stem_conv_wgrad(const float* in, const float* dy, const int32_t* nbr, int64_t n, int k3, int cin, int cout,
                       float* dw, void* workspace, size_t workspace_bytes, void* stream);

/* tcgen05 gather-GEMM, first stage of the xCPE conv: prod[r, :] = in[pair_in[r], :] @ w[tap(r)]^T for r < p_pad
 * (bf16 in, fp32 accumulate in TMEM, bf16 out).  w is [k^3, cout, cin] bf16 (K-major); tile_tap [p_pad/256] int32
 * gives the tap of every 256-row tile (every tap segment is padded to 256 rows).  cin, cout multiples of 16 / 32.
 * Persistent CTAs, 256-row tiles: two M=128 accumulators share every W stage. */
int ss_subm_conv_gemm256(const void* in_bf16, const int32_t* pair_in, const void* w_bf16, const int32_t* tile_tap,
                         int64_t p_pad, int k3, int cin, int cout, void* prod_bf16, void* stream);

/* The same stage on CTA pairs (tcgen05 cta_group::2, csrc/conv_gemm3.cu): identical arguments and
 * results; cout >= 256 (the 256-column slab is split between the two CTAs).  Double-buffered TMEM accumulators: the
 * store of tile i overlaps the MMAs of tile i + 1. */
int ss_subm_conv_gemm_pair(const void* in_bf16, const int32_t* pair_in, const void* w_bf16, const int32_t* tile_tap,
                           int64_t p_pad, int k3, int cin, int cout, void* prod_bf16, void* stream);

/* out[p,:] = bias + sum_t prod[ypos[t][p], :]  (fp32 accumulate) -> bf16/fp32 */
int ss_subm_conv_reduce(const void* prod_bf16, const int32_t* ypos, const float* bias, int64_t n, int k3, int cout,
                        void* out, int out_is_bf16, void* stream);

/* The gather-sum fused with the Block's next two steps (point_transformer_v3m1_base.py:318-326, conv Linear folded into the
 * taps): z = bias + sum_t prod[ypos[t][p], :] (fp32, never stored), res_out = res + LN(z; g0, b0) (fp32, may alias res),
 * norm_out = LN(res_out; g1, b1) (bf16).  channels % 8 == 0, <= 1024. */
int ss_subm_conv_reduce_add_ln(const void* prod_bf16, const int32_t* ypos, const float* bias, const float* res, const float* g0,
                               const float* b0, const float* g1, const float* b1, float eps, int64_t n, int k3, int channels,
                               float* res_out, void* norm_out_bf16, void* stream);

/* The xCPE conv of a Block in ONE launch (cout >= 256, k3 <= 32): ss_subm_conv_gemm_pair + ss_subm_conv_reduce_add_ln,
 * bit-identical to that pair of calls.  Twelve reducer warps per CTA walk the output voxels along `order_row` (the
 * serialized order the pair lists were built along), wait on per-tile completion counters and sum the product rows
 * while they are still in L2; the GEMM takes the 256-row product tiles in `tile_order` (a permutation of the p_pad / 256
 * tiles: by the rank of each tile's first output, ops.kmap_pairs; tile_pos is its inverse) and never runs more than
 * ~48 MB of products ahead of the tiles the reducers have asked for, so the products are read back from L2.  prod_bf16
 * [p_pad, cout] and tile_flags [p_pad / 256 + 2] are scratch; ypos_rank [n, 32] from ss_kmap_pairs.  Replaces spconv.SubMConv3d + Linear + LayerNorm + residual +
 * LayerNorm of point_transformer_v3m1_base.py:277-287,318-326. */
int ss_subm_conv_fused_add_ln(const void* in_bf16, const int32_t* pair_in, const void* w_bf16, const int32_t* tile_tap,
                              const int32_t* tile_order, const int32_t* tile_pos, int64_t p_pad, int k3, int cin, int cout,
                              void* prod_bf16, int32_t* tile_flags, const int64_t* order_row, const int32_t* ypos_rank,
                              const float* bias,
                              const float* res, const float* g0, const float* b0, const float* g1, const float* b1, float eps,
                              int64_t n, float* res_out, void* norm_out_bf16, void* stream);

/* ------------------------------------------------------------------------------------------------
 * SerializedAttention (point_transformer_v3m1_base.py:114-222). */

/* Patch table [max_patches] of int32x4 (q_begin, q_end, kv_begin, kv_end) in sorted positions, built
 * on the device from offset (no host sync); unused entries are zero.  max_patches >= n/K + n_batch. */
int ss_patch_table(const int64_t* offset, int n_batch, int patch_size, int max_patches, int32_t* table,
                   int32_t* n_patches_dev, void* stream);

/* qkv [n, 3*H*d] laid out (3, H, d) per row; order_row [n] = serialized_order[order_index].
 * out[order_row[j], h*d:(h+1)*d] = softmax(q_j K^T * scale) V over the patch of sorted position j. */
int ss_patch_attention_simt(const void* qkv, int in_is_bf16, const int64_t* order_row, const int32_t* table,
                            int max_patches, int patch_size, int heads, int head_dim, float scale, void* out,
                            int out_is_bf16, void* stream);

/* tcgen05 / TMEM kernel: bf16 in/out, head_dim in {16, 32, 48}, patch_size <= 1024 (any sequence length). */
int ss_patch_attention(const void* qkv_bf16, const int64_t* order_row, const int32_t* table, int max_patches,
                       int patch_size, int heads, int head_dim, float scale, void* out_bf16, void* stream);

/* Training forward: the same kernel, additionally writes lse2 [heads, n] fp32 = log2-domain log-sum-exp of the scaled
 * scores of every query, indexed by SORTED position (lse2[h * n + j]); consumed by ss_patch_attention_backward. */
int ss_patch_attention_lse(const void* qkv_bf16, const int64_t* order_row, const int32_t* table, int max_patches,
                           int patch_size, int heads, int head_dim, float scale, void* out_bf16, float* lse2, int64_t n,
                           void* stream);

/* Backward of the patch attention including the adjoints of the order / inverse gathers (autograd of
 * flash_attn_varlen_qkvpacked_func + the two index ops, point_transformer_v3m1_base.py:181-216):
 * dqkv [n, 3*H*d] bf16 = d loss / d qkv given dout [n, H*d] bf16, the forward's out and lse2.  tcgen05 kernels
 * (csrc/attention_bwd.cu); workspace from ss_patch_attention_backward_workspace_bytes. */
size_t ss_patch_attention_backward_workspace_bytes(int64_t n, int heads, int head_dim);
int ss_patch_attention_backward(const void* qkv_bf16, const void* out_bf16, const void* dout_bf16, const float* lse2,
                                const int64_t* order_row, const int32_t* table, int max_patches, int patch_size,
                                int heads, int head_dim, float scale, int64_t n, void* dqkv_bf16, void* workspace,
                                size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * MLP (point_transformer_v3m1_base.py:225-248): out = act(x W^T + b) on CTA pairs (tcgen05 cta_group::2), the
 * activation fused into the TMEM epilogue.  x [n, cin] bf16, w [cout, cin] bf16 (nn.Linear layout), bias fp32 [cout] or
 * NULL, act: 0 none, 1 exact GELU.  cin % 16 == 0, cout % 32 == 0. */
int ss_linear_act_bf16(const void* x_bf16, const void* w_bf16, const float* bias, int64_t n, int cin, int cout, int act,
                       void* out_bf16, void* stream);

/* The same GEMM with the residual add of a Block fused into the epilogue (point_transformer_v3m1_base.py:334-336,
 * `point.feat = shortcut + drop_path(mlp(point.feat))` in eval): out_f32[n, cout] = res + (x W^T + b) in fp32 (res may be
 * out_f32 itself: in place), and, if out_bf16 is not NULL, the bf16 copy of the result next to it. */
int ss_linear_residual_bf16(const void* x_bf16, const void* w_bf16, const float* bias, const float* res, int64_t n, int cin,
                            int cout, float* out_f32, void* out_bf16, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Row-wise fusions around the GEMMs of a Block (point_transformer_v3m1_base.py:318-338). */

/* y = res + f(delta), f = LayerNorm(g0,b0) if g0 else identity; res_out = y (fp32, may alias res);
 * norm_out = LayerNorm(y; g1,b1) if g1 else cast(y).  Any of res / res_out / norm_out may be NULL. */
int ss_add_layernorm(const float* res, const void* delta, int delta_is_bf16, const float* g0, const float* b0,
                     const float* g1, const float* b1, float eps, int64_t n, int channels, float* res_out,
                     void* norm_out, int norm_is_bf16, void* stream);

/* out = F.normalize(res + delta, p=2, dim=1, eps) (fp32) in one pass: the last Block's residual add
 * (point_transformer_v3m1_base.py:334-336) fused with the L2 normalisation LangPretrainer applies to the backbone output
 * (models/default.py:98).  out_bf16 (nullable) receives the same rows rounded to bf16 (the operand of ss_lang_head_tc).
 * channels % 8 == 0, <= 1024; 16-byte aligned pointers. */
int ss_add_l2_normalize(const float* res, const void* delta, int delta_is_bf16, float eps, int64_t n, int channels,
                        float* out, void* out_bf16, void* stream);

/* out = act(x * scale[c] + shift[c]) (scale/shift nullable).  act: 0 none, 1 GELU(erf). */
int ss_affine_act(const void* x, int in_is_bf16, const float* scale, const float* shift, int act, int64_t n,
                  int channels, void* out, int out_is_bf16, void* stream);

/* F.normalize(x, p=2, dim=1, eps)  (pointcept/models/default.py:98) */
int ss_l2_normalize(const void* x, int in_is_bf16, int64_t n, int channels, float eps, void* out, int out_is_bf16,
                    void* stream);

/* ------------------------------------------------------------------------------------------------
 * Language head and losses. */

/* logits = feat @ text^T; probs = sigmoid.  mode 0: max_prob [n], label [n] (-1 if max < threshold)
 * (pointcept/engines/hooks/evaluator.py:793-800); mode 1: probs_accum[idx[p] or p, :] += probs
 * (pointcept/engines/test.py:335-349).  normalize != 0 applies F.normalize to the row first. */
int ss_lang_head(const void* feat, int feat_is_bf16, const float* text, int64_t n, int channels, int n_classes,
                 int normalize, float threshold, int mode, const int64_t* idx, float* max_prob, int64_t* label,
                 float* probs_accum, void* stream);

/* Same head on the tensor cores (tcgen05 GEMM, max/argmax or accumulate fused into the TMEM epilogue):
 * bf16 feat [n, channels] and bf16 text [n_classes <= 256, channels], fp32 accumulate. */
int ss_lang_head_tc(const void* feat_bf16, const void* text_bf16, int64_t n, int channels, int n_classes,
                    float threshold, int mode, const int64_t* idx, float* max_prob, int64_t* label, float* probs_accum,
                    void* stream);

/* ------------------------------------------------------------------------------------------------
 * Training path adjoints.  LayerNorm backward (torch nn.LayerNorm in the reference's Block,
 * point_transformer_v3m1_base.py:277-338): dx = d/dx of LN(x; gamma, beta) contracted with dy, in x's dtype;
 * dgamma / dbeta (fp32, [channels]) are ACCUMULATED: zero them first.  channels % 8 == 0, <= 1024. */
/* GELU (erf form, nn.GELU()) backward on bf16 tensors: dx = dy * (Phi(x) + x phi(x)); n_elements % 8 == 0. */
int ss_gelu_backward_bf16(const void* x, const void* dy, int64_t n_elements, void* dx, void* stream);

/* out[c] = sum_r x[r, c]: bias gradient of a Linear layer (autograd of F.linear, db = dy.sum(0)), bf16 rows, fp32
 * deterministic two-stage sum.  channels % 8 == 0. */
size_t ss_colsum_workspace_bytes(int channels);
int ss_colsum_bf16(const void* x_bf16, int64_t n, int channels, float* out, void* workspace, size_t workspace_bytes,
                   void* stream);

/* Weight gradient of the 3^3 submanifold conv on the tensor cores: dw[t][co][ci] += sum over the pairs r of the K chunks
 * of tap t of dy[pair_out[r]][co] * x[pair_in[r]][ci].  chunks: [n_chunks] int32x4 (tap, k_begin, k_end, 0) over the pair
 * lists of ss_kmap_pairs (exact ranges: padding rows excluded); dw fp32 [k3, cout, cin], ACCUMULATED (zero it first);
 * cin, cout multiples of 32. */
int ss_subm_conv_wgrad(const void* x_bf16, const void* dy_bf16, const int32_t* pair_in, const int64_t* pair_out,
                       const int32_t* chunks, int n_chunks, int k3, int cin, int cout, float* dw, void* stream);
int ss_layernorm_backward(const void* x, int x_is_bf16, const void* dy, int dy_is_bf16, const float* gamma, float eps,
                          int64_t n, int channels, void* dx, float* dgamma, float* dbeta, void* stream);

/* ------------------------------------------------------------------------------------------------
 * SphereCrop (pointcept/datasets/transform.py:1419-1535, modes "random" / "center"): order[j] = index of the j-th
 * nearest point to `center3` (HOST pointer, 3 floats), distances in numpy's fp32 arithmetic, ties by ascending
 * index; dist_bits_sorted[j] = bit pattern of that squared distance (low 32 bits). */
size_t ss_sphere_crop_workspace_bytes(int64_t n);
int ss_sphere_crop_order(const float* coord, int64_t n, const float* center3, int64_t* order, uint64_t* dist_bits_sorted,
                         void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Zero-shot evaluation tail: neighbour voting (pointcept/utils/misc.py:17-95: cKDTree k-NN + majority vote) and
 * the confusion-matrix update (pointcept/engines/hooks/evaluator.py:830-834).  `origin3` is a HOST pointer to the
 * three floats of the grid origin; everything else is device memory.  Grid: nx*ny*nz cells of edge `cell`. */

/* cell_id[i] = linear cell of reference point i, cell_count[cell]++ (cell_count zero-initialised by the caller). */
int ss_vote_bin_count(const float* pts_xyz, int64_t m, const float* origin3, float cell, int nx, int ny, int nz,
                      int32_t* cell_count, int64_t* cell_id, void* stream);
/* Counting-sort fill: binned[cell_start[cell] + slot] = (x, y, z, label bits); cursor zero-initialised. */
int ss_vote_bin_fill(const float* pts_xyz, const int32_t* labels, const int64_t* cell_id, int64_t m,
                     const int64_t* cell_start, int32_t* cursor, void* binned_xyzl, void* stream);
/* out[q] = majority label of the k nearest reference points of query q (fp64 distances, ties to the smallest
 * label, no valid label -> ignore_label).  k <= 64. */
int ss_knn_vote(const void* binned_xyzl, const int64_t* cell_start, const float* origin3, float cell, int nx, int ny,
                int nz, const float* query_xyz, int64_t nq, int k, int ignore_label, int num_classes, int32_t* out,
                void* stream);
/* confusion[gt, pred] += 1, or fn_ignore[gt] += 1 where pred == ignore_index (int64 counters, atomics). */
int ss_confusion_update(const int64_t* gt, const int64_t* pred, int64_t n, int num_classes, int64_t ignore_index,
                        int64_t* confusion, int64_t* fn_ignore, void* stream);

/* acc3 = {sum_valid (1 - cos), sum_valid ||pred - target||^2, n_valid} as doubles
 * (pointcept/models/losses/misc.py:247-295).  target_dtype: 0 fp32, 1 bf16, 2 fp16. */
int ss_cos_l2_loss(const void* pred, int pred_is_bf16, const void* target, int target_dtype, const uint8_t* mask,
                   int64_t n, int channels, double* acc3, void* stream);

/* sums [(label*2 + half), channels] (fp32) and counts [n_classes*2] over valid points
 * (pointcept/models/losses/misc.py:355-389). */
int ss_class_half_sums(const void* pred, int pred_is_bf16, const uint8_t* mask, const int64_t* segment,
                       const int64_t* half, int64_t n, int channels, int n_classes, float* sums, int32_t* counts,
                       void* stream);

/* ------------------------------------------------------------------------------------------------
 * Adjoints of the pooling / unpooling reductions and of the language losses (training; csrc/pool_loss_bwd.cu). */

/* d src of the segment mean (reduce = 1) / sum (reduce = 0) of SerializedPooling (autograd through
 * torch_scatter.segment_csr, point_transformer_v3m1_base.py:416-418): dsrc[p,:] = dout[cluster[p],:] (/ count). */
int ss_segment_mean_bwd(const void* dout, int dout_is_bf16, const int64_t* cluster, const int64_t* seg_start, int64_t n,
                        int channels, int reduce, void* dsrc, int dsrc_is_bf16, void* stream);

/* d child of SerializedUnpooling's gather `point.feat[inverse]` (ref :478): dchild[m,:] = sum over the members of cluster
 * m (positions seg_start[m] .. seg_start[m+1] of order0) of dout[p,:]; deterministic segment sum. */
int ss_unpool_gather_add_bwd(const void* dout, int dout_is_bf16, const int64_t* order0, const int64_t* seg_start, int64_t m,
                             int channels, void* dchild, int dchild_is_bf16, void* stream);

/* d pred of w_cos * sum_valid(1 - cos) / n_valid + w_l2 * sum_valid ||pred - target||^2 / n_valid (losses/misc.py:254-295)
 * times *grad_out (device scalar, NULL = 1).  acc3 = the double[3] ss_cos_l2_loss wrote (n_valid is read from it on the
 * device).  Rows with mask == 0 get zeros.  dpred fp32 [n, channels]. */
int ss_cos_l2_loss_bwd(const void* pred, int pred_is_bf16, const void* target, int target_dtype, const uint8_t* mask,
                       int64_t n, int channels, const double* acc3, const float* grad_out, float w_cos, float w_l2,
                       float* dpred, void* stream);

/* d pred of ss_class_half_sums: dpred[p,:] = dsums[segment[p] * 2 + half[p], :] for valid rows, else 0. */
int ss_class_half_sums_bwd(const float* dsums, const uint8_t* mask, const int64_t* segment, const int64_t* half, int64_t n,
                           int channels, int n_classes, float* dpred, void* stream);

/* Library / build identification; number of kernels this library has launched in the process. */
const char* ss_version(void);
uint64_t ss_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SCENESPLAT_B200_H_ */
