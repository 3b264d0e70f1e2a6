/* st_b200.h -- C ABI of libst_b200.so: the sm_100a kernels behind smart-tree's hot path.
 *
 * smart-tree itself holds no native code: its arithmetic lives in three un-vendored
 * third-party packages (spconv, FRNN, cugraph/cudf/cupy).  Each group of
 * entry points below replaces the native calls the reference reaches through one Python
 * call site; the call site is cited as file:line under /root/reference/.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host
 *   - the CALLER allocates every output and workspace (sizes via *_workspace_bytes or the
 *     stated worst-case bounds); the library never allocates, never frees, holds no state
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it; functions that
 *     return a count through an *_host pointer synchronise that stream before returning
 *   - return value: 0 = ok, negative = error; st_last_error() gives the thread-local text
 *   - features are fp32 row-major; coordinates int32 (batch, z, y, x) with
 *     0 <= z,y,x <= 65533 and 0 <= batch < 32768 (checked by st_hash_build); graph indices int32
 */
#ifndef ST_B200_H
#define ST_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ST_OK 0
#define ST_ERR_ARG (-1)
#define ST_ERR_CUDA (-2)
#define ST_ERR_WORKSPACE (-3)
#define ST_ERR_UNSUPPORTED (-4)

#define ST_ACT_RELU 1

int st_version(void);
const char *st_last_error(void);
/* 0 if `device` is sm_100 or newer, ST_ERR_UNSUPPORTED otherwise. */
int st_device_check(int device);
int st_sm_count(int device);

/* ------------------------------------------------------------------ K1 voxelise
 * replaces spconv.pytorch.utils.PointToVoxel(...).generate_voxel_with_id with
 * max_num_points_per_voxel=1          smart_tree/dataset/dataset.py:199-216
 * Batched over blocks: point i belongs to block point_block[i] (or block 0 if NULL); block b
 * has range lo = block_lo[b*3..], grid = block_grid[b*3..] (xyz order; grid =
 * round((hi-lo)/vsize) computed by the caller).  coordinate c = floor((p-lo)/vsize) (fp32);
 * points with c<0 or c>=grid are dropped (pc_voxel_id -1).  The first point (lowest index)
 * of a voxel is its representative; voxels are numbered in first-appearance order.
 *   points[n, ld] (xyz first) ; pc_voxel_id[n] ; rep_point[n] ; coords[n,4]=(b,z,y,x)
 *   *n_voxels_host receives M.                                                        */
size_t st_voxelize_workspace_bytes(int64_t n);
int st_voxelize(const float *points, int64_t n, int ld, const int32_t *point_block,
                const float *block_lo, const int32_t *block_grid, int32_t n_blocks, float vsize,
                int32_t *pc_voxel_id, int32_t *rep_point, int32_t *coords,
                int64_t *n_voxels_host, void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------ block tiling (inference)
 * replaces SingleTreeInference.compute_blocks + cube_filter (pure torch in the reference: one O(N)
 * mask, one device->host copy per block)   smart_tree/dataset/dataset.py:166-190 ; util/maths.py:145-155
 * st_block_list : block id = torch.div(xyz, block_size, rounding_mode="floor"); blocks holding more
 *                 than min_points points, sorted by (x,y,z) id as torch.unique(dim=0) does;
 *                 block_ids[cap,3] (floats), kept_keys[cap] (opaque, for the next two calls).
 * st_block_count / st_block_emit : every (block, point) pair with the point inside the block's
 *                 enlarged cube  centre - half_cube <= p < centre + half_cube, centre = id*block_size
 *                 + half_block (fp32, as the reference computes it); `reach` = how many neighbouring
 *                 blocks a point can spill into per axis (ceil(buffer/block_size)).  Pairs come out
 *                 block-major with ascending point index inside a block, together with each block's
 *                 member bounding box block_lo / block_hi [n_blocks,3] (the voxeliser's range).      */
size_t st_block_workspace_bytes(int64_t n, int64_t n_pairs);
int st_block_list(const float *xyz, int64_t n, float block_size, int min_points, float *block_ids,
                  uint64_t *kept_keys, int32_t cap, int64_t *n_blocks_host, void *workspace,
                  size_t workspace_bytes, void *stream);
int st_block_count(const float *xyz, int64_t n, const uint64_t *kept_keys, int32_t n_blocks,
                   float block_size, float half_block, float half_cube, int reach, int32_t *offsets,
                   int64_t *n_pairs_host, void *workspace, size_t workspace_bytes, void *stream);
int st_block_emit(const float *xyz, int64_t n, const uint64_t *kept_keys, int32_t n_blocks,
                  float block_size, float half_block, float half_cube, int reach,
                  const int32_t *offsets, int64_t n_pairs, int64_t *point_index, int32_t *point_block,
                  float *block_lo, float *block_hi, void *workspace, size_t workspace_bytes,
                  void *stream);

/* ------------------------------------------------------------------ K2 coordinate table + sub-manifold map
 * replaces spconv ops.get_indice_pairs(subm=True) reached from every SubMConv3d.forward
 *                                      smart_tree/model/model_blocks.py:24-32,123-144,258-282
 * Hash table: `capacity` slots (power of two >= 2n, see st_hash_capacity), keys u64, vals i32.
 * `status` (optional, device int32): set to 1 if any coordinate lies outside the range the 64-bit keys can hold
 * (0 <= z,y,x <= 65533, 0 <= batch < 32768); such rows are left out of the table -- the caller must treat it as an error.
 * nbr[27, n]: row of the active voxel at coords[i] + (kz-1,ky-1,kx-1), k=(kz*3+ky)*3+kx, or -1.
 * spatial_shape (optional, device int32[3] = (z,y,x)): the reference declares spatial_shape = max(coords), not max+1
 * (smart_tree/model/sparse.py:15-19), and spconv bound-checks neighbour LOCATIONS against it, so voxels on the max
 * faces are invisible as neighbours.  NULL = unbounded grid (the default of this library);
 * non-NULL reproduces the clip: a neighbour with any coordinate >= spatial_shape is reported absent (the centre tap
 * is never clipped).  Only the clip is modelled, not the index aliasing of the out-of-range voxels.             */
int64_t st_hash_capacity(int64_t n);
int st_hash_build(const int32_t *coords, int64_t n, uint64_t *keys, int32_t *vals,
                  int64_t capacity, int32_t *status, void *stream);
int st_subm_map(const int32_t *coords, int64_t n, const uint64_t *keys, const int32_t *vals,
                int64_t capacity, const int32_t *spatial_shape, int32_t *nbr, void *stream);

/* ------------------------------------------------------------------ K3 strided / inverse conv maps
 * replaces spconv generate_conv_inds for SparseConv3d(k=3,s=2,p=1,indice_key) and the reuse of
 * that rulebook by SparseInverseConv3d   smart_tree/model/model_blocks.py:57-70,90-100
 * Step 1: out_coords = sorted unique { (b, (p+1-k)/2) } ; *n_out_host receives M (M <= 8n;
 *         out_coords must hold 8n rows).  Step 2 (after the caller built the hash of
 *         out_coords): down[27,M] = input row feeding output o through tap k (p = 2o-1+k),
 *         up[27,n] = output row fed by input p through tap k; -1 where absent.           */
/* out_shape (optional, device int32[3]): strict_spconv_bounds as above -- outputs with any coordinate >= out_shape
 * (= (spatial_shape - 1) / 2 + 1 of the input level, the shape spconv hands to the next level) are not created.    */
size_t st_strided_coords_workspace_bytes(int64_t n);
int st_strided_coords(const int32_t *coords, int64_t n, int morton_order, const int32_t *out_shape,
                      int32_t *out_coords, int64_t *n_out_host, void *workspace, size_t workspace_bytes,
                      void *stream);
/* Row order is free inside the network (every encoder is undone by its inverse conv), so the engine
 * keeps each level in (batch, Z-order): morton_order=1 above sorts the new level that way, and
 * st_morton_perm gives the permutation that does the same for the caller's level-0 rows
 * (perm[k] = caller row of the k-th voxel).  Gathers of spatial neighbours then hit L1/L2 lines that
 * a CTA has just touched instead of random rows.                                              */
size_t st_morton_workspace_bytes(int64_t n);
int st_morton_perm(const int32_t *coords, int64_t n, int32_t *perm, void *workspace,
                   size_t workspace_bytes, void *stream);
int st_strided_maps(const int32_t *coords, int64_t n, int64_t n_out, const uint64_t *out_keys,
                    const int32_t *out_vals, int64_t out_capacity, int32_t *down, int32_t *up,
                    void *stream);

/* Row gather dst[r,:] = src[idx[r],:] (rows of row_bytes bytes, a multiple of 4; int32 indices): the row
 * permutations of the path (Z-order rows of the network, voxel representatives; torch indexing in the reference:
 * smart_tree/dataset/dataset.py:218-226, model/sparse.py:40-61).                                          */
int st_gather_rows(const void *src, const int32_t *idx, int64_t n, int row_bytes, void *dst, void *stream);

/* Devoxelise: per-voxel predictions back to every input point.  The reference's voxeliser
 * returns pc_voxel_id and drops it (smart_tree/dataset/dataset.py:214); this is the gather it omits.  Pair t is
 * (point pair_point[t], block pair_block[t]) with voxel row pair_voxel[t] (-1 = dropped) as st_block_emit /
 * st_voxelize produce them; a point takes the prediction of the block whose inner half-open cube
 * [centre - size/2, centre + size/2) contains it (util/maths.py:135-155).  Outputs are initialised here:
 * point_medial[n,3] = 0, point_class[n] = -1, point_voxel[n] = -1 for points without a prediction.          */
int st_devoxelize(const float *xyz, int64_t n_points, const int64_t *pair_point, const int32_t *pair_block,
                  const int32_t *pair_voxel, int64_t n_pairs, const float *block_centres, float block_size,
                  const float *voxel_medial, const int32_t *voxel_class, float *point_medial,
                  int32_t *point_class, int32_t *point_voxel, void *stream);

/* ------------------------------------------------------------------ K4/K5/K6 gather convolution, fused epilogue
 * replaces spconv Fsp.indice_subm_conv / indice_conv / indice_inverse_conv (ConvAlgo.Native)
 * followed by nn.BatchNorm1d(eval), the residual add, nn.ReLU and torch.cat
 *                                      smart_tree/model/model_blocks.py:33-34,68-69,99-100,148-156,238-240
 *   out[i, :] = act( scale * sum_k W[k] . in[map[k, i], :] + shift + residual[i, :] )
 * map[ntaps, n_out] (NULL with ntaps==1 means the identity map: a 1x1 conv / nn.Linear),
 * w[ntaps, cin, cout] (the spconv weight [cout,kz,ky,kx,cin] transposed by the caller),
 * scale/shift/residual may be NULL; in_ld/out_ld/res_ld are row strides in floats so that
 * a conv can read from / write into a column slice of a wider buffer (the skip concat).
 * Optional second input (fused ResBlock identity 1x1 conv): adds w2[cin2,cout] . in2[i,:]
 * after the affine, before the activation.                                             */
int st_conv_gather(const float *in, int in_ld, const int32_t *map, int64_t n_out, int ntaps,
                   const float *w, int cin, int cout, const float *scale, const float *shift,
                   const float *residual, int res_ld, const float *in2, int in2_ld,
                   const float *w2, int cin2, float *out, int out_ld, int act, void *stream);

/* Stem: the 1x1 input conv + BN + ReLU (smart_tree/model/model.py:31-38, SubMConv3d k=1 fast path) fused with the
 * row permutation that puts the network's rows into Z-order: out[i,:] = act(scale * (W . in[row_index[i], :cin]) + shift).
 * w[cin, cout]; cin <= 8, cout in {8, 16}; `in` may be a column slice of a wider array (in_ld floats per row);
 * row_index may be NULL (identity).  ST_ERR_UNSUPPORTED for other shapes (use st_conv_gather with ntaps = 1).       */
int st_stem_conv(const float *in, int in_ld, const int32_t *row_index, int64_t n, const float *w, int cin,
                 int cout, const float *scale, const float *shift, float *out, int out_ld, int act, void *stream);

/* Inverse (decoder) conv on parity-sorted rows, fp32 FMA kernel: same plan (st_inverse_plan / st_strided_maps_inv) and
 * result as st_conv_gather_tc_inv below; w[ntaps, cin, cout] plain.  For the narrow decoder of level 0 (16 -> 8), where a
 * fine voxel has 3.4 of 27 taps on average, the CTA reads only the map rows of the taps its tiles' masks name.
 *                                      replaces spconv SparseInverseConv3d  smart_tree/model/model_blocks.py:91-100 */
int st_conv_gather_inv(const float *in, int in_ld, const int32_t *up_sorted, const int32_t *row_index,
                       const uint32_t *tile_mask, int64_t n_out, int ntaps, const float *w, int cin, int cout,
                       const float *scale, const float *shift, float *out, int out_ld, int act, void *stream);

/* Brick variant of the sub-manifold 3x3x3 conv for the narrow level (cin 8 or 16, cout 8: level 0 of the UNet), no
 * gather map.  Rows must be in (batch, Z-order) -- st_morton_perm order -- so that the voxels of an aligned 4x4x4 brick
 * are contiguous rows; st_brick_plan_build derives, once per level, the brick table, a 1-byte window-cell code per row and
 * a halo list per brick (occupied cells of the 6x6x6 window outside the brick, found through the level's hash table).
 * st_conv_brick then stages each brick's window in shared memory and computes the same result as st_conv_gather with
 * the level's st_subm_map (unbounded grid: no strict_spconv_bounds), including the fused epilogue.
 * st_brick_plan_info: [0] bricks, [1] halo entries, [2] status -- nonzero if the rows are not strictly increasing in
 * (batch, Z-order) (the plan is then unusable).      replaces spconv SubMConv3d  smart_tree/model/model_blocks.py:8-38,107-156 */
size_t st_brick_plan_bytes(int64_t n);
int st_brick_plan_build(const int32_t *coords, int64_t n, const uint64_t *keys, const int32_t *vals, int64_t capacity,
                        void *plan, size_t plan_bytes, void *stream);
int st_brick_plan_info(const void *plan, int64_t n, int32_t *info_host);
int st_conv_brick(const float *in, int in_ld, const void *plan, int64_t n, const float *w /* [27, cin, 8] */, int cin,
                  int cout, const float *scale, const float *shift, const float *residual, int res_ld, const float *in2,
                  int in2_ld, const float *w2, int cin2, float *out, int out_ld, int act, void *stream);

/* Tensor-core variant (tcgen05.mma kind::tf32 with a 3xTF32 split, accumulators in TMEM).  Same
 * contract as st_conv_gather except that `wprep` is the weight tensor pre-arranged by
 * st_conv_tc_prepare (st_conv_tc_weight_floats(ntaps,cin,cout) floats; -1 = unsupported channel
 * counts: cin must divide 32 or be a multiple of 32, cout a multiple of 8) and `map` is mandatory.   */
int64_t st_conv_tc_weight_floats(int ntaps, int cin, int cout);
int st_conv_tc_prepare(const float *w, int ntaps, int cin, int cout, float *wprep, void *stream);
/* Fused ResBlock identity (model_blocks.py:148-156: ReLU(BN(conv(t)) + identity_1x1(x))): st_conv_tc_prepare_fused
 * appends w2[cin2,cout] / scale as extra K stages to the prepared weights, and st_conv_gather_tc called with
 * in2 != NULL and w2 == NULL walks them through the identity map, so the 1x1 conv runs on the tensor cores
 * inside the same accumulation instead of as a per-thread FMA loop in the epilogue.  cin2 % 16 == 0;
 * every scale[c] must be non-zero (the caller checks; otherwise pass w2 to keep the epilogue form).           */
int64_t st_conv_tc_weight_floats_fused(int ntaps, int cin, int cout, int cin2);
int st_conv_tc_prepare_fused(const float *w, int ntaps, int cin, int cout, const float *w2, int cin2,
                             const float *scale, float *wprep, void *stream);
int st_conv_gather_tc(const float *in, int in_ld, const int32_t *map, int64_t n_out, int ntaps,
                      const float *wprep, int cin, int cout, const float *scale, const float *shift,
                      const float *residual, int res_ld, const float *in2, int in2_ld,
                      const float *w2, int cin2, float *out, int out_ld, int act, void *stream);

/* Inverse (decoder) conv on parity-sorted rows       replaces spconv SparseInverseConv3d  model_blocks.py:91-100
 * A fine voxel p receives from coarse voxel o through tap k iff p = 2o - 1 + k: per axis the tap is fixed by the
 * parity of p, so at most 8 of the 27 taps (3.4 on average) can be non-empty and which ones depends only on the
 * parity class of p.  st_inverse_plan sorts the fine level's rows by parity class (stable), permutes the `up` map of
 * st_strided_maps accordingly (up_sorted[ntaps, n]) and records per 128-row tile the taps that occur (tile_mask);
 * st_conv_gather_tc_inv (contract of st_conv_gather_tc, no residual / second input) walks only the K stages that
 * have work in each tile and writes launch row j to output row row_index[j].                                  */
size_t st_inverse_plan_workspace_bytes(int64_t n);
int st_inverse_plan(const int32_t *coords, const int32_t *up, int64_t n, int ntaps, int32_t *row_index,
                    int32_t *up_sorted, uint32_t *tile_mask, void *workspace, size_t workspace_bytes, void *stream);
/* st_strided_maps and st_inverse_plan in one call (the engine's path): `down` as st_strided_maps, the `up` map
 * only in parity-sorted launch order (up_sorted / row_index / tile_mask).                                      */
size_t st_strided_maps_inv_workspace_bytes(int64_t n);
int st_strided_maps_inv(const int32_t *coords, int64_t n, int64_t n_out, const uint64_t *out_keys,
                        const int32_t *out_vals, int64_t out_capacity, int32_t *down,
                        int32_t *row_index, int32_t *up_sorted, uint32_t *tile_mask, void *workspace,
                        size_t workspace_bytes, void *stream);
int st_conv_gather_tc_inv(const float *in, int in_ld, const int32_t *up_sorted, const int32_t *row_index,
                          const uint32_t *tile_mask, int64_t n_out, int ntaps, const float *wprep, int cin, int cout,
                          const float *scale, const float *shift, float *out, int out_ld, int act, void *stream);

/* Tile-plan variant of the tensor-core conv: the DISTINCT source rows of every 128-row output tile are
 * staged in shared memory once (rows are in Z-order, so a tile's 27 x 128 map entries name only ~2 x 128
 * distinct rows) and the A fragments are gathered from there instead of through the L1.  The plan is
 * built once per gather map (spconv builds its rulebook per indice_key the same way:
 * smart_tree/model/model_blocks.py:24-32,58-67,91-98) and shared by every conv that uses the map:
 *   plan = [ per tile: split flag + distinct-row counts | 16-bit local gather map [28][128] | row lists ]
 * st_conv_plan_bytes(n_out) bytes, 256-byte aligned, caller-allocated.  n_in = rows of the source array.
 * st_conv_gather_tp has the contract of st_conv_gather_tc with `plan` in place of `map`
 * (st_conv_tp_supported: 3x3x3 maps, cin in {8,16,32}, cout a multiple of 8).                      */
size_t st_conv_plan_bytes(int64_t n_out);
int st_conv_plan_build(const int32_t *map, int64_t n_out, int ntaps, int64_t n_in, void *plan,
                       size_t plan_bytes, void *stream);
int st_conv_tp_supported(int ntaps, int cin, int cout);
int st_conv_gather_tp(const float *in, int in_ld, const void *plan, int64_t n_out, int ntaps,
                      const float *wprep, int cin, int cout, const float *scale, const float *shift,
                      const float *residual, int res_ld, const float *in2, int in2_ld,
                      const float *w2, int cin2, float *out, int out_ld, int act, void *stream);

/* Fused heads: the three SparseFC stacks (8->8,BN,ReLU, 8->4,BN,ReLU, 4->{1,3,2}), F.normalize,
 * exp(radius)*direction and argmax     smart_tree/model/model.py:83-85 ; model_blocks.py:258-282 ;
 *                                      smart_tree/model/model_inference.py:87-88
 * params: packed fp32 block described in smart-tree_b200/engine.py (pack_heads).
 * Outputs (any may be NULL): radius[n] (log radius), direction[n,3] (unit), class_logits[n,2],
 * medial_vector[n,3], class_l[n] (int32 argmax, first maximum).                         */
int st_heads_fused(const float *in, int in_ld, int64_t n, const float *params,
                   const int32_t *out_index /* optional: result of row i is written to row out_index[i] */,
                   float *radius, float *direction, float *class_logits, float *medial_vector,
                   int32_t *class_l, void *stream);

/* ------------------------------------------------------------------ K7 fixed-radius kNN
 * replaces frnn.frnn_grid_points(p1, p2, len1, len2, K, r, return_sorted=True)
 *                                      smart_tree/skeleton/graph.py:15-24
 * K nearest of dst[m,3] within radius r of each src[n,3]: d2 = (dx*dx+dy*dy)+dz*dz < r*r,
 * ascending (d2, index); idx[n,K] (-1 padded), d2[n,K] (-1 padded).  K <= 32.
 * query_radius (optional, [n]): additionally restrict query i to d <= query_radius[i]
 * (exactly the reference's later `idxs[dists > radii] = -1`).                           */
size_t st_knn_workspace_bytes(int64_t m);
int st_knn(const float *src, int64_t n, const float *dst, int64_t m, int K, float r,
           const float *query_radius, int32_t *idx, float *d2, void *workspace,
           size_t workspace_bytes, void *stream);

/* outlier_removal                      smart_tree/skeleton/filter.py:6-11
 * keep[i] = 1 iff the nb nearest neighbours (self included) within r=max radius all exist and
 * have sqrt(d2) < radii[i].                                                             */
int st_outlier_mask(const float *points, int64_t n, const float *radii, float r_max, int nb,
                    uint8_t *keep, void *workspace, size_t workspace_bytes, void *stream);

/* nn_graph + make_edges                smart_tree/skeleton/graph.py:36-40,52-60
 * From a kNN result: drop neighbour if sqrt(d2) > radii[i]; emit (i, j, sqrt(d2)) for j > 0
 * (sic), row-major order.  edges[n*K,2], weights[n*K]; *n_edges_host receives E.         */
size_t st_edges_workspace_bytes(int64_t n, int K);
int st_edges_from_knn(const int32_t *idx, const float *d2, int64_t n, int K, const float *radii,
                      int32_t *edges, float *weights, int64_t *n_edges_host, void *workspace,
                      size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------ K8 connected components
 * replaces cugraph.connected_components on Graph(directed=False, renumber=False)
 *                                      smart_tree/data_types/graph.py:32-51
 * label[v] = smallest vertex id of v's component; size[v] = component size (at every v).   */
int st_connected_components(const int32_t *edges, int64_t n_edges, int64_t n, int32_t *label,
                            int32_t *size, void *stream);

/* ------------------------------------------------------------------ K9 shortest paths
 * replaces cugraph.sssp (twice)        smart_tree/skeleton/shortest_path.py:12-21 ;
 *                                      smart_tree/skeleton/skeletonize.py:73-85
 * Undirected weighted graph given as CSR over arcs (both directions, self loops removed):
 * row_ptr[n+1], col[A], w[A].  sources[s]: dist 0 there.  dist = least fixed point of
 * d[v] = min_u fl32(d[u]+w); pred[v] = lowest u with fl32(d[u]+w)==d[v] (-1 at sources and
 * unreachable, dist FLT_MAX there).  st_csr_build makes the CSR from an edge list.        */
size_t st_csr_workspace_bytes(int64_t n, int64_t n_edges);
int st_csr_build(const int32_t *edges, const float *weights, int64_t n_edges,
                 const int32_t *vertex_map /* optional: edge endpoints are renumbered through it, < 0 drops the edge */,
                 int64_t n, int32_t *row_ptr, int32_t *col, float *w, int64_t *n_arcs_host,
                 void *workspace, size_t workspace_bytes, void *stream);
/* orig_id (optional, [n]): the caller may number the GRAPH in any order that suits the kernel -- spatial (Z) order makes the
 * vertex range of a CTA a compact blob, so that most hops of a shortest-path chain stay in its shared memory -- and pass the
 * map back to its own numbering: vertex v of the graph is vertex orig_id[v] of the caller.  dist / pred are then written in
 * the CALLER's numbering (dist[orig_id[v]], pred values are caller ids) and the tie rule "lowest predecessor" is evaluated
 * on caller ids, i.e. the result does not depend on the graph's numbering.  `sources` are graph vertices.                */
int st_sssp(const int32_t *row_ptr, const int32_t *col, const float *w, int64_t n,
            const int32_t *sources, int32_t n_sources,
            float delta /* threshold step of the distance-ordered schedule; any value gives the same result; <=0: default */,
            float *dist, int32_t *pred,
            int32_t *sweeps_host, void *ctl_workspace /* 4608 + 16n B, device */, const int32_t *orig_id, void *stream);
/* pred_graph + second sssp             smart_tree/skeleton/shortest_path.py:46-55
 * tree_dist[v] = tree_dist[pred[v]] + ||p_v - p_pred(v)||, 0 at roots (pred<0 & reachable
 * flag), FLT_MAX where unreachable[v] != 0.                                              */
int st_tree_distances(const float *points, const int32_t *pred, const uint8_t *is_root,
                      int64_t n, float *tree_dist, void *ctl_workspace /* 256 + 8n B, device */,
                      void *stream);

/* ------------------------------------------------------------------ K10 greedy branch extraction
 * replaces sample_tree / trace_route / select_path_points (torch + one FRNN call per branch)
 *                                      smart_tree/skeleton/path.py:9-140
 * Batched over components: component c owns vertices [comp_off[c], comp_off[c+1]); pred is
 * component-local (-1 at the root); all quirks of SURVEY Appendix E are reproduced.
 * The uniform grid used to claim points near a path has cell size `cell_size` (enlarged if it
 * would exceed the cell budget); any positive value gives identical results.
 * Outputs are SEGMENT-LOCAL: component c writes into [comp_off[c], comp_off[c+1]) of
 * path_vertices[n] (component-local vertex ids, branch after branch, root side first),
 * branch_len[n] and branch_parent[n] (one entry per emitted branch, in emission order = branch
 * id); comp_n_branches[c] / comp_n_path[c] give how many entries of the segment are valid.   */
size_t st_sample_tree_workspace_bytes(int64_t n, int32_t n_comp);
int st_sample_tree(const float *medial_pts, const float *radii, const int32_t *pred,
                   const float *tree_dist, const int32_t *comp_off, int32_t n_comp, int64_t n,
                   float cell_size, int32_t *path_vertices, int32_t *branch_len,
                   int32_t *branch_parent, int32_t *comp_n_branches, int32_t *comp_n_path,
                   void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------ K11 point -> tapered tube projection (repair)
 * replaces pts_to_nearest_tube_gpu     smart_tree/util/queries.py:89-133
 * For query q (one per branch): tubes [tube_off[q], tube_off[q+1]) of a[.,3], b[.,3], r1, r2;
 * t = clip((p-a).(b-a)/(b-a).(b-a), 0, 1); picks argmin |dist - r(t)| (first minimum);
 * out_vec[q,3] = projection - p ; out_idx[q] (tube index local to the query); out_r[q].   */
int st_points_to_tubes(const float *pts, int64_t n_q, const float *a, const float *b,
                       const float *r1, const float *r2, const int32_t *tube_off,
                       float *out_vec, int32_t *out_idx, float *out_r, void *stream);

/* TreeSkeleton.repair for a whole skeleton in one launch   smart_tree/data_types/tree.py:73-92
 * nodes[R,4] = (xyz, radius) of all branches; branch b owns rows row[b]+1 .. row[b]+len[b] and row[b] is a
 * spare row that receives its connection point (and the first node's radius).  Branches are listed sorted
 * by depth below the first repaired level: level_off[n_levels+1]; parent[b] indexes the same arrays
 * (any branch, not only listed ones: give all branches, levels cover the repaired ones first ... see
 * smart-tree_b200/data_types/tree.py); parent_repaired[b] != 0 iff the parent's spare row is already part
 * of its polyline when b is visited (the reference visits parents before children).              */
int st_repair_branches(float *nodes, const int32_t *row, const int32_t *len, const int32_t *parent,
                       const uint8_t *parent_repaired, const int32_t *level_off, int32_t n_levels,
                       void *stream);

/* ------------------------------------------------------------------ K12 skeleton finishing, all components, one launch
 * replaces the branch assembly of sample_tree   smart_tree/skeleton/path.py:118-129
 *          TreeSkeleton.prune / repair / smooth smart_tree/data_types/tree.py:94-121, 73-92, 123-134
 *          (DisjointTreeSkeleton.prune touches the first skeleton only, tree.py:164-168)
 * Inputs are st_sample_tree's outputs (segment-local) plus the component-grouped medial points / radii.
 * out (int32 words, 16-byte aligned, st_finish_skeletons_out_ints(n) words):
 *   header[4]   = { B = total branches, R = total node rows (= path vertices + B), 0, 0 }
 *   bmeta[B][4] = { spare row index, node count, parent branch id (component-local), flags }
 *                 flags: 1 = kept by prune, 2 = spare row holds the repair connection point, 4 = smoothed
 *   nodes[R][4] = float (x, y, z, raw radius); branch b owns rows row+1 .. row+len, `row` is its spare row
 *   smooth[R]   = float radius after the box filter (== raw radius where the branch was not smoothed)
 * Components appear in order; branches in emission order (= branch id).  smooth_kernel: 0 = off, else odd.
 * depth_workspace: n int32.                                                                             */
size_t st_finish_skeletons_out_ints(int64_t n);
int st_finish_skeletons(const float *medial_pts, const float *radii, const int32_t *comp_off,
                        int32_t n_comp, int64_t n, const int32_t *path_vertices,
                        const int32_t *branch_len, const int32_t *branch_parent,
                        const int32_t *comp_n_branches, const int32_t *comp_n_path, int prune_first,
                        float min_radius, float min_length, int repair, int smooth_kernel,
                        int32_t *depth_workspace, int32_t *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* ST_B200_H */
