/*
 * molkgnn_b200 -- C-ABI of the B200-native MolKGNN molecular-kernel convolution.
 *
 * Drop-in boundary for ONE hot path of LanceKnight/MolKGNN: the KernelConv / BaseKernelSetConv /
 * KernelSetConv stack driven by MolGCN, plus the degree-bucket pre-transform that feeds it.
 * The reference has no native code; every entry point below cites the Python interface
 * it replaces (paths relative to the reference repository).  Plain pointers and sizes only: all pointers
 * are DEVICE pointers unless the name ends in _host; the caller owns every buffer (the library never
 * allocates device memory); every call is asynchronous on `stream` (a cudaStream_t passed as void*)
 * unless stated.  Return value: 0 = ok, <0 = error (molkgnn_last_error() gives the text, thread-local).
 *
 * Layouts
 *   activations   x[N, ldx] fp32 row-major, ldx % 4 == 0, columns F..ldx-1 zero; xnorm[N] = ||x_row||_2
 *   buckets       degree d in 1..4; bucket rows r = 0..n_d-1 are the nodes of out-degree d in ascending node id
 *                 (wrapper.py:599-600); sel[boff_d + r] = node id; nei[eoff_d + r*d + j] = j-th neighbour in edge
 *                 order (wrapper.py:567-572); ehat[(eoff_d + r*d + j) * 8 + c] = bond attribute row
 *                 edge_attr[2*(eid/2)] (wrapper.py:586-591) divided by max(norm, 1e-8), zero padded to 8
 *   scores        compact: sc_d[r * L_d + k] per degree bucket (sc + scoff[d-1]);  dense: sc[node * ld + koff_d + k]
 *   argmax        uint8 per (r,k), same compact indexing: bits 0..6 = permutation index into the table of
 *                 kernels.py:109-128, bit 7 = chirality sign negative (kernels.py:396-400)
 *   packed params see molkgnn_packed_floats(): per degree, L2-normalised kernel rows in the order the kernels stage
 *                 them in shared memory
 */
#ifndef MOLKGNN_B200_H_
#define MOLKGNN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOLKGNN_MAX_DEG 4
#define MOLKGNN_EDGE_PAD 8      /* bond-attribute rows are padded to 8 floats (edge_attr_dim <= 8) */
#define MOLKGNN_COS_EPS 1e-8f   /* torch.nn.CosineSimilarity eps, kernels.py:189 */
#define MOLKGNN_TILE_NODES 128  /* nodes per molecule tile = one tcgen05 N / TMEM column block */

/* Degree-bucket plan of one collated batch.  Device arrays are filled by molkgnn_bucket_build();
 * n/boff/eoff are host copies of the per-degree counts (index d-1). */
typedef struct molkgnn_plan {
    int32_t N, E;
    int32_t n[4];              /* nodes per degree bucket */
    int32_t boff[4];           /* prefix of n      : bucket row offset into sel / tsign base for d=4 */
    int32_t eoff[4];           /* prefix of n_d * d: offset into nei / nei_eid / ehat rows */
    int32_t* deg;              /* [N]   out-degree                                   (wrapper.py:574-576) */
    int32_t* pos;              /* [N]   row of the node inside its bucket */
    int32_t* sel;              /* [N]   selected_index_deg1..4 concatenated          (wrapper.py:599-600) */
    int32_t* nei;              /* [E]   nei_index_deg1..4 concatenated               (wrapper.py:567-572,611) */
    int32_t* nei_eid;          /* [E]   edge id of (r,j) */
    float*   ehat;             /* [E,8] normalised neighbour bond attributes          (wrapper.py:578-593) */
    int8_t*  tsign;            /* [n_4] sign of p2.(p0 x p1) of the calibrated neighbour positions (kernels.py:336,356) */
    int32_t* in_cnt;           /* [N]   in-degree */
    int32_t* in_src;           /* [N,4] sources of the in-edges, edge order           (KernelLayer.py:119, PyG aggr='add') */
    int32_t* in_j;             /* [N,4] position of this node inside the source's neighbour list */
    /* molecule tiles (tile kernels): consecutive node ranges that no edge crosses, each <= MOLKGNN_TILE_NODES nodes.
     * Filled by molkgnn_bucket_build(); n_tiles == 0 means "no tiling" (a molecule larger than the tile, or a plan
     * built from bucket tensors) and the bucket-order kernels are used instead. */
    int32_t* tile_start;       /* [n_tiles + 1] first node of every tile; tile_start[n_tiles] = N (capacity N/32 + 4) */
    int32_t n_tiles;
    int32_t tile_max_nodes;    /* largest tile */
    int32_t tile_max_deg[4];   /* largest number of degree-d nodes in one tile */
    void*    tile_meta;        /* [N/32 + 4] records of molkgnn_tile_meta_bytes() bytes: per-tile node lists, local neighbour
                                  ids, bucket rows, in-lists (csrc/tile.cuh TileMetaG) */
    float*   ehat_node;        /* [E,8] the rows of ehat in node order (tile-contiguous) */
    int32_t* node_tile;        /* [N]   tile << 8 | row of the node inside its tile */
    int32_t* tile_order;       /* [N/32 + 4] balanced schedule of the tile-major backward kernels (nullable): the tiles of
                                  persistent CTA b of a tile_grid-CTA launch, CTA after CTA (csrc/tile.cuh TileWalk) */
    int32_t tile_grid;         /* CTAs the schedule was built for (0 = none: round robin) */
} molkgnn_plan_t;

/* One KernelSetConv layer (kernels.py:754-781): raw parameters of the four KernelConv modules + packed workspace. */
typedef struct molkgnn_layer {
    int32_t F;                 /* node_attr_dim of this layer */
    int32_t Fp;                /* F rounded up to a multiple of 4 (row stride of the packed kernel rows) */
    int32_t Fe;                /* edge_attr_dim (<= 8) */
    int32_t L[4];              /* kernels per degree (kernels.py:760) */
    int32_t koff[4];           /* column offset of degree block in the [N,K] score matrix (kernels.py:725-727) */
    int32_t K;                 /* sum(L) */
    const float* x_center[4];            /* [L,F]      kernels.py:58  */
    const float* x_support[4];           /* [L,d,F]    kernels.py:61  */
    const float* edge_attr_support[4];   /* [L,d,Fe]   kernels.py:65  */
    const float* p_support[4];           /* [L,d,3]    kernels.py:69  */
    const float* w_support[4];           /* scalar support_attr_sc_weight       kernels.py:79 */
    const float* w_center[4];            /* scalar center_attr_sc_weight        kernels.py:76 */
    const float* w_edge[4];              /* scalar edge_attr_support_sc_weight  kernels.py:82 */
    float* packed[4];                    /* workspace, molkgnn_packed_floats(d, L, Fp) floats each, 16-byte aligned */
    void* tile_img;                      /* workspace, molkgnn_tile_img_bytes(layer) bytes, 128-byte aligned (nullable:
                                            disables the tile kernels for this layer) */
} molkgnn_layer_t;

/* Parameter gradients of one layer, in the reference's own parameter layouts (what autograd would return). */
typedef struct molkgnn_layer_grads {
    float* x_center[4];
    float* x_support[4];
    float* edge_attr_support[4];
    float* w_support[4];
    float* w_center[4];
    float* w_edge[4];
} molkgnn_layer_grads_t;

const char* molkgnn_last_error(void);
int molkgnn_version(void);
/* number of kernels this library has launched so far in the process (bench.py reports it as gpu_launches) */
int64_t molkgnn_launch_count(void);
/* number of SMs of the current device (grid sizing), <0 on error */
int molkgnn_num_sms(void);

/* ---- degree bucketing: replaces ToXAndPAndEdgeAttrForDeg.__call__ (wrapper.py:559-672) + PyG collation ---- */
/* size of one per-tile metadata record (plan->tile_meta) */
int64_t molkgnn_tile_meta_bytes(void);
/* sizeof of the structs of this header as compiled into the library: 0 molkgnn_plan_t, 1 molkgnn_layer_t,
 * 2 molkgnn_layer_grads_t, 3 molkgnn_stack_layout_t (-1 otherwise).  A binding checks its mirror against these. */
int64_t molkgnn_struct_bytes(int32_t which);
/* Tile schedule of the tile-major backward kernels (plan->tile_order): 0 = round robin, 1 = balanced schedule for plans where
 * a few persistent CTAs would walk one tile more than the rest (default; env MOLKGNN_TILE_ORDER), 2 = always.  Returns the
 * previous mode.  Takes effect for plans built afterwards. */
int molkgnn_set_tile_order(int mode);
/* bytes of scratch needed by molkgnn_bucket_build */
int64_t molkgnn_bucket_scratch_bytes(int32_t N, int32_t E);
/* Builds the plan from a collated batch.  edge_index is the PyG [2,E] int64 tensor (row 0 = source).  SYNCHRONISES
 * the stream once to bring the four bucket sizes to the host (plan->n/boff/eoff).  Fails (-3) if a node has
 * out-degree outside 1..4 or in-degree > 4 (the reference silently mis-shapes its output, kernels.py:743-747). */
int molkgnn_bucket_build(molkgnn_plan_t* plan, const int64_t* edge_index, const float* p, int32_t p_dim,
                         const float* edge_attr, int32_t Fe, void* scratch, void* stream);
/* The same pass in two halves, so that the host can queue other work (parameter packing, its own bookkeeping) behind the
 * counting kernels before it blocks: _begin enqueues the counting kernels and returns; _finish (same arguments, same stream) brings the
 * bucket sizes to the host (the one stream synchronisation), validates them and enqueues the assignment kernels. */
int molkgnn_bucket_build_begin(molkgnn_plan_t* plan, const int64_t* edge_index, const float* p, int32_t p_dim,
                               const float* edge_attr, int32_t Fe, void* scratch, void* stream);
int molkgnn_bucket_build_finish(molkgnn_plan_t* plan, const int64_t* edge_index, const float* p, int32_t p_dim,
                                const float* edge_attr, int32_t Fe, void* scratch, void* stream);
/* _finish for a batch that carries the reference's precomputed per-degree DATA tensors (kernels.py:628-645).  The reference
 * convolves those raw tensors (BaseKernelSetConv.forward, kernels.py:679: nei_edge_attr_deg*; KernelConv, kernels.py:356:
 * nei_p_deg* - p_focal_deg*), never the edge_attr handed to MolGCN.forward -- which MolKGNNNet batch-normalises first
 * (MolKGNNNet.py:116-119).  Topology still comes from edge_index (bit-exact with the index tensors, wrapper.py:595-635);
 * the bond rows of bucket row r, slot j are nei_edge_attr[d-1][(r*d + j)*Fe ..], n_rows[d-1] = rows the caller holds
 * (must equal n_d * d, checked); p_focal4 [n_4,3] / nei_p4 [n_4,4,3] (nullable: use p) feed the chirality sign. */
int molkgnn_bucket_build_finish_ref(molkgnn_plan_t* plan, const int64_t* edge_index, const float* p, int32_t p_dim,
                                    const float* const nei_edge_attr[4], const int64_t n_rows[4], int32_t Fe,
                                    const float* p_focal4, const float* nei_p4, void* scratch, void* stream);
/* Writes the reference-format attributes of degree d (int64 indices, raw fp32 gathers), bit-exact with
 * wrapper.py:595-635 after PyG collation: selected_index[n_d], nei_index[n_d*d], p_focal[n_d,p_dim],
 * nei_p[n_d,d,p_dim], nei_edge_attr[n_d,d,Fe].  Any output pointer may be NULL. */
int molkgnn_bucket_export(const molkgnn_plan_t* plan, int32_t d, const float* p, int32_t p_dim,
                          const float* edge_attr, int32_t Fe, int64_t* selected_index, int64_t* nei_index,
                          float* p_focal, float* nei_p, float* nei_edge_attr, void* stream);
/* Builds a plan from reference-format bucket tensors (the attributes a PyG batch already carries,
 * kernels.py:628-645) instead of edge_index.  nei_p/p_focal/nei_edge_attr per degree as in the reference. */
int molkgnn_plan_from_buckets(molkgnn_plan_t* plan, const int64_t* const selected_index[4],
                              const int64_t* const nei_index[4], const float* const p_focal[4],
                              const float* const nei_p[4], int32_t p_dim, const float* const nei_edge_attr[4],
                              int32_t Fe, void* stream);

/* ---- activations ---- */
/* out[N,ldo] = x[N,ldx] zero padded (skipped if out == x), norm[N] = ||row||.  kernels.py:189-190 (norm half of cosine) */
int molkgnn_pad_norm(const float* x, int32_t N, int32_t F, int32_t ldx, float* out, int32_t ldo, float* norm,
                     void* stream);

/* ---- parameters ---- */
int64_t molkgnn_packed_floats(int32_t d, int32_t L, int32_t Fp);
/* Normalises the kernel rows of one layer into layer->packed (cosine denominators of kernels.py:189-190 on the
 * kernel side), evaluates the softmax mixing weights (kernels.py:402-412) and the support chirality signs
 * (kernels.py:338-341) for every permutation. */
int molkgnn_param_pack(const molkgnn_layer_t* layer, void* stream);
/* The same for nl layers in three launches (one per pack kernel, blockIdx.y = layer).  what: bit 0 = normalised rows,
 * weights, signs; bit 1 = operand images of the bucket-order tensor-core forward; bit 2 = kernel-block images of the
 * molecule-tile kernels.  molkgnn_param_pack(layer) == molkgnn_param_pack_layers(layer, 1, 7). */
int molkgnn_param_pack_layers(const molkgnn_layer_t* layers, int32_t nl, int32_t what, void* stream);
/* bytes of the fp16 (hi, lo) tensor-core operand images of the whole kernel set used by the tile kernels: resident-operand
 * layout for layers of <= 8 kernel blocks and <= 112 features; stage-major (forward) + K-step-major (backward) layouts for WIDE
 * layers (<= 16 degree-pure blocks, <= 512 features: csrc/conv_fwd_wide.cu, conv_bwd_wide.cu); 0 if neither applies */
int64_t molkgnn_tile_img_bytes(const molkgnn_layer_t* layer);
/* Normalised fp16 (hi, lo) images of the activations in tile order, the tensor-core operand of the tile kernels:
 * ximg holds plan->n_tiles records of molkgnn_tile_ximg_bytes() bytes (wide layers: the forward's and the backward's layouts,
 * all tiles of the first, then all tiles of the second).  Built once per layer, read by forward and backward. */
int64_t molkgnn_tile_ximg_bytes(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer);
int molkgnn_tile_ximg_build(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer, const float* x, int32_t ldx,
                            const float* xnorm, void* ximg, void* stream);
/* molkgnn_pad_norm + molkgnn_tile_ximg_build of the raw layer-0 input in one pass over x (node_attr_dim <= 64): x [N, ldx]
 * with F valid columns, out [N, Fp] zero padded, norm [N] (bitwise what molkgnn_pad_norm writes), ximg as above. */
int molkgnn_tile_ximg_build_raw(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer, const float* x, int32_t ldx,
                                float* out, float* norm, void* ximg, void* stream);

/* ---- forward: KernelConv.calculate_total_score for the four buckets (kernels.py:353-425, 610-751) ---- */
int64_t molkgnn_conv_fwd_smem_bytes(const molkgnn_layer_t* layer);
/* sc_mode 0: compact per-degree blocks at sc + scoff[d-1] (n_d*L_d floats each, scoff in floats);
 * sc_mode 1: dense sc[node*ld_sc + koff_d + k] (only the L_d entries of the node's own block are written).
 * argmax (compact, byte offsets = scoff) always receives the permutation actually used + chirality bit;
 * argmax_tile (nullable, molkgnn_tile_argmax_bytes() bytes): tile-ordered copy for the backward (tile kernel only);
 * argmax_free (nullable) receives the free-running arg-max; argmax_in (nullable) forces the permutation
 * (parity harness / replay).  counter: 8 int32 of scratch per call.  ximg (nullable): the activations' fp16 images
 * from molkgnn_tile_ximg_build(); with them, a plan that carries molecule tiles and an eligible layer the molecule-tile
 * tcgen05 kernel runs, otherwise the bucket-order kernels. */
int molkgnn_conv_fwd(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer, const float* x, int32_t ldx,
                     const float* xnorm, int32_t is_last_layer, float* sc, int32_t sc_mode, int32_t ld_sc,
                     const int64_t scoff[4], uint8_t* argmax, uint8_t* argmax_free, const uint8_t* argmax_in,
                     int32_t* counter, const void* ximg, uint8_t* argmax_tile, void* stream);
/* bytes of the optional tile-ordered copy of the arg-max (argmax_tile above / below): n_tiles slots of the fullest tile's
 * (node, kernel) pair count; the molecule-tile forward writes it, the molecule-tile backward fetches a tile's slot with one
 * bulk copy instead of per-pair reads of the bucket-order array.  0 if the plan / layer is not eligible. */
int64_t molkgnn_tile_argmax_bytes(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer);

/* Selects the forward kernel: 3 = layer-fused molecule-tile tcgen05 kernel for molkgnn_stack_fwd (default: ONE launch for all
 * layers, csrc/stack_fwd_fused.cu; needs a tiled plan and eligible layers, else -- and for the per-layer entry point -- 2),
 * 2 = per-layer molecule-tile tcgen05 kernel (needs ximg, a tiled plan and an eligible layer, else falls back to 1),
 * 1 = bucket-order tcgen05 kernel (falls back to 0 for layers whose kernel set does not fit shared memory), 0 = fp32 SIMT
 * kernel.  Returns the previous setting (-1 = not yet chosen). */
int molkgnn_set_fwd_path(int path);
/* the forward kernel currently selected (0 .. 3; resolves the MOLKGNN_FWD environment override on first use) */
int molkgnn_get_fwd_path(void);

/* ---- batch assembly from a packed molecule store: replaces the PyG DataLoader collate the reference relies on
 * (data.py:136-229; PyG Batch.from_data_list: attributes concatenated along __cat_dim__, keys containing "index" offset by
 * the number of nodes in front of the graph).  Store: x [sumN,F], p [sumN,P], edge_attr [sumE,Fe], edge_index [2,sumE] with
 * molecule-LOCAL node ids (row stride E_total), node_ptr / edge_ptr [M_total+1], optional targets y [M_total,Y].  ids [M]: the
 * molecules of the batch (any order, repeats allowed).  Outputs (caller-sized from the host copies of the pointers): x_out
 * [N_b,F], p_out [N_b,P], ea_out [E_b,Fe], ei_out [2,E_b] (row stride E_b, batch-global ids), batch_out [N_b], ptr_out [M+1],
 * y_out [M,Y]; node_off / edge_off [M+1] scratch; err: device word, bit 0 = id out of range (checked by the caller). ---- */
int molkgnn_collate(const int64_t* ids, int32_t M, int64_t M_total, const int64_t* node_ptr, const int64_t* edge_ptr,
                    const float* x, int32_t F, const float* p, int32_t P, const float* edge_attr, int32_t Fe,
                    const int64_t* edge_index, int64_t E_total, const float* y, int32_t Y, int64_t* node_off,
                    int64_t* edge_off, float* x_out, float* p_out, float* ea_out, int64_t* ei_out, int64_t E_b,
                    int64_t* batch_out, int64_t* ptr_out, float* y_out, int32_t* err, void* stream);

/* global_add_pool of the caller of the stack (MolKGNNNet.py:59,144-146): out[g, :] = sum of rows ptr[g] .. ptr[g+1]-1 of z
 * [N, C] (row stride ldz), added in node order -- deterministic, no atomics.  ptr [B+1] as produced by molkgnn_collate. */
int molkgnn_segment_sum(const float* z, int32_t C, int32_t ldz, const int64_t* ptr, int32_t B, float* out, void* stream);

/* ---- propagate: MolGCN.forward line `h = self.propagate(edge_index, sim_sc)` (KernelLayer.py:119-123) ---- */
/* h[i, koff_d + k] = sum over in-edges (j -> i) in edge order of sc_{deg j}[pos j, k]; columns K..ldh-1 zeroed;
 * hnorm[i] = ||h_i|| (nullable).  ximg (nullable; needs a tiled plan and ldh <= 112): additionally writes the normalised
 * fp16 (hi, lo) images of h in tile order, i.e. what molkgnn_tile_ximg_build() would produce for the next layer
 * (node_attr_dim = K, Fp = ldh), so that the next layer's activations are never re-read for the conversion. */
int molkgnn_propagate_fwd(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer, const float* sc,
                          const int64_t scoff[4], float* h, int32_t ldh, float* hnorm, void* ximg, void* stream);

/* ---- backward (the reference uses autograd over kernels.py:353-425 and KernelLayer.py:119) ---- */
int64_t molkgnn_conv_bwd_partial_floats(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer);
/* floats of the `coef` scratch: sum_d n_d*L_d for the bucket-order kernels, the tile-ordered coefficient + arg-max arrays
 * (n_tiles slots of the fullest tile's size) for the molecule-tile path */
int64_t molkgnn_conv_bwd_coef_floats(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer);
/* grad_mode 0: g[n,k] = grad[n*ldg + koff_d + k]                 (grad w.r.t. the dense score matrix)
 * grad_mode 1: g[n,k] = sum_{i in nei(n)} grad[i*ldg + koff_d + k] (grad w.r.t. the propagated h: fuses propagate^T)
 * coef: scratch of molkgnn_conv_bwd_coef_floats() floats, 16-byte aligned.  partials: scratch of
 * molkgnn_conv_bwd_partial_floats() floats.  grad_x (nullable) [N,ldgx] receives dL/dx including the cosine
 * normalisation Jacobian; columns F..ldgx-1 zeroed.  grads (nullable members skipped) receives dL/dparam.
 * phases: bit 0 = k_bwd_w (coef + per-CTA partial sums), bit 1 = parameter finalize, bit 2 = k_bwd_x (grad_x);
 * 7 runs everything; the split exists so that a profiler can time the three kernels separately.
 *
 * Molecule-tile tensor-core path (default when eligible): runs when ximg (molkgnn_tile_ximg_build of this layer's x, or
 * the images molkgnn_propagate_fwd wrote) is given, the plan carries tiles, the layer is eligible and ldgx == Fp.
 * scratch: N * roundup(Fp,16) floats (partial input gradients handed between the two launches of a 4-block layer).
 * Then phase bit 0 runs the whole backward (coefficients + bond gradients, parameter partial sums AND grad_x), bit 2 is
 * a no-op. */
int molkgnn_conv_bwd(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer, const float* x, int32_t ldx,
                     const float* xnorm, const float* grad, int32_t ldg, int32_t grad_mode, const uint8_t* argmax,
                     const int64_t scoff[4], float* coef, float* partials, float* grad_x, int32_t ldgx,
                     const molkgnn_layer_grads_t* grads, int32_t phases, const void* ximg, float* scratch,
                     const uint8_t* argmax_tile, void* stream);
/* 1 = molecule-tile tensor-core backward when eligible (default), 0 = bucket-order SIMT kernels; returns the old value */
int molkgnn_set_bwd_path(int path);
/* how often each path ran so far: forward tile / forward other / backward tile / backward other */
void molkgnn_path_counts(int64_t out[4]);

/* ---- the whole conv stack in one call: MolGCN.forward / its autograd backward (KernelLayer.py:107-120) ----
 * The per-layer entry points above stay the public building blocks; these two issue the same launches for every layer
 * from native code (one host call per pass instead of ~15 per layer) out of ONE caller-owned workspace. */
#define MOLKGNN_MAX_LAYERS 16
#define MOLKGNN_STACK_KEEP_SC 1      /* flags: keep every layer's compact scores (else one buffer is reused) */
#define MOLKGNN_STACK_WANT_FREE 2    /* flags: also record the free-running arg-max of every layer */
#define MOLKGNN_STACK_PACKED 4       /* flags: molkgnn_param_pack() already ran for every layer on this stream */
#define MOLKGNN_STACK_INFERENCE 8    /* flags: no backward follows (torch.no_grad): the layer-fused forward writes only h_out */
typedef struct molkgnn_stack_layout {
    int64_t fwd_bytes;                          /* forward workspace: lives from stack_fwd to stack_bwd */
    int64_t bwd_bytes;                          /* backward scratch: only during stack_bwd */
    int64_t grad_floats;                        /* flat parameter-gradient buffer, floats */
    /* byte offsets into the forward workspace (128-byte aligned); -1 = absent */
    int64_t h[MOLKGNN_MAX_LAYERS];              /* input activations of layer i, [N, Fp_i] fp32 (layer 0: padded x) */
    int64_t hnorm[MOLKGNN_MAX_LAYERS + 1];      /* their row norms [N]; [nl] = norms of the output */
    int64_t ximg[MOLKGNN_MAX_LAYERS];           /* fp16 (hi, lo) tile images of layer i's input */
    int64_t sc[MOLKGNN_MAX_LAYERS];             /* compact scores of layer i */
    int64_t argmax[MOLKGNN_MAX_LAYERS];         /* uint8 per (node, kernel) pair, compact */
    int64_t argmax_free[MOLKGNN_MAX_LAYERS];
    int64_t argmax_tile[MOLKGNN_MAX_LAYERS];    /* tile-ordered copy of the arg-max (molecule-tile path), -1 = absent */
    int64_t counter;
    int64_t sc_elems[MOLKGNN_MAX_LAYERS];       /* sum_d n_d * L_d of layer i */
    int64_t scoff[MOLKGNN_MAX_LAYERS][4];       /* element offsets of the per-degree blocks inside sc / argmax */
    /* byte offsets into the backward scratch */
    int64_t coef, partials, scratch, gx[2];
    /* float offsets into the flat gradient buffer, per layer and degree (d-1); g_w = 3 floats: support, centre, edge */
    int64_t g_x_center[MOLKGNN_MAX_LAYERS][4], g_x_support[MOLKGNN_MAX_LAYERS][4],
            g_edge_attr_support[MOLKGNN_MAX_LAYERS][4], g_w[MOLKGNN_MAX_LAYERS][4];
    int64_t partials_alt;                       /* second partial-copy buffer (backward scratch): layers alternate, so that
                                                   the parameter finalisation of layer i runs beside layer i-1's kernels */
} molkgnn_stack_layout_t;

/* Sizes and offsets for a built plan (plan->n / n_tiles valid) and nl layers (layers[i].F must chain: F_{i+1} = K_i). */
int molkgnn_stack_layout(const molkgnn_plan_t* plan, const molkgnn_layer_t* layers, int32_t nl, int32_t flags,
                         molkgnn_stack_layout_t* out);
/* Forward of the stack: parameter packing of every layer, pad + norm of x [N, ldx] (F = layers[0].F columns), then per
 * layer conv (all four buckets) and propagate.  h_out [N, ldh] (ldh = roundup4(K_last)) receives the output of the last
 * propagate.  argmax_in: nullable array of nl nullable pointers (forced permutations, parity harness).  tile_fwd (HOST,
 * nullable, nl ints): receives per layer whether the molecule-tile forward ran, i.e. whether the tile-ordered arg-max copy
 * in the workspace is valid; hand it to molkgnn_stack_bwd. */
int molkgnn_stack_fwd(const molkgnn_plan_t* plan, const molkgnn_layer_t* layers, int32_t nl,
                      const molkgnn_stack_layout_t* lay, int32_t flags, const float* x, int32_t ldx, void* workspace,
                      float* h_out, int32_t ldh, const uint8_t* const* argmax_in, int32_t* tile_fwd, void* stream);
/* Backward of the stack from grad_h [N, ldg] (gradient w.r.t. h_out).  grad_x (nullable) [N, Fp_0].  grad_flat
 * (nullable): flat parameter-gradient buffer laid out by molkgnn_stack_layout (g_* offsets). */
int molkgnn_stack_bwd(const molkgnn_plan_t* plan, const molkgnn_layer_t* layers, int32_t nl,
                      const molkgnn_stack_layout_t* lay, void* workspace, void* bwd_scratch, const float* grad_h,
                      int32_t ldg, float* grad_x, float* grad_flat, const int32_t* tile_fwd, void* stream);

/* ---- diagnostics ---- */
/* CUDA-event profiler of the library's own launches (bench.py: live per-kernel durations on the launching stream).
 * enable(1) clears and starts recording one (start, end) event pair per launch group; read() synchronises the device,
 * writes one line "name launches total_ms" per group name into buf and clears the records.  Returns the previous state /
 * the text length (<0 on error).  Not thread safe; off by default. */
int molkgnn_profile_enable(int on);
/* restrict recording to the launch groups of one name (NULL = all): keeps the event overhead out of a timed region */
int molkgnn_profile_only(const char* name);
int molkgnn_profile_read(char* buf, int cap);
/* Known-answer test of the tcgen05 (UMMA) plumbing: D[128,N] (fp32) = A * B^T on the tensor cores, one CTA.
 * A, B: fp16.  a_mn = 0: A is [128,K] row-major (K-major operand); a_mn = 1: A is [K,128] row-major (MN-major
 * operand); same for B with N.  swap = 1 exchanges the two stride fields of the descriptors (diagnostic only). */
int molkgnn_tc_selftest(const void* A, const void* B, float* D, int32_t N, int32_t K, int32_t a_mn, int32_t b_mn,
                        int32_t swap, void* stream);

/* ---- data-parallel step: one-shot all-reduce of the flat kernel-parameter gradient buffer over NVLink peer memory ----------
 * (csrc/oneshot.cu; the reference has no counterpart -- it trains on one GPU; replaces the ncclAllReduce of SURVEY 8(e) for
 * latency-sized buffers).  Every rank creates an exchange buffer, the 64-byte CUDA IPC handles are exchanged by the host
 * (molkgnn_b200/dp.py: torch.distributed all_gather), every rank opens its peers' buffers, and then one kernel per step sums
 * the W copies in rank order (bitwise identical results on all ranks).  Ranks must call _allreduce the same number of times. */
int molkgnn_oneshot_create(int32_t rank, int32_t world, int64_t bytes, void** handle);
int molkgnn_oneshot_ipc_handle(void* handle, void* out64);
int molkgnn_oneshot_open(void* handle, const void* handles /* world x 64 bytes in rank order */);
int molkgnn_oneshot_allreduce(void* handle, float* flat, int64_t n, int32_t average, void* stream);
int molkgnn_oneshot_error(void* handle);      /* 0, or r + 1 if rank r's flag timed out since the last call; synchronises */
/* queues a copy of the error word (same encoding) into 4 bytes of page-locked host memory behind the work on `stream`; no synchronisation */
int molkgnn_oneshot_error_async(void* handle, void* pinned_host4, void* stream);
int molkgnn_oneshot_destroy(void* handle);

#ifdef __cplusplus
}
#endif
#endif /* MOLKGNN_B200_H_ */
