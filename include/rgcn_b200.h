/*
 * rgcn_b200.h — C ABI of the B200-native RGCN relational message-passing engine.
 *
 * Drop-in boundary for the per-relation aggregation path of thiviyanT/torch-rgcn
 * (torch_rgcn/layers.py RelationalGraphConvolutionNC.forward :222-308,
 *  RelationalGraphConvolutionLP.forward :450-565, and the helpers in
 *  torch_rgcn/utils.py).  The reference has no FFI of its own: the path sits
 * behind a Python nn.Module.  The thin shim in torch_rgcn_b200/_lib.py binds
 * these entry points with ctypes and passes raw device pointers; no torch type
 * crosses this boundary.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the comment says "host";
 *   - pointers are borrowed for the duration of the call; nothing is allocated
 *     or freed behind the caller's back: scratch memory is caller-provided and
 *     sized by the matching *_workspace_bytes() query;
 *   - all work is enqueued on `stream` (a cudaStream_t cast to void*); calls are
 *     asynchronous and re-entrant per stream;
 *   - return value: 0 = ok, negative = error (see enum), message in
 *     rgcn_last_error() (thread-local);
 *   - indices: node ids and edge counts fit int32 inside the engine
 *     (N, nnz < 2^31); triples arrive as int64 (torch.long) like the reference.
 */
#ifndef RGCN_B200_H
#define RGCN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RGCN_ABI_VERSION 15
#define RGCN_CHUNK_EDGES 1024   /* edges per relation-major work chunk (r_chunkptr) */
#define RGCN_TILE_ROWS_PER_ITEM 256   /* rows per phase-2 work item of the tiled kernels */
#define RGCN_SPAN_EDGES 1024          /* edges per phase-1 work item (span) of the tiled kernels */
#define RGCN_MAX_RING_DEPTH 64        /* upper bound on rgcn_graph.ring_depth */
#define RGCN_LONG_ROW 512             /* rows with more edges are listed in d_long / s_long and processed cooperatively */
#define RGCN_FUSE_TILE 16             /* entries per MMA tile of the fused row-block lists (one relation per tile) */
#define RGCN_FUSE_REC_WORDS 36        /* int32 words per tile record of rgcn_fused.rec (144 bytes) */
#define RGCN_FUSE_AHEAD 8             /* a tile record names the relation of the tile this many places later */
#define RGCN_FUSE_MAX_ITEM_TILES (1 << 20)  /* upper bound on rgcn_graph.fuse_item_tiles */
#define RGCN_MAX_PEERS 8              /* ranks of one NVLink domain a row-sharded forward can store to */

typedef void* rgcn_stream_t;

enum rgcn_status {
    RGCN_OK = 0,
    RGCN_ERR_ARG = -1,          /* bad argument / shape */
    RGCN_ERR_UNSUPPORTED = -2,  /* combination the engine (like the reference) does not support */
    RGCN_ERR_CUDA = -3,         /* CUDA runtime error */
    RGCN_ERR_WORKSPACE = -4     /* workspace too small */
};

/* how per-edge weights are derived (reference layers.py:263-273 / :498-510) */
enum rgcn_norm {
    RGCN_NORM_ROW = 0,          /* vertical stacking: 1 / #{edges with the same (p, s)} */
    RGCN_NORM_COL_SWAPPED = 1,  /* horizontal stacking: 1 / permuted column counts, the literal
                                   cat([sums[n:2n], sums[:n], sums[-i:]]) rule */
    RGCN_NORM_EXPLICIT = 2      /* caller supplies val_in (e.g. a relation shard of a larger graph) */
};

enum rgcn_weight_form {
    RGCN_W_DENSE = 0,           /* weights (R', I, O)                         layers.py:155 */
    RGCN_W_BASIS = 1,           /* bases (B, I, O), comps (R', B)             layers.py:160-161, :242 */
    RGCN_W_BLOCK = 2,           /* blocks (Rb, nb, I/nb, O/nb) [+ blocks_self] layers.py:169-170, :375-378 */
    RGCN_W_DIAG = 3             /* weights (R', I), O == I                    layers.py:147-151 */
};

enum rgcn_dtype { RGCN_F32 = 0, RGCN_BF16 = 1 };

const char* rgcn_last_error(void);   /* host string, valid until the next failing call on this thread */
int rgcn_abi_version(void);

/* ------------------------------------------------------------------------------------------
 * Helper kernels — device versions of torch_rgcn/utils.py.
 * ---------------------------------------------------------------------------------------- */

/* utils.py:127-141  add_inverse_and_self: out is (2E+N, 3) = [triples; (o, p+R, s); (v, 2R, v)] */
int rgcn_add_inverse_and_self(const int64_t* triples, int64_t num_triples, int64_t num_nodes, int64_t num_rels,
                              int64_t* out, rgcn_stream_t stream);

/* utils.py:100-107  generate_inverses: out is (E, 3) = (o, p+R, s) */
int rgcn_generate_inverses(const int64_t* triples, int64_t num_triples, int64_t num_rels, int64_t* out,
                           rgcn_stream_t stream);

/* layers.py:481-487 + utils.py:110-124: the edge list the LP layer walks,
 * out is (3E + n_self, 3) = [T; inverse(T); T; (v, 2R, v) for v in self_nodes].
 * self_nodes lists the nodes whose self-loop survived edge dropout (all nodes in eval mode). */
int rgcn_lp_triples_plus(const int64_t* triples, int64_t num_triples, int64_t num_rels,
                         const int64_t* self_nodes, int64_t num_self, int64_t* out, rgcn_stream_t stream);

/* utils.py:143-166  stack_matrices: indices_out is (nnz, 2); vertical: (p*N+s, o), horizontal: (s, p*N+o).
 * bounds_out (2 x int64, device) receives max(indices[:,0]) and max(indices[:,1]) for the reference's asserts
 * (utils.py:163-164); may be NULL. */
int rgcn_stack_matrices(const int64_t* triples, int64_t nnz, int64_t num_nodes, int64_t num_rels, int vertical,
                        int64_t* indices_out, int64_t* bounds_out, rgcn_stream_t stream);

/* utils.py:71-97  sum_sparse: sums_out[k] = sum of values over the row (row_normalisation != 0) or column
 * of entry k.  table_ws: scratch of `rows` (or `cols`) floats. */
int rgcn_sum_sparse(const int64_t* indices, const float* values, int64_t nnz, int64_t rows, int64_t cols,
                    int row_normalisation, float* table_ws, float* sums_out, rgcn_stream_t stream);

/* utils.py:168-196  block_diag: blocks (R, nb, bi, bo) -> out (R, nb*bi, nb*bo), zero off the diagonal */
int rgcn_block_diag(const float* blocks, int64_t num_rels, int64_t num_blocks, int64_t block_in, int64_t block_out,
                    float* out, rgcn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Graph plan: the sorted edge lists the propagation kernels walk.
 * Replaces stack_matrices + sum_sparse + the COO constructor on every forward
 * (layers.py:255-279 / :490-516).  All arrays are caller-allocated device memory.
 * ---------------------------------------------------------------------------------------- */
/* Optional super-tiling of the rows (destination rows for the forward, source rows for the backward) so that
 * the per-edge messages of one tile fit a small ring that stays in L2.  Tile of row r = rowptr[r] / tile_edges.
 * Edges are sorted by (tile, relation, row); tile k owns positions [rowptr[tilerow[k]], rowptr[tilerow[k+1]]). */
typedef struct rgcn_tiling {
    int32_t* tilerow;       /* T+1: first row of every tile (tilerow[T] = N) */
    int32_t* row;           /* nnz: tile-side endpoint (forward: subject s, backward: object o) */
    int32_t* col;           /* nnz: the other endpoint (the row that is gathered) */
    int32_t* rel;           /* nnz: relation */
    int32_t* slot;          /* nnz: position of the edge in the row-major CSR of the tile side */
    float* val;             /* nnz */
    int32_t* stepptr;       /* T+lag+1: work-queue prefix: step j = spans of tile j, then row blocks of tile j-lag,
                               lag = ring_depth / 2 (a tile is summed only after `lag` later tiles were queued) */
    int32_t* slotneed;      /* T: row blocks of the earlier tiles that share ring slot k % ring_depth */
    int32_t* items;         /* 8 x int32 per work-queue item (see rgcn_tile_item), capacity = rgcn_tile_items_bound() */
} rgcn_tiling;

/* one entry of the tiled kernels' in-order work queue */
typedef struct rgcn_tile_item {
    int32_t kind;           /* 0 = transform span (<= RGCN_SPAN_EDGES consecutive edges of a tile), 1 = row-sum block */
    int32_t tile;
    int32_t a;              /* span: unused          | row block: first row */
    int32_t b;              /* span: first edge      | row block: end row   */
    int32_t c;              /* span: number of edges | row block: unused    */
    int32_t slot_bias;      /* first CSR position of the tile (message row 0 of its ring slot) */
    int32_t need;           /* counter value to wait for before starting */
    int32_t pad;
} rgcn_tile_item;

/* Optional fused row-block lists (bf16 features, 16x16 blocks in groups of four): rows are cut into blocks of `fuse_rows`
 * consecutive rows whose fp32 output tile lives in shared memory; the edges of a block are sorted by
 * (relation, row parity, row) and every (block, relation) run is dealt over whole 16-entry tiles, so that one MMA
 * tile never mixes relations.  Tiles are numbered block by block; a block with more than fuse_item_tiles tiles is
 * split into several work items, which then add their partial sums into the output with atomics ("shared"
 * items).  Work items are consecutive in tile order: item q covers tiles [items[4q+1], items[4q+2]). */
typedef struct rgcn_fused {
    int32_t* col;           /* cap: row of the gathered matrix per entry (tile * 16 + slot), -1 = padding */
    int32_t* rec;           /* cap / 16 tile records of RGCN_FUSE_REC_WORDS words, streamed to shared memory as they
                               are.  Slot s of a tile: word 4 (s % 8) + 2 (s / 8) = byte offset of the entry's row
                               in the block's accumulators, 256 * local row + 64 * (local row & 1), plus (low 4 bits)
                               the entry's rank among the entries of the same row in this tile; next word = bits of
                               the fp32 edge weight, 0 = padding.  Word 32: relation of the tile RGCN_FUSE_AHEAD
                               places later | (largest rank in this tile) << 24 (non-zero: entries share rows and
                               the kernel adds them rank by rank).  Word 33: relation of this tile.  Word 34:
                               largest rank. */
    int32_t* blk_tile;      /* NB + 1: first tile of every row block, NB = ceil(N / fuse_rows) */
    int32_t* items;         /* 4 x int32 per work item {block, first tile, end tile, shared};
                               capacity rgcn_fused_items_bound() */
    int32_t* meta;          /* 8 x int32: [0] work items, [1] tiles, [2] 1 if the tiles do not fit cap (list unusable),
                               [3] row blocks that were split, [4] tiles in which entries share a row */
} rgcn_fused;

typedef struct rgcn_graph {
    int64_t num_nodes;
    int64_t num_rels;       /* R' = number of relation ids the layer sees */
    int64_t nnz;            /* rows of triples_plus */
    /* destination-major CSR (rows = subject s), sorted by (s, p, o) — forward walk */
    int32_t* d_rowptr;      /* N+1 */
    int32_t* d_src;         /* nnz: object o */
    int32_t* d_rel;         /* nnz */
    float* d_val;           /* nnz */
    /* source-major CSR (rows = object o), sorted by (o, p, s) — backward-to-features walk */
    int32_t* s_rowptr;      /* N+1 */
    int32_t* s_dst;         /* nnz: subject s */
    int32_t* s_rel;         /* nnz */
    float* s_val;           /* nnz */
    /* relation-major list, sorted by (p, o, s) — relation-batched kernels (gathers walk X rows in order) */
    int32_t* r_relptr;      /* R'+1 */
    int32_t* r_dst;         /* nnz */
    int32_t* r_src;         /* nnz */
    float* r_val;           /* nnz */
    int32_t* r_dslot;       /* nnz: position of relation-major edge k in the destination-major list */
    int32_t* r_sslot;       /* nnz: position of relation-major edge k in the source-major list */
    int32_t* r_chunkptr;    /* R'+1: relation p owns chunks [r_chunkptr[p], r_chunkptr[p+1]) of RGCN_CHUNK_EDGES edges */
    float* val;             /* nnz, in the caller's edge order: the reference's `vals` (layers.py:273) */
    int32_t* status;        /* 8 x int32: [0] = number of triples with s, p or o out of range (utils.py:163-164),
                               [1] / [2] = largest destination / source tile in edges, [3] = kernel watchdog flag,
                               [4] / [5] = number of long destination / source rows, [6] = edges of the largest
                               relation */
    int32_t* d_long;        /* nnz / RGCN_LONG_ROW + 1: destination rows with more than RGCN_LONG_ROW edges */
    int32_t* s_long;        /* nnz / RGCN_LONG_ROW + 1: source rows with more than RGCN_LONG_ROW edges */
    int64_t num_long_dst;   /* host copy of status[4], or -1 if the caller did not read it back (kernels then launch
                               the upper bound nnz / RGCN_LONG_ROW of CTAs and exit early) */
    int64_t num_long_src;   /* host copy of status[5], or -1 */
    int64_t max_rel_edges;  /* host copy of status[6], or 0 if not read back: lets the dense tensor-core forward walk the
                               relation chunks quantile by quantile (destination ranges stay L2-resident) */
    int64_t tile_edges;     /* 0: no tiling (ft / bt unused) */
    int64_t num_tiles;      /* T = (nnz - 1) / tile_edges + 1 (trailing tiles may be empty) */
    int64_t tile_capacity;  /* host copy of max(status[1], status[2]), filled by the caller after the build */
    int64_t ring_depth;     /* message tiles in flight (2 .. RGCN_MAX_RING_DEPTH); set by the caller with tile_edges */
    rgcn_tiling ft;         /* forward tiling (destination rows) */
    rgcn_tiling bt;         /* backward tiling (source rows) */
    int64_t fuse_rows;      /* 0: no fused row-block lists (ff / fb unused); else rows per block (multiple of 16) */
    int64_t fuse_cap;       /* entries allocated per list (multiple of 16) */
    int64_t fuse_item_tiles;/* largest work item in tiles, 1 .. RGCN_FUSE_MAX_ITEM_TILES */
    int64_t fuse_dirs;      /* which lists to build: bit 0 = ff (forward), bit 1 = fb (feature gradient); the arrays of a
                               list that is not built may be NULL */
    int64_t fuse_items[2];  /* host copies of ff / fb meta[0] filled by the caller after the build;
                               0 = list unusable (overflow or not read back): the kernels fall back */
    int64_t fuse_split[2];  /* host copies of ff / fb meta[3] */
    int64_t fuse_tiles[2];  /* host copies of ff / fb meta[1] */
    rgcn_fused ff;          /* forward lists (blocks of destination rows, gathers X[o]) */
    rgcn_fused fb;          /* backward lists (blocks of source rows, gathers grad_out[s]) */
} rgcn_graph;

size_t rgcn_graph_workspace_bytes(int64_t nnz, int64_t num_nodes, int64_t num_rels, int64_t tile_edges,
                                  int64_t fuse_rows);
/* upper bound on the number of work-queue items of one tiling (host-side sizing of rgcn_tiling.items) */
int64_t rgcn_tile_items_bound(int64_t nnz, int64_t num_nodes, int64_t num_rels, int64_t tile_edges);
/* length of rgcn_tiling.stepptr for a given tiling */
int64_t rgcn_tile_steps_len(int64_t nnz, int64_t tile_edges, int64_t ring_depth);

/* host-side sizing of rgcn_fused.items */
int64_t rgcn_fused_items_bound(int64_t num_nodes, int64_t fuse_rows, int64_t fuse_cap, int64_t fuse_item_tiles);

/* n_general / n_self are the (n, i) of the horizontal permutation: NC ((nnz-N)/2, N), LP (|T|, |T|+|self|).
 * val_in (nnz floats, caller order) is read only for RGCN_NORM_EXPLICIT. */
int rgcn_graph_build(const int64_t* triples, int64_t nnz, int64_t num_nodes, int64_t num_rels,
                     int norm, int64_t n_general, int64_t n_self, const float* val_in,
                     rgcn_graph* graph, void* workspace, size_t workspace_bytes, rgcn_stream_t stream);
/* graph->tile_edges > 0 asks rgcn_graph_build to also fill graph->ft / graph->bt (arrays caller-allocated);
 * graph->fuse_rows > 0 asks for graph->ff / graph->fb (fuse_cap and fuse_item_tiles set by the caller). */

/* ------------------------------------------------------------------------------------------
 * Propagation: out[s] = bias + sum_e val_e * T_{p_e}(X[o_e]) and its gradients.
 * ---------------------------------------------------------------------------------------- */
typedef struct rgcn_params {
    int32_t form;           /* enum rgcn_weight_form */
    int32_t featureless;    /* X = identity: in_dim == num_nodes, messages are weight rows (layers.py:286-288) */
    int64_t in_dim;
    int64_t out_dim;
    int64_t num_bases;
    int64_t num_blocks;
    int64_t num_block_rels; /* relations stored in `blocks`: R' (NC) or R'-1 (LP, self relation dense) */
    const float* weights;   /* DENSE (R', I, O) | DIAG (R', I) */
    const float* bases;     /* (B, I, O) */
    const float* comps;     /* (R', B) */
    const float* blocks;    /* (num_block_rels, nb, I/nb, O/nb) */
    const float* blocks_self; /* (I, O) dense weight of relation R'-1, or NULL   layers.py:378, :544 */
    const float* bias;      /* (O) or NULL */
    const float* self_mask; /* (N, O) or NULL: dropout mask on the transformed features of relation R'-1
                               ('schlichtkrull-dropout', layers.py:545-546) */
    int32_t out_dtype;      /* rgcn_forward only: enum rgcn_dtype of `out`.  RGCN_BF16 is accepted on the fused row-block
                               path only (bf16 features, a usable ff list without split blocks): a row-sharded layer
                               writes its rows straight into the bf16 all-gather buffer */
    int32_t pad_;
    int64_t row_lo;         /* rgcn_forward, fused row-block path: write output rows [row_lo, row_hi) only (multiples of */
    int64_t row_hi;         /* fuse_rows; 0, 0 = all rows; the plan must hold no edge into other rows).  rgcn_backward,
                               bf16 tensor-core path with a bf16 feature gradient: only the feature-gradient rows
                               [row_lo, row_hi) are written (the plan must hold no edge out of other rows) */
    void* peer_out[8];      /* (RGCN_MAX_PEERS entries) rgcn_forward only, fused row-block path with a bf16 output: when num_peer_out > 0
                               every output row is stored to the same offset of ALL these (N, 64) bf16 buffers instead
                               of `out` -- the symmetric exchange buffers of the ranks of a row-sharded layer, mapped
                               over NVLink (peer-to-peer stores issued by the kernel's flush: the all-gather of the
                               output rows is part of the kernel).  The caller synchronises the ranks afterwards. */
    int32_t num_peer_out;
    int32_t pad2_;
} rgcn_params;

typedef struct rgcn_grads { /* NULL = gradient not wanted; buffers are overwritten, not accumulated */
    void* features;         /* (N, I), fp32 unless features_dtype says bf16 */
    float* weights;
    float* bases;
    float* comps;
    float* blocks;
    float* blocks_self;
    float* bias;
    int32_t features_dtype; /* enum rgcn_dtype of the `features` buffer (RGCN_BF16 only with bf16 input features) */
} rgcn_grads;

size_t rgcn_forward_workspace_bytes(const rgcn_graph* graph, const rgcn_params* params, int feature_dtype);
int rgcn_forward(const rgcn_graph* graph, const rgcn_params* params, const void* features, int feature_dtype,
                 float* out, void* workspace, size_t workspace_bytes, rgcn_stream_t stream);

size_t rgcn_backward_workspace_bytes(const rgcn_graph* graph, const rgcn_params* params, int feature_dtype);
int rgcn_backward(const rgcn_graph* graph, const rgcn_params* params, const void* features, int feature_dtype,
                  const float* grad_out, const rgcn_grads* grads, void* workspace, size_t workspace_bytes,
                  rgcn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * DistMult decoder of the link-prediction models.
 * triples: (B, 3) int64 rows (s, p, o); nodes (N, dim) fp32; relations (R, dim) fp32.
 * status (1 x int32, device, may be NULL) is incremented once per triple with an index out of range; such
 * triples score 0 and contribute nothing (the reference raises an IndexError / device assert).
 * ---------------------------------------------------------------------------------------- */
/* layers.py:86-98  DistMult.forward: scores[b] = sum_d nodes[s,d] relations[p,d] nodes[o,d] (+ the three biases,
 * all NULL or all given) */
int rgcn_distmult_forward(const int64_t* triples, int64_t num_triples, const float* nodes, int64_t num_nodes,
                          const float* relations, int64_t num_rels, int64_t dim, const float* sbias,
                          const float* pbias, const float* obias, float* scores, int32_t* status,
                          rgcn_stream_t stream);
/* gradients of the above (autograd upstream); NULL = not wanted, buffers are overwritten */
int rgcn_distmult_backward(const int64_t* triples, int64_t num_triples, const float* nodes, int64_t num_nodes,
                           const float* relations, int64_t num_rels, int64_t dim, const float* grad_scores,
                           float* g_nodes, float* g_relations, float* g_sbias, float* g_pbias, float* g_obias,
                           rgcn_stream_t stream);
/* layers.py:77-84  DistMult.s_penalty: out[0] = mean(nodes[s]^2) + mean(relations[p]^2) + mean(nodes[o]^2),
 * computed from occurrence counts of the nodes / relations in the batch (kept in the workspace for the backward) */
size_t rgcn_distmult_penalty_workspace_bytes(int64_t num_nodes, int64_t num_rels);
int rgcn_distmult_penalty(const int64_t* triples, int64_t num_triples, const float* nodes, int64_t num_nodes,
                          const float* relations, int64_t num_rels, int64_t dim, float* out, int32_t* status,
                          void* workspace, size_t workspace_bytes, rgcn_stream_t stream);
/* workspace: as left by rgcn_distmult_penalty for the same batch; grad: device scalar (d loss / d penalty) */
int rgcn_distmult_penalty_backward(const void* workspace, int64_t num_triples, const float* nodes, int64_t num_nodes,
                                   const float* relations, int64_t num_rels, int64_t dim, const float* grad,
                                   float* g_nodes, float* g_relations, rgcn_stream_t stream);
/* utils/misc.py:174-189  negative_sampling's masked assignment: batch (count, 3) int64 in place,
 * batch[i, head_mask[i] ? 0 : 2] = corruptions[i] */
int rgcn_corrupt_triples(int64_t* batch, const uint8_t* head_mask, const int64_t* corruptions, int64_t count,
                         rgcn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Filtered ranking evaluation: utils/misc.py:29-110 (generate_true_dict, filter_scores,
 * evaluate) with DistMult scores, on node embeddings the caller computed once.
 * ---------------------------------------------------------------------------------------- */
/* Sorted lists of the known true completions, replacing the Python dictionaries of generate_true_dict:
 * head != 0: keys[e] = p * N + o, vals[e] = s (true heads of (p, o));  head == 0: keys = p * N + s, vals = o.
 * Entries are sorted by (key, value); all_triples (M, 3) int64; keys / vals have M entries. */
size_t rgcn_rank_filter_workspace_bytes(int64_t num_true);
int rgcn_rank_build_filter(const int64_t* all_triples, int64_t num_true, int64_t num_nodes, int64_t num_rels, int head,
                           uint64_t* keys, int32_t* vals, int32_t* status, void* workspace, size_t workspace_bytes,
                           rgcn_stream_t stream);
/* ranks[i] = #{candidates scoring above the target} + (#{ties, target included} - 1) / 2 + 1 over all num_nodes
 * head (head != 0) or tail completions of queries[i], other known true completions removed when filter lists are
 * given (num_true > 0).  Scores are DistMult scores (biases all NULL or all given). */
size_t rgcn_rank_workspace_bytes(int64_t num_queries, int64_t dim);
int rgcn_rank_triples(const int64_t* queries, int64_t num_queries, int head, const float* nodes, int64_t num_nodes,
                      const float* relations, int64_t num_rels, int64_t dim, const float* sbias, const float* pbias,
                      const float* obias, const uint64_t* filter_keys, const int32_t* filter_vals, int64_t num_true,
                      int64_t* ranks, int32_t* status, void* workspace, size_t workspace_bytes, rgcn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Per-step graph construction for link prediction: utils/misc.py:125-172 (edge_neighborhood),
 * :121-123 (uniform_sampling) and experiments/predict_links.py:143-148 (general edge dropout).
 * ---------------------------------------------------------------------------------------- */
/* Vertex adjacency of the training triples (E, 3) int64 in the reference's order (utils/misc.py:129-132): entries of
 * vertex v = adj[adj_ptr[v] .. adj_ptr[v+1]), by edge index, the subject's entry before the object's; entry k is the
 * int32 pair adj[2k] = edge index, adj[2k+1] = other end.  adj_ptr (N + 1), adj (2 E, 2), 8-byte aligned.  Built once
 * per training set. */
size_t rgcn_sampler_build_workspace_bytes(int64_t num_edges);
int rgcn_sampler_build(const int64_t* triples, int64_t num_edges, int64_t num_nodes, int32_t* adj_ptr, int32_t* adj,
                       int32_t* status, void* workspace, size_t workspace_bytes, rgcn_stream_t stream);
/* Edge-neighbourhood sample of sample_size <= num_edges distinct edges: out_edges[i] = index of the i-th pick.
 * uniforms (sample_size, 2) fp32 in [0, 1): [i, 0] picks the vertex (inverse CDF over count x seen, or over count > 0
 * while no seen vertex has edges left), [i, 1] the not-yet-picked incident edge.  status must be zero on entry and
 * stays zero on success. */
size_t rgcn_sample_workspace_bytes(int64_t num_edges, int64_t num_nodes);
int rgcn_sample_edge_neighborhood(const int32_t* adj_ptr, const int32_t* adj, int64_t num_edges, int64_t num_nodes, const float* uniforms, int64_t sample_size,
                                  int32_t* out_edges, int32_t* status, void* workspace, size_t workspace_bytes,
                                  rgcn_stream_t stream);
/* out[k] = triples[index[k]] for k < n (index int32, or int64 when index_is_int64): picked edge numbers, a uniform
 * sample or the tail of a dropout permutation -> the (n, 3) graph.  Out-of-range indices are counted in status. */
int rgcn_take_triples(const int64_t* triples, int64_t num_rows, const void* index, int index_is_int64, int64_t n,
                      int64_t* out, int32_t* status, rgcn_stream_t stream);

/* out[i] = (float) in[i] for i < n: widens the bf16 exchange buffer of a row-sharded layer to the fp32 output the layer
 * returns (streaming 16-byte loads / stores).  n must be a multiple of 8, pointers 16-byte aligned. */
int rgcn_widen_rows(const void* in_bf16, int64_t n, float* out, rgcn_stream_t stream);

/* number of kernels the engine has launched on this process since load (bench.py's gpu_launches) */
int64_t rgcn_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Relation sharding (new capability; the reference is single-device).
 * Greedy longest-processing-time bin packing of relations onto `world` ranks by edge count.
 * rel_nnz and rel_to_rank are HOST arrays of length num_rels.
 * ---------------------------------------------------------------------------------------- */
int rgcn_shard_plan(const int64_t* rel_nnz, int64_t num_rels, int32_t world, int32_t* rel_to_rank);

#ifdef __cplusplus
}
#endif
#endif /* RGCN_B200_H */
