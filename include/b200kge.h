/*
 * b200kge.h — C ABI of the B200-native KGE scoring engine (libb200kge.so).
 *
 * This is the drop-in boundary for ONE path of uma-pi1/kge (LibKGE): embedding-row gather +
 * relational scorer forward (+ fused BCE/KL loss, rank/tie counting, negative-sample scoring)
 * behind KgeModel.score_spo/score_sp/score_po/score_sp_po and RelationalScorer.score_emb.
 * The reference has no FFI for this path (it is PyTorch tensor expressions); every entry point
 * below names the reference function (file:line under /root/reference) whose arithmetic it
 * replaces.  INTEGRATION.md shows the ctypes binding a LibKGE maintainer would add.
 *
 * Conventions
 *  - Plain C: raw device pointers, sizes, a cudaStream_t passed as void*.  No torch types.
 *  - All float data is fp32, row-major.  Index arrays are int64 (LibKGE collates `.long()`,
 *    train_1vsAll.py:34) and live on the device unless the name says `_host`.
 *  - A "rows view" (b200kge_rows_t) is how an embedding operand is passed: row i of the operand is
 *        base + (idx ? idx[i] : i) * ld
 *    so the same entry point serves KgeModel.score_* (base = embedding table, idx = batch indexes:
 *    the LookupEmbedder gather lookup_embedder.py:96-97 is fused) and RelationalScorer.score_emb
 *    (base = already-gathered [n,D] matrix, idx = NULL).  idx == NULL with rows == vocab is
 *    "embed_all" (lookup_embedder.py:99-112) without the table copy.
 *  - Outputs and workspace are caller-allocated; nothing is retained across calls; the library
 *    keeps no global device state and is re-entrant per stream.
 *  - Every function returns 0 on success or a negative b200kge_status; b200kge_last_error() gives a
 *    thread-local message.  CUDA allocation failures are reported with the literal text
 *    "CUDA out of memory" so LibKGE's sub-batch auto-tuner (train.py:384-413) keeps working.
 */
#ifndef B200KGE_H_
#define B200KGE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200KGE_VERSION 100

typedef void* b200kge_stream_t; /* cudaStream_t */

typedef enum {
  B200KGE_OK = 0,
  B200KGE_ERR_INVALID = -1,     /* bad argument (ValueError on the Python side)            */
  B200KGE_ERR_UNSUPPORTED = -2, /* valid but not handled by this build (e.g. D % 4 != 0)   */
  B200KGE_ERR_CUDA = -3,        /* CUDA runtime error; message holds cudaGetErrorString    */
  B200KGE_ERR_WORKSPACE = -4,   /* workspace too small; see b200kge_workspace_bytes        */
  B200KGE_ERR_NO_DEVICE = -5    /* no sm_100 device: the library never falls back to a CPU */
} b200kge_status;

/* Scorers on the path (kge/model/<name>.py). */
typedef enum {
  B200KGE_COMPLEX = 0,  /* complex.py:18-43   */
  B200KGE_DISTMULT = 1, /* distmult.py:13-25  */
  B200KGE_SIMPLE = 2,   /* simple.py:13-33    */
  B200KGE_CP = 3,       /* cp.py:13-30        */
  B200KGE_RESCAL = 4,   /* rescal.py:14-52    */
  B200KGE_TRANSE = 5,   /* transe.py:15-37    */
  B200KGE_ROTATE = 6    /* rotate.py:20-69    */
} b200kge_model;

/* `combine` of RelationalScorer.score_emb (kge_model.py:151-181) for the 1-vs-N forms. */
typedef enum {
  B200KGE_SP_ = 0, /* "sp_": out[i,j] = score(s_i, p_i, cand_j) */
  B200KGE__PO = 1  /* "_po": out[i,j] = score(cand_j, p_i, o_i) */
} b200kge_combine;

/* Which kernel family computes dot-product scorers. */
typedef enum {
  B200KGE_PREC_AUTO = 0,   /* F16X3 for dot-product scorers with 32 <= K <= 1024 and n >= 16, else fp32 SIMT */
  B200KGE_PREC_FP32 = 1,   /* CUDA-core fp32 FFMA (bit-for-bit fp32 products)                    */
  B200KGE_PREC_3XTF32 = 2, /* tcgen05 tensor cores, hi/lo split, fp32-equivalent (~3e-6 of rms)  */
  B200KGE_PREC_TF32 = 3,   /* tcgen05 single pass (~4e-3 of rms; does NOT meet the 1e-4 bar)     */
  B200KGE_PREC_TF32_BF16X2 = 4, /* tf32 hi*hi + two bf16 cross terms: 8 MMAs per 32-wide K chunk instead
                              of 12, operand error ~2^-20 (below the accumulator's)               */
  B200KGE_PREC_F16X3 = 5   /* operands split ONCE per call into row-scaled fp16 hi/lo planes (22 significant
                              bits), hi*hi + hi*lo + lo*hi on the f16 tensor pipe: 6 MMA slots per 32
                              reduction elements, no shared-memory round trip in the main loop    */
} b200kge_precision;

typedef enum {
  B200KGE_LOSS_BCE = 1, /* BCEWithLogitsKgeLoss, reduction sum, + offset  loss.py:153-159 */
  B200KGE_LOSS_KL = 2   /* KLDivWithSoftmaxKgeLoss (CE for index labels)   loss.py:198-213 */
} b200kge_loss;

typedef struct {
  const float* base;  /* device */
  const int64_t* idx; /* device, may be NULL */
  int64_t rows;       /* number of rows of the operand (length of idx, or table rows) */
  int64_t ld;         /* row stride in floats */
  int32_t dim;        /* row width in floats */
} b200kge_rows_t;

/* Labels of a 1-vs-N loss: exactly one of idx / dense is non-NULL.
 * idx   [n]     position of the single 1 per row (1vsAll; loss.py:105-117 makes it one-hot)
 * dense [n,ldl] label matrix (KvsAll multi-hot incl. label smoothing train_KvsAll.py:242-266,
 *               negative sampling [1,0,...] train_negative_sampling.py:128-137)            */
typedef struct {
  const int64_t* idx;
  const float* dense;
  int64_t ldl;
} b200kge_labels_t;

/* Library / device ---------------------------------------------------------------------------- */
int b200kge_version(void);
const char* b200kge_last_error(void);
/* 0 if the current device is sm_100 (B200), else B200KGE_ERR_NO_DEVICE. */
int b200kge_device_ok(void);
/* Number of kernel launches issued by this library on the calling thread since the last reset
 * (bench.py reports it as gpu_launches). */
int64_t b200kge_launch_count(int reset);

/* Profiling aid (bench.py): when enabled, the dominant pairwise kernel of every 1-vs-N call on this
 * thread is bracketed by CUDA events recorded on the stream it is launched on;
 * b200kge_profile_last_ms synchronises on the closing event and returns that kernel's duration. */
int b200kge_profile_enable(int on);
int b200kge_profile_last_ms(float* ms);

/* Bytes of device workspace sufficient for any call below with n query rows (per direction), m
 * candidate rows and entity width D.  cand_has_idx != 0 reserves room to gather an index subset
 * of candidates for the tensor-core path. */
size_t b200kge_workspace_bytes(int model, int64_t n, int64_t m, int32_t D, int cand_has_idx);

/* Row-wise triples ---------------------------------------------------------------------------- */
/* out[i] = score(s_i, p_i, o_i).  Replaces KgeModel.score_spo kge_model.py:663-680 and
 * score_emb(combine="spo") of every in-scope scorer.  TransE adds eps=1e-6 to the difference like
 * F.pairwise_distance (transe.py:18). */
int b200kge_score_spo(int model, float l_norm, const b200kge_rows_t* s, const b200kge_rows_t* p,
                      const b200kge_rows_t* o, int64_t n, float* out, b200kge_stream_t stream);

/* 1-vs-N ---------------------------------------------------------------------------------------
 * out[i*ldo + j], i < n, j < cand->rows.  `q` are the per-row entity operands (subjects for sp_,
 * objects for _po), `p` the per-row relation operands, `cand` the candidate entities (all of them,
 * a contiguous chunk, or an index subset).  Replaces KgeModel.score_sp / score_po
 * kge_model.py:682-725 and score_emb(combine in {"sp_","_po"}). */
int b200kge_score_1vsN(int model, int combine, float l_norm, int precision,
                       const b200kge_rows_t* q, const b200kge_rows_t* p,
                       const b200kge_rows_t* cand, int64_t n, float* out, int64_t ldo,
                       void* workspace, size_t workspace_bytes, b200kge_stream_t stream);

/* out is [n, 2m] = [sp_ scores | _po scores], m = cand->rows.  Replaces KgeModel.score_sp_po
 * kge_model.py:749-789 (one launch sequence, no torch.cat copy). */
int b200kge_score_sp_po(int model, float l_norm, int precision, const b200kge_rows_t* s,
                        const b200kge_rows_t* p, const b200kge_rows_t* o,
                        const b200kge_rows_t* cand, int64_t n, float* out, int64_t ldo,
                        void* workspace, size_t workspace_bytes, b200kge_stream_t stream);

/* Fused 1-vs-N score + loss: the [n,m] scores never reach HBM.
 * loss_out[0] (device float) receives the SUM-reduced loss of the block exactly as
 * KgeLoss.__call__ returns it (the caller divides by batch size, train_1vsAll.py:65);
 * row_loss_out (optional, [n]) receives the per-row terms.  Replaces score_sp/score_po followed by
 * BCEWithLogitsKgeLoss / KLDivWithSoftmaxKgeLoss (train_1vsAll.py:64-65,75-76,
 * train_KvsAll.py:275-289). */
int b200kge_score_1vsN_loss(int model, int combine, float l_norm, int precision,
                            const b200kge_rows_t* q, const b200kge_rows_t* p,
                            const b200kge_rows_t* cand, int64_t n, const b200kge_labels_t* labels,
                            int loss_kind, float offset, float* loss_out, float* row_loss_out,
                            void* workspace, size_t workspace_bytes, b200kge_stream_t stream);

/* Fused 1-vs-N score + rank/tie counting against a chunk of candidates (additive over chunks,
 * eval_entity_ranking.py:222-229,310-313).  true_score[i] is the score of row i's true answer;
 * filter (optional) is the reference's dense label chunk [n, ldf] holding +inf at known-true
 * columns (own answer zeroed, :287-290) and is SUBTRACTED from the scores before comparing
 * (:561-566).  rank/ties are int64 [n] and are ACCUMULATED INTO (caller zeroes them before the
 * first chunk).  Replaces score_sp_po + _filter_and_rank + _get_ranks_and_num_ties :533-596. */
int b200kge_score_1vsN_rank(int model, int combine, float l_norm, int precision,
                            const b200kge_rows_t* q, const b200kge_rows_t* p,
                            const b200kge_rows_t* cand, int64_t n, const float* true_score,
                            const float* filter, int64_t ldf, float rtol, float atol,
                            int64_t* rank, int64_t* ties, void* workspace, size_t workspace_bytes,
                            b200kge_stream_t stream);

/* Both directions of the ranking of one batch in ONE launch sequence: rows 0..n-1 are the sp_ queries (ranked
 * against true_score[0..n)), rows n..2n-1 the _po queries (true_score[n..2n)); rank/ties are int64 [2n] in the
 * same order and are ACCUMULATED INTO; filter (optional) is [2n, ldf] in the same row order.  This is what
 * EntityRankingJob does per chunk with score_sp_po + _filter_and_rank + _get_ranks_and_num_ties
 * (eval_entity_ranking.py:222-229,533-596) without materialising the [n, 2m] scores.  CP (whose directions read
 * different candidate columns) returns B200KGE_ERR_UNSUPPORTED: call b200kge_score_1vsN_rank per direction. */
int b200kge_rank_sp_po(int model, float l_norm, int precision, const b200kge_rows_t* s,
                       const b200kge_rows_t* p, const b200kge_rows_t* o, const b200kge_rows_t* cand,
                       int64_t n, const float* true_score, const float* filter, int64_t ldf, float rtol,
                       float atol, int64_t* rank, int64_t* ties, void* workspace, size_t workspace_bytes,
                       b200kge_stream_t stream);

/* b200kge_rank_sp_po with the filter as CSR instead of a dense [2n, m] matrix: stacked row r lists the (sorted)
 * candidate columns filter_col[filter_off[r] .. filter_off[r+1]) that hold known answers; they are excluded from the
 * counts exactly like the reference's "+inf label subtracted" (eval_entity_ranking.py:489-531,561-566), except the
 * row's own answer own_col[r] (may be NULL; :287-290).  Columns are positions in `cand`, which must be a plain
 * table or chunk (cand->idx == NULL).  The CSR is consumed inside the scoring kernels' epilogues (pre-split tensor-core
 * kernels: one cursor per thread = row; CUDA-core kernel: one forward-moving cursor per owned row); CP and the in-kernel
 * split precision modes return B200KGE_ERR_UNSUPPORTED before anything is launched — pass the dense filter to
 * b200kge_rank_sp_po instead. */
int b200kge_rank_sp_po_csr(int model, float l_norm, int precision, const b200kge_rows_t* s,
                           const b200kge_rows_t* p, const b200kge_rows_t* o, const b200kge_rows_t* cand,
                           int64_t n, const float* true_score, const int64_t* filter_off,
                           const int64_t* filter_col, const int64_t* own_col, float rtol, float atol,
                           int64_t* rank, int64_t* ties, void* workspace, size_t workspace_bytes,
                           b200kge_stream_t stream);

/* b200kge_score_sp_po FUSED WITH THE ALL-GATHER of an entity-sharded table: `cand` is this rank's shard; the two
 * halves are written at out[i*ldo + j] (sp_) and out[i*ldo + col_block + j] (_po), j < cand->rows, AND at the same
 * offsets into each of the n_peers (<= 7) buffers peer_out[g] — peer-mapped device pointers to the other ranks'
 * symmetric output buffers (NVLink / NVSwitch).  With out = base + lo (lo = first global row of the shard),
 * ldo = 2 * E_total and col_block = E_total every rank's [n, 2E_total] logits matrix (kge_model.py:749-789 layout)
 * is complete once all ranks have passed a barrier: the kernel's own epilogue stores replace ncclAllGather and the
 * re-layout copy.  CP is not offered. */
int b200kge_score_sp_po_bcast(int model, float l_norm, int precision, const b200kge_rows_t* s,
                              const b200kge_rows_t* p, const b200kge_rows_t* o, const b200kge_rows_t* cand,
                              int64_t n, float* out, float* const* peer_out, int n_peers, int64_t ldo,
                              int64_t col_block, void* workspace, size_t workspace_bytes,
                              b200kge_stream_t stream);

/* Entity-sharded tables: this rank owns global rows [lo, lo + shard->rows) of the entity table.
 * out[i, :] = shard row (idx[i] - lo) if the rank owns global id idx[i], else zeros — the contribution of this rank
 * to the query-row exchange (sum over ranks == the gathered rows, exactly: every other rank adds zeros).  One
 * kernel, no host synchronisation (the reference's LookupEmbedder.embed, lookup_embedder.py:96-97, on a
 * partitioned table). */
int b200kge_shard_gather_rows(const b200kge_rows_t* shard, int64_t lo, const int64_t* idx, int64_t n,
                              float* out, int64_t ldo, b200kge_stream_t stream);

/* Dense-score epilogues (for callers that already hold a score matrix) ------------------------ */
/* KgeLoss on a dense [n,m] score matrix: loss.py:153-159 (BCE) / :198-213 (KL). */
int b200kge_loss_dense(const float* scores, int64_t lds, int64_t n, int64_t m,
                       const b200kge_labels_t* labels, int loss_kind, float offset,
                       float* loss_out, float* row_loss_out, void* workspace,
                       size_t workspace_bytes, b200kge_stream_t stream);

/* _get_ranks_and_num_ties on a dense [n,m] score matrix (eval_entity_ranking.py:571-596), with the
 * optional filter subtraction of _filter_and_rank (:561-566).  Bit-exact integer outputs;
 * ACCUMULATES INTO rank/ties. */
int b200kge_rank_dense(const float* scores, int64_t lds, int64_t n, int64_t m,
                       const float* true_score, const float* filter, int64_t ldf, float rtol,
                       float atol, int64_t* rank, int64_t* ties, b200kge_stream_t stream);

/* Negative sampling -----------------------------------------------------------------------------
 * out[i*ldo + k] = score of triple i with slot `slot` (0=S,1=P,2=O) replaced by neg[i*K + k].
 * The gather of the sampled rows is fused with the per-negative dot/distance.  If
 * with_positive != 0, column 0 of out receives score_spo of the positive triple and negatives go
 * to columns 1..K (the [n,1+K] assembly of train_negative_sampling.py:139-148).
 * Replaces BatchNegativeSample.score sampler.py:263-344 (both `triple` and `batch`
 * implementations give the same numbers; this computes them directly). */
int b200kge_ns_score(int model, float l_norm, const b200kge_rows_t* s, const b200kge_rows_t* p,
                     const b200kge_rows_t* o, const b200kge_rows_t* slot_table, int slot,
                     const int64_t* neg, int64_t n, int64_t K, int with_positive, float* out,
                     int64_t ldo, b200kge_stream_t stream);

/* On-device uniform negative sampling: out[i*K + k] ~ U{0, ..., vocab-1}, the device counterpart of
 * KgeUniformSampler._sample (kge/util/sampler.py:588-596: torch.randint on the CPU + a host->device copy of the ids).
 * Counter-based Philox4x32-10: the result depends on (seed, offset, position) only; use a fresh `offset` per call
 * (e.g. a batch counter) for independent draws.  No filtering of positives (sampler.py default filtering off). */
int b200kge_sample_uniform(uint64_t seed, uint64_t offset, int64_t vocab, int64_t n, int64_t K, int64_t* out,
                           b200kge_stream_t stream);

/* One whole 1vsAll forward step (train_1vsAll.py:48-82) for a batch of triples [n,3] (int64,
 * row-major s,p,o): fused score_sp+loss and score_po+loss against the whole entity table, both
 * directions stacked into one launch of 2n query rows where the model allows it.  loss_out[0]
 * (device) receives  (loss(score_sp, o) + loss(score_po, s)) / n.  `ent`/`rel` are the
 * device-resident tables (idx must be NULL). */
int b200kge_train_1vsall_forward(int model, float l_norm, int precision,
                                 const b200kge_rows_t* ent, const b200kge_rows_t* rel,
                                 const int64_t* triples, int64_t n, int loss_kind, float offset,
                                 float* loss_out, void* workspace, size_t workspace_bytes,
                                 b200kge_stream_t stream);

/* Host-buffer form of the same step (end-to-end measurement, embedding in a host-side loop):
 * copies triples_host [n,3] to the device (triples.to(device), train_1vsAll.py:59), runs
 * b200kge_train_1vsall_forward, copies the scalar back to *loss_host (.item(), :66,77) and
 * synchronises the stream. */
int b200kge_train_1vsall_forward_host(int model, float l_norm, int precision,
                                      const b200kge_rows_t* ent, const b200kge_rows_t* rel,
                                      const int64_t* triples_host, int64_t n, int loss_kind,
                                      float offset, float* loss_host, void* workspace,
                                      size_t workspace_bytes, b200kge_stream_t stream);

/* ---- host-side label plumbing (CPU; no device involved) ------------------------------------------
 * The key -> all-values index that KvsAll training and filtered entity ranking build their label and
 * filter coordinates from, in CSR form (row offsets + column ids) — what the device epilogues take
 * instead of coord_to_sparse_tensor(...).to_dense() (kge/job/util.py:32-60).
 *
 * b200kge_kvsall_index_build replaces KvsAllIndex.__init__ (kge/indexing.py:19-55,178-194):
 *   triples [n,3] (host), key columns (key_col0, key_col1) and value column = a permutation of 0,1,2.
 *   keys_out [n,2], offsets_out [n+1], values_out [n] are caller-allocated at capacity n;
 *   *num_keys receives the number of distinct keys.  Keys ascend lexicographically (np.unique axis=0),
 *   values ascend within a key, duplicate triples keep their duplicate values — element for element
 *   the reference's _keys / _values_offset / _values. */
int b200kge_kvsall_index_build(const int64_t* triples, int64_t n, int key_col0, int key_col1,
                               int value_col, int64_t* keys_out, int64_t* offsets_out,
                               int64_t* values_out, int64_t* num_keys);

/* KvsAllIndex.get_all (kge/indexing.py:113-166) and get_sp_po_coords_from_spo_batch
 * (kge/job/util.py:6-30): for each of nq query keys [nq,2] the values of that key (none if the key
 * is absent), as CSR: offsets_out [nq+1] (offsets_out[nq] = nnz), cols_out [nnz] = value +
 * col_shift (col_shift = num_entities for the po half of a [n, 2E] label matrix).  Call with
 * cols_out = NULL to size, then again with the buffer. */
int b200kge_kvsall_lookup(const int64_t* keys, const int64_t* offsets, const int64_t* values,
                          int64_t num_keys, const int64_t* query_keys, int64_t nq,
                          int64_t col_shift, int64_t* offsets_out, int64_t* cols_out);

/* The collate step of KvsAll training for one query type (kge/job/train_KvsAll.py:116-203): example
 * ids are key indexes; queries_out [nb,2] receives their keys, offsets_out [nb+1] / cols_out [nnz]
 * their labels as CSR (cols_out = NULL to size). */
int b200kge_kvsall_gather(const int64_t* keys, const int64_t* offsets, const int64_t* values,
                          int64_t num_keys, const int64_t* examples, int64_t nb,
                          int64_t* queries_out, int64_t* offsets_out, int64_t* cols_out);

/* ---- gradients, penalties, CSR labels (validated on a B200 in round 2) -------------
 *
 * b200kge_gemm_nt: C[M,N] = A[M,K] * B[N,K]^T, fp32 in / fp32 out, computed on the f16 tensor pipe from
 * hi/lo fp16 planes split once in HBM (presplit.cu + pairwise_tc3.cu) — the building block of the
 * backward GEMMs; fp32-equivalent (operand error ~5e-7 of the result's rms).  Reductions longer than 512 run
 * split-K: 512-element segments accumulated in fp32, which bounds the tensor core's accumulator error. */
size_t b200kge_gemm_nt_workspace_bytes(int64_t M, int64_t N, int64_t K);
int b200kge_gemm_nt(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t M, int64_t N,
                      int64_t K, float* C, int64_t ldc, void* workspace, size_t workspace_bytes,
                      b200kge_stream_t stream);

/* Backward of b200kge_train_1vsall_forward (loss.backward() at kge/job/train_1vsAll.py:70,81) with BCE or KL, for the
 * dot family (tensor-core GEMMs, below) and for TransE (l_norm 1, 2) / RotatE (l_norm 1) (CUDA-core row-gradient passes,
 * grad_distance.cu: dQ_i = sum_j G_ij s'(Q_i - T_j), dT_j = sum_i G_ij s'(T_j - Q_i)): dense gradients of the entity table d_ent [E, lde] and of the relation table d_rel
 * [R, ldr] of  (loss(score_sp, o) + loss(score_po, s)) / n.  Both buffers are overwritten (the reference
 * accumulates into .grad; add them there).  Recompute-based: scores, G = n dL/dz (sigmoid(z+off) - y | softmax(z) - y), two tensor-core
 * GEMMs (dT = G^T Q, dQ = G T), row-wise unfold of dQ through the relation fold (grad.cu). */
size_t b200kge_train_1vsall_backward_workspace_bytes(int model, int64_t n, int64_t E, int32_t D);
int b200kge_train_1vsall_backward(int model, float l_norm, const b200kge_rows_t* ent, const b200kge_rows_t* rel,
                                    const int64_t* triples, int64_t n, int loss_kind, float offset,
                                    float* d_ent, int64_t lde, float* d_rel, int64_t ldr,
                                    void* workspace, size_t workspace_bytes, b200kge_stream_t stream);

/* Backward of b200kge_score_1vsN over the whole entity table for the dot family (the unfused route: a job computes a
 * dense [n, E] score matrix, its loss, and autograd hands back grad_scores = dL/dscores [n, ldg]): dense gradients of
 * the entity table d_ent [E, lde] and the relation table d_rel [R, ldr], both OVERWRITTEN.  Fold of the n query rows,
 * fp16 hi/lo planes of grad_scores and of its transpose, two split-K tensor-core GEMMs (dT = G^T Q, dQ = G T) and the
 * row-wise unfold — no cuBLAS, no [n, E, D] intermediate. */
size_t b200kge_score_1vsN_backward_workspace_bytes(int model, int64_t n, int64_t E, int32_t D);
int b200kge_score_1vsN_backward(int model, int combine, float l_norm, const b200kge_rows_t* ent, const b200kge_rows_t* rel,
                                const int64_t* q_idx, const int64_t* p_idx, int64_t n, const float* grad_scores,
                                int64_t ldg, float* d_ent, int64_t lde, float* d_rel, int64_t ldr, void* workspace,
                                size_t workspace_bytes, b200kge_stream_t stream);

/* Backward of b200kge_score_1vsN_loss_csr / batch_size (loss_value.backward() at kge/job/train_KvsAll.py:294) for the
 * dot family: dense gradients d_ent [E, lde], d_rel [R, ldr] (OVERWRITTEN) of  sum_i loss(score row i, y_i) / batch_size
 * with y = (1 - eps) * count + (eps > 0 ? 1/E : 0) from the CSR labels — recompute, G planes written with the label-free
 * value everywhere and patched at the nnz listed entries, two split-K tensor-core GEMMs, unfold.  Workspace:
 * b200kge_score_1vsN_backward_workspace_bytes. */
int b200kge_score_1vsN_loss_csr_backward(int model, int combine, const b200kge_rows_t* ent, const b200kge_rows_t* rel,
                                         const int64_t* q_idx, const int64_t* p_idx, int64_t n, const int64_t* csr_off,
                                         const int64_t* csr_col, float label_smoothing, int loss_kind, float offset,
                                         int64_t batch_size, float* d_ent, int64_t lde, float* d_rel, int64_t ldr,
                                         void* workspace, size_t workspace_bytes, b200kge_stream_t stream);

/* KvsAll loss with CSR multi-hot labels (kge/job/train_KvsAll.py:242-300 without the densified label matrix):
 * row i's labels are the columns csr_col[csr_off[i] .. csr_off[i+1]) (sorted; a repeated column counts as often
 * as it appears, like duplicate triples in the reference), optionally smoothed: y = (1 - eps) * count + 1/m.
 * *loss_out = sum_i loss(score row i, y_i) (BCE with offset | KL), row_loss_out (optional) the per-row terms.
 * On the pre-split tensor-core path (dot family) ONE fused pass scores, reduces the label-free loss terms and emits
 * the nnz listed scores from its epilogue (per-thread cursor into the row's sorted segment); elsewhere the listed
 * scores come from the row-wise triple kernel.  cand must be a plain table; eps > 0 needs a dot-family model.  Sizes: b200kge_score_1vsN_loss_csr_workspace_bytes. */
size_t b200kge_score_1vsN_loss_csr_workspace_bytes(int model, int64_t n, int64_t m, int32_t D, int64_t nnz);
int b200kge_score_1vsN_loss_csr(int model, int combine, float l_norm, int precision,
                                  const b200kge_rows_t* q, const b200kge_rows_t* p,
                                  const b200kge_rows_t* cand, int64_t n, const int64_t* csr_off,
                                  const int64_t* csr_col, int64_t nnz, float label_smoothing, int loss_kind,
                                  float offset, float* loss_out, float* row_loss_out, void* workspace,
                                  size_t workspace_bytes, b200kge_stream_t stream);

/* Backward of one slot of a negative-sampling batch with BCE (kge/job/train_negative_sampling.py:113-164): the
 * [n, 1+K] block of the slot (column 0 = the positive triple, label 1; columns 1.. = the sampled ids neg [n,K],
 * label 0), loss summed and divided by batch_size.  ADDS into d_ent [E, lde] and d_rel [R, ldr] (zero them before
 * the first slot).  slot 0 (S) or 2 (O); TransE with l_norm 1 or 2, RotatE with l_norm 1, and the dot family.
 * workspace: n * round_up(K_folded, 32) floats. */
int b200kge_ns_backward(int model, float l_norm, const b200kge_rows_t* ent, const b200kge_rows_t* rel,
                          const int64_t* triples, int slot, const int64_t* neg, int64_t n, int64_t K,
                          float offset, int64_t batch_size, float* d_ent, int64_t lde, float* d_rel,
                          int64_t ldr, void* workspace, size_t workspace_bytes, b200kge_stream_t stream);

/* LookupEmbedder.penalty (kge/model/embedder/lookup_embedder.py:123-177) on the rows view `rows` (the whole
 * table, or the batch's unique rows through rows->idx with their `counts`, NULL = all ones):
 *   *out = scale * sum_r counts[r] * sum_k |x_rk|^p        (complex_abs: x -> sqrt(re^2 + im^2 + 1e-14): "n3",
 *   which the reference accepts in complex space only, lookup_embedder.py:29-34)
 * scale = regularize_weight / p (unweighted) or regularize_weight / p / len(indexes) (weighted).
 * workspace: (ceil(rows / 8) + 1) floats.  Deterministic (fixed-order two-stage sum). */
int b200kge_lookup_penalty(const b200kge_rows_t* rows, const float* counts, float p, int complex_abs,
                             float scale, float* out, void* workspace, size_t workspace_bytes,
                             b200kge_stream_t stream);

/* LookupEmbedder._normalize_embeddings (:64-69): rows of weight [rows, dim] (row stride ld) scaled in place to
 * unit Lp norm (torch.nn.functional.normalize, eps 1e-12). */
int b200kge_normalize_rows(float* weight, int64_t ld, int64_t rows, int32_t dim, float p,
                             b200kge_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B200KGE_H_ */
