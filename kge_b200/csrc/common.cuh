// common.cuh — shared host/device helpers of libb200kge (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/b200kge.h"

namespace b200kge {

// ---------------------------------------------------------------------------------------------
// Host-side error plumbing (capi.cu owns the storage).
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int check_cuda(cudaError_t e, const char* what);
// profiling brackets around the dominant kernel (no-ops unless b200kge_profile_enable(1))
void profile_begin(cudaStream_t st);
void profile_end(cudaStream_t st);

#define B2K_CUDA(expr)                                             \
  do {                                                             \
    int _s = ::b200kge::check_cuda((expr), #expr);                 \
    if (_s != 0) return _s;                                        \
  } while (0)

#define B2K_LAUNCH_CHECK(name)                                     \
  do {                                                             \
    ::b200kge::count_launch();                                     \
    int _s = ::b200kge::check_cuda(cudaGetLastError(), name);      \
    if (_s != 0) return _s;                                        \
  } while (0)

// ---------------------------------------------------------------------------------------------
// Device view of b200kge_rows_t.
struct Rows {
  const float* base;
  const int64_t* idx;
  int64_t rows;
  int64_t ld;
  int dim;
  __host__ __device__ __forceinline__ const float* row(int64_t i) const {
    return base + (idx ? idx[i] : i) * ld;
  }
};

inline Rows to_rows(const b200kge_rows_t* r) {
  Rows v;
  v.base = r->base; v.idx = r->idx; v.rows = r->rows; v.ld = r->ld; v.dim = r->dim;
  return v;
}

// How a pair (query row, candidate row) is reduced over the feature dimension.
enum PairOp : int {
  PAIR_DOT = 0,      // sum q*t                          (ComplEx, DistMult, SimplE, CP, RESCAL)
  PAIR_L1 = 1,       // -sum |q-t|                       (TransE l_norm=1)
  PAIR_L2 = 2,       // -sqrt(sum (q-t)^2)               (TransE l_norm=2)
  PAIR_LP = 3,       // -(sum |q-t|^p)^(1/p)             (TransE other p)
  PAIR_CMOD_L1 = 4,  // -sum_k |q_k - t_k| complex       (RotatE l_norm=1)
  PAIR_CMOD_LP = 5   // -(sum_k |q_k-t_k|^p)^(1/p)       (RotatE other p)
};

// Folded problem: score(i, j) = pair(Q[i, 0:K], cand_j[col_off : col_off+K]).
struct Folded {
  int pair_op;
  int K;        // reduction length in floats (complex pair ops: K = 2h, re at k, im at k+h)
  int col_off;  // first candidate column used
};

__host__ __device__ inline int relation_dim(int model, int D) {
  if (model == B200KGE_CP || model == B200KGE_ROTATE) return D / 2;
  if (model == B200KGE_RESCAL) return D * D;
  return D;
}

// ---------------------------------------------------------------------------------------------
// Epilogues.  A kernel computes x = score(row, col) for a tile and hands every valid element to
// one of these functors; per-row state lives in registers and is flushed once per (row, chunk).
enum EpiKind : int { EPI_STORE = 0, EPI_BCE = 1, EPI_KL = 2, EPI_RANK = 3 };

struct FinalizeArgs;
struct EpiParams {
  // EPI_STORE
  float* out;
  int64_t ldo;
  // losses
  const int64_t* label_idx;
  const float* label_dense;
  int64_t ldl;
  float offset;
  float* part;        // [n][nchunks][F] partial row states
  int nchunks;
  // rank
  const float* true_score;
  const float* filter;
  int64_t ldf;
  float rtol, atol;
  unsigned long long* rank;
  unsigned long long* ties;
  // row remapping for the fused sp_po launch: logical query row r (0..2n) maps to output row
  // r % n_rows_out and column block (r / n_rows_out) * col_block
  int64_t n_rows_out;
  int64_t col_block;
  // EPI_STORE, GEMM use (pairwise_tc3.cu split-K): add into `out` (zeroed by the caller) instead of overwriting it
  int accumulate_out;
  // CSR side input: row r lists the sorted columns csr_col[csr_off[r] .. csr_off[r+1]) (duplicates
  // allowed).  Losses: the raw scores at the listed columns are written to csr_out[t] (t = position in csr_col) —
  // the multi-hot label terms are then sums over nnz values, no [n, m] label matrix exists (train_KvsAll.py:242-266);
  // with csr_extra the score of column 0 of row r goes to csr_out[csr_nnz + r] (KL: lse_i - z_i0 comes back from
  // the one-hot-at-0 pass).  Rank: listed columns are filtered (score -> -inf, eval_entity_ranking.py:561-566)
  // except the row's own answer csr_skip[r] (:287-290).  Candidate columns must be table rows (cand.idx == NULL).
  const int64_t* csr_off;
  const int64_t* csr_col;
  float* csr_out;
  int64_t csr_nnz;
  int csr_extra;
  const int64_t* csr_skip;
  // EPI_STORE fused with the all-gather of an entity-sharded table: every score is also stored, at the
  // same offset, into the symmetric output buffers of the n_peers other ranks (peer-mapped pointers over
  // NVLink / NVSwitch) — the shard's logits land in place in everybody's [n, 2E] matrix, no NCCL all-gather, no
  // re-layout copy
  float* out_peer[7];
  int n_peers;
  int store_vec4;      // set by launch_pairwise_simt: every destination row start is 16-byte aligned
};

// lower bound of `key` in the sorted segment col[lo, hi)
__device__ __forceinline__ int64_t csr_lower_bound(const int64_t* __restrict__ col, int64_t lo, int64_t hi, int64_t key) {
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (__ldg(col + mid) < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

#define B2K_NEG_HUGE (-3.0e38f)

__device__ __forceinline__ float softplus_f(float z) {
  // max(z,0) + log1p(exp(-|z|)), matching torch's BCEWithLogits formulation (loss.py:150)
  // branch-free (a data-dependent branch here diverges per element and cost +40 % kernel time):
  // log1p(e) ~= e for tiny e (where 1+e would round to 1), else log(1+e); both arms are already
  // computed, so the compiler emits a select.
  const float e = __expf(-fabsf(z));
  const float lg = __logf(1.0f + e);
  const float l = (e < 1e-5f) ? e : lg;
  return fmaxf(z, 0.0f) + l;
}

template <int KIND> struct RowState {};

template <> struct RowState<EPI_STORE> {
  static constexpr int F = 0;
  __device__ __forceinline__ void init() {}
};

// BCE: a = sum softplus(z), b = sum y*z            loss.py:153-159
template <> struct RowState<EPI_BCE> {
  static constexpr int F = 2;
  float a, b;
  __device__ __forceinline__ void init() { a = 0.f; b = 0.f; }
  __device__ __forceinline__ void combine(const RowState& o) { a += o.a; b += o.b; }
};

// KL / CE: online (m, s) for logsumexp, plus label sums   loss.py:198-213
//   y_sum = sum y, yx = sum y*x, ylogy = sum y*log(y)
template <> struct RowState<EPI_KL> {
  static constexpr int F = 5;
  float m, s, y_sum, yx, ylogy;
  __device__ __forceinline__ void init() { m = B2K_NEG_HUGE; s = 0.f; y_sum = 0.f; yx = 0.f; ylogy = 0.f; }
  __device__ __forceinline__ void combine(const RowState& o) {
    float mn = fmaxf(m, o.m);
    s = s * __expf(m - mn) + o.s * __expf(o.m - mn);
    m = mn;
    y_sum += o.y_sum; yx += o.yx; ylogy += o.ylogy;
  }
};

// rank / ties counters     eval_entity_ranking.py:571-596
template <> struct RowState<EPI_RANK> {
  static constexpr int F = 0;
  unsigned int greater, close;
  __device__ __forceinline__ void init() { greater = 0u; close = 0u; }
  __device__ __forceinline__ void combine(const RowState& o) { greater += o.greater; close += o.close; }
};

// torch.isclose(x, t, rtol, atol) for fp32 operands (equal_nan=False), evaluated in fp32 exactly
// as ATen does: (x == t) | (isfinite(|x-t|) & (|x-t| <= atol + |rtol*t|)).
__device__ __forceinline__ bool isclose_f(float x, float t, float rtol, float atol) {
  float allowed = __fadd_rn(atol, fabsf(__fmul_rn(rtol, t)));
  float actual = fabsf(__fsub_rn(x, t));
  return (x == t) || (isfinite(actual) && actual <= allowed);
}

template <int KIND>
__device__ __forceinline__ void epi_elem(const EpiParams& P, RowState<KIND>& st, int64_t row,
                                         int64_t col, float x, float row_aux) {
  if constexpr (KIND == EPI_STORE) {
    int64_t r = row, cb = 0;
    if (P.n_rows_out > 0 && row >= P.n_rows_out) { r = row - P.n_rows_out; cb = P.col_block; }
    const int64_t at = r * P.ldo + cb + col;
    P.out[at] = x;
    for (int g = 0; g < P.n_peers; ++g) P.out_peer[g][at] = x;
  } else if constexpr (KIND == EPI_BCE) {
    float z = x + P.offset;
    st.a += softplus_f(z);
    if (P.label_dense) {
      st.b = fmaf(P.label_dense[row * P.ldl + col], z, st.b);
    } else if (col == (int64_t)__float_as_int(row_aux)) {
      st.b += z;
    }
  } else if constexpr (KIND == EPI_KL) {
    float mn = fmaxf(st.m, x);
    st.s = st.s * __expf(st.m - mn) + __expf(x - mn);
    st.m = mn;
    if (P.label_dense) {
      float y = P.label_dense[row * P.ldl + col];
      if (y != 0.f) {
        st.y_sum += y;
        st.yx = fmaf(y, x, st.yx);
        st.ylogy = fmaf(y, __logf(y), st.ylogy);
      }
    } else if (col == (int64_t)__float_as_int(row_aux)) {
      st.y_sum += 1.0f;
      st.yx += x;
    }
  } else if constexpr (KIND == EPI_RANK) {
    float v = x;
    if (P.filter) v = __fsub_rn(v, P.filter[row * P.ldf + col]);   // :561-566
    if (isnan(v)) v = -INFINITY;                                    // :583-584
    bool c = isclose_f(v, row_aux, P.rtol, P.atol);
    st.close += c ? 1u : 0u;
    st.greater += (!c && v > row_aux) ? 1u : 0u;
  }
}

// Per-row auxiliary scalar loaded once per (thread,row): label index (as int bits; rows are
// < 2^31 candidates per call) or the NaN-cleaned true score.
template <int KIND>
__device__ __forceinline__ float epi_row_aux(const EpiParams& P, int64_t row) {
  if constexpr (KIND == EPI_BCE || KIND == EPI_KL) {
    return P.label_idx ? __int_as_float((int)P.label_idx[row]) : __int_as_float(-1);
  } else if constexpr (KIND == EPI_RANK) {
    float t = P.true_score[row];
    return isnan(t) ? -INFINITY : t;                                // :585-586
  } else {
    return 0.f;
  }
}

template <int KIND>
__device__ __forceinline__ void epi_flush(const EpiParams& P, const RowState<KIND>& st,
                                          int64_t row, int chunk) {
  if constexpr (KIND == EPI_BCE) {
    float* p = P.part + (row * P.nchunks + chunk) * 2;
    p[0] = st.a; p[1] = st.b;
  } else if constexpr (KIND == EPI_KL) {
    float* p = P.part + (row * P.nchunks + chunk) * 5;
    p[0] = st.m; p[1] = st.s; p[2] = st.y_sum; p[3] = st.yx; p[4] = st.ylogy;
  } else if constexpr (KIND == EPI_RANK) {
    // integer atomics: order-independent, hence bit-exact
    if (st.greater) atomicAdd(P.rank + row, (unsigned long long)st.greater);
    if (st.close) atomicAdd(P.ties + row, (unsigned long long)st.close);
  }
}

// Combine the states of the `width` (power of two <= 32) adjacent lanes that share a row.
template <int KIND>
__device__ __forceinline__ void epi_lane_reduce(RowState<KIND>& st, int width) {
  if constexpr (KIND != EPI_STORE) {
    constexpr int W = sizeof(RowState<KIND>) / 4;
    for (int off = width >> 1; off > 0; off >>= 1) {
      RowState<KIND> o;
      uint32_t* src = reinterpret_cast<uint32_t*>(&st);
      uint32_t* dst = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
      for (int w = 0; w < W; ++w) dst[w] = __shfl_xor_sync(0xffffffffu, src[w], off);
      st.combine(o);
    }
  }
}


// ---------------------------------------------------------------------------------------------
// Deterministic finaliser of the fused losses: partial[n][nchunks][F] -> row loss -> scalar.
// Fixed reduction order at every level (lanes over chunks -> shuffle tree, rows per warp in
// order, warps per block in order, blocks in index order by the last block): the result does not
// depend on scheduling.
struct FinalizeArgs {
  const float* part;
  int nchunks;
  int64_t n;
  float* loss_out;
  float* row_loss_out;
  float scale;
  int accumulate;
  unsigned int* ticket;   // zero-initialised counter for the last-block-done protocol
  float* block_sums;      // [gridDim.x] scratch
};

template <int LOSS>
__device__ __forceinline__ float finalize_row(const float* __restrict__ part, int nchunks, int64_t r) {
  if constexpr (LOSS == B200KGE_LOSS_BCE) {
    float a = 0.f, b = 0.f;
    for (int c = 0; c < nchunks; ++c) {
      const float* p = part + (r * nchunks + c) * 2;
      a += p[0]; b += p[1];
    }
    return a - b;                       // sum softplus(z) - sum y*z
  } else {
    RowState<EPI_KL> st;
    st.init();
    for (int c = 0; c < nchunks; ++c) {
      const float* p = part + (r * nchunks + c) * 5;
      RowState<EPI_KL> o;
      o.m = p[0]; o.s = p[1]; o.y_sum = p[2]; o.yx = p[3]; o.ylogy = p[4];
      st.combine(o);
    }
    const float lse = st.m + logf(st.s);
    // KLDiv(log_softmax(x), y / max(||y||_1, 1e-12)), reduction sum   loss.py:209-213
    const float yc = fmaxf(st.y_sum, 1e-12f);
    const float w = st.y_sum / yc;
    return (st.y_sum > 0.f) ? (st.ylogy / yc - w * logf(yc) - st.yx / yc + lse * w) : 0.f;
  }
}


// ---------------------------------------------------------------------------------------------
// Internal kernels' host launchers (one per .cu file).
int launch_fold_queries(int model, int combine, const Rows& q, const Rows& p, int64_t n,
                        int64_t row0, float* Q, int64_t ldq, cudaStream_t st);
// one launch: unpack triples [n,3], fold sp_ rows (0..n) and _po rows (n..2n) into Q, write the
// stacked labels [o ; s] and zero the finalisation ticket
int launch_prep_1vsall(int model, const Rows& ent, const Rows& rel, const int64_t* triples, int64_t n,
                       float* Q, int64_t ldq, int64_t* labels2n, unsigned int* ticket, cudaStream_t st);
int launch_gather_rows(const Rows& src, int col_off, int K, float* dst, int64_t ldd,
                       cudaStream_t st);
int launch_pairwise_simt(int epi_kind, int pair_op, float l_norm, const float* Q, int64_t ldq,
                         int64_t nq, const Rows& cand, int col_off, int K, const EpiParams& P,
                         cudaStream_t st);
int pairwise_simt_nchunks(int64_t nq, int64_t m);
// tcgen05 path: returns B200KGE_ERR_UNSUPPORTED if the shape cannot be served.
bool tc_supported(int pair_op, int K, const Rows& cand, int col_off);
int tc_nchunks(int64_t nq, int64_t m);
int launch_pairwise_tc(int epi_kind, int passes, const float* Q, int64_t ldq,
                       int64_t nq, const float* T, int64_t ldt, int64_t m, int K,
                       const EpiParams& P, cudaStream_t st);
// scratch: >= 1024 bytes of device memory (128 block sums + ticket); ticket_zeroed: the caller already
// zeroed the ticket word at scratch+512 on this stream (else a 4-byte memset is enqueued)
int launch_loss_finalize(int loss_kind, const float* part, int nchunks, int64_t n, float* loss_out,
                         float* row_loss_out, float scale, int accumulate, void* scratch,
                         int ticket_zeroed, cudaStream_t st);
int launch_spo(int model, float l_norm, const Rows& s, const Rows& p, const Rows& o, int64_t n,
               float* out, int64_t out_stride, cudaStream_t st);
int launch_ns(int model, float l_norm, const Rows& s, const Rows& p, const Rows& o,
              const Rows& table, int slot, const int64_t* neg, int64_t n, int64_t K, float* out,
              int64_t ldo, int col0, cudaStream_t st);
int launch_sample_uniform(uint64_t seed, uint64_t offset, int64_t vocab, int64_t total, int64_t* out, cudaStream_t st);
int launch_loss_dense(int loss_kind, const float* scores, int64_t lds, int64_t n, int64_t m,
                      const EpiParams& P, cudaStream_t st);
int loss_dense_nchunks(int64_t m);
int launch_rank_dense(const float* scores, int64_t lds, int64_t n, int64_t m, const EpiParams& P,
                      cudaStream_t st);

}  // namespace b200kge
