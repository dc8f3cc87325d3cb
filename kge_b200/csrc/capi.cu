// capi.cu — extern "C" entry points of libb200kge (see include/b200kge.h for the contract).
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include "fold.cuh"
#include "tc_common.cuh"

namespace b200kge {

static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches += n; }

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  if (e == cudaErrorMemoryAllocation)
    set_error("CUDA out of memory (%s)", what);  // literal matched by LibKGE train.py:384-413
  else
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return B200KGE_ERR_CUDA;
}

static thread_local int g_prof_on = 0;
static thread_local cudaEvent_t g_ev0 = nullptr, g_ev1 = nullptr;
static thread_local int g_prof_valid = 0;

void profile_begin(cudaStream_t st) {
  if (!g_prof_on) return;
  if (!g_ev0) { cudaEventCreate(&g_ev0); cudaEventCreate(&g_ev1); }
  cudaEventRecord(g_ev0, st);
}
void profile_end(cudaStream_t st) {
  if (!g_prof_on || !g_ev0) return;
  cudaEventRecord(g_ev1, st);
  g_prof_valid = 1;
}

namespace {

// Which pre-split kernel scores a block with n query rows PER DIRECTION and reduction length K: the CTA-pair kernel
// (pairwise_tc4.cu) for n >= 128 and K > 448 (two 128-row halves per cluster tile; 65.8 vs 67.6 us at the FB15k-237
// headline shape, K = 512), the 1-CTA kernel (pairwise_tc3.cu) otherwise (short reductions make the pair's per-tile
// cross-CTA hand-over visible: RESCAL d=200 KvsAll 0.186 vs 0.146 ms).  Depends on n and K only — never on the
// candidate count or on stacking — so every call of one batch (true scores on the unique targets, chunk scores, fused
// forms) runs the same kernel.  B200KGE_TC_VERSION=3 | 4 forces one of them.
bool use_pair_kernel(int64_t n, int K) {
  const char* env_v = getenv("B200KGE_TC_VERSION");
  if (env_v && atoi(env_v) == 3) return false;
  if (env_v && atoi(env_v) == 4) return true;
  return n >= 128 && K > 448;
}

// bump allocator over the caller's workspace
struct Arena {
  uint8_t* base;
  size_t cap, off;
  void* take(size_t bytes) {
    size_t a = (off + 255) & ~size_t(255);
    if (a + bytes > cap) return nullptr;
    off = a + bytes;
    return base + a;
  }
};

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

int validate_model(int model, const Rows& ent_like, const Rows& rel_like) {
  if (model < 0 || model > B200KGE_ROTATE) { set_error("unknown model %d", model); return B200KGE_ERR_INVALID; }
  const int D = ent_like.dim;
  if (D <= 0) { set_error("entity dim must be positive"); return B200KGE_ERR_INVALID; }
  if ((model == B200KGE_COMPLEX || model == B200KGE_SIMPLE || model == B200KGE_CP || model == B200KGE_ROTATE) && (D & 1)) {
    set_error("model %d requires embeddings of even dimensionality (got %d)", model, D);
    return B200KGE_ERR_INVALID;
  }
  const int want = relation_dim(model, D);
  if (rel_like.dim != want) {
    set_error("relation dim %d does not match model %d with entity dim %d (expected %d)", rel_like.dim, model, D, want);
    return B200KGE_ERR_INVALID;
  }
  return 0;
}

int validate_norm(int model, float l_norm) {
  if ((model == B200KGE_TRANSE || model == B200KGE_ROTATE) && !(l_norm > 0.f && l_norm < 1e30f)) {
    set_error("l_norm must be a positive finite number (got %g)", (double)l_norm);
    return B200KGE_ERR_INVALID;
  }
  return 0;
}

// One direction (or the stacked sp+po pair) of a 1-vs-N problem, with any epilogue.
struct Block {
  int model, combine;     // combine of the first n rows; stacked => second n rows use the other one
  const Rows* q0; const Rows* q1;  // per-row entity operands of the two halves (q1 null if not stacked)
  const Rows* p;
  const Rows* cand;
  int64_t n;
  const float* Qpre = nullptr;   // already-folded queries [nq, round_up(K,32)] (skips the fold launches)
};

// nchunks_out / part_out (optional): number of per-row partial chunks and the partial buffer the loss
// epilogues wrote (input of launch_loss_finalize).
int run_block(const Block& B, float l_norm, int precision, int epi_kind, EpiParams P, Arena& ws,
              cudaStream_t st, int* nchunks_out, float** part_out = nullptr) {
  const int64_t n = B.n, nq = B.q1 ? 2 * n : n, m = B.cand->rows;
  const int D = B.q0->dim;
  Folded f0 = folded_problem(B.model, B.combine, D, l_norm);
  Folded f1 = f0;
  if (B.q1) f1 = folded_problem(B.model, 1 - B.combine, D, l_norm);
  const bool cols_differ = B.q1 && (f0.col_off != f1.col_off);   // CP: halves read different columns
  const int K = f0.K;
  const int64_t ldq = round_up(K, 32);

  // Path selection depends on the model, K, the table and the PER-DIRECTION row count n only — never on whether
  // the two directions are stacked — so score_sp / score_po / score_sp_po / the fused forms of one batch all run
  // the same arithmetic (EntityRankingJob compares them, eval_entity_ranking.py:192-203,242-274).
  //   AUTO    -> F16X3 (pre-split fp16 planes, pairwise_tc3.cu) for dot-product scorers with 32 <= K <= 1024 and
  //              n >= 16; fp32 SIMT otherwise (beyond K = 1024 the tensor core's fp32 accumulator error, which
  //              grows with the reduction length — 2.8e-4 of rms at K = 14541 — leaves too little margin)
  //   F16X3   -> pairwise_tc4.cu (CTA pair) for n >= 128 and K > 448, else pairwise_tc3.cu (use_pair_kernel)
  //   TF32_BF16X2 / 3XTF32 / TF32 -> pairwise_tc.cu (in-kernel split of raw fp32 tiles; needs TMA-able tables)
  int tc_kind = 0;       // 0 SIMT, 1 in-kernel split (pairwise_tc.cu), 3 pre-split planes
  if (f0.pair_op == PAIR_DOT && precision != B200KGE_PREC_FP32 && !cols_differ) {
    if (precision == B200KGE_PREC_AUTO) {
      if (K >= 32 && K <= 1024 && n >= 16 && m < (1ll << 31)) tc_kind = 3;
    } else if (precision == B200KGE_PREC_F16X3) {
      if (K < 16 || m >= (1ll << 31)) { set_error("the pre-split tensor-core path needs K >= 16"); return B200KGE_ERR_UNSUPPORTED; }
      tc_kind = 3;
    } else {
      if (!tc_supported(f0.pair_op, K, *B.cand, f0.col_off)) { set_error("tensor-core path needs K>=32, 16-byte aligned tables with ld%%4==0"); return B200KGE_ERR_UNSUPPORTED; }
      tc_kind = 1;
    }
  } else if (precision != B200KGE_PREC_AUTO && precision != B200KGE_PREC_FP32) {
    if (f0.pair_op != PAIR_DOT) { set_error("tensor-core precision modes apply to dot-product scorers only"); return B200KGE_ERR_UNSUPPORTED; }
  }
  { const char* env_v = getenv("B200KGE_TC_VERSION");      // experiments: 1 forces the in-kernel split
    if (env_v && atoi(env_v) == 1 && tc_kind == 3 && tc_supported(f0.pair_op, K, *B.cand, f0.col_off)) {
      tc_kind = 1; precision = B200KGE_PREC_TF32_BF16X2;
    } }

  if (P.csr_off && (cols_differ || B.cand->idx || tc_kind == 1 || (tc_kind == 0 && epi_kind != EPI_RANK))) {
    // CSR side inputs: the pre-split tensor-core epilogue (losses and rank) and the CUDA-core kernel's rank epilogue.
    // Nothing has been launched or taken from the workspace yet: callers fall back to their dense / composed form.
    set_error("this CSR side input is not consumed by the kernel serving this call (losses: pre-split tensor-core path; "
              "rank: that path or the CUDA-core kernel; plain candidate table)");
    return B200KGE_ERR_UNSUPPORTED;
  }

  if (cols_differ) {
    // run the two halves as separate blocks (CP reads different candidate columns per direction)
    Block h0 = B; h0.q1 = nullptr;
    Block h1 = B; h1.q0 = B.q1; h1.q1 = nullptr; h1.combine = 1 - B.combine;
    EpiParams P0 = P, P1 = P;
    P0.n_rows_out = 0; P1.n_rows_out = 0;
    if (epi_kind == EPI_STORE) { P1.out = P.out + P.col_block; }
    else { set_error("stacked fused epilogues are not available for CP"); return B200KGE_ERR_UNSUPPORTED; }
    int rc = run_block(h0, l_norm, precision, epi_kind, P0, ws, st, nchunks_out);
    if (rc) return rc;
    return run_block(h1, l_norm, precision, epi_kind, P1, ws, st, nchunks_out);
  }

  if (tc_kind) {
    int rc = 0;
    const float* Q = B.Qpre;
    if (!Q) {
      float* Qw = (float*)ws.take((size_t)nq * ldq * 4);
      if (!Qw) { set_error("workspace too small for folded queries"); return B200KGE_ERR_WORKSPACE; }
      rc = launch_fold_queries(B.model, B.combine, *B.q0, *B.p, n, 0, Qw, ldq, st);
      if (rc) return rc;
      if (B.q1) { rc = launch_fold_queries(B.model, 1 - B.combine, *B.q1, *B.p, n, n, Qw, ldq, st); if (rc) return rc; }
      Q = Qw;
    }
    if (tc_kind == 3) {
      // pre-split fp16 path (presplit.cu + pairwise_tc3.cu | pairwise_tc4.cu): one launch derives the hi/lo planes of
      // the folded queries and of the (gathered) candidate rows, one launch scores them.
      const bool pair = use_pair_kernel(n, K);
      const int Kp = (int)round_up(K, 64);
      SplitSet SQ{Q, ldq, nullptr, 0, nq, nq, K, Kp, nullptr, nullptr, nullptr};
      SplitSet ST{B.cand->base, B.cand->ld, B.cand->idx, f0.col_off, m, m + 32, K, Kp, nullptr, nullptr, nullptr};
      SQ.hi = ws.take((size_t)nq * Kp * 2); SQ.lo = ws.take((size_t)nq * Kp * 2);
      SQ.inv_scale = (float*)ws.take((size_t)nq * 4);
      ST.hi = ws.take((size_t)m * Kp * 2); ST.lo = ws.take((size_t)m * Kp * 2);
      ST.inv_scale = (float*)ws.take((size_t)(m + 32) * 4);
      if (!SQ.hi || !SQ.lo || !SQ.inv_scale || !ST.hi || !ST.lo || !ST.inv_scale) {
        set_error("workspace too small for the pre-split operand planes");
        return B200KGE_ERR_WORKSPACE;
      }
      const int nch3 = pair ? tc4_nchunks(nq, m) : tc3_nchunks(nq, m);
      if (epi_kind == EPI_BCE || epi_kind == EPI_KL) {
        const int F = (epi_kind == EPI_BCE) ? 2 : 5;
        P.part = (float*)ws.take((size_t)nq * nch3 * F * 4);
        if (!P.part) { set_error("workspace too small for loss partials"); return B200KGE_ERR_WORKSPACE; }
        if (part_out) *part_out = P.part;
      }
      P.nchunks = nch3;
      if (nchunks_out) *nchunks_out = nch3;
      if ((rc = launch_presplit(ST, SQ, st))) return rc;
      if (pair) return launch_pairwise_tc4(epi_kind, SQ, ST, P, st);
      return launch_pairwise_tc3(epi_kind, SQ, ST, P, st);
    }
    const int passes = (precision == B200KGE_PREC_TF32) ? 1 : (precision == B200KGE_PREC_3XTF32 ? 3 : 2);
    const float* T = B.cand->base + f0.col_off;
    int64_t ldt = B.cand->ld;
    if (B.cand->idx) {
      float* G = (float*)ws.take((size_t)m * ldq * 4);
      if (!G) { set_error("workspace too small to gather the candidate subset"); return B200KGE_ERR_WORKSPACE; }
      rc = launch_gather_rows(*B.cand, f0.col_off, K, G, ldq, st);
      if (rc) return rc;
      T = G; ldt = ldq;
    }
    const int nch = tc_nchunks(nq, m);
    if (epi_kind == EPI_BCE || epi_kind == EPI_KL) {
      const int F = (epi_kind == EPI_BCE) ? 2 : 5;
      P.part = (float*)ws.take((size_t)nq * nch * F * 4);
      if (!P.part) { set_error("workspace too small for loss partials"); return B200KGE_ERR_WORKSPACE; }
      if (part_out) *part_out = P.part;
    }
    P.nchunks = nch;
    if (nchunks_out) *nchunks_out = nch;
    return launch_pairwise_tc(epi_kind, passes, Q, ldq, nq, T, ldt, m, K, P, st);
  }

  int rc = 0;
  const float* Q = B.Qpre;
  if (!Q) {
    float* Qw = (float*)ws.take((size_t)nq * ldq * 4);
    if (!Qw) { set_error("workspace too small for folded queries"); return B200KGE_ERR_WORKSPACE; }
    rc = launch_fold_queries(B.model, B.combine, *B.q0, *B.p, n, 0, Qw, ldq, st);
    if (rc) return rc;
    if (B.q1) { rc = launch_fold_queries(B.model, 1 - B.combine, *B.q1, *B.p, n, n, Qw, ldq, st); if (rc) return rc; }
    Q = Qw;
  }
  const int nch = pairwise_simt_nchunks(nq, m);
  if (epi_kind == EPI_BCE || epi_kind == EPI_KL) {
    const int F = (epi_kind == EPI_BCE) ? 2 : 5;
    P.part = (float*)ws.take((size_t)nq * nch * F * 4);
    if (!P.part) { set_error("workspace too small for loss partials"); return B200KGE_ERR_WORKSPACE; }
    if (part_out) *part_out = P.part;
  }
  P.nchunks = nch;
  if (nchunks_out) *nchunks_out = nch;
  return launch_pairwise_simt(epi_kind, f0.pair_op, l_norm, Q, ldq, nq, *B.cand, f0.col_off, K, P, st);
}

EpiParams empty_epi() {
  EpiParams P;
  memset(&P, 0, sizeof(P));
  return P;
}

int check_1vsN_args(int model, int combine, const b200kge_rows_t* q, const b200kge_rows_t* p,
                    const b200kge_rows_t* cand, int64_t n) {
  if (!q || !p || !cand) { set_error("null operand"); return B200KGE_ERR_INVALID; }
  if (combine != B200KGE_SP_ && combine != B200KGE__PO) {
    set_error("cannot handle combine=%d", combine);   // ValueError in kge_model.py:211
    return B200KGE_ERR_INVALID;
  }
  if (n < 0 || q->rows < n || p->rows < n) { set_error("operand has fewer than n=%lld rows", (long long)n); return B200KGE_ERR_INVALID; }
  if (cand->dim != q->dim) { set_error("candidate dim %d != query entity dim %d", cand->dim, q->dim); return B200KGE_ERR_INVALID; }
  return validate_model(model, to_rows(q), to_rows(p));
}

__global__ void unpack_triples_kernel(const int64_t* __restrict__ tri, int64_t n, int64_t* s, int64_t* p,
                                      int64_t* o, int64_t* labels2n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) {
    const int64_t a = tri[3 * i], b = tri[3 * i + 1], c = tri[3 * i + 2];
    s[i] = a; p[i] = b; o[i] = c;
    labels2n[i] = c;       // sp_ rows are labelled with the object      train_1vsAll.py:64-65
    labels2n[n + i] = a;   // _po rows are labelled with the subject     train_1vsAll.py:75-76
  }
}

__global__ void pack_triples_kernel(const int64_t* __restrict__ q, const int64_t* __restrict__ p, int64_t n, int combine,
                                    int64_t* __restrict__ tri) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  tri[3 * i + 1] = p[i];
  tri[3 * i + (combine == B200KGE_SP_ ? 0 : 2)] = q[i];
  tri[3 * i + (combine == B200KGE_SP_ ? 2 : 0)] = 0;
}

__global__ void __launch_bounds__(128)
shard_gather_rows_kernel(Rows shard, int64_t lo, const int64_t* __restrict__ idx, float* __restrict__ out, int64_t ldo) {
  const int64_t i = blockIdx.x;
  const int64_t g = idx[i] - lo;
  const bool mine = g >= 0 && g < shard.rows;
  const float* __restrict__ src = shard.base + (mine ? g : 0) * shard.ld;
  float* __restrict__ dst = out + i * ldo;
  for (int k = threadIdx.x; k < shard.dim; k += blockDim.x) dst[k] = mine ? src[k] : 0.f;
}

}  // namespace
}  // namespace b200kge

using namespace b200kge;

extern "C" {

int b200kge_version(void) { return B200KGE_VERSION; }
const char* b200kge_last_error(void) { return g_err; }
int64_t b200kge_launch_count(int reset) {
  int64_t v = g_launches;
  if (reset) g_launches = 0;
  return v;
}

int b200kge_profile_enable(int on) { g_prof_on = on; g_prof_valid = 0; return 0; }
int b200kge_profile_last_ms(float* ms) {
  if (!ms) { set_error("null operand"); return B200KGE_ERR_INVALID; }
  if (!g_prof_valid) { set_error("no profiled kernel yet"); return B200KGE_ERR_INVALID; }
  B2K_CUDA(cudaEventSynchronize(g_ev1));
  B2K_CUDA(cudaEventElapsedTime(ms, g_ev0, g_ev1));
  return 0;
}

int b200kge_device_ok(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
    cudaGetLastError();
    set_error("no CUDA device available: libb200kge has no CPU fallback");
    return B200KGE_ERR_NO_DEVICE;
  }
  if (major != 10) { set_error("device compute capability %d.x is not sm_100 (B200)", major); return B200KGE_ERR_NO_DEVICE; }
  return 0;
}

size_t b200kge_workspace_bytes(int model, int64_t n, int64_t m, int32_t D, int cand_has_idx) {
  (void)model;
  const int64_t ldq = round_up(D, 32);
  const int64_t nq = 2 * n;
  size_t b = 0;
  b += 2 * ((size_t)nq * ldq * 4 + 256);                 // Qhi/Qlo (or Q)
  if (cand_has_idx) b += (size_t)m * ldq * 4 + 256;      // gathered candidate subset
  int64_t nch = pairwise_simt_nchunks(nq, m);
  if (nch < 320) nch = 320;                              // tensor-core kernels: <= 2 * #SMs chunks per row
  b += (size_t)nq * nch * 5 * 4 + 256;                   // loss partials
  b += (size_t)n * 3 * 8 + (size_t)n * 5 * 8 + 4096;     // host entry: triples, s/p/o, labels, scalar, finaliser scratch
  { // pre-split fp16 planes + row scales (the default tensor-core path)
    const int64_t Kp = round_up(D, 64);
    b += 2 * ((size_t)nq * Kp * 2 + 256) + 2 * ((size_t)m * Kp * 2 + 256) + (size_t)(nq + m + 32) * 4 + 512;
  }
  return b + 4096;
}

int b200kge_score_spo(int model, float l_norm, const b200kge_rows_t* s, const b200kge_rows_t* p,
                      const b200kge_rows_t* o, int64_t n, float* out, b200kge_stream_t stream) {
  if (!s || !p || !o || (!out && n > 0)) { set_error("null operand"); return B200KGE_ERR_INVALID; }
  int rc = validate_model(model, to_rows(s), to_rows(p)); if (rc) return rc;
  if ((rc = validate_norm(model, l_norm))) return rc;
  if (s->rows < n || p->rows < n || o->rows < n) { set_error("operand has fewer than n rows"); return B200KGE_ERR_INVALID; }
  return launch_spo(model, l_norm, to_rows(s), to_rows(p), to_rows(o), n, out, 1, (cudaStream_t)stream);
}

int b200kge_score_1vsN(int model, int combine, float l_norm, int precision,
                       const b200kge_rows_t* q, const b200kge_rows_t* p,
                       const b200kge_rows_t* cand, int64_t n, float* out, int64_t ldo,
                       void* workspace, size_t workspace_bytes, b200kge_stream_t stream) {
  int rc = check_1vsN_args(model, combine, q, p, cand, n); if (rc) return rc;
  if ((rc = validate_norm(model, l_norm))) return rc;
  Rows Q = to_rows(q), Pr = to_rows(p), C = to_rows(cand);
  Arena ws{(uint8_t*)workspace, workspace_bytes, 0};
  EpiParams P = empty_epi();
  P.out = out; P.ldo = ldo;
  Block B{model, combine, &Q, nullptr, &Pr, &C, n};
  return run_block(B, l_norm, precision, EPI_STORE, P, ws, (cudaStream_t)stream, nullptr);
}

int b200kge_score_sp_po(int model, float l_norm, int precision, const b200kge_rows_t* s,
                        const b200kge_rows_t* p, const b200kge_rows_t* o,
                        const b200kge_rows_t* cand, int64_t n, float* out, int64_t ldo,
                        void* workspace, size_t workspace_bytes, b200kge_stream_t stream) {
  int rc = check_1vsN_args(model, B200KGE_SP_, s, p, cand, n); if (rc) return rc;
  if ((rc = check_1vsN_args(model, B200KGE__PO, o, p, cand, n))) return rc;
  if ((rc = validate_norm(model, l_norm))) return rc;
  Rows Sr = to_rows(s), Pr = to_rows(p), Or = to_rows(o), C = to_rows(cand);
  Arena ws{(uint8_t*)workspace, workspace_bytes, 0};
  EpiParams P = empty_epi();
  P.out = out; P.ldo = ldo; P.n_rows_out = n; P.col_block = cand->rows;
  Block B{model, B200KGE_SP_, &Sr, &Or, &Pr, &C, n};
  return run_block(B, l_norm, precision, EPI_STORE, P, ws, (cudaStream_t)stream, nullptr);
}

int b200kge_score_sp_po_bcast(int model, float l_norm, int precision, const b200kge_rows_t* s,
                              const b200kge_rows_t* p, const b200kge_rows_t* o, const b200kge_rows_t* cand,
                              int64_t n, float* out, float* const* peer_out, int n_peers, int64_t ldo,
                              int64_t col_block, void* workspace, size_t workspace_bytes,
                              b200kge_stream_t stream) {
  int rc = check_1vsN_args(model, B200KGE_SP_, s, p, cand, n); if (rc) return rc;
  if ((rc = check_1vsN_args(model, B200KGE__PO, o, p, cand, n))) return rc;
  if ((rc = validate_norm(model, l_norm))) return rc;
  if (n_peers < 0 || n_peers > 7 || (n_peers > 0 && !peer_out)) { set_error("0..7 peer buffers"); return B200KGE_ERR_INVALID; }
  if (col_block < cand->rows) { set_error("col_block smaller than the number of candidates"); return B200KGE_ERR_INVALID; }
  Rows Sr = to_rows(s), Pr = to_rows(p), Or = to_rows(o), C = to_rows(cand);
  Arena ws{(uint8_t*)workspace, workspace_bytes, 0};
  EpiParams P = empty_epi();
  P.out = out; P.ldo = ldo; P.n_rows_out = n; P.col_block = col_block;
  P.n_peers = n_peers;
  for (int g = 0; g < n_peers; ++g) P.out_peer[g] = peer_out[g];
  if (model == B200KGE_CP) { set_error("the broadcast store is not offered for CP"); return B200KGE_ERR_UNSUPPORTED; }
  Block B{model, B200KGE_SP_, &Sr, &Or, &Pr, &C, n};
  return run_block(B, l_norm, precision, EPI_STORE, P, ws, (cudaStream_t)stream, nullptr);
}

int b200kge_score_1vsN_loss(int model, int combine, float l_norm, int precision,
                            const b200kge_rows_t* q, const b200kge_rows_t* p,
                            const b200kge_rows_t* cand, int64_t n, const b200kge_labels_t* labels,
                            int loss_kind, float offset, float* loss_out, float* row_loss_out,
                            void* workspace, size_t workspace_bytes, b200kge_stream_t stream) {
  int rc = check_1vsN_args(model, combine, q, p, cand, n); if (rc) return rc;
  if ((rc = validate_norm(model, l_norm))) return rc;
  if (!labels || (!labels->idx) == (!labels->dense)) { set_error("exactly one of labels.idx / labels.dense must be given"); return B200KGE_ERR_INVALID; }
  if (loss_kind != B200KGE_LOSS_BCE && loss_kind != B200KGE_LOSS_KL) { set_error("unknown loss kind %d", loss_kind); return B200KGE_ERR_INVALID; }
  if (!loss_out) { set_error("loss_out is null"); return B200KGE_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  Rows Q = to_rows(q), Pr = to_rows(p), C = to_rows(cand);
  Arena ws{(uint8_t*)workspace, workspace_bytes, 0};
  EpiParams P = empty_epi();
  P.label_idx = labels->idx; P.label_dense = labels->dense; P.ldl = labels->ldl;
  P.offset = (loss_kind == B200KGE_LOSS_BCE) ? offset : 0.f;
  Block B{model, combine, &Q, nullptr, &Pr, &C, n};
  int nch = 0;
  const int epi = (loss_kind == B200KGE_LOSS_BCE) ? EPI_BCE : EPI_KL;
  float* part = nullptr;
  rc = run_block(B, l_norm, precision, epi, P, ws, st, &nch, &part);
  if (rc) return rc;
  if (n == 0 || C.rows == 0) { B2K_CUDA(cudaMemsetAsync(loss_out, 0, 4, st)); return 0; }
  void* scratch = ws.take(1024);
  if (!scratch) { set_error("workspace too small"); return B200KGE_ERR_WORKSPACE; }
  return launch_loss_finalize(loss_kind, part, nch, n, loss_out, row_loss_out, 1.0f, 0, scratch, 0, st);
}

int b200kge_score_1vsN_rank(int model, int combine, float l_norm, int precision,
                            const b200kge_rows_t* q, const b200kge_rows_t* p,
                            const b200kge_rows_t* cand, int64_t n, const float* true_score,
                            const float* filter, int64_t ldf, float rtol, float atol,
                            int64_t* rank, int64_t* ties, void* workspace, size_t workspace_bytes,
                            b200kge_stream_t stream) {
  int rc = check_1vsN_args(model, combine, q, p, cand, n); if (rc) return rc;
  if ((rc = validate_norm(model, l_norm))) return rc;
  if (!true_score || !rank || !ties) { set_error("null rank operand"); return B200KGE_ERR_INVALID; }
  Rows Q = to_rows(q), Pr = to_rows(p), C = to_rows(cand);
  Arena ws{(uint8_t*)workspace, workspace_bytes, 0};
  EpiParams P = empty_epi();
  P.true_score = true_score; P.filter = filter; P.ldf = ldf; P.rtol = rtol; P.atol = atol;
  P.rank = reinterpret_cast<unsigned long long*>(rank);
  P.ties = reinterpret_cast<unsigned long long*>(ties);
  Block B{model, combine, &Q, nullptr, &Pr, &C, n};
  return run_block(B, l_norm, precision, EPI_RANK, P, ws, (cudaStream_t)stream, nullptr);
}

int b200kge_rank_sp_po(int model, float l_norm, int precision, const b200kge_rows_t* s,
                       const b200kge_rows_t* p, const b200kge_rows_t* o, const b200kge_rows_t* cand,
                       int64_t n, const float* true_score, const float* filter, int64_t ldf, float rtol,
                       float atol, int64_t* rank, int64_t* ties, void* workspace, size_t workspace_bytes,
                       b200kge_stream_t stream) {
  int rc = check_1vsN_args(model, B200KGE_SP_, s, p, cand, n); if (rc) return rc;
  if ((rc = check_1vsN_args(model, B200KGE__PO, o, p, cand, n))) return rc;
  if ((rc = validate_norm(model, l_norm))) return rc;
  if (!true_score || !rank || !ties) { set_error("null rank operand"); return B200KGE_ERR_INVALID; }
  Rows Sr = to_rows(s), Pr = to_rows(p), Or = to_rows(o), C = to_rows(cand);
  Arena ws{(uint8_t*)workspace, workspace_bytes, 0};
  EpiParams P = empty_epi();
  P.true_score = true_score; P.filter = filter; P.ldf = ldf; P.rtol = rtol; P.atol = atol;
  P.rank = reinterpret_cast<unsigned long long*>(rank);
  P.ties = reinterpret_cast<unsigned long long*>(ties);
  Block B{model, B200KGE_SP_, &Sr, &Or, &Pr, &C, n};
  return run_block(B, l_norm, precision, EPI_RANK, P, ws, (cudaStream_t)stream, nullptr);
}

int b200kge_rank_sp_po_csr(int model, float l_norm, int precision, const b200kge_rows_t* s,
                           const b200kge_rows_t* p, const b200kge_rows_t* o, const b200kge_rows_t* cand,
                           int64_t n, const float* true_score, const int64_t* filter_off,
                           const int64_t* filter_col, const int64_t* own_col, float rtol, float atol,
                           int64_t* rank, int64_t* ties, void* workspace, size_t workspace_bytes,
                           b200kge_stream_t stream) {
  int rc = check_1vsN_args(model, B200KGE_SP_, s, p, cand, n); if (rc) return rc;
  if ((rc = check_1vsN_args(model, B200KGE__PO, o, p, cand, n))) return rc;
  if ((rc = validate_norm(model, l_norm))) return rc;
  if (!true_score || !rank || !ties || !filter_off) { set_error("null rank operand"); return B200KGE_ERR_INVALID; }
  Rows Sr = to_rows(s), Pr = to_rows(p), Or = to_rows(o), C = to_rows(cand);
  Arena ws{(uint8_t*)workspace, workspace_bytes, 0};
  EpiParams P = empty_epi();
  P.true_score = true_score; P.rtol = rtol; P.atol = atol;
  P.rank = reinterpret_cast<unsigned long long*>(rank);
  P.ties = reinterpret_cast<unsigned long long*>(ties);
  P.csr_off = filter_off; P.csr_col = filter_col; P.csr_skip = own_col;
  Block B{model, B200KGE_SP_, &Sr, &Or, &Pr, &C, n};
  return run_block(B, l_norm, precision, EPI_RANK, P, ws, (cudaStream_t)stream, nullptr);
}

int b200kge_shard_gather_rows(const b200kge_rows_t* shard, int64_t lo, const int64_t* idx, int64_t n,
                              float* out, int64_t ldo, b200kge_stream_t stream) {
  if (!shard || (!idx && n > 0) || (!out && n > 0)) { set_error("null operand"); return B200KGE_ERR_INVALID; }
  if (shard->idx) { set_error("shard must be a plain table (idx == NULL)"); return B200KGE_ERR_INVALID; }
  if (ldo < shard->dim) { set_error("output row stride smaller than the row width"); return B200KGE_ERR_INVALID; }
  if (n <= 0) return 0;
  shard_gather_rows_kernel<<<(unsigned)n, 128, 0, (cudaStream_t)stream>>>(to_rows(shard), lo, idx, out, ldo);
  B2K_LAUNCH_CHECK("shard_gather_rows_kernel");
  return 0;
}

int b200kge_loss_dense(const float* scores, int64_t lds, int64_t n, int64_t m,
                       const b200kge_labels_t* labels, int loss_kind, float offset,
                       float* loss_out, float* row_loss_out, void* workspace,
                       size_t workspace_bytes, b200kge_stream_t stream) {
  if (!scores || !labels || !loss_out) { set_error("null operand"); return B200KGE_ERR_INVALID; }
  if ((!labels->idx) == (!labels->dense)) { set_error("exactly one of labels.idx / labels.dense must be given"); return B200KGE_ERR_INVALID; }
  if (loss_kind != B200KGE_LOSS_BCE && loss_kind != B200KGE_LOSS_KL) { set_error("unknown loss kind %d", loss_kind); return B200KGE_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0 || m == 0) { B2K_CUDA(cudaMemsetAsync(loss_out, 0, 4, st)); return 0; }
  Arena ws{(uint8_t*)workspace, workspace_bytes, 0};
  const int nch = loss_dense_nchunks(m);
  const int F = (loss_kind == B200KGE_LOSS_BCE) ? 2 : 5;
  EpiParams P = empty_epi();
  P.label_idx = labels->idx; P.label_dense = labels->dense; P.ldl = labels->ldl;
  P.offset = (loss_kind == B200KGE_LOSS_BCE) ? offset : 0.f;
  P.nchunks = nch;
  P.part = (float*)ws.take((size_t)n * nch * F * 4);
  if (!P.part) { set_error("workspace too small for loss partials (need %zu bytes)", (size_t)n * nch * F * 4); return B200KGE_ERR_WORKSPACE; }
  void* scratch = ws.take(1024);
  if (!scratch) { set_error("workspace too small"); return B200KGE_ERR_WORKSPACE; }
  int rc = launch_loss_dense(loss_kind, scores, lds, n, m, P, st);
  if (rc) return rc;
  return launch_loss_finalize(loss_kind, P.part, nch, n, loss_out, row_loss_out, 1.0f, 0, scratch, 0, st);
}

int b200kge_rank_dense(const float* scores, int64_t lds, int64_t n, int64_t m,
                       const float* true_score, const float* filter, int64_t ldf, float rtol,
                       float atol, int64_t* rank, int64_t* ties, b200kge_stream_t stream) {
  if (!scores || !true_score || !rank || !ties) { set_error("null operand"); return B200KGE_ERR_INVALID; }
  EpiParams P = empty_epi();
  P.true_score = true_score; P.filter = filter; P.ldf = ldf; P.rtol = rtol; P.atol = atol;
  P.rank = reinterpret_cast<unsigned long long*>(rank);
  P.ties = reinterpret_cast<unsigned long long*>(ties);
  return launch_rank_dense(scores, lds, n, m, P, (cudaStream_t)stream);
}

int b200kge_ns_score(int model, float l_norm, const b200kge_rows_t* s, const b200kge_rows_t* p,
                     const b200kge_rows_t* o, const b200kge_rows_t* slot_table, int slot,
                     const int64_t* neg, int64_t n, int64_t K, int with_positive, float* out,
                     int64_t ldo, b200kge_stream_t stream) {
  if (n == 0 || (K == 0 && !with_positive)) return 0;          // nothing to score
  if (!s || !p || !o || !slot_table || (!neg && n * K > 0) || !out) { set_error("null operand"); return B200KGE_ERR_INVALID; }
  if (slot < 0 || slot > 2) { set_error("slot must be 0 (S), 1 (P) or 2 (O)"); return B200KGE_ERR_INVALID; }
  int rc = validate_model(model, to_rows(s), to_rows(p)); if (rc) return rc;
  if ((rc = validate_norm(model, l_norm))) return rc;
  if (slot_table->idx) { set_error("slot_table must be a plain table (idx == NULL)"); return B200KGE_ERR_INVALID; }
  if (slot_table->dim != (slot == 1 ? p->dim : s->dim)) { set_error("slot_table width does not match the slot"); return B200KGE_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  const int col0 = with_positive ? 1 : 0;
  if (with_positive) {
    rc = launch_spo(model, l_norm, to_rows(s), to_rows(p), to_rows(o), n, out, ldo, st);
    if (rc) return rc;
  }
  return launch_ns(model, l_norm, to_rows(s), to_rows(p), to_rows(o), to_rows(slot_table), slot, neg, n, K,
                   out, ldo, col0, st);
}

int b200kge_sample_uniform(uint64_t seed, uint64_t offset, int64_t vocab, int64_t n, int64_t K, int64_t* out,
                           b200kge_stream_t stream) {
  if (vocab <= 0) { set_error("vocabulary size must be positive"); return B200KGE_ERR_INVALID; }
  if (n < 0 || K < 0 || (!out && n * K > 0)) { set_error("null operand"); return B200KGE_ERR_INVALID; }
  return launch_sample_uniform(seed, offset, vocab, n * K, out, (cudaStream_t)stream);
}

int b200kge_train_1vsall_forward(int model, float l_norm, int precision,
                                 const b200kge_rows_t* ent, const b200kge_rows_t* rel,
                                 const int64_t* triples, int64_t n, int loss_kind, float offset,
                                 float* loss_out, void* workspace, size_t workspace_bytes,
                                 b200kge_stream_t stream) {
  if (!ent || !rel || !triples || !loss_out) { set_error("null operand"); return B200KGE_ERR_INVALID; }
  if (ent->idx || rel->idx) { set_error("ent/rel must be plain tables"); return B200KGE_ERR_INVALID; }
  int rc = validate_model(model, to_rows(ent), to_rows(rel)); if (rc) return rc;
  if ((rc = validate_norm(model, l_norm))) return rc;
  if (loss_kind != B200KGE_LOSS_BCE && loss_kind != B200KGE_LOSS_KL) { set_error("unknown loss kind %d", loss_kind); return B200KGE_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  if (n <= 0) { B2K_CUDA(cudaMemsetAsync(loss_out, 0, 4, st)); return 0; }
  Arena ws{(uint8_t*)workspace, workspace_bytes, 0};
  Rows E = to_rows(ent), R = to_rows(rel);
  const int epi = (loss_kind == B200KGE_LOSS_BCE) ? EPI_BCE : EPI_KL;
  const float scale = 1.0f / (float)n;       // "/ batch_size"   train_1vsAll.py:65,76
  Folded f0 = folded_problem(model, B200KGE_SP_, E.dim, l_norm), f1 = folded_problem(model, B200KGE__PO, E.dim, l_norm);
  {
    // Pre-split tensor-core path (dot family except CP, whose directions read different table columns): the whole
    // step is THREE launches — prologue (gather + both folds + operand split of queries and table + labels), the
    // scorer with the loss reduction in its epilogue, and the fixed-order finaliser.
    const char* env_v = getenv("B200KGE_TC_VERSION");
    const int tcv = (env_v && atoi(env_v) == 1) ? 1 : (use_pair_kernel(n, f0.K) ? 4 : 3);
    const int K = f0.K;
    const bool presplit = f0.col_off == f1.col_off && f0.pair_op == PAIR_DOT && model != B200KGE_CP &&
                          (precision == B200KGE_PREC_AUTO || precision == B200KGE_PREC_F16X3) && K >= 32 && K <= 1024 &&
                          n >= 16 && E.rows < (1ll << 31) && (tcv == 3 || tcv == 4) &&
                          ((size_t)round_up(K, 64) + (model == B200KGE_RESCAL ? E.dim : 0)) * 4 <= 48 * 1024;
    if (presplit) {
      const int64_t nq = 2 * n, m = E.rows;
      const int Kp = (int)round_up(K, 64);
      SplitSet SQ{nullptr, 0, nullptr, 0, nq, nq, K, Kp, nullptr, nullptr, nullptr};
      SplitSet ST{E.base, E.ld, nullptr, f0.col_off, m, m + 32, K, Kp, nullptr, nullptr, nullptr};
      SQ.hi = ws.take((size_t)nq * Kp * 2); SQ.lo = ws.take((size_t)nq * Kp * 2);
      SQ.inv_scale = (float*)ws.take((size_t)nq * 4);
      ST.hi = ws.take((size_t)m * Kp * 2); ST.lo = ws.take((size_t)m * Kp * 2);
      ST.inv_scale = (float*)ws.take((size_t)(m + 32) * 4);
      int64_t* lab = (int64_t*)ws.take((size_t)n * 2 * 8);
      uint8_t* scratch = (uint8_t*)ws.take(1024);
      const int nch = tcv == 4 ? tc4_nchunks(nq, m) : tc3_nchunks(nq, m);
      const int F = (loss_kind == B200KGE_LOSS_BCE) ? 2 : 5;
      float* part = (float*)ws.take((size_t)nq * nch * F * 4);
      if (!SQ.hi || !SQ.lo || !SQ.inv_scale || !ST.hi || !ST.lo || !ST.inv_scale || !lab || !scratch || !part) {
        set_error("workspace too small");
        return B200KGE_ERR_WORKSPACE;
      }
      unsigned int* ticket = reinterpret_cast<unsigned int*>(scratch + 512);
      if ((rc = launch_prep_split_1vsall(model, E, R, triples, n, SQ, ST, lab, ticket, st))) return rc;
      EpiParams P = empty_epi();
      P.label_idx = lab;
      P.offset = (loss_kind == B200KGE_LOSS_BCE) ? offset : 0.f;
      P.part = part; P.nchunks = nch;
      // (finalising inside the scorer — last CTA per query tile, then last tile — was measured: +8 us in the kernel
      // against 5.8 us for the separate fixed-order finaliser, profiles/r2_summary.md; the finaliser stays separate)
      if ((rc = tcv == 4 ? launch_pairwise_tc4(epi, SQ, ST, P, st) : launch_pairwise_tc3(epi, SQ, ST, P, st))) return rc;
      return launch_loss_finalize(loss_kind, part, nch, nq, loss_out, nullptr, scale, 0, scratch, 1, st);
    }
  }
  if (f0.col_off == f1.col_off) {
    // sp_ and _po rows stacked into ONE problem of 2n query rows against the same table:
    // prologue (unpack + both folds + labels) = 1 launch, scoring + loss + finalisation = 1 launch
    const int64_t ldq = round_up(f0.K, 32);
    float* Q = (float*)ws.take((size_t)(2 * n) * ldq * 4);
    int64_t* lab = (int64_t*)ws.take((size_t)n * 2 * 8);
    uint8_t* scratch = (uint8_t*)ws.take(1024);      // finaliser scratch: block sums + ticket (+512)
    if (!Q || !lab || !scratch) { set_error("workspace too small"); return B200KGE_ERR_WORKSPACE; }
    rc = launch_prep_1vsall(model, E, R, triples, n, Q, ldq, lab, reinterpret_cast<unsigned int*>(scratch + 512), st);
    if (rc) return rc;
    Rows S = E; S.idx = lab; S.rows = n;           // placeholders: operands are pre-folded
    Rows Pr = R; Pr.idx = lab; Pr.rows = n;
    EpiParams P = empty_epi();
    P.label_idx = lab;
    P.offset = (loss_kind == B200KGE_LOSS_BCE) ? offset : 0.f;
    Block B{model, B200KGE_SP_, &S, &S, &Pr, &E, n};
    B.Qpre = Q;
    int nch = 0;
    float* part = nullptr;
    rc = run_block(B, l_norm, precision, epi, P, ws, st, &nch, &part);
    if (rc) return rc;
    return launch_loss_finalize(loss_kind, part, nch, 2 * n, loss_out, nullptr, scale, 0, scratch, 1, st);
  }
  int64_t* sidx = (int64_t*)ws.take((size_t)n * 8);
  int64_t* pidx = (int64_t*)ws.take((size_t)n * 8);
  int64_t* oidx = (int64_t*)ws.take((size_t)n * 8);
  int64_t* lab = (int64_t*)ws.take((size_t)n * 2 * 8);
  if (!sidx || !pidx || !oidx || !lab) { set_error("workspace too small"); return B200KGE_ERR_WORKSPACE; }
  unpack_triples_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(triples, n, sidx, pidx, oidx, lab);
  B2K_LAUNCH_CHECK("unpack_triples_kernel");
  Rows S = E; S.idx = sidx; S.rows = n;
  Rows O = E; O.idx = oidx; O.rows = n;
  Rows Pr = R; Pr.idx = pidx; Pr.rows = n;
  EpiParams P = empty_epi();
  P.label_idx = lab;
  P.offset = (loss_kind == B200KGE_LOSS_BCE) ? offset : 0.f;
  for (int dir = 0; dir < 2; ++dir) {   // CP: the two directions read different table columns
    Arena w2 = ws;
    EpiParams Pd = P;
    Pd.label_idx = lab + dir * n;
    Block B{model, dir, dir == 0 ? &S : &O, nullptr, &Pr, &E, n};
    int nch = 0;
    float* part = nullptr;
    rc = run_block(B, l_norm, precision, epi, Pd, w2, st, &nch, &part);
    if (rc) return rc;
    void* scratch = w2.take(1024);
    if (!scratch) { set_error("workspace too small"); return B200KGE_ERR_WORKSPACE; }
    rc = launch_loss_finalize(loss_kind, part, nch, n, loss_out, nullptr, scale, dir, scratch, 0, st);
    if (rc) return rc;
  }
  return 0;
}

int b200kge_train_1vsall_forward_host(int model, float l_norm, int precision,
                                      const b200kge_rows_t* ent, const b200kge_rows_t* rel,
                                      const int64_t* triples_host, int64_t n, int loss_kind,
                                      float offset, float* loss_host, void* workspace,
                                      size_t workspace_bytes, b200kge_stream_t stream) {
  if (!triples_host || !loss_host) { set_error("null operand"); return B200KGE_ERR_INVALID; }
  if (n <= 0) { *loss_host = 0.f; return 0; }
  cudaStream_t st = (cudaStream_t)stream;
  Arena ws{(uint8_t*)workspace, workspace_bytes, 0};
  int64_t* tri = (int64_t*)ws.take((size_t)n * 3 * 8);
  float* loss_dev = (float*)ws.take(256);
  if (!tri || !loss_dev) { set_error("workspace too small"); return B200KGE_ERR_WORKSPACE; }
  // triples.to(device)   train_1vsAll.py:59
  B2K_CUDA(cudaMemcpyAsync(tri, triples_host, (size_t)n * 3 * 8, cudaMemcpyHostToDevice, st));
  size_t used = (ws.off + 255) & ~size_t(255);
  int rc = b200kge_train_1vsall_forward(model, l_norm, precision, ent, rel, tri, n, loss_kind, offset, loss_dev,
                                        ws.base + used, workspace_bytes - used, stream);
  if (rc) return rc;
  // .item()   train_1vsAll.py:66,77
  B2K_CUDA(cudaMemcpyAsync(loss_host, loss_dev, 4, cudaMemcpyDeviceToHost, st));
  B2K_CUDA(cudaStreamSynchronize(st));
  return 0;
}

}  // extern "C"

// ==================================================================================================
// Pre-split fp16 GEMM, the analytic backward of the 1vsAll step for the dot family (grad.cu), penalties, row
// normalisation, negative-sampling backward, CSR-label losses: SURVEY 8f rows, validated on a B200 in round 2.
namespace {

bool take_planes(Arena& ws, SplitSet& S) {
  S.hi = ws.take((size_t)S.rows * S.Kp * 2);
  S.lo = ws.take((size_t)S.rows * S.Kp * 2);
  S.inv_scale = (float*)ws.take((size_t)S.rows_pad * 4);
  return S.hi && S.lo && S.inv_scale;
}
size_t planes_bytes(int64_t rows, int64_t rows_pad, int64_t Kp) {
  return 2 * ((size_t)rows * Kp * 2 + 256) + (size_t)rows_pad * 4 + 256;
}

// C[M,N] = A B^T on planes: A = "queries" (rows M), B = "table" (rows N, inv_scale padded to N+32)
// Long reductions are split into 512-element segments accumulated in fp32 (C is zeroed here first).
int gemm_planes(const SplitSet& A, const SplitSet& B, float* C, int64_t ldc, cudaStream_t st) {
  EpiParams P = empty_epi();
  P.out = C; P.ldo = ldc;
  if (A.Kp > 512) {
    P.accumulate_out = 1;
    cudaError_t e = cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)B.rows * 4, (size_t)A.rows, st);
    if (e != cudaSuccess) return check_cuda(e, "cudaMemset2DAsync(gemm output)");
  }
  return launch_pairwise_tc3(EPI_STORE, A, B, P, st);
}

// One block of the backward up to dQ: nq folded query rows Q [nq, ldq] with one-hot labels lab [nq] against
// the candidate columns [off, off+K) of the entity table; stores dT into those columns of d_ent and dQ
// [nq, ldq] into the caller's buffer.  dir as in launch_unfold.
int backward_block(int model, const Rows& E, const Rows& R, const int64_t* triples, int64_t n, int dir,
                   const float* Q, int64_t ldq, const int64_t* lab, int col_off, int K, int loss_kind, float offset,
                   float* d_ent, int64_t lde, float* dQ, Arena ws, cudaStream_t st,
                   const float* Gdense = nullptr, int64_t ldg = 0, const int64_t* csr_off = nullptr,
                   const int64_t* csr_col = nullptr, float csr_a = 1.f, float csr_b = 0.f, float inv_batch = 0.f) {
  const int64_t nq = dir < 0 ? 2 * n : n, m = E.rows;
  const int64_t ldz = round_up(m, 4), Ep = round_up(m, 64), Np = round_up(nq, 64);
  const int64_t ldE = round_up(m, 4), ldN = round_up(nq, 4);
  int rc;
  SplitSet SG{nullptr, 0, nullptr, 0, nq, nq, (int)m, (int)Ep, nullptr, nullptr, nullptr};
  SplitSet SGT{nullptr, 0, nullptr, 0, m, m, (int)nq, (int)Np, nullptr, nullptr, nullptr};
  if (Gdense) {
    // the caller's dL/dz (autograd through a dense score matrix): planes of G and of its transpose, row scaled
    float* Gt = (float*)ws.take((size_t)m * ldN * 4);
    if (!Gt) { set_error("workspace too small for the transposed gradient"); return B200KGE_ERR_WORKSPACE; }
    if ((rc = launch_transpose(Gdense, ldg, nq, m, Gt, ldN, st))) return rc;
    SG.src = Gdense; SG.ld = ldg;
    SGT.src = Gt; SGT.ld = ldN;
    if (!take_planes(ws, SG) || !take_planes(ws, SGT)) { set_error("workspace too small for the gradient planes"); return B200KGE_ERR_WORKSPACE; }
    if ((rc = launch_presplit(SGT, SG, st))) return rc;
  } else {
  // 1. scores through the validated scorer (plain-store epilogue)
  float* z = (float*)ws.take((size_t)nq * ldz * 4);
  if (!z) { set_error("workspace too small for the score matrix"); return B200KGE_ERR_WORKSPACE; }
  {
    Rows S = E; S.idx = lab; S.rows = n;      // placeholders: operands are pre-folded
    Rows Pr = R; Pr.idx = lab; Pr.rows = n;
    EpiParams P = empty_epi();
    P.out = z; P.ldo = ldz;
    Block B{model, dir <= 0 ? B200KGE_SP_ : B200KGE__PO, &S, dir < 0 ? &S : nullptr, &Pr, &E, n};
    B.Qpre = Q;
    if ((rc = run_block(B, 1.0f, B200KGE_PREC_AUTO, EPI_STORE, P, ws, st, nullptr))) return rc;
  }
  // 2. G = sigmoid(z + off) - y as planes, both layouts
  if (!take_planes(ws, SG) || !take_planes(ws, SGT)) { set_error("workspace too small for the gradient planes"); return B200KGE_ERR_WORKSPACE; }
  float* row_stat = nullptr;
  if (loss_kind == B200KGE_LOSS_KL) {
    row_stat = (float*)ws.take((size_t)nq * 2 * 4);
    if (!row_stat) { set_error("workspace too small for the row statistics"); return B200KGE_ERR_WORKSPACE; }
  }
  if (csr_off) {
    if ((rc = launch_grad_planes_csr(z, ldz, nq, m, csr_off, csr_col, csr_a, csr_b, row_stat,
                                     loss_kind == B200KGE_LOSS_KL ? 0.f : offset, inv_batch, SG.hi, SG.lo, Ep, SGT.hi,
                                     SGT.lo, Np, SG.inv_scale, SGT.inv_scale, st))) return rc;
  } else
  if ((rc = launch_grad_planes(z, ldz, nq, m, lab, nullptr, 0, row_stat, loss_kind == B200KGE_LOSS_KL ? 0.f : offset,
                               1.0f / (float)n, SG.hi, SG.lo, Ep, SGT.hi, SGT.lo, Np, SG.inv_scale, SGT.inv_scale,
                               st))) return rc;
  }
  // 3. transposed operands T^T [K, E] and Q^T [K, nq], then their planes
  float* Tt = (float*)ws.take((size_t)K * ldE * 4);
  float* Qt = (float*)ws.take((size_t)K * ldN * 4);
  if (!Tt || !Qt) { set_error("workspace too small for the transposed operands"); return B200KGE_ERR_WORKSPACE; }
  if ((rc = launch_transpose(E.base + col_off, E.ld, m, K, Tt, ldE, st))) return rc;
  if ((rc = launch_transpose(Q, ldq, nq, K, Qt, ldN, st))) return rc;
  SplitSet STt{Tt, ldE, nullptr, 0, K, K + 32, (int)m, (int)Ep, nullptr, nullptr, nullptr};
  SplitSet SQt{Qt, ldN, nullptr, 0, K, K + 32, (int)nq, (int)Np, nullptr, nullptr, nullptr};
  if (!take_planes(ws, STt) || !take_planes(ws, SQt)) { set_error("workspace too small for the operand planes"); return B200KGE_ERR_WORKSPACE; }
  if ((rc = launch_presplit(STt, SQt, st))) return rc;
  // 4. dT = G^T Q -> entity-table gradient columns [off, off+K) (overwrites);  dQ = G T
  if ((rc = gemm_planes(SGT, SQt, d_ent + col_off, lde, st))) return rc;
  return gemm_planes(SG, STt, dQ, ldq, st);
  // 5. (caller) unfold dQ into the rows of the batch — after EVERY block has stored its dT columns
}

size_t backward_block_bytes(int64_t nq, int64_t m, int64_t K, int64_t ldq) {
  const int64_t Ep = round_up(m, 64), Np = round_up(nq, 64);
  size_t b = (size_t)nq * round_up(m, 4) * 4 + 256;
  b += b200kge_workspace_bytes(0, nq, m, (int32_t)K, 0);
  b += planes_bytes(nq, nq, Ep) + planes_bytes(m, m, Np);
  b += (size_t)K * round_up(m, 4) * 4 + (size_t)K * round_up(nq, 4) * 4 + 2 * (size_t)nq * ldq * 4 + (size_t)nq * 8 + 5 * 256;
  b += planes_bytes(K, K + 32, Ep) + planes_bytes(K, K + 32, Np);
  return b;
}

}  // namespace

extern "C" {

size_t b200kge_gemm_nt_workspace_bytes(int64_t M, int64_t N, int64_t K) {
  const int64_t Kp = round_up(K, 64);
  return planes_bytes(M, M, Kp) + planes_bytes(N, N + 32, Kp) + 1024;
}

int b200kge_gemm_nt(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t M, int64_t N, int64_t K,
                      float* C, int64_t ldc, void* workspace, size_t workspace_bytes, b200kge_stream_t stream) {
  if (!A || !B || !C) { set_error("null operand"); return B200KGE_ERR_INVALID; }
  if (M < 0 || N < 0 || K <= 0 || N >= (1ll << 31)) { set_error("bad GEMM shape"); return B200KGE_ERR_INVALID; }
  if (M == 0 || N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  Arena ws{(uint8_t*)workspace, workspace_bytes, 0};
  const int Kp = (int)round_up(K, 64);
  SplitSet SA{A, lda, nullptr, 0, M, M, (int)K, Kp, nullptr, nullptr, nullptr};
  SplitSet SB{B, ldb, nullptr, 0, N, N + 32, (int)K, Kp, nullptr, nullptr, nullptr};
  if (!take_planes(ws, SA) || !take_planes(ws, SB)) { set_error("workspace too small for the operand planes"); return B200KGE_ERR_WORKSPACE; }
  int rc = launch_presplit(SB, SA, st);
  if (rc) return rc;
  return gemm_planes(SA, SB, C, ldc, st);
}

size_t b200kge_train_1vsall_backward_workspace_bytes(int model, int64_t n, int64_t E, int32_t D) {
  const int64_t K = (model == B200KGE_CP) ? D / 2 : D;
  const int64_t nq = 2 * n, ldq = round_up(K, 32);
  if (model == B200KGE_TRANSE || model == B200KGE_ROTATE)   // Q, dQ, labels, z, G, G^T, z^T, row stats + scorer workspace
    return 2 * (size_t)nq * ldq * 4 + (size_t)nq * 8 + 2 * (size_t)nq * round_up(E, 4) * 4 + 2 * (size_t)E * round_up(nq, 4) * 4 +
           (size_t)nq * 8 + 16 * 256 + b200kge_workspace_bytes(model, n, E, D, 0);
  return (size_t)nq * ldq * 4 + (size_t)n * 5 * 8 + 4096 + backward_block_bytes(nq, E, K, ldq);
}

int b200kge_train_1vsall_backward(int model, float l_norm, const b200kge_rows_t* ent, const b200kge_rows_t* rel,
                                    const int64_t* triples, int64_t n, int loss_kind, float offset, float* d_ent,
                                    int64_t lde, float* d_rel, int64_t ldr, void* workspace, size_t workspace_bytes,
                                    b200kge_stream_t stream) {
  if (!ent || !rel || !triples || !d_ent || !d_rel) { set_error("null operand"); return B200KGE_ERR_INVALID; }
  if (ent->idx || rel->idx) { set_error("ent/rel must be plain tables"); return B200KGE_ERR_INVALID; }
  int rc = validate_model(model, to_rows(ent), to_rows(rel)); if (rc) return rc;
  if ((rc = validate_norm(model, l_norm))) return rc;
  if (loss_kind != B200KGE_LOSS_BCE && loss_kind != B200KGE_LOSS_KL) { set_error("unknown loss kind %d", loss_kind); return B200KGE_ERR_INVALID; }
  if (lde < ent->dim || ldr < rel->dim) { set_error("gradient leading dimensions are smaller than the table widths"); return B200KGE_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  Rows E = to_rows(ent), R = to_rows(rel);
  B2K_CUDA(cudaMemsetAsync(d_rel, 0, (size_t)R.rows * ldr * 4, st));
  if (n <= 0) { B2K_CUDA(cudaMemsetAsync(d_ent, 0, (size_t)E.rows * lde * 4, st)); return 0; }
  Arena ws{(uint8_t*)workspace, workspace_bytes, 0};
  if (model == B200KGE_TRANSE || model == B200KGE_ROTATE) {
    // distance family (grad_distance.cu): scores by the CUDA-core scorer, dense G, two row-gradient passes, unfold
    Folded f = folded_problem(model, B200KGE_SP_, E.dim, l_norm);
    if (f.pair_op != PAIR_L1 && f.pair_op != PAIR_L2 && f.pair_op != PAIR_CMOD_L1) {
      set_error("the distance-family backward covers l_norm 1 and 2 (TransE) and 1 (RotatE)");
      return B200KGE_ERR_UNSUPPORTED;
    }
    const int64_t nq = 2 * n, m = E.rows, ldq = round_up(f.K, 32), ldz = round_up(m, 4), ldN = round_up(nq, 4);
    float* Q = (float*)ws.take((size_t)nq * ldq * 4);
    float* dQ = (float*)ws.take((size_t)nq * ldq * 4);
    int64_t* lab = (int64_t*)ws.take((size_t)nq * 8);
    float* z = (float*)ws.take((size_t)nq * ldz * 4);
    float* G = (float*)ws.take((size_t)nq * ldz * 4);
    float* Gt = (float*)ws.take((size_t)m * ldN * 4);
    float* row_stat = loss_kind == B200KGE_LOSS_KL ? (float*)ws.take((size_t)nq * 2 * 4) : nullptr;
    if (!Q || !dQ || !lab || !z || !G || !Gt || (loss_kind == B200KGE_LOSS_KL && !row_stat)) {
      set_error("workspace too small");
      return B200KGE_ERR_WORKSPACE;
    }
    if ((rc = launch_prep_1vsall(model, E, R, triples, n, Q, ldq, lab, nullptr, st))) return rc;
    {
      Rows S = E; S.idx = lab; S.rows = n;      // placeholders: operands are pre-folded
      Rows Pr = R; Pr.idx = lab; Pr.rows = n;
      EpiParams P = empty_epi();
      P.out = z; P.ldo = ldz;
      Block B{model, B200KGE_SP_, &S, &S, &Pr, &E, n};
      B.Qpre = Q;
      if ((rc = run_block(B, l_norm, B200KGE_PREC_AUTO, EPI_STORE, P, ws, st, nullptr))) return rc;
    }
    if (row_stat && (rc = launch_row_lse(z, ldz, nq, m, lab, row_stat, st))) return rc;
    if ((rc = launch_grad_dense(z, ldz, nq, m, lab, row_stat, loss_kind == B200KGE_LOSS_KL ? 0.f : offset, 1.0f / (float)n,
                                f.pair_op == PAIR_L2, G, ldz, st))) return rc;
    if ((rc = launch_transpose(G, ldz, nq, m, Gt, ldN, st))) return rc;
    // each pass reads its weights transposed ([column, row]): the other pass's orientation
    if ((rc = launch_pair_rowgrad(f.pair_op, Q, ldq, nq, E.base, E.ld, m, f.K, Gt, ldN, dQ, ldq, st))) return rc;
    if ((rc = launch_pair_rowgrad(f.pair_op, E.base, E.ld, m, Q, ldq, nq, f.K, G, ldz, d_ent, lde, st))) return rc;
    return launch_unfold_distance(model, E, R, triples, n, -1, dQ, ldq, d_ent, lde, d_rel, ldr, st);
  }
  Folded f0 = folded_problem(model, B200KGE_SP_, E.dim, 1.0f), f1 = folded_problem(model, B200KGE__PO, E.dim, 1.0f);
  const int64_t ldq = round_up(f0.K, 32);
  if (f0.col_off == f1.col_off) {
    float* Q = (float*)ws.take((size_t)(2 * n) * ldq * 4);
    int64_t* lab = (int64_t*)ws.take((size_t)n * 2 * 8);
    if (!Q || !lab) { set_error("workspace too small"); return B200KGE_ERR_WORKSPACE; }
    float* dQ = (float*)ws.take((size_t)(2 * n) * ldq * 4);
    if (!dQ) { set_error("workspace too small"); return B200KGE_ERR_WORKSPACE; }
    if ((rc = launch_prep_1vsall(model, E, R, triples, n, Q, ldq, lab, nullptr, st))) return rc;
    if ((rc = backward_block(model, E, R, triples, n, -1, Q, ldq, lab, f0.col_off, f0.K, loss_kind, offset, d_ent, lde, dQ, ws, st))) return rc;
    return launch_unfold(model, E, R, triples, n, -1, dQ, ldq, d_ent, lde, d_rel, ldr, st);
  }
  // CP: the two directions pair with different halves of the candidate columns
  int64_t* sidx = (int64_t*)ws.take((size_t)n * 8);
  int64_t* pidx = (int64_t*)ws.take((size_t)n * 8);
  int64_t* oidx = (int64_t*)ws.take((size_t)n * 8);
  int64_t* lab = (int64_t*)ws.take((size_t)n * 2 * 8);
  float* Q = (float*)ws.take((size_t)n * ldq * 4);
  float* dQ2 = (float*)ws.take((size_t)(2 * n) * ldq * 4);
  if (!sidx || !pidx || !oidx || !lab || !Q || !dQ2) { set_error("workspace too small"); return B200KGE_ERR_WORKSPACE; }
  unpack_triples_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(triples, n, sidx, pidx, oidx, lab);
  B2K_LAUNCH_CHECK("unpack_triples_kernel");
  Rows S = E; S.idx = sidx; S.rows = n;
  Rows O = E; O.idx = oidx; O.rows = n;
  Rows Pr = R; Pr.idx = pidx; Pr.rows = n;
  for (int dir = 0; dir < 2; ++dir) {
    const Folded& f = dir == 0 ? f0 : f1;
    if ((rc = launch_fold_queries(model, dir, dir == 0 ? S : O, Pr, n, 0, Q, ldq, st))) return rc;
    if ((rc = backward_block(model, E, R, triples, n, dir, Q, ldq, lab + dir * n, f.col_off, f.K, loss_kind, offset, d_ent,
                             lde, dQ2 + (size_t)dir * n * ldq, ws, st))) return rc;
  }
  // both halves of the dense column gradient are stored: now add the batch rows' own gradients
  for (int dir = 0; dir < 2; ++dir)
    if ((rc = launch_unfold(model, E, R, triples, n, dir, dQ2 + (size_t)dir * n * ldq, ldq, d_ent, lde, d_rel, ldr, st))) return rc;
  return 0;
}


size_t b200kge_score_1vsN_backward_workspace_bytes(int model, int64_t n, int64_t E, int32_t D) {
  const int64_t K = (model == B200KGE_CP) ? D / 2 : D;
  const int64_t ldq = round_up(K, 32);
  if (model == B200KGE_TRANSE || model == B200KGE_ROTATE)   // Q, dQ, triples, z + W (L2), W^T, scorer workspace
    return 2 * (size_t)n * ldq * 4 + (size_t)n * 4 * 8 + 2 * (size_t)n * round_up(E, 4) * 4 + (size_t)E * round_up(n, 4) * 4 +
           16 * 256 + b200kge_workspace_bytes(model, n, E, D, 0);
  return 2 * (size_t)n * ldq * 4 + (size_t)n * 4 * 8 + (size_t)E * round_up(n, 4) * 4 + 8192 + backward_block_bytes(n, E, K, ldq);
}

int b200kge_score_1vsN_backward(int model, int combine, float l_norm, const b200kge_rows_t* ent, const b200kge_rows_t* rel,
                                const int64_t* q_idx, const int64_t* p_idx, int64_t n, const float* grad_scores,
                                int64_t ldg, float* d_ent, int64_t lde, float* d_rel, int64_t ldr, void* workspace,
                                size_t workspace_bytes, b200kge_stream_t stream) {
  if (!ent || !rel || !q_idx || !p_idx || !grad_scores || !d_ent || !d_rel) { set_error("null operand"); return B200KGE_ERR_INVALID; }
  if (ent->idx || rel->idx) { set_error("ent/rel must be plain tables"); return B200KGE_ERR_INVALID; }
  if (combine != B200KGE_SP_ && combine != B200KGE__PO) { set_error("cannot handle combine=%d", combine); return B200KGE_ERR_INVALID; }
  int rc = validate_model(model, to_rows(ent), to_rows(rel)); if (rc) return rc;
  if ((rc = validate_norm(model, l_norm))) return rc;
  if (lde < ent->dim || ldr < rel->dim || ldg < ent->rows) { set_error("leading dimensions too small"); return B200KGE_ERR_INVALID; }
  const bool distance = (model == B200KGE_TRANSE || model == B200KGE_ROTATE);
  cudaStream_t st = (cudaStream_t)stream;
  Rows E = to_rows(ent), R = to_rows(rel);
  Folded f = folded_problem(model, combine, E.dim, l_norm);
  if (distance && f.pair_op != PAIR_L1 && f.pair_op != PAIR_L2 && f.pair_op != PAIR_CMOD_L1) {
    set_error("the distance-family backward covers l_norm 1 and 2 (TransE) and 1 (RotatE)");
    return B200KGE_ERR_UNSUPPORTED;
  }
  B2K_CUDA(cudaMemsetAsync(d_rel, 0, (size_t)R.rows * ldr * 4, st));
  B2K_CUDA(cudaMemsetAsync(d_ent, 0, (size_t)E.rows * lde * 4, st));
  if (n <= 0) return 0;
  Arena ws{(uint8_t*)workspace, workspace_bytes, 0};
  const int64_t ldq = round_up(f.K, 32);
  float* Q = (float*)ws.take((size_t)n * ldq * 4);
  float* dQ = (float*)ws.take((size_t)n * ldq * 4);
  int64_t* tri = (int64_t*)ws.take((size_t)n * 3 * 8);
  if (!Q || !dQ || !tri) { set_error("workspace too small"); return B200KGE_ERR_WORKSPACE; }
  // the unfold works on [n,3] triples: (q, p, .) for sp_, (., p, q) for _po
  pack_triples_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(q_idx, p_idx, n, combine, tri);
  B2K_LAUNCH_CHECK("pack_triples_kernel");
  Rows A = E; A.idx = q_idx; A.rows = n;
  Rows Pr = R; Pr.idx = p_idx; Pr.rows = n;
  if ((rc = launch_fold_queries(model, combine, A, Pr, n, 0, Q, ldq, st))) return rc;
  if (distance) {
    // grad_distance.cu: W = dL/dscores (L2: divided by the recomputed scores), then the two row-gradient passes
    const int64_t m = E.rows, ldz = round_up(m, 4), ldN = round_up(n, 4);
    const float* W = grad_scores;
    int64_t ldw = ldg;
    if (f.pair_op == PAIR_L2) {
      float* z = (float*)ws.take((size_t)n * ldz * 4);
      float* Wd = (float*)ws.take((size_t)n * ldz * 4);
      if (!z || !Wd) { set_error("workspace too small"); return B200KGE_ERR_WORKSPACE; }
      EpiParams P = empty_epi();
      P.out = z; P.ldo = ldz;
      Block B{model, combine, &A, nullptr, &Pr, &E, n};
      B.Qpre = Q;
      if ((rc = run_block(B, l_norm, B200KGE_PREC_AUTO, EPI_STORE, P, ws, st, nullptr))) return rc;
      if ((rc = launch_div_scores(grad_scores, ldg, z, ldz, n, m, Wd, ldz, st))) return rc;
      W = Wd; ldw = ldz;
    }
    float* Wt = (float*)ws.take((size_t)m * ldN * 4);
    if (!Wt) { set_error("workspace too small"); return B200KGE_ERR_WORKSPACE; }
    if ((rc = launch_transpose(W, ldw, n, m, Wt, ldN, st))) return rc;
    if ((rc = launch_pair_rowgrad(f.pair_op, Q, ldq, n, E.base, E.ld, m, f.K, Wt, ldN, dQ, ldq, st))) return rc;
    if ((rc = launch_pair_rowgrad(f.pair_op, E.base, E.ld, m, Q, ldq, n, f.K, W, ldw, d_ent, lde, st))) return rc;
    return launch_unfold_distance(model, E, R, tri, n, combine, dQ, ldq, d_ent, lde, d_rel, ldr, st);
  }
  if ((rc = backward_block(model, E, R, tri, n, combine, Q, ldq, nullptr, f.col_off, f.K, B200KGE_LOSS_BCE, 0.f, d_ent, lde, dQ,
                           ws, st, grad_scores, ldg))) return rc;
  return launch_unfold(model, E, R, tri, n, combine, dQ, ldq, d_ent, lde, d_rel, ldr, st);
}

int b200kge_score_1vsN_loss_csr_backward(int model, int combine, const b200kge_rows_t* ent, const b200kge_rows_t* rel,
                                         const int64_t* q_idx, const int64_t* p_idx, int64_t n, const int64_t* csr_off,
                                         const int64_t* csr_col, float label_smoothing, int loss_kind, float offset,
                                         int64_t batch_size, float* d_ent, int64_t lde, float* d_rel, int64_t ldr,
                                         void* workspace, size_t workspace_bytes, b200kge_stream_t stream) {
  if (!ent || !rel || !q_idx || !p_idx || !csr_off || !d_ent || !d_rel) { set_error("null operand"); return B200KGE_ERR_INVALID; }
  if (ent->idx || rel->idx) { set_error("ent/rel must be plain tables"); return B200KGE_ERR_INVALID; }
  if (combine != B200KGE_SP_ && combine != B200KGE__PO) { set_error("cannot handle combine=%d", combine); return B200KGE_ERR_INVALID; }
  int rc = validate_model(model, to_rows(ent), to_rows(rel)); if (rc) return rc;
  if (model > B200KGE_RESCAL) { set_error("the tensor-core backward covers the dot family only (model %d)", model); return B200KGE_ERR_UNSUPPORTED; }
  if (loss_kind != B200KGE_LOSS_BCE && loss_kind != B200KGE_LOSS_KL) { set_error("unknown loss kind %d", loss_kind); return B200KGE_ERR_INVALID; }
  if (batch_size <= 0 || !(label_smoothing >= 0.f && label_smoothing < 1.f)) { set_error("bad batch_size / label_smoothing"); return B200KGE_ERR_INVALID; }
  if (lde < ent->dim || ldr < rel->dim) { set_error("leading dimensions too small"); return B200KGE_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  Rows E = to_rows(ent), R = to_rows(rel);
  B2K_CUDA(cudaMemsetAsync(d_rel, 0, (size_t)R.rows * ldr * 4, st));
  B2K_CUDA(cudaMemsetAsync(d_ent, 0, (size_t)E.rows * lde * 4, st));
  if (n <= 0) return 0;
  Arena ws{(uint8_t*)workspace, workspace_bytes, 0};
  Folded f = folded_problem(model, combine, E.dim, 1.0f);
  const int64_t ldq = round_up(f.K, 32);
  float* Q = (float*)ws.take((size_t)n * ldq * 4);
  float* dQ = (float*)ws.take((size_t)n * ldq * 4);
  int64_t* tri = (int64_t*)ws.take((size_t)n * 3 * 8);
  if (!Q || !dQ || !tri) { set_error("workspace too small"); return B200KGE_ERR_WORKSPACE; }
  pack_triples_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(q_idx, p_idx, n, combine, tri);
  B2K_LAUNCH_CHECK("pack_triples_kernel");
  Rows A = E; A.idx = q_idx; A.rows = n;
  Rows Pr = R; Pr.idx = p_idx; Pr.rows = n;
  if ((rc = launch_fold_queries(model, combine, A, Pr, n, 0, Q, ldq, st))) return rc;
  const float a = 1.0f - label_smoothing, b = label_smoothing > 0.f ? 1.0f / (float)E.rows : 0.f;
  if ((rc = backward_block(model, E, R, tri, n, combine, Q, ldq, q_idx /* placeholder index vector */, f.col_off, f.K, loss_kind,
                           offset, d_ent, lde, dQ, ws, st, nullptr, 0, csr_off, csr_col, a, b, 1.0f / (float)batch_size))) return rc;
  return launch_unfold(model, E, R, tri, n, combine, dQ, ldq, d_ent, lde, d_rel, ldr, st);
}

int b200kge_lookup_penalty(const b200kge_rows_t* rows, const float* counts, float p, int complex_abs, float scale,
                             float* out, void* workspace, size_t workspace_bytes, b200kge_stream_t stream) {
  if (!rows || !out) { set_error("null operand"); return B200KGE_ERR_INVALID; }
  if (!(p > 0.f)) { set_error("p must be positive (got %g)", (double)p); return B200KGE_ERR_INVALID; }
  if (complex_abs && (rows->dim & 1)) { set_error("complex-space penalty needs an even embedding width"); return B200KGE_ERR_INVALID; }
  return launch_penalty(to_rows(rows), counts, p, complex_abs, scale, (float*)workspace, workspace_bytes / 4, out,
                        (cudaStream_t)stream);
}

int b200kge_normalize_rows(float* weight, int64_t ld, int64_t rows, int32_t dim, float p, b200kge_stream_t stream) {
  if (!weight && rows > 0) { set_error("null operand"); return B200KGE_ERR_INVALID; }
  return launch_normalize_rows(weight, ld, rows, dim, p, (cudaStream_t)stream);
}


int b200kge_ns_backward(int model, float l_norm, const b200kge_rows_t* ent, const b200kge_rows_t* rel,
                          const int64_t* triples, int slot, const int64_t* neg, int64_t n, int64_t K, float offset,
                          int64_t batch_size, float* d_ent, int64_t lde, float* d_rel, int64_t ldr, void* workspace,
                          size_t workspace_bytes, b200kge_stream_t stream) {
  if (!ent || !rel || !triples || (!neg && n * K > 0) || !d_ent || !d_rel) { set_error("null operand"); return B200KGE_ERR_INVALID; }
  if (ent->idx || rel->idx) { set_error("ent/rel must be plain tables"); return B200KGE_ERR_INVALID; }
  int rc = validate_model(model, to_rows(ent), to_rows(rel)); if (rc) return rc;
  if ((rc = validate_norm(model, l_norm))) return rc;
  if (batch_size <= 0) { set_error("batch_size must be positive"); return B200KGE_ERR_INVALID; }
  if (lde < ent->dim || ldr < rel->dim) { set_error("gradient leading dimensions are smaller than the table widths"); return B200KGE_ERR_INVALID; }
  Rows E = to_rows(ent), R = to_rows(rel);
  Folded f = folded_problem(model, B200KGE_SP_, E.dim, l_norm);
  const int64_t ldq = round_up(f.K, 32);
  Arena ws{(uint8_t*)workspace, workspace_bytes, 0};
  float* dQ = (float*)ws.take((size_t)n * ldq * 4);
  if (!dQ && n > 0) { set_error("workspace too small (need n * round_up(K,32) floats)"); return B200KGE_ERR_WORKSPACE; }
  return launch_ns_backward(model, l_norm, E, R, triples, slot, neg, n, K, offset, 1.0f / (float)batch_size, d_ent, lde,
                            d_rel, ldr, dQ, ldq, (cudaStream_t)stream);
}


size_t b200kge_score_1vsN_loss_csr_workspace_bytes(int model, int64_t n, int64_t m, int32_t D, int64_t nnz) {
  const int64_t ldq = round_up(D, 32), tot = nnz + n;
  size_t b = b200kge_workspace_bytes(model, n, m, D, 0);
  b += (size_t)n * 8 + 3 * ((size_t)tot * 8 + 256) + (size_t)tot * 4 + 4 * ((size_t)n * 4 + 256) + 2048;
  b += (size_t)n * ldq * 4 + ((size_t)((m + 1023) / 1024) + 1) * ldq * 4 + 512;
  return b;
}

int b200kge_score_1vsN_loss_csr(int model, int combine, float l_norm, int precision, const b200kge_rows_t* q,
                                  const b200kge_rows_t* p, const b200kge_rows_t* cand, int64_t n,
                                  const int64_t* csr_off, const int64_t* csr_col, int64_t nnz, float label_smoothing,
                                  int loss_kind, float offset, float* loss_out, float* row_loss_out, void* workspace,
                                  size_t workspace_bytes, b200kge_stream_t stream) {
  int rc = check_1vsN_args(model, combine, q, p, cand, n); if (rc) return rc;
  if ((rc = validate_norm(model, l_norm))) return rc;
  if (!csr_off || (!csr_col && nnz > 0) || !loss_out || nnz < 0) { set_error("null operand"); return B200KGE_ERR_INVALID; }
  if (cand->idx) { set_error("CSR labels address the columns of a plain candidate table (cand.idx must be NULL)"); return B200KGE_ERR_INVALID; }
  if (loss_kind != B200KGE_LOSS_BCE && loss_kind != B200KGE_LOSS_KL) { set_error("unknown loss kind %d", loss_kind); return B200KGE_ERR_INVALID; }
  if (!(label_smoothing >= 0.f && label_smoothing < 1.f)) { set_error("label_smoothing must be in [0, 1)"); return B200KGE_ERR_INVALID; }
  const bool dot = model <= B200KGE_RESCAL;
  if (label_smoothing > 0.f && !dot) { set_error("label smoothing with CSR labels is available for the dot family"); return B200KGE_ERR_UNSUPPORTED; }
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0 || cand->rows == 0) { B2K_CUDA(cudaMemsetAsync(loss_out, 0, 4, st)); return 0; }
  Rows Q = to_rows(q), Pr = to_rows(p), C = to_rows(cand);
  const int64_t m = C.rows, tot = nnz + (loss_kind == B200KGE_LOSS_KL ? n : 0);
  Arena ws{(uint8_t*)workspace, workspace_bytes, 0};
  int64_t* lab = (int64_t*)ws.take((size_t)n * 8);
  int64_t* qsel = (int64_t*)ws.take((size_t)tot * 8 + 8);
  int64_t* psel = (int64_t*)ws.take((size_t)tot * 8 + 8);
  int64_t* esel = (int64_t*)ws.take((size_t)tot * 8 + 8);
  float* zpos = (float*)ws.take((size_t)tot * 4 + 8);
  float* fused = (float*)ws.take((size_t)n * 4);
  float* rows = row_loss_out ? row_loss_out : (float*)ws.take((size_t)n * 4);
  float* total = (float*)ws.take(256);
  void* scratch = ws.take(1024);
  if (!lab || !qsel || !psel || !esel || !zpos || !fused || !rows || !total || !scratch) { set_error("workspace too small"); return B200KGE_ERR_WORKSPACE; }
  // 1. label-free fused pass: BCE with no label (index -1) -> sum_j softplus;  KL with the one-hot label at
  //    column 0 -> lse_i - z_i0.  On the pre-split tensor-core path the SAME pass also emits the scores of the listed
  //    columns (and z_i0) from its epilogue (per-thread cursor into the row's sorted CSR segment, tc_common.cuh):
  //    DRAM traffic = table + queries + nnz * 12 bytes.
  B2K_CUDA(cudaMemsetAsync(lab, loss_kind == B200KGE_LOSS_BCE ? 0xFF : 0, (size_t)n * 8, st));
  bool emitted = false;
  {
    const int epi = (loss_kind == B200KGE_LOSS_BCE) ? EPI_BCE : EPI_KL;
    for (int attempt = 0; attempt < 2; ++attempt) {
      EpiParams P = empty_epi();
      P.label_idx = lab;
      P.offset = (loss_kind == B200KGE_LOSS_BCE) ? offset : 0.f;
      if (attempt == 0) {
        P.csr_off = csr_off; P.csr_col = csr_col; P.csr_out = zpos; P.csr_nnz = nnz;
        P.csr_extra = (loss_kind == B200KGE_LOSS_KL) ? 1 : 0;
      }
      Block B{model, combine, &Q, nullptr, &Pr, &C, n};
      int nch = 0;
      float* part = nullptr;
      Arena w2 = ws;
      rc = run_block(B, l_norm, precision, epi, P, w2, st, &nch, &part);
      if (rc == B200KGE_ERR_UNSUPPORTED && attempt == 0) continue;      // not the pre-split path: compose below
      if (rc) return rc;
      emitted = (attempt == 0);
      if ((rc = launch_loss_finalize(loss_kind, part, nch, n, total, fused, 1.0f, 0, scratch, 0, st))) return rc;
      break;
    }
  }
  // 2. otherwise: scores of the listed columns (and of column 0 for KL) through the row-wise triple kernel
  if (!emitted) {
    if ((rc = launch_csr_expand(csr_off, csr_col, n, nnz, loss_kind == B200KGE_LOSS_KL ? 1 : 0, Q.idx, Pr.idx, qsel, psel,
                                esel, st))) return rc;
    if (tot > 0) {
      Rows Qs = Q; Qs.idx = qsel; Qs.rows = tot;
      Rows Ps = Pr; Ps.idx = psel; Ps.rows = tot;
      Rows Es = C; Es.idx = esel; Es.rows = tot;
      if (combine == B200KGE_SP_) rc = launch_spo(model, l_norm, Qs, Ps, Es, tot, zpos, 1, st);
      else                        rc = launch_spo(model, l_norm, Es, Ps, Qs, tot, zpos, 1, st);
      if (rc) return rc;
    }
  }
  // 3. label smoothing: sum_j z_ij = Q_i . colsum(T)   (dot family)
  float* zsum = nullptr;
  if (label_smoothing > 0.f) {
    Folded f = folded_problem(model, combine, Q.dim, l_norm);
    const int64_t ldq = round_up(f.K, 32);
    float* Qf = (float*)ws.take((size_t)n * ldq * 4);
    float* cs = (float*)ws.take(((size_t)((m + 1023) / 1024) + 1) * ldq * 4);
    zsum = (float*)ws.take((size_t)n * 4);
    if (!Qf || !cs || !zsum) { set_error("workspace too small"); return B200KGE_ERR_WORKSPACE; }
    if ((rc = launch_fold_queries(model, combine, Q, Pr, n, 0, Qf, ldq, st))) return rc;
    if ((rc = launch_row_score_sums(Qf, ldq, n, C.base + f.col_off, C.ld, m, f.K, cs, zsum, st))) return rc;
  }
  // 4. per-row combination and the scalar
  const float a = 1.0f - label_smoothing, b = label_smoothing > 0.f ? 1.0f / (float)m : 0.f;
  if ((rc = launch_csr_rows(loss_kind, csr_off, csr_col, zpos, n, nnz, fused, zsum, a, b, (float)m,
                            loss_kind == B200KGE_LOSS_BCE ? offset : 0.f, rows, st))) return rc;
  return launch_rows_sum(rows, n, 1.0f, loss_out, st);
}

}  // extern "C"
