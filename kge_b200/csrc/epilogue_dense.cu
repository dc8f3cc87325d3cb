// epilogue_dense.cu — loss / rank reductions over an already-materialised score matrix, and the
// deterministic finaliser of the fused kernels' per-(row, chunk) partial states.
//
//  * loss_dense_kernel : KgeLoss on dense [n,m] scores — BCEWithLogitsKgeLoss loss.py:153-159,
//                        KLDivWithSoftmaxKgeLoss loss.py:198-213 — one pass, no one-hot label
//                        matrix (loss.py:105-117) and no log_softmax buffer.
//  * rank_dense_kernel : EntityRankingJob._get_ranks_and_num_ties eval_entity_ranking.py:571-596
//                        (+ filter subtraction :561-566) in one pass; integer counts, bit-exact.
//  * loss_finalize_kernel: fixed-order reduction partial[n][nchunks][F] -> row losses -> scalar.
#include "common.cuh"

namespace b200kge {

namespace {

constexpr int DN_THREADS = 256, DN_CHUNK = 4096;

template <int KIND>
__device__ __forceinline__ void block_reduce_flush(const EpiParams& P, RowState<KIND>& st,
                                                   int64_t row, int chunk) {
  constexpr int W = sizeof(RowState<KIND>) / 4;
  __shared__ uint32_t red[DN_THREADS / 32][W > 0 ? W : 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  epi_lane_reduce<KIND>(st, 32);
  if (lane == 0) {
    const uint32_t* s = reinterpret_cast<const uint32_t*>(&st);
#pragma unroll
    for (int w = 0; w < W; ++w) red[warp][w] = s[w];
  }
  __syncthreads();
  if (warp == 0) {
    RowState<KIND> r;
    r.init();
    if (lane < DN_THREADS / 32) {
      uint32_t* d = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
      for (int w = 0; w < W; ++w) d[w] = red[lane][w];
    }
    epi_lane_reduce<KIND>(r, 8);
    if (lane == 0) epi_flush<KIND>(P, r, row, chunk);
  }
}

template <int KIND>
__global__ void __launch_bounds__(DN_THREADS)
dense_epilogue_kernel(const float* __restrict__ scores, int64_t lds, int64_t m, EpiParams P) {
  const int64_t row = blockIdx.y;
  const int chunk = blockIdx.x;
  const int64_t c0 = (int64_t)chunk * DN_CHUNK;
  const int64_t c1 = (c0 + DN_CHUNK < m) ? c0 + DN_CHUNK : m;
  RowState<KIND> st;
  st.init();
  const float aux = epi_row_aux<KIND>(P, row);
  const float* __restrict__ x = scores + row * lds;
  for (int64_t c = c0 + threadIdx.x; c < c1; c += DN_THREADS) epi_elem<KIND>(P, st, row, c, x[c], aux);
  block_reduce_flush<KIND>(P, st, row, chunk);
}

// One warp per row: lanes hold the row's chunk partials (coalesced), a fixed shuffle tree combines
// them; lane 0 accumulates the warp's rows in order.  Blocks publish their sums; the last block to
// finish adds them in index order.
constexpr int FIN_BLOCKS = 128, FIN_THREADS = 256;

template <int LOSS>
__global__ void __launch_bounds__(FIN_THREADS)
loss_finalize_kernel(FinalizeArgs A) {
  __shared__ float wsum[FIN_THREADS / 32];
  __shared__ int is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = FIN_THREADS / 32;
  float acc = 0.f;
  for (int64_t r = (int64_t)blockIdx.x * nw + warp; r < A.n; r += (int64_t)gridDim.x * nw) {
    float rl;
    if constexpr (LOSS == B200KGE_LOSS_BCE) {
      const float* __restrict__ p = A.part + r * A.nchunks * 2;
      float a = 0.f, b = 0.f;
      for (int c = lane; c < A.nchunks; c += 32) { a += p[2 * c]; b += p[2 * c + 1]; }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, off);
        b += __shfl_xor_sync(0xffffffffu, b, off);
      }
      rl = a - b;                       // sum softplus(z) - sum y*z
    } else {
      RowState<EPI_KL> st;
      st.init();
      for (int c = lane; c < A.nchunks; c += 32) {
        const float* __restrict__ p = A.part + (r * A.nchunks + c) * 5;
        RowState<EPI_KL> o;
        o.m = p[0]; o.s = p[1]; o.y_sum = p[2]; o.yx = p[3]; o.ylogy = p[4];
        st.combine(o);
      }
      epi_lane_reduce<EPI_KL>(st, 32);
      const float lse = st.m + logf(st.s);
      // KLDiv(log_softmax(x), y / max(||y||_1, 1e-12)), reduction sum   loss.py:209-213
      const float yc = fmaxf(st.y_sum, 1e-12f);
      const float w = st.y_sum / yc;
      rl = (st.y_sum > 0.f) ? (st.ylogy / yc - w * logf(yc) - st.yx / yc + lse * w) : 0.f;
    }
    if (lane == 0) {
      if (A.row_loss_out) A.row_loss_out[r] = rl;
      acc += rl;
    }
  }
  if (lane == 0) wsum[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float bs = 0.f;
    for (int w = 0; w < nw; ++w) bs += wsum[w];
    A.block_sums[blockIdx.x] = bs;
    __threadfence();
    is_last = (atomicAdd(A.ticket, 1u) == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (is_last && warp == 0) {
    // last block: warp 0 adds the (<= 128) block sums — lane l takes blocks l, l+32, ... in order, then a
    // fixed shuffle tree: deterministic
    __threadfence();
    float tot = 0.f;
    for (unsigned b = lane; b < gridDim.x; b += 32) tot += __ldcg(A.block_sums + b);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, off);
    if (lane == 0) {
      A.loss_out[0] = (A.accumulate ? A.loss_out[0] : 0.f) + A.scale * tot;
      *A.ticket = 0u;
    }
  }
}

}  // namespace

int loss_dense_nchunks(int64_t m) { return (int)((m + DN_CHUNK - 1) / DN_CHUNK); }

int launch_loss_dense(int loss_kind, const float* scores, int64_t lds, int64_t n, int64_t m,
                      const EpiParams& P, cudaStream_t st) {
  if (n == 0 || m == 0) return 0;
  if (n > 65535) { set_error("too many rows for one launch (%lld)", (long long)n); return B200KGE_ERR_UNSUPPORTED; }
  dim3 grid((unsigned)loss_dense_nchunks(m), (unsigned)n);
  if (loss_kind == B200KGE_LOSS_BCE) dense_epilogue_kernel<EPI_BCE><<<grid, DN_THREADS, 0, st>>>(scores, lds, m, P);
  else if (loss_kind == B200KGE_LOSS_KL) dense_epilogue_kernel<EPI_KL><<<grid, DN_THREADS, 0, st>>>(scores, lds, m, P);
  else { set_error("unknown loss kind %d", loss_kind); return B200KGE_ERR_INVALID; }
  B2K_LAUNCH_CHECK("dense_epilogue_kernel(loss)");
  return 0;
}

int launch_rank_dense(const float* scores, int64_t lds, int64_t n, int64_t m, const EpiParams& P,
                      cudaStream_t st) {
  if (n == 0 || m == 0) return 0;
  if (n > 65535) { set_error("too many rows for one launch (%lld)", (long long)n); return B200KGE_ERR_UNSUPPORTED; }
  dim3 grid((unsigned)loss_dense_nchunks(m), (unsigned)n);
  dense_epilogue_kernel<EPI_RANK><<<grid, DN_THREADS, 0, st>>>(scores, lds, m, P);
  B2K_LAUNCH_CHECK("dense_epilogue_kernel(rank)");
  return 0;
}

int launch_loss_finalize(int loss_kind, const float* part, int nchunks, int64_t n, float* loss_out,
                         float* row_loss_out, float scale, int accumulate, void* scratch,
                         int ticket_zeroed, cudaStream_t st) {
  float* block_sums = reinterpret_cast<float*>(scratch);
  unsigned int* ticket = reinterpret_cast<unsigned int*>(reinterpret_cast<uint8_t*>(scratch) + 512);
  if (!ticket_zeroed) B2K_CUDA(cudaMemsetAsync(ticket, 0, 4, st));
  int64_t want = (n + FIN_THREADS / 32 - 1) / (FIN_THREADS / 32);
  const int grid = (int)(want < 1 ? 1 : (want > FIN_BLOCKS ? FIN_BLOCKS : want));
  FinalizeArgs A{part, nchunks, n, loss_out, row_loss_out, scale, accumulate, ticket, block_sums};
  if (loss_kind == B200KGE_LOSS_BCE) loss_finalize_kernel<B200KGE_LOSS_BCE><<<grid, FIN_THREADS, 0, st>>>(A);
  else if (loss_kind == B200KGE_LOSS_KL) loss_finalize_kernel<B200KGE_LOSS_KL><<<grid, FIN_THREADS, 0, st>>>(A);
  else { set_error("unknown loss kind %d", loss_kind); return B200KGE_ERR_INVALID; }
  B2K_LAUNCH_CHECK("loss_finalize_kernel");
  return 0;
}

}  // namespace b200kge
