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

// One thread walks the chunks of a row in order (deterministic), then a fixed smem tree sums rows.
template <int LOSS>
__global__ void __launch_bounds__(1024)
loss_finalize_kernel(const float* __restrict__ part, int nchunks, int64_t n,
                     float* __restrict__ loss_out, float* __restrict__ row_loss_out, float scale,
                     int accumulate) {
  __shared__ float red[1024];
  float local = 0.f;
  for (int64_t r = threadIdx.x; r < n; r += blockDim.x) {
    float rl;
    if constexpr (LOSS == B200KGE_LOSS_BCE) {
      float a = 0.f, b = 0.f;
      for (int c = 0; c < nchunks; ++c) {
        const float* p = part + (r * nchunks + c) * 2;
        a += p[0]; b += p[1];
      }
      rl = a - b;                       // sum softplus(z) - sum y*z
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
      rl = (st.y_sum > 0.f) ? (st.ylogy / yc - w * logf(yc) - st.yx / yc + lse * w) : 0.f;
    }
    if (row_loss_out) row_loss_out[r] = rl;
    local += rl;
  }
  red[threadIdx.x] = local;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss_out[0] = (accumulate ? loss_out[0] : 0.f) + scale * red[0];
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

int launch_loss_finalize(int loss_kind, const float* part, int nchunks, int64_t n,
                         const int64_t* /*label_idx*/, float* loss_out, float* row_loss_out,
                         float scale, int accumulate, cudaStream_t st) {
  if (loss_kind == B200KGE_LOSS_BCE)
    loss_finalize_kernel<B200KGE_LOSS_BCE><<<1, 1024, 0, st>>>(part, nchunks, n, loss_out, row_loss_out, scale, accumulate);
  else if (loss_kind == B200KGE_LOSS_KL)
    loss_finalize_kernel<B200KGE_LOSS_KL><<<1, 1024, 0, st>>>(part, nchunks, n, loss_out, row_loss_out, scale, accumulate);
  else { set_error("unknown loss kind %d", loss_kind); return B200KGE_ERR_INVALID; }
  B2K_LAUNCH_CHECK("loss_finalize_kernel");
  return 0;
}

}  // namespace b200kge
