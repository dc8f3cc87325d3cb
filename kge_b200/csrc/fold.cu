// fold.cu — prologue kernels: fused embedding-row gather + relation folding.
//
// Every in-scope scorer's 1-vs-N form factors as  score(i, j) = pair(Q_i, cand_j[cols])  where Q_i
// depends only on the per-row operands (subject-or-object row, relation row).  This kernel
// gathers those rows by index straight from the embedding tables (LookupEmbedder.embed,
// lookup_embedder.py:96-97) and builds Q without materialising the gathered rows or the
// reference's concatenated operands (complex.py:26-32).  Folds:
//   ComplEx  sp_: [s_re*p_re - s_im*p_im | s_im*p_re + s_re*p_im]   _po: [p_re*o_re + p_im*o_im | p_re*o_im - p_im*o_re]
//   DistMult a*p
//   SimplE   sp_: 1/2 [s_t*p_b | s_h*p_f]                           _po: 1/2 [o_t*p_f | o_h*p_b]
//   CP       sp_: s[:h]*p  (vs cand[:, h:])                          _po: o[h:]*p (vs cand[:, :h])
//   RESCAL   sp_: s^T M_p                                            _po: M_p o
//   TransE   sp_: s + p                                              _po: o - p
//   RotatE   sp_: s * e^{i theta}                                    _po: conj(e^{i theta}) * o
#include "fold.cuh"

namespace b200kge {

__device__ __forceinline__ void store_q(float* Q, float* Qhi, float* Qlo, int64_t off, float v) {
  if (Q) Q[off] = v;
  if (Qhi) {
    // tf32 hi/lo split: hi = truncation to tf32 (what the tensor core does to a raw fp32 operand:
    // it ignores the low 13 mantissa bits — verified on B200, profiles/r1_notes.md), lo = the exact
    // remainder rounded-to-nearest to tf32 (halves the dominant error term vs letting the
    // hardware truncate it).
    float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    uint32_t lo_bits;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo_bits) : "f"(v - hi));
    float lo = __uint_as_float(lo_bits);
    Qhi[off] = hi;
    Qlo[off] = lo;
  }
}

template <int MODEL>
__global__ void __launch_bounds__(128)
fold_kernel(int combine, Rows qa, Rows pr, int64_t row0, float* __restrict__ Q, int64_t ldq,
            float* __restrict__ Qhi, float* __restrict__ Qlo, int K) {
  const int64_t i = blockIdx.x;
  const float* __restrict__ a = qa.row(i);
  const float* __restrict__ p = pr.row(i);
  const int D = qa.dim;
  const int h = D >> 1;
  const int64_t obase = (row0 + i) * ldq;
  const bool sp = (combine == B200KGE_SP_);

  if constexpr (MODEL == B200KGE_RESCAL) {
    extern __shared__ float sh[];
    for (int k = threadIdx.x; k < D; k += blockDim.x) sh[k] = a[k];
    __syncthreads();
    fold_rescal_block(sp, sh, p, D, [&](int k, float v) { store_q(Q, Qhi, Qlo, obase + k, v); });
  } else {
    for (int k = threadIdx.x; k < K; k += blockDim.x)
      store_q(Q, Qhi, Qlo, obase + k, fold_element<MODEL>(sp, a, p, k, h));
  }
  // zero the padding columns [K, ldq) so padded K-chunks contribute nothing
  for (int64_t k = K + threadIdx.x; k < ldq; k += blockDim.x) store_q(Q, Qhi, Qlo, obase + k, 0.f);
}

int launch_fold_queries(int model, int combine, const Rows& q, const Rows& p, int64_t n,
                        int64_t row0, float* Q, int64_t ldq, float* Qhi, float* Qlo,
                        cudaStream_t st) {
  if (n == 0) return 0;
  const int D = q.dim;
  int K = D;
  if (model == B200KGE_CP) K = D / 2;
  dim3 grid((unsigned)n), block(128);
  switch (model) {
    case B200KGE_COMPLEX:
      fold_kernel<B200KGE_COMPLEX><<<grid, block, 0, st>>>(combine, q, p, row0, Q, ldq, Qhi, Qlo, K); break;
    case B200KGE_DISTMULT:
      fold_kernel<B200KGE_DISTMULT><<<grid, block, 0, st>>>(combine, q, p, row0, Q, ldq, Qhi, Qlo, K); break;
    case B200KGE_SIMPLE:
      fold_kernel<B200KGE_SIMPLE><<<grid, block, 0, st>>>(combine, q, p, row0, Q, ldq, Qhi, Qlo, K); break;
    case B200KGE_CP:
      fold_kernel<B200KGE_CP><<<grid, block, 0, st>>>(combine, q, p, row0, Q, ldq, Qhi, Qlo, K); break;
    case B200KGE_RESCAL:
      fold_kernel<B200KGE_RESCAL><<<grid, block, D * sizeof(float), st>>>(combine, q, p, row0, Q, ldq, Qhi, Qlo, K); break;
    case B200KGE_TRANSE:
      fold_kernel<B200KGE_TRANSE><<<grid, block, 0, st>>>(combine, q, p, row0, Q, ldq, Qhi, Qlo, K); break;
    case B200KGE_ROTATE:
      fold_kernel<B200KGE_ROTATE><<<grid, block, 0, st>>>(combine, q, p, row0, Q, ldq, Qhi, Qlo, K); break;
    default:
      set_error("unknown model %d", model);
      return B200KGE_ERR_INVALID;
  }
  B2K_LAUNCH_CHECK("fold_kernel");
  return 0;
}

// Gather candidate rows (index subset) into a dense [m, ldd] block holding only the K columns the
// pair op reads; used by the tensor-core path, whose TMA loads need a regular 2-D table.
__global__ void __launch_bounds__(256)
gather_rows_kernel(Rows src, int col_off, int K, float* __restrict__ dst, int64_t ldd) {
  const int64_t r = blockIdx.x;
  const float* __restrict__ s = src.row(r) + col_off;
  for (int k = threadIdx.x; k < ldd; k += blockDim.x) dst[r * ldd + k] = (k < K) ? s[k] : 0.f;
}

int launch_gather_rows(const Rows& src, int col_off, int K, float* dst, int64_t ldd,
                       cudaStream_t st) {
  if (src.rows == 0) return 0;
  gather_rows_kernel<<<(unsigned)src.rows, 256, 0, st>>>(src, col_off, K, dst, ldd);
  B2K_LAUNCH_CHECK("gather_rows_kernel");
  return 0;
}

}  // namespace b200kge
