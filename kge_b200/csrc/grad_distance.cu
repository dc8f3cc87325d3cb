// grad_distance.cu — SURVEY §8 f-1 for the distance family: backward of the 1-vs-N scores of TransE (L1, L2;
// transe.py:20-35) and RotatE (L1 of complex moduli; rotate.py:42-65) through the folded form score(i, j) = pair(Q_i, T_j).
//
// Every pair op here is a function of the difference d = q - t, so with s'(d) = d pair / d q:
//   dQ_i = sum_j G_ij s'(Q_i - T_j)            dT_j = sum_i G_ij s'(T_j - Q_i)        (s' is odd in d)
// i.e. ONE kernel, "row gradient": dA[r, :] = sum_c W[r, c] * s'(A_r - B_c), called with (A, B, W) = (Q, T, G) and
// with (T, Q, G^T).  No tensor cores (sign / normalise per element, like the forward: CUDA-core bound, 4 instructions
// per (r, c, k) and pass); a CTA owns RB rows of A and walks ALL columns, so dA is written once, without atomics.
//   L1      s = -sum_k |d_k|             s'_k = -sign(d_k)
//   L2      s = -||d||                   s'_k = -d_k / ||d|| = d_k / s        (s = Z[r, c], the stored score; 0 if s = 0)
//   CMOD L1 s = -sum_k |d_k| (complex)   s'_(re,im),k = -(d_re, d_im)_k / |d_k|   (0 if |d_k| = 0)
// G = n dL/dz comes from grad_dense_kernel (BCE: sigmoid(z + off) - y; KL: w softmax(z) - y / sum y), fp32, dense.
#include "common.cuh"

namespace b200kge {

namespace {

constexpr int RB = 16;       // rows of A per CTA
constexpr int KL = 16;       // k lanes per row
constexpr int KC = 64;       // reduction elements per chunk (CMOD: 32 complex elements = 32 re + 32 im)
constexpr int CT = 128;      // columns per shared-memory tile

template <int PAIR>
__global__ void __launch_bounds__(RB * KL)
pair_rowgrad_kernel(const float* __restrict__ A, int64_t lda, int64_t ra, const float* __restrict__ B, int64_t ldb,
                    int64_t rb, int K, const float* __restrict__ W, int64_t ldw, const float* __restrict__ Z,
                    int64_t ldz, float* __restrict__ dA, int64_t ldda) {
  __shared__ float Bs[CT][KC + 1];
  __shared__ float Ws[RB][CT + 1];
  const int tid = threadIdx.x, r = tid / KL, kl = tid % KL;
  const int64_t row = (int64_t)blockIdx.x * RB + r;
  const bool row_ok = row < ra;
  constexpr bool CMOD = (PAIR == PAIR_CMOD_L1);
  const int h = K >> 1;
  const int span = CMOD ? h : K;                 // elements walked by the chunk loop
  constexpr int EPC = CMOD ? KC / 2 : KC;        // elements per chunk
  constexpr int PER = EPC / KL;                  // elements per thread and chunk (4 | 2)
  for (int e0 = 0; e0 < span; e0 += EPC) {
    float a0[PER], a1[PER], acc0[PER], acc1[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int e = e0 + kl + KL * j;
      a0[j] = (row_ok && e < span) ? A[row * lda + e] : 0.f;
      a1[j] = (CMOD && row_ok && e < span) ? A[row * lda + h + e] : 0.f;
      acc0[j] = 0.f; acc1[j] = 0.f;
    }
    for (int64_t c0 = 0; c0 < rb; c0 += CT) {
      __syncthreads();
      // B tile: CT columns x this chunk's elements (CMOD: re at [0, EPC), im at [EPC, 2 EPC))
      for (int t = tid; t < CT * KC; t += RB * KL) {
        const int c = t / KC, x = t % KC;
        const int64_t col = c0 + c;
        int e; int64_t off;
        if (CMOD) { e = e0 + (x % EPC); off = (x < EPC) ? e : h + e; } else { e = e0 + x; off = e; }
        Bs[c][x] = (col < rb && e < span) ? __ldg(B + col * ldb + off) : 0.f;
      }
      for (int t = tid; t < RB * CT; t += RB * KL) {
        const int rr = t / CT, c = t % CT;
        const int64_t row2 = (int64_t)blockIdx.x * RB + rr, col = c0 + c;
        float w = (row2 < ra && col < rb) ? __ldg(W + row2 * ldw + col) : 0.f;
        if (PAIR == PAIR_L2 && w != 0.f) {
          const float z = __ldg(Z + row2 * ldz + col);       // z = -||d||
          w = (z != 0.f) ? w / z : 0.f;
        }
        Ws[rr][c] = w;
      }
      __syncthreads();
#pragma unroll 4
      for (int c = 0; c < CT; ++c) {
        const float w = Ws[r][c];
#pragma unroll
        for (int j = 0; j < PER; ++j) {
          if constexpr (PAIR == PAIR_L1) {
            const float d = a0[j] - Bs[c][kl + KL * j];
            acc0[j] -= w * ((d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f));
          } else if constexpr (PAIR == PAIR_L2) {
            acc0[j] = fmaf(w, a0[j] - Bs[c][kl + KL * j], acc0[j]);     // w already holds G / z
          } else {
            const float dre = a0[j] - Bs[c][kl + KL * j], dim = a1[j] - Bs[c][EPC + kl + KL * j];
            const float m2 = fmaf(dim, dim, dre * dre);
            const float inv = (m2 > 0.f) ? rsqrtf(m2) : 0.f;
            const float wi = w * inv;
            acc0[j] = fmaf(-wi, dre, acc0[j]);
            acc1[j] = fmaf(-wi, dim, acc1[j]);
          }
        }
      }
    }
    if (row_ok) {
#pragma unroll
      for (int j = 0; j < PER; ++j) {
        const int e = e0 + kl + KL * j;
        if (e < span) {
          dA[row * ldda + e] = acc0[j];
          if (CMOD) dA[row * ldda + h + e] = acc1[j];
        }
      }
    }
  }
}

// G[i, e] = n dL/dz_ie for one-hot labels (1vsAll): BCE sigmoid(z + off) - y | KL softmax(z) - y  (row_stat[2i] = lse_i),
// times inv_n.  grid = (ceil(E / 256), nq)
__global__ void __launch_bounds__(256)
grad_dense_kernel(const float* __restrict__ z, int64_t ldz, int64_t E, const int64_t* __restrict__ label_idx,
                  const float* __restrict__ row_stat, float offset, float inv_n, float* __restrict__ G, int64_t ldg) {
  const int64_t i = blockIdx.y, e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const float x = z[i * ldz + e] + offset;
  const float y = (label_idx[i] == e) ? 1.f : 0.f;
  float g;
  if (row_stat) {
    const float ys = row_stat[2 * i + 1], yc = fmaxf(ys, 1e-12f);
    g = (ys / yc) * expf(x - row_stat[2 * i]) - y / yc;
  } else {
    g = 1.0f / (1.0f + expf(-x)) - y;
  }
  G[i * ldg + e] = g * inv_n;
}

template <int PAIR>
int launch_rowgrad_t(const float* A, int64_t lda, int64_t ra, const float* B, int64_t ldb, int64_t rb, int K, const float* W,
                     int64_t ldw, const float* Z, int64_t ldz, float* dA, int64_t ldda, cudaStream_t st) {
  if (ra == 0 || rb == 0) return 0;
  pair_rowgrad_kernel<PAIR><<<(unsigned)((ra + RB - 1) / RB), RB * KL, 0, st>>>(A, lda, ra, B, ldb, rb, K, W, ldw, Z, ldz, dA, ldda);
  B2K_LAUNCH_CHECK("pair_rowgrad_kernel");
  return 0;
}

}  // namespace

int launch_pair_rowgrad(int pair_op, const float* A, int64_t lda, int64_t ra, const float* B, int64_t ldb, int64_t rb, int K,
                        const float* W, int64_t ldw, const float* Z, int64_t ldz, float* dA, int64_t ldda, cudaStream_t st) {
  switch (pair_op) {
    case PAIR_L1:      return launch_rowgrad_t<PAIR_L1>(A, lda, ra, B, ldb, rb, K, W, ldw, Z, ldz, dA, ldda, st);
    case PAIR_L2:      return launch_rowgrad_t<PAIR_L2>(A, lda, ra, B, ldb, rb, K, W, ldw, Z, ldz, dA, ldda, st);
    case PAIR_CMOD_L1: return launch_rowgrad_t<PAIR_CMOD_L1>(A, lda, ra, B, ldb, rb, K, W, ldw, Z, ldz, dA, ldda, st);
  }
  set_error("the distance-family backward covers L1, L2 (TransE) and the L1 of complex moduli (RotatE)");
  return B200KGE_ERR_UNSUPPORTED;
}

int launch_grad_dense(const float* z, int64_t ldz, int64_t nq, int64_t E, const int64_t* label_idx, const float* row_stat,
                      float offset, float inv_n, float* G, int64_t ldg, cudaStream_t st) {
  if (nq == 0 || E == 0) return 0;
  if (nq > 65535) { set_error("too many rows for one launch (%lld)", (long long)nq); return B200KGE_ERR_UNSUPPORTED; }
  dim3 grid((unsigned)((E + 255) / 256), (unsigned)nq);
  grad_dense_kernel<<<grid, 256, 0, st>>>(z, ldz, E, label_idx, row_stat, offset, inv_n, G, ldg);
  B2K_LAUNCH_CHECK("grad_dense_kernel");
  return 0;
}

}  // namespace b200kge
