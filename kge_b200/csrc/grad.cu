// grad.cu — EXPERIMENTAL ("next"): device pieces of the backward of the fused 1vsAll step for
// the dot family.  Not on any default path and not validated on hardware yet; the math is pinned on the CPU
// (oracle/kge_fold.py against autograd and against gradients of the live reference,
// tests/test_fold_algebra.py), the kernels below transcribe it.
//
//   z  = Q T^T                       recomputed with the validated scorer (plain-store epilogue)
//   G  = sigmoid(z + off) - y        grad_planes_kernel: written ONCE as fp16 hi/lo planes in both layouts,
//                                    G [nq, Ep] and G^T [E, Np] (scale 2^14; 1/n rides in the row scale)
//   dT = G^T Q   [E, K]              pre-split fp16 GEMM (pairwise_tc3.cu, store epilogue) on G^T and Q^T planes
//   dQ = G  T    [nq, K]             same on G and T^T planes
//   (da, dp) = unfold(a, p, dQ)      unfold_kernel: row-wise vector-Jacobian products of the relation fold,
//                                    atomically added into the entity / relation gradient tables
// This first version trades HBM traffic for simplicity (scores and both G layouts are materialised, operands
// are transposed through HBM); fusing the G pass into the scorer's epilogue and a split-K dQ are the
// follow-ups once it is parity-green.
#include <cuda_fp16.h>
#include "fold.cuh"
#include "tc_common.cuh"

namespace b200kge {

namespace {

// ---------------------------------------------------------------------------------------------
// dst[c, r] = src[r, c]   (fp32, 32x32 tiles through shared memory); dst columns [R, ldd) are zeroed
__global__ void __launch_bounds__(256)
transpose_kernel(const float* __restrict__ src, int64_t lds, int64_t R, int64_t C, float* __restrict__ dst,
                 int64_t ldd) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8
  const int64_t r0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t r = r0 + ty + 8 * k, c = c0 + tx;
    tile[ty + 8 * k][tx] = (r < R && c < C) ? __ldg(src + r * lds + c) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t c = c0 + ty + 8 * k, r = r0 + tx;
    if (c < C && r < ldd) dst[c * ldd + r] = tile[tx][ty + 8 * k];   // r >= R carries the zeros loaded above
  }
}

__device__ __forceinline__ void split_store(__half* __restrict__ hi, __half* __restrict__ lo, int64_t pos, float g) {
  const float s = g * 16384.f;
  const __half h = __float2half_rn(s);
  hi[pos] = h;
  lo[pos] = __float2half_rn(s - __half2float(h));
}

// KL needs the row's log-sum-exp and label mass first: row_stat[2i] = logsumexp_j z_ij, row_stat[2i+1] = sum_j y_ij
// (loss.py:198-213).  One block per row.
__global__ void __launch_bounds__(256)
row_lse_kernel(const float* __restrict__ z, int64_t ldz, int64_t E, const int64_t* __restrict__ label_idx,
               const float* __restrict__ label_dense, int64_t ldl, float* __restrict__ row_stat) {
  __shared__ float red[3][8];
  const int64_t i = blockIdx.x;
  const float* __restrict__ zr = z + i * ldz;
  float mx = -INFINITY, ys = 0.f;
  for (int64_t e = threadIdx.x; e < E; e += blockDim.x) {
    mx = fmaxf(mx, zr[e]);
    if (label_dense) ys += label_dense[i * ldl + e];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    ys += __shfl_xor_sync(0xffffffffu, ys, o);
  }
  if (lane == 0) { red[0][warp] = mx; red[1][warp] = ys; }
  __syncthreads();
  mx = red[0][0]; ys = red[1][0];
#pragma unroll
  for (int w = 1; w < 8; ++w) { mx = fmaxf(mx, red[0][w]); ys += red[1][w]; }
  float se = 0.f;
  for (int64_t e = threadIdx.x; e < E; e += blockDim.x) se += expf(zr[e] - mx);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
  if (lane == 0) red[2][warp] = se;
  __syncthreads();
  if (threadIdx.x == 0) {
    se = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) se += red[2][w];
    row_stat[2 * i] = mx + logf(se);
    row_stat[2 * i + 1] = label_idx ? ((label_idx[i] >= 0 && label_idx[i] < E) ? 1.f : 0.f) : ys;
  }
}

// G = dL/dz * n as fp16 hi/lo planes, row-major [nq, Ep] and transposed [E, Np]; pads zeroed.
//   BCE (row_stat == nullptr): G = sigmoid(z + off) - y          KL: G = w * exp(z - lse) - y / yc,  yc = max(sum y, 1e-12), w = sum y / yc
// grid = (Ep/32, Np/32), block = 32 x 8.
__global__ void __launch_bounds__(256)
grad_planes_kernel(const float* __restrict__ z, int64_t ldz, int64_t nq, int64_t E,
                   const int64_t* __restrict__ label_idx, const float* __restrict__ label_dense, int64_t ldl,
                   const float* __restrict__ row_stat,
                   float offset, float inv_n, __half* __restrict__ g_hi, __half* __restrict__ g_lo, int64_t Ep,
                   __half* __restrict__ gt_hi, __half* __restrict__ gt_lo, int64_t Np,
                   float* __restrict__ g_scale, float* __restrict__ gt_scale) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t e0 = (int64_t)blockIdx.x * 32, i0 = (int64_t)blockIdx.y * 32;
  const float inv = inv_n * (1.0f / 16384.f);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t i = i0 + ty + 8 * k, e = e0 + tx;
    float g = 0.f;
    if (i < nq && e < E) {
      const float x = __ldg(z + i * ldz + e) + offset;
      const float y = label_idx ? ((label_idx[i] == e) ? 1.f : 0.f) : __ldg(label_dense + i * ldl + e);
      if (row_stat) {   // labels are normalised by their row sum first (loss.py:209-213)
        const float ys = row_stat[2 * i + 1], yc = fmaxf(ys, 1e-12f);
        g = (ys / yc) * expf(x - row_stat[2 * i]) - y / yc;
      }
      else g = 1.0f / (1.0f + expf(-x)) - y;
    }
    tile[ty + 8 * k][tx] = g;
    if (i < nq) {
      split_store(g_hi, g_lo, i * Ep + e, g);                 // e < Ep by construction of the grid
      if (blockIdx.x == 0 && tx == 0) g_scale[i] = inv;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t e = e0 + ty + 8 * k, i = i0 + tx;
    if (e < E) {
      split_store(gt_hi, gt_lo, e * Np + i, tile[tx][ty + 8 * k]);   // i < Np by construction of the grid
      if (blockIdx.y == 0 && tx == 0) gt_scale[e] = inv;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Row-wise unfold: block b handles query row b.  dir < 0: rows [0,n) are sp_ (a = subject), rows [n,2n) are
// _po (a = object) — the stacked layout of prep_1vsall_kernel; dir = 0 / 1: all rows sp_ / _po.
template <int MODEL>
__global__ void __launch_bounds__(128)
unfold_kernel(Rows ent, Rows rel, const int64_t* __restrict__ tri, int64_t n, int dir,
              const float* __restrict__ dQ, int64_t ldq, float* __restrict__ d_ent, int64_t lde,
              float* __restrict__ d_rel, int64_t ldr) {
  const int64_t b = blockIdx.x;
  const bool sp = dir < 0 ? (b < n) : (dir == 0);
  const int64_t i = (dir < 0 && b >= n) ? b - n : b;
  const int64_t si = tri[3 * i], pi = tri[3 * i + 1], oi = tri[3 * i + 2];
  const int64_t ai = sp ? si : oi;
  const float* __restrict__ a = ent.base + ai * ent.ld;
  const float* __restrict__ p = rel.base + pi * rel.ld;
  const float* __restrict__ g = dQ + b * ldq;
  float* __restrict__ da = d_ent + ai * lde;
  float* __restrict__ dp = d_rel + pi * ldr;
  const int D = ent.dim, h = D >> 1;

  if constexpr (MODEL == B200KGE_RESCAL) {
    // sp_: q = a^T M  => da_i = sum_j M[i,j] g_j,  dM[i,j] = a_i g_j
    // _po: q = M a    => da_j = sum_i g_i M[i,j],  dM[i,j] = g_i a_j
    extern __shared__ float sh[];
    float* sa = sh;
    float* sg = sh + D;
    for (int k = threadIdx.x; k < D; k += blockDim.x) { sa[k] = a[k]; sg[k] = g[k]; }
    __syncthreads();
    if (sp) {
      const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
      for (int r = warp; r < D; r += nw) {
        float acc = 0.f;
        for (int j = lane; j < D; j += 32) acc = fmaf(p[(int64_t)r * D + j], sg[j], acc);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (lane == 0) atomicAdd(da + r, acc);
      }
    } else {
      for (int j = threadIdx.x; j < D; j += blockDim.x) {
        float acc = 0.f;
        for (int r = 0; r < D; ++r) acc = fmaf(sg[r], p[(int64_t)r * D + j], acc);
        atomicAdd(da + j, acc);
      }
    }
    for (int idx = threadIdx.x; idx < D * D; idx += blockDim.x) {
      const int r = idx / D, j = idx - r * D;
      atomicAdd(dp + idx, sp ? sa[r] * sg[j] : sg[r] * sa[j]);
    }
  } else if constexpr (MODEL == B200KGE_COMPLEX) {
    for (int k = threadIdx.x; k < h; k += blockDim.x) {
      const float a_re = a[k], a_im = a[k + h], p_re = p[k], p_im = p[k + h], g_re = g[k], g_im = g[k + h];
      if (sp) {   // Q_re = a_re p_re - a_im p_im ; Q_im = a_im p_re + a_re p_im
        atomicAdd(da + k, g_re * p_re + g_im * p_im);
        atomicAdd(da + k + h, -g_re * p_im + g_im * p_re);
        atomicAdd(dp + k, g_re * a_re + g_im * a_im);
        atomicAdd(dp + k + h, -g_re * a_im + g_im * a_re);
      } else {    // Q_re = p_re a_re + p_im a_im ; Q_im = p_re a_im - p_im a_re
        atomicAdd(da + k, g_re * p_re - g_im * p_im);
        atomicAdd(da + k + h, g_re * p_im + g_im * p_re);
        atomicAdd(dp + k, g_re * a_re + g_im * a_im);
        atomicAdd(dp + k + h, g_re * a_im - g_im * a_re);
      }
    }
  } else if constexpr (MODEL == B200KGE_DISTMULT) {
    for (int k = threadIdx.x; k < D; k += blockDim.x) {
      atomicAdd(da + k, g[k] * p[k]);
      atomicAdd(dp + k, g[k] * a[k]);
    }
  } else if constexpr (MODEL == B200KGE_SIMPLE) {
    for (int k = threadIdx.x; k < h; k += blockDim.x) {
      const float a_h = a[k], a_t = a[k + h], p_f = p[k], p_b = p[k + h];
      const float g0 = 0.5f * g[k], g1 = 0.5f * g[k + h];
      if (sp) {   // Q = 1/2 [a_t p_b | a_h p_f]
        atomicAdd(da + k, g1 * p_f);
        atomicAdd(da + k + h, g0 * p_b);
        atomicAdd(dp + k, g1 * a_h);
        atomicAdd(dp + k + h, g0 * a_t);
      } else {    // Q = 1/2 [a_t p_f | a_h p_b]
        atomicAdd(da + k, g1 * p_b);
        atomicAdd(da + k + h, g0 * p_f);
        atomicAdd(dp + k, g0 * a_t);
        atomicAdd(dp + k + h, g1 * a_h);
      }
    }
  } else {  // CP: sp_ Q = a[:h] p (vs cand[:, h:]);  _po Q = a[h:] p (vs cand[:, :h])
    const int ao = sp ? 0 : h;
    for (int k = threadIdx.x; k < h; k += blockDim.x) {
      atomicAdd(da + ao + k, g[k] * p[k]);
      atomicAdd(dp + k, g[k] * a[ao + k]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Lp / N3 penalty of embedding rows (lookup_embedder.py:123-177): sum over the selected rows of
// count_r * sum_k |x_rk|^p  (N3 in complex space: x -> sqrt(re^2 + im^2 + 1e-14), p = 3), times scale.
// One block per PEN_ROWS rows, per-block partial sums, last block (ticket) adds them in block order.
constexpr int PEN_ROWS = 8;

__device__ __forceinline__ float pow_p(float a, float p, int ip) {
  if (ip == 1) return a;
  if (ip == 2) return a * a;
  if (ip == 3) return a * a * a;
  return powf(a, p);
}

__global__ void __launch_bounds__(256)
penalty_kernel(Rows tab, const float* __restrict__ counts, float p, int complex_abs, float scale,
               float* __restrict__ partial, unsigned int* __restrict__ ticket, float* __restrict__ out) {
  __shared__ float red[8];
  __shared__ bool last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * PEN_ROWS + warp;
  const int ip = (p == 1.f) ? 1 : (p == 2.f) ? 2 : (p == 3.f) ? 3 : 0;
  float acc = 0.f;
  if (r < tab.rows) {
    const float* __restrict__ x = tab.row(r);
    if (complex_abs) {
      const int h = tab.dim >> 1;
      for (int k = lane; k < h; k += 32) {
        const float re = x[k], im = x[k + h];
        acc += pow_p(sqrtf(re * re + im * im + 1e-14f), p, ip);
      }
    } else {
      for (int k = lane; k < tab.dim; k += 32) acc += pow_p(fabsf(x[k]), p, ip);
    }
    if (counts) acc *= counts[r];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < PEN_ROWS; ++w) t += red[w];
    partial[blockIdx.x] = t;
    __threadfence();
    last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    float t = 0.f;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) t += __ldcg(partial + b);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    __syncthreads();
    if (lane == 0) red[warp] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) tot += red[w];
      *out = tot * scale;
      *ticket = 0u;
    }
  }
}

// rows scaled to unit Lp norm in place (F.normalize, eps = 1e-12; lookup_embedder.py:64-69). One warp per row.
__global__ void __launch_bounds__(256)
normalize_rows_kernel(float* __restrict__ w, int64_t ld, int64_t rows, int dim, float p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 8 + warp;
  if (r >= rows) return;
  float* __restrict__ x = w + r * ld;
  const int ip = (p == 1.f) ? 1 : (p == 2.f) ? 2 : (p == 3.f) ? 3 : 0;
  float acc = 0.f;
  for (int k = lane; k < dim; k += 32) acc += pow_p(fabsf(x[k]), p, ip);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  const float nrm = (ip == 1) ? acc : (ip == 2) ? sqrtf(acc) : powf(acc, 1.0f / p);
  const float inv = 1.0f / fmaxf(nrm, 1e-12f);
  for (int k = lane; k < dim; k += 32) x[k] *= inv;
}

}  // namespace

int launch_penalty(const Rows& tab, const float* counts, float p, int complex_abs, float scale, float* scratch,
                   size_t scratch_floats, float* out, cudaStream_t st) {
  const int64_t blocks = (tab.rows + PEN_ROWS - 1) / PEN_ROWS;
  if (blocks == 0) { return check_cuda(cudaMemsetAsync(out, 0, 4, st), "cudaMemsetAsync(penalty)"); }
  if ((size_t)blocks + 1 > scratch_floats || blocks >= (1ll << 31)) { set_error("workspace too small for the penalty partials"); return B200KGE_ERR_WORKSPACE; }
  unsigned int* ticket = reinterpret_cast<unsigned int*>(scratch + blocks);
  cudaError_t e = cudaMemsetAsync(ticket, 0, 4, st);
  if (e != cudaSuccess) return check_cuda(e, "cudaMemsetAsync(ticket)");
  penalty_kernel<<<(unsigned)blocks, 256, 0, st>>>(tab, counts, p, complex_abs, scale, scratch, ticket, out);
  B2K_LAUNCH_CHECK("penalty_kernel");
  return 0;
}

int launch_normalize_rows(float* w, int64_t ld, int64_t rows, int dim, float p, cudaStream_t st) {
  if (rows == 0 || !(p > 0.f)) return 0;
  normalize_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(w, ld, rows, dim, p);
  B2K_LAUNCH_CHECK("normalize_rows_kernel");
  return 0;
}

int launch_transpose(const float* src, int64_t lds, int64_t R, int64_t C, float* dst, int64_t ldd, cudaStream_t st) {
  if (R == 0 || C == 0) return 0;
  dim3 grid((unsigned)((ldd + 31) / 32), (unsigned)((C + 31) / 32));
  transpose_kernel<<<grid, 256, 0, st>>>(src, lds, R, C, dst, ldd);
  B2K_LAUNCH_CHECK("transpose_kernel");
  return 0;
}

int launch_grad_planes(const float* z, int64_t ldz, int64_t nq, int64_t E, const int64_t* label_idx,
                       const float* label_dense, int64_t ldl, float* row_stat /* [2*nq] scratch: KL; null: BCE */,
                       float offset, float inv_n, void* g_hi, void* g_lo,
                       int64_t Ep, void* gt_hi, void* gt_lo, int64_t Np, float* g_scale, float* gt_scale,
                       cudaStream_t st) {
  if (nq == 0 || E == 0) return 0;
  if (row_stat) {
    row_lse_kernel<<<(unsigned)nq, 256, 0, st>>>(z, ldz, E, label_idx, label_dense, ldl, row_stat);
    B2K_LAUNCH_CHECK("row_lse_kernel");
  }
  dim3 grid((unsigned)(Ep / 32), (unsigned)(Np / 32));
  grad_planes_kernel<<<grid, 256, 0, st>>>(z, ldz, nq, E, label_idx, label_dense, ldl, row_stat, offset, inv_n,
                                            (__half*)g_hi, (__half*)g_lo, Ep, (__half*)gt_hi, (__half*)gt_lo, Np,
                                            g_scale, gt_scale);
  B2K_LAUNCH_CHECK("grad_planes_kernel");
  return 0;
}

int launch_unfold(int model, const Rows& ent, const Rows& rel, const int64_t* triples, int64_t n, int dir,
                  const float* dQ, int64_t ldq, float* d_ent, int64_t lde, float* d_rel, int64_t ldr, cudaStream_t st) {
  if (n == 0) return 0;
  const int64_t nq = dir < 0 ? 2 * n : n;
  dim3 grid((unsigned)nq), block(128);
  const int D = ent.dim;
#define B2K_UNFOLD(M, SM) case M: unfold_kernel<M><<<grid, block, SM, st>>>(ent, rel, triples, n, dir, dQ, ldq, d_ent, lde, d_rel, ldr); break;
  switch (model) {
    B2K_UNFOLD(B200KGE_COMPLEX, 0) B2K_UNFOLD(B200KGE_DISTMULT, 0) B2K_UNFOLD(B200KGE_SIMPLE, 0)
    B2K_UNFOLD(B200KGE_CP, 0) B2K_UNFOLD(B200KGE_RESCAL, 2 * D * sizeof(float))
    default: set_error("the analytic backward covers the dot family only (model %d)", model); return B200KGE_ERR_UNSUPPORTED;
  }
#undef B2K_UNFOLD
  B2K_LAUNCH_CHECK("unfold_kernel");
  return 0;
}

}  // namespace b200kge
