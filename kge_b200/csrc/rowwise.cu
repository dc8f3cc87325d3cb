// rowwise.cu — row-wise triple scoring (combine="spo") and fused negative-sample scoring.
//
//  * spo_kernel: one warp per triple; gathers the three rows by index and reduces the model's
//    trilinear form / distance in registers.  Replaces KgeModel.score_spo (kge_model.py:663-680) +
//    score_emb(combine="spo") of complex.py:34-35, distmult.py:16-17, simple.py:21-23,
//    cp.py:21-22, rescal.py:27-35, transe.py:17-18, rotate.py:30-41.
//  * ns_kernel: BatchNegativeSample.score (sampler.py:263-344) for the S and O slots: the CTA folds
//    (other entity, relation) of its positive triple into q once, then each warp gathers sampled
//    rows and reduces pair(q, row) — the gather of the sampled indexes is fused with the
//    per-negative dot/distance; nothing of size [n*K, D] is materialised (the reference's `triple`
//    implementation gathers 3 x [n*K, D]).  The P slot goes through spo_kernel with row divisors.
#include "fold.cuh"

namespace b200kge {

namespace {

struct RowsDiv {
  Rows r;
  int64_t div;  // logical row t reads operand row t / div
  __device__ __forceinline__ const float* row(int64_t t) const { return r.row(div > 1 ? t / div : t); }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

template <int MODEL>
__global__ void __launch_bounds__(256)
spo_kernel(RowsDiv S, RowsDiv Pr, RowsDiv O, int64_t n, float l_norm, float* __restrict__ out,
           int64_t out_ld, int64_t out_div, int64_t col0) {
  const int lane = threadIdx.x & 31;
  const int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= n) return;
  const float* __restrict__ s = S.row(t);
  const float* __restrict__ p = Pr.row(t);
  const float* __restrict__ o = O.row(t);
  const int D = S.r.dim, h = D >> 1;
  float acc = 0.f;
  if constexpr (MODEL == B200KGE_COMPLEX) {
    for (int k = lane; k < h; k += 32) {
      const float s_re = s[k], s_im = s[k + h], p_re = p[k], p_im = p[k + h], o_re = o[k], o_im = o[k + h];
      acc += s_re * p_re * o_re + s_im * p_re * o_im + s_re * p_im * o_im - s_im * p_im * o_re;
    }
  } else if constexpr (MODEL == B200KGE_DISTMULT) {
    for (int k = lane; k < D; k += 32) acc = fmaf(s[k] * p[k], o[k], acc);
  } else if constexpr (MODEL == B200KGE_SIMPLE) {
    for (int k = lane; k < h; k += 32)
      acc += 0.5f * (s[k] * p[k] * o[k + h] + s[k + h] * p[k + h] * o[k]);
  } else if constexpr (MODEL == B200KGE_CP) {
    for (int k = lane; k < h; k += 32) acc = fmaf(s[k] * p[k], o[k + h], acc);
  } else if constexpr (MODEL == B200KGE_RESCAL) {
    const int64_t dd = (int64_t)D * D;
    for (int64_t e = lane; e < dd; e += 32) {
      const int r = (int)(e / D), c = (int)(e - (int64_t)r * D);
      acc = fmaf(s[r] * p[e], o[c], acc);
    }
  } else if constexpr (MODEL == B200KGE_TRANSE) {
    for (int k = lane; k < D; k += 32) {
      const float d = fabsf(((s[k] + p[k]) - o[k]) + 1e-6f);  // F.pairwise_distance eps, transe.py:18
      if (l_norm == 1.0f) acc += d;
      else if (l_norm == 2.0f) acc = fmaf(d, d, acc);
      else acc += __powf(d, l_norm);
    }
  } else {  // ROTATE
    for (int k = lane; k < h; k += 32) {
      float sn, c;
      sincosf(p[k], &sn, &c);
      const float q_re = s[k] * c - s[k + h] * sn, q_im = s[k] * sn + s[k + h] * c;
      const float d_re = q_re - o[k], d_im = q_im - o[k + h];
      const float m2 = fmaf(d_im, d_im, d_re * d_re);
      if (l_norm == 1.0f) acc += sqrtf(m2);
      else acc += __powf(m2, 0.5f * l_norm);
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    if constexpr (MODEL == B200KGE_TRANSE || MODEL == B200KGE_ROTATE) {
      if (l_norm == 1.0f) acc = -acc;
      else if (l_norm == 2.0f) acc = -sqrtf(acc);
      else acc = -powf(acc, 1.0f / l_norm);
    }
    // logical triple t lands at out[(t / out_div) * out_ld + col0 + t % out_div]
    int64_t r = t, c = 0;
    if (out_div > 1) { r = t / out_div; c = t - r * out_div; }
    out[r * out_ld + col0 + c] = acc;
  }
}

int launch_spo_div(int model, float l_norm, const RowsDiv& s, const RowsDiv& p, const RowsDiv& o,
                   int64_t n, float* out, int64_t out_ld, int64_t out_div, int64_t col0,
                   cudaStream_t st) {
  if (n == 0) return 0;
  const int wpb = 8;
  const int64_t blocks = (n + wpb - 1) / wpb;
  if (blocks > 2147483647LL) { set_error("too many triples"); return B200KGE_ERR_UNSUPPORTED; }
  dim3 grid((unsigned)blocks), block(wpb * 32);
#define B2K_SPO(M) case M: spo_kernel<M><<<grid, block, 0, st>>>(s, p, o, n, l_norm, out, out_ld, out_div, col0); break;
  switch (model) {
    B2K_SPO(B200KGE_COMPLEX) B2K_SPO(B200KGE_DISTMULT) B2K_SPO(B200KGE_SIMPLE) B2K_SPO(B200KGE_CP)
    B2K_SPO(B200KGE_RESCAL) B2K_SPO(B200KGE_TRANSE) B2K_SPO(B200KGE_ROTATE)
    default: set_error("unknown model %d", model); return B200KGE_ERR_INVALID;
  }
#undef B2K_SPO
  B2K_LAUNCH_CHECK("spo_kernel");
  return 0;
}

// ---------------------------------------------------------------------------------------------
constexpr int NS_WARPS = 8, NS_PER_BLOCK = 256;

__device__ __forceinline__ float sqrt_approx(float x) {
  float y;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// pair(q, t) partial sum of one lane over a row, 16-byte loads: lane handles float4 groups lane, lane + 32, ...
// (complex pair ops: the re group at k4 pairs with the im group at k4 + hk/4)
template <int PAIR>
__device__ __forceinline__ float ns_row_partial(const float4* __restrict__ q4, const float4* __restrict__ t4, int n4, int lane) {
  float acc = 0.f;
  if constexpr (PAIR == PAIR_CMOD_L1) {
    const int h4 = n4 >> 1;
    for (int k = lane; k < h4; k += 32) {
      const float4 tr = __ldg(t4 + k), ti = __ldg(t4 + k + h4), qr = q4[k], qi = q4[k + h4];
      float dr, di;
      dr = qr.x - tr.x; di = qi.x - ti.x; acc += sqrt_approx(fmaf(di, di, dr * dr));
      dr = qr.y - tr.y; di = qi.y - ti.y; acc += sqrt_approx(fmaf(di, di, dr * dr));
      dr = qr.z - tr.z; di = qi.z - ti.z; acc += sqrt_approx(fmaf(di, di, dr * dr));
      dr = qr.w - tr.w; di = qi.w - ti.w; acc += sqrt_approx(fmaf(di, di, dr * dr));
    }
  } else {
    for (int k = lane; k < n4; k += 32) {
      const float4 t = __ldg(t4 + k), q = q4[k];
      if constexpr (PAIR == PAIR_DOT) {
        acc = fmaf(q.x, t.x, acc); acc = fmaf(q.y, t.y, acc); acc = fmaf(q.z, t.z, acc); acc = fmaf(q.w, t.w, acc);
      } else if constexpr (PAIR == PAIR_L1) {
        acc += fabsf(q.x - t.x); acc += fabsf(q.y - t.y); acc += fabsf(q.z - t.z); acc += fabsf(q.w - t.w);
      } else {  // PAIR_L2
        float d;
        d = q.x - t.x; acc = fmaf(d, d, acc); d = q.y - t.y; acc = fmaf(d, d, acc);
        d = q.z - t.z; acc = fmaf(d, d, acc); d = q.w - t.w; acc = fmaf(d, d, acc);
      }
    }
  }
  return acc;
}

// the warp's rows kk, kk + NS_WARPS, ... of this CTA's range, TWO at a time (both rows' loads are in flight together)
template <int PAIR>
__device__ __forceinline__ void ns_rows_vec(const float* q, const Rows& table, int col_off, int K, const int64_t* __restrict__ neg_row,
                                            int64_t k0, int64_t kend, int warp, int lane, float* __restrict__ out_row) {
  const float4* q4 = reinterpret_cast<const float4*>(q);
  const int n4 = K >> 2;
  for (int64_t kk = k0 + warp; kk < kend; kk += 2 * NS_WARPS) {
    const int64_t kb = kk + NS_WARPS;
    const bool two = kb < kend;
    const int64_t e0 = __ldg(neg_row + kk), e1 = two ? __ldg(neg_row + kb) : e0;
    const float4* t0 = reinterpret_cast<const float4*>(table.base + e0 * table.ld + col_off);
    const float4* t1 = reinterpret_cast<const float4*>(table.base + e1 * table.ld + col_off);
    float a0 = ns_row_partial<PAIR>(q4, t0, n4, lane);
    float a1 = two ? ns_row_partial<PAIR>(q4, t1, n4, lane) : 0.f;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, off);
      a1 += __shfl_xor_sync(0xffffffffu, a1, off);
    }
    if (lane == 0) {
      if constexpr (PAIR == PAIR_L1 || PAIR == PAIR_CMOD_L1) { a0 = -a0; a1 = -a1; }
      else if constexpr (PAIR == PAIR_L2) { a0 = -sqrtf(a0); a1 = -sqrtf(a1); }
      out_row[kk] = a0;
      if (two) out_row[kb] = a1;
    }
  }
}

template <int MODEL>
__global__ void __launch_bounds__(NS_WARPS * 32)
ns_kernel(Rows A, Rows Pr, Rows table, int sp, const int64_t* __restrict__ neg, int64_t Kneg,
          Folded f, float l_norm, float* __restrict__ out, int64_t ldo, int col0, int vec_ok) {
  extern __shared__ __align__(16) float sh[];  // q[K] (+ entity row for RESCAL)
  const int64_t i = blockIdx.x;
  const int D = A.dim, h = D >> 1, K = f.K;
  const float* __restrict__ a = A.row(i);
  const float* __restrict__ p = Pr.row(i);
  float* q = sh;
  if constexpr (MODEL == B200KGE_RESCAL) {
    float* sa = sh + ((K + 3) & ~3);
    for (int k = threadIdx.x; k < D; k += blockDim.x) sa[k] = a[k];
    __syncthreads();
    fold_rescal_block(sp != 0, sa, p, D, [&](int k, float v) { q[k] = v; });
  } else {
    for (int k = threadIdx.x; k < K; k += blockDim.x) q[k] = fold_element<MODEL>(sp != 0, a, p, k, h);
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t k0 = (int64_t)blockIdx.y * NS_PER_BLOCK;
  const int64_t kend = (k0 + NS_PER_BLOCK < Kneg) ? k0 + NS_PER_BLOCK : Kneg;
  const int64_t* __restrict__ neg_row = neg + i * Kneg;
  float* __restrict__ out_row = out + i * ldo + col0;
  if (vec_ok) {
    // 16-byte loads, two rows in flight per warp (the scalar form below gathered at 3.4 TB/s from an L2-resident table)
    switch (f.pair_op) {
      case PAIR_DOT:     ns_rows_vec<PAIR_DOT>(q, table, f.col_off, K, neg_row, k0, kend, warp, lane, out_row); return;
      case PAIR_L1:      ns_rows_vec<PAIR_L1>(q, table, f.col_off, K, neg_row, k0, kend, warp, lane, out_row); return;
      case PAIR_L2:      ns_rows_vec<PAIR_L2>(q, table, f.col_off, K, neg_row, k0, kend, warp, lane, out_row); return;
      case PAIR_CMOD_L1: ns_rows_vec<PAIR_CMOD_L1>(q, table, f.col_off, K, neg_row, k0, kend, warp, lane, out_row); return;
      default: break;
    }
  }
  const int hk = K >> 1;
  for (int64_t kk = k0 + warp; kk < kend; kk += NS_WARPS) {
    const int64_t e = neg_row[kk];
    const float* __restrict__ t = table.base + e * table.ld + f.col_off;
    float acc = 0.f;
    if (f.pair_op == PAIR_DOT) {
      for (int k = lane; k < K; k += 32) acc = fmaf(q[k], t[k], acc);
    } else if (f.pair_op == PAIR_L1) {
      for (int k = lane; k < K; k += 32) acc += fabsf(q[k] - t[k]);
    } else if (f.pair_op == PAIR_L2) {
      for (int k = lane; k < K; k += 32) { const float d = q[k] - t[k]; acc = fmaf(d, d, acc); }
    } else if (f.pair_op == PAIR_LP) {
      for (int k = lane; k < K; k += 32) acc += __powf(fabsf(q[k] - t[k]), l_norm);
    } else {
      for (int k = lane; k < hk; k += 32) {
        const float d_re = q[k] - t[k], d_im = q[k + hk] - t[k + hk];
        const float m2 = fmaf(d_im, d_im, d_re * d_re);
        acc += (f.pair_op == PAIR_CMOD_L1) ? sqrtf(m2) : __powf(m2, 0.5f * l_norm);
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      if (f.pair_op == PAIR_L1 || f.pair_op == PAIR_CMOD_L1) acc = -acc;
      else if (f.pair_op == PAIR_L2) acc = -sqrtf(acc);
      else if (f.pair_op != PAIR_DOT) acc = -powf(acc, 1.0f / l_norm);
      out_row[kk] = acc;
    }
  }
}

}  // namespace

int launch_spo(int model, float l_norm, const Rows& s, const Rows& p, const Rows& o, int64_t n,
               float* out, int64_t out_stride, cudaStream_t st) {
  RowsDiv S{s, 1}, P{p, 1}, O{o, 1};
  return launch_spo_div(model, l_norm, S, P, O, n, out, out_stride, 1, 0, st);
}

int launch_ns(int model, float l_norm, const Rows& s, const Rows& p, const Rows& o,
              const Rows& table, int slot, const int64_t* neg, int64_t n, int64_t K, float* out,
              int64_t ldo, int col0, cudaStream_t st) {
  if (n == 0 || K == 0) return 0;
  if (slot == 1) {
    // P slot: score(s_i, neg[i,k], o_i) for every (i,k) through the row-wise kernel: logical triple
    // t = i*K + k reads s/o row t/K and relation row neg[t]  (num_samples.p defaults to 0, so this
    // path is rare; config-default.yaml:346-349).
    Rows pneg = table; pneg.idx = neg; pneg.rows = n * K;
    RowsDiv S{s, K}, P{pneg, 1}, O{o, K};
    return launch_spo_div(model, l_norm, S, P, O, n * K, out, ldo, K, col0, st);
  }
  const int sp = (slot == 2) ? 1 : 0;  // O slot: fold (s,p) and score against sampled objects
  const Rows& a = sp ? s : o;
  Folded f = folded_problem(model, sp ? B200KGE_SP_ : B200KGE__PO, a.dim, l_norm);
  size_t smem = (size_t)((f.K + 3) & ~3) * sizeof(float) + (model == B200KGE_RESCAL ? (size_t)a.dim * sizeof(float) : 0);
  // 16-byte loads: rows of the sampled table start 16-byte aligned and the reduction splits into float4 groups
  const bool cm = (f.pair_op == PAIR_CMOD_L1 || f.pair_op == PAIR_CMOD_LP);
  const int vec_ok = (table.ld % 4 == 0 && f.col_off % 4 == 0 && f.K % (cm ? 8 : 4) == 0 &&
                      (reinterpret_cast<uintptr_t>(table.base) & 15) == 0) ? 1 : 0;
  const int64_t by = (K + NS_PER_BLOCK - 1) / NS_PER_BLOCK;
  if (by > 65535) { set_error("too many negatives per row (%lld)", (long long)K); return B200KGE_ERR_UNSUPPORTED; }
  dim3 grid((unsigned)n, (unsigned)by), block(NS_WARPS * 32);
#define B2K_NS(M) case M: ns_kernel<M><<<grid, block, smem, st>>>(a, p, table, sp, neg, K, f, l_norm, out, ldo, col0, vec_ok); break;
  switch (model) {
    B2K_NS(B200KGE_COMPLEX) B2K_NS(B200KGE_DISTMULT) B2K_NS(B200KGE_SIMPLE) B2K_NS(B200KGE_CP)
    B2K_NS(B200KGE_RESCAL) B2K_NS(B200KGE_TRANSE) B2K_NS(B200KGE_ROTATE)
    default: set_error("unknown model %d", model); return B200KGE_ERR_INVALID;
  }
#undef B2K_NS
  B2K_LAUNCH_CHECK("ns_kernel");
  return 0;
}

// ---------------------------------------------------------------------------------------------
// On-device uniform negative sampling: KgeUniformSampler._sample (sampler.py:588-596) is
// torch.randint(vocabulary_size, (n, num_samples)) on the CPU inside DataLoader workers, followed by a host->device copy
// of n*K int64 per slot; here the ids are drawn where they are consumed.  Counter-based Philox4x32-10 (Salmon et al.,
// SC'11): element e of the call uses counter (e / 2, offset) under key (seed), so results depend on (seed, offset, e)
// only — reproducible, no generator state.  Each 64-bit draw r maps to floor(r * vocab / 2^64) (bias <= vocab / 2^64).
namespace {

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

__global__ void __launch_bounds__(256)
sample_uniform_kernel(uint64_t seed, uint64_t offset, uint64_t vocab, int64_t total, int64_t* __restrict__ out) {
  const int64_t pair = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;      // two ids per Philox block
  if (2 * pair >= total) return;
  uint32_t c[4] = {(uint32_t)pair, (uint32_t)((uint64_t)pair >> 32), (uint32_t)offset, (uint32_t)(offset >> 32)};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  const uint64_t r0 = ((uint64_t)c[1] << 32) | c[0], r1 = ((uint64_t)c[3] << 32) | c[2];
  out[2 * pair] = (int64_t)__umul64hi(r0, vocab);
  if (2 * pair + 1 < total) out[2 * pair + 1] = (int64_t)__umul64hi(r1, vocab);
}

}  // namespace

int launch_sample_uniform(uint64_t seed, uint64_t offset, int64_t vocab, int64_t total, int64_t* out, cudaStream_t st) {
  if (total <= 0) return 0;
  const int64_t pairs = (total + 1) / 2;
  sample_uniform_kernel<<<(unsigned)((pairs + 255) / 256), 256, 0, st>>>(seed, offset, (uint64_t)vocab, total, out);
  B2K_LAUNCH_CHECK("sample_uniform_kernel");
  return 0;
}

}  // namespace b200kge
