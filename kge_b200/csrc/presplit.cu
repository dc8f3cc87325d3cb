// presplit.cu — operand split for the pre-split fp16 tensor-core kernels (pairwise_tc3.cu, pairwise_tc4.cu): the
// default path of the dot family (B200KGE_PREC_AUTO / F16X3).
//
// A fp32 value x of row r is represented as  x = inv_scale[r] * (hi + lo),  hi = fp16_rn(x * 2^s),
// lo = fp16_rn(x * 2^s - hi),  2^s chosen per row so that max|x * 2^s| lies in [2^13, 2^14): 22 significant
// bits, no fp16 overflow, lo in the normal range for every element within 2^-13 of the row maximum.  The
// three products hi*hi + hi*lo + lo*hi on the f16 tensor pipe (fp32 accumulate) then reproduce the fp32 GEMM
// of the reference (torch.mm in complex.py:37,39 etc.) to ~5e-7 of the score rms before accumulation
// round-off — the same accuracy class as the in-kernel tf32+bf16 split, at 6 instead of 8 MMA slots per 32
// reduction elements and without any shared-memory round trip in the main loop.
//
// One warp per row, two passes over the row (|max|, then convert; the second read hits L1).  HBM-bound:
// 4 B read + 4 B written per element.
#include <cuda_fp16.h>
#include "tc_common.cuh"

namespace b200kge {

namespace {

constexpr int PS_WARPS = 8;

__device__ __forceinline__ void presplit_row(const SplitSet& S, int64_t r, int lane) {
  if (r >= S.rows) {
    if (r < S.rows_pad && lane == 0) S.inv_scale[r] = 0.f;
    return;
  }
  const int64_t src_row = S.idx ? S.idx[r] : r;
  const float* __restrict__ x = S.src + src_row * S.ld + S.col_off;
  const int K = S.K, Kp = S.Kp;
  const bool vec = (K % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  float amax = 0.f;
  bool bad = false;
  if (vec) {
    for (int k = lane * 4; k < K; k += 128) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + k));
      amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
      bad |= !(isfinite(v.x) && isfinite(v.y) && isfinite(v.z) && isfinite(v.w));
    }
  } else {
    for (int k = lane; k < K; k += 32) {
      const float v = __ldg(x + k);
      amax = fmaxf(amax, fabsf(v));
      bad |= !isfinite(v);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  bad = __any_sync(0xffffffffu, bad);
  int e = 13;                                    // => scale 1 for all-zero or non-finite rows
  if (amax > 0.f && !bad) {
    e = ilogbf(amax);
    e = e < -60 ? -60 : (e > 60 ? 60 : e);
  }
  const float mul = scalbnf(1.f, 13 - e), inv = scalbnf(1.f, e - 13);
  __half* __restrict__ hi = reinterpret_cast<__half*>(S.hi) + r * Kp;
  __half* __restrict__ lo = reinterpret_cast<__half*>(S.lo) + r * Kp;
  if (vec) {
    for (int k = lane * 4; k < Kp; k += 128) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < K) v = __ldg(reinterpret_cast<const float4*>(x + k));
      const float s[4] = {v.x * mul, v.y * mul, v.z * mul, v.w * mul};
      __half h[4], l[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        h[i] = __float2half_rn(s[i]);
        l[i] = __float2half_rn(s[i] - __half2float(h[i]));
      }
      uint2 ph, pl;
      ph.x = (uint32_t)__half_as_ushort(h[0]) | ((uint32_t)__half_as_ushort(h[1]) << 16);
      ph.y = (uint32_t)__half_as_ushort(h[2]) | ((uint32_t)__half_as_ushort(h[3]) << 16);
      pl.x = (uint32_t)__half_as_ushort(l[0]) | ((uint32_t)__half_as_ushort(l[1]) << 16);
      pl.y = (uint32_t)__half_as_ushort(l[2]) | ((uint32_t)__half_as_ushort(l[3]) << 16);
      *reinterpret_cast<uint2*>(hi + k) = ph;     // Kp % 64 == 0 and 256-byte aligned planes: 8-byte aligned
      *reinterpret_cast<uint2*>(lo + k) = pl;
    }
  } else {
    for (int k = lane; k < Kp; k += 32) {
      const float s = (k < K) ? __ldg(x + k) * mul : 0.f;
      const __half h = __float2half_rn(s);
      hi[k] = h;
      lo[k] = __float2half_rn(s - __half2float(h));
    }
  }
  if (lane == 0) S.inv_scale[r] = inv;
}

__global__ void __launch_bounds__(PS_WARPS * 32)
presplit_kernel(const SplitSet A, const SplitSet B, const int blocks_a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((int)blockIdx.x < blocks_a) presplit_row(A, (int64_t)blockIdx.x * PS_WARPS + warp, lane);
  else presplit_row(B, (int64_t)(blockIdx.x - blocks_a) * PS_WARPS + warp, lane);
}

}  // namespace

int launch_presplit(const SplitSet& A, const SplitSet& B, cudaStream_t st) {
  const int64_t ba = (A.rows_pad + PS_WARPS - 1) / PS_WARPS, bb = (B.rows_pad + PS_WARPS - 1) / PS_WARPS;
  if (ba + bb == 0) return 0;
  if (ba + bb >= (1ll << 31)) { set_error("too many rows for the operand split"); return B200KGE_ERR_INVALID; }
  presplit_kernel<<<(unsigned)(ba + bb), PS_WARPS * 32, 0, st>>>(A, B, (int)ba);
  B2K_LAUNCH_CHECK("presplit_kernel");
  return 0;
}

}  // namespace b200kge
