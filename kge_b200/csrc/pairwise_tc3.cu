// pairwise_tc3.cu — tcgen05 1-vs-N scorer on PRE-SPLIT fp16 operand planes: THE default kernel of the dot family
// (B200KGE_PREC_AUTO / F16X3; measured 0.069 ms at the FB15k-237 headline shape against 0.118 ms for the in-kernel
// split of pairwise_tc.cu, profiles/r2_summary.md).
//
//   S[q, e] = qs[q] * ts[e] * sum_k (Qh[q,k]*Th[e,k] + Qh[q,k]*Tl[e,k] + Ql[q,k]*Th[e,k])
//
// Qh/Ql, Th/Tl: fp16 hi/lo planes of the row-scaled folded queries / entity table written once per call by
// presplit.cu (qs, ts: per-row powers of two).  Why (analysis of the default kernel, profiles/r1_notes.md):
// at M=128 x N=256 every tcgen05.mma streams 12 KB of operands from shared memory in 128 clk (96 of the
// SM's 128 B/clk), so the in-kernel split of pairwise_tc.cu (raw tile written by TMA, read and re-written
// by the splitter warps: 144 KB of extra shared-memory traffic per 32-wide K chunk on top of the MMAs'
// 96 KB) is what bounds it.  Each table tile is consumed by every query tile (16x at n=1024, both
// directions), so the split is done ONCE in HBM instead (presplit.cu: 4 B read + 4 B written per element,
// the planes take exactly the bytes of the raw table) and this kernel is a pure TMA -> MMA -> epilogue
// pipeline: per 64-wide K chunk 96 KB land by TMA and 12 MMAs (6 f16 slots per 32 elements instead of 8)
// read 144 KB; nothing else touches shared memory in the main loop.
//
// Pipeline: ring of 4 slots of 48 KB = (query plane box 16 KB | table plane box 32 KB).  K chunk c uses
// slots 2c%4 (hi planes) and 2c%4+1 (lo planes): the hi*hi MMAs start when only the hi half has landed.
//   full[s]   TMA bytes landed                          -> MMA
//   free[s]   MMAs reading slot s retired (commit)      -> TMA producer
// CTA = 12 warps, one CTA per SM, persistent over (query tile, range of entity tiles):
//   warp 0 TMA producer | warp 1 MMA issuer + TMEM alloc | warps 2-3 idle | warps 4-11 epilogue
// (warp numbering as in pairwise_tc.cu so the epilogue code is shared verbatim).  Tile = 128 queries x <=256
// entities, 2 TMEM accumulators of 256 columns.
#include "tc_common.cuh"

namespace b200kge {

namespace {

constexpr int TM = 128;             // queries per tile  (UMMA M)
constexpr int TN = 256;             // max entities per tile (UMMA N)
using tc::EPI_WARPS;
using tc::STG_LD;
constexpr int NTHREADS3 = 12 * 32;
constexpr int STG_BYTES = EPI_WARPS * 32 * STG_LD * 4;
constexpr int PIPE_BYTES = 192 * 1024;
constexpr int TMEM_COLS = 512;

// TKH = halfs per K chunk: 64 (128-byte rows, 128-byte swizzle, 4 slots of 48 KB — the default) or 32 (64-byte
// rows, 64-byte swizzle, 8 slots of 24 KB: same bytes in flight in twice as many, half as large transactions;
// B200KGE_TC3_TK=32).  Which granularity feeds the MMAs better is a measurement for the next round.
template <int TKH> struct Cfg3 {
  static constexpr int A_BYTES = TM * TKH * 2;
  static constexpr int B_BYTES = TN * TKH * 2;
  static constexpr int SLOT_BYTES = A_BYTES + B_BYTES;
  static constexpr int NSLOT = PIPE_BYTES / SLOT_BYTES;      // 4 | 8
  static constexpr int SMEM_BYTES = 1024 /*align slack*/ + PIPE_BYTES + STG_BYTES + 256 /*barriers*/;
};
template <int TKH>
__device__ __forceinline__ uint64_t make_desc3(uint32_t addr) {
  return (TKH == 64) ? ptx::umma_desc_sw128(addr) : ptx::umma_desc_sw64(addr);
}

struct Tc3Params {
  int64_t nq, m;
  int nk;           // K chunks of 64 halfs (planes are zero padded to nk * 64)
  int last_ksteps;  // 16-element MMA steps of the LAST chunk that hold real data (1..4): zero padding is not multiplied
  int q_tiles, e_tiles, echunks;
  int tn;           // entities per tile (multiple of 16, <= TN)
  int ksplit;       // > 1: split-K GEMM mode (EPI_STORE only): the reduction is cut into `ksplit` segments of `kseg`
  int kseg;         //      K chunks; every (tile, segment) is its own work item and ADDS into the zeroed output
  const float* q_scale;   // [nq]
  const float* t_scale;   // [m + 32], zero beyond m
  EpiParams epi;
};

template <int EPI, int TKH>
__global__ void __launch_bounds__(NTHREADS3, 1)
pairwise_tc3_kernel(const __grid_constant__ CUtensorMap tmQh, const __grid_constant__ CUtensorMap tmQl,
                    const __grid_constant__ CUtensorMap tmTh, const __grid_constant__ CUtensorMap tmTl,
                    const Tc3Params prm) {
  using C = Cfg3<TKH>;
  constexpr int NSLOT = C::NSLOT, A_BYTES = C::A_BYTES, SLOT_BYTES = C::SLOT_BYTES;
  constexpr int NPAIR = NSLOT / 2;      // K chunks in flight: chunk c owns slots 2*(c % NPAIR), +1 on use c / NPAIR
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* stg = reinterpret_cast<float*>(smem + NSLOT * SLOT_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NSLOT * SLOT_BYTES + STG_BYTES);
  uint64_t* full = bars;                  // [NSLOT]
  uint64_t* free_ = bars + NSLOT;         // [NSLOT]
  uint64_t* tfull = bars + 2 * NSLOT;     // [2]
  uint64_t* tempty = tfull + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk = prm.nk;
  const int total_work = prm.q_tiles * prm.echunks * prm.ksplit;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmQh);
    ptx::prefetch_tensormap(&tmQl);
    ptx::prefetch_tensormap(&tmTh);
    ptx::prefetch_tensormap(&tmTl);
    for (int s = 0; s < NSLOT; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&free_[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&tfull[b], 1);
      ptx::mbar_init(&tempty[b], EPI_WARPS);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<TMEM_COLS>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  int k0 = 0, k1 = nk;      // K-chunk range of the current work item (split-K mode: one segment)
  auto work_range = [&](int w, int& qt, int& et0, int& et1, int& ec) {
    if (prm.ksplit > 1) {
      const int ks = w % prm.ksplit;
      w /= prm.ksplit;
      k0 = ks * prm.kseg;
      k1 = (k0 + prm.kseg < nk) ? k0 + prm.kseg : nk;
    }
    qt = w / prm.echunks;
    ec = w - qt * prm.echunks;
    const int base = prm.e_tiles / prm.echunks, rem = prm.e_tiles % prm.echunks;
    et0 = ec * base + (ec < rem ? ec : rem);
    et1 = et0 + base + (ec < rem ? 1 : 0);
  };

  if (warp == 0) {
    // ================================ TMA producer =========================================
    // the whole warp walks the loops (uniform control flow, operands in uniform registers), one elected lane issues —
    // see pairwise_tc4.cu: inside `if (lane == 0)` ptxas wraps every UTMALDG / UTCHMMA in a register-broadcast loop
    {
      const bool issuer = ptx::elect_one();
      const uint32_t tx = (uint32_t)(A_BYTES + prm.tn * TKH * 2);
      uint32_t c = 0;      // K-chunk counter
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        int qt, et0, et1, ec;
        work_range(w, qt, et0, et1, ec);
        for (int et = et0; et < et1; ++et) {
          for (int kc = k0; kc < k1; ++kc, ++c) {
            const uint32_t par = ((c / NPAIR) & 1) ^ 1;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int s = 2 * (int)(c % NPAIR) + h;
              ptx::mbar_wait_bounded(&free_[s], par);
              uint8_t* sp = smem + s * SLOT_BYTES;
              if (issuer) {
                ptx::mbar_arrive_expect_tx(&full[s], tx);
                ptx::tma_load_2d(sp, h ? &tmQl : &tmQh, &full[s], kc * TKH, qt * TM);
                ptx::tma_load_2d(sp + A_BYTES, h ? &tmTl : &tmTh, &full[s], kc * TKH, et * prm.tn);
              }
              __syncwarp();
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ============================================
    {
      const bool issuer = ptx::elect_one();
      const uint32_t idesc = ptx::umma_idesc_f16(TM, prm.tn);
      uint32_t c = 0, it = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        int qt, et0, et1, ec;
        work_range(w, qt, et0, et1, ec);
        for (int et = et0; et < et1; ++et, ++it) {
          const int b = it & 1;
          ptx::mbar_wait_bounded(&tempty[b], ((it >> 1) & 1) ^ 1);   // epilogue drained this accumulator
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(b * TN);
          for (int kc = k0; kc < k1; ++kc, ++c) {
            const int sh = 2 * (int)(c % NPAIR), sl = sh + 1;
            const uint32_t par = (c / NPAIR) & 1;
            const uint32_t a_hi = ptx::smem_u32(smem + sh * SLOT_BYTES), b_hi = a_hi + A_BYTES;
            const uint32_t a_lo = ptx::smem_u32(smem + sl * SLOT_BYTES), b_lo = a_lo + A_BYTES;
            const bool tail = (kc == nk - 1) && prm.last_ksteps < TKH / 16;      // zero-padded steps are not multiplied
            ptx::mbar_wait_bounded(&full[sh], par);
            ptx::tc_fence_after();
            if (issuer) {
              if (!tail) {
#pragma unroll
                for (int k = 0; k < TKH / 16; ++k)
                  ptx::umma_bf16(d_tmem, make_desc3<TKH>(a_hi + k * 32), make_desc3<TKH>(b_hi + k * 32), idesc,
                                 (kc > k0 || k > 0) ? 1u : 0u);
              } else {
                for (int k = 0; k < prm.last_ksteps; ++k)
                  ptx::umma_bf16(d_tmem, make_desc3<TKH>(a_hi + k * 32), make_desc3<TKH>(b_hi + k * 32), idesc,
                                 (kc > k0 || k > 0) ? 1u : 0u);
              }
            }
            ptx::mbar_wait_bounded(&full[sl], par);
            ptx::tc_fence_after();
            if (issuer) {
              if (!tail) {
#pragma unroll
                for (int k = 0; k < TKH / 16; ++k) {
                  ptx::umma_bf16(d_tmem, make_desc3<TKH>(a_hi + k * 32), make_desc3<TKH>(b_lo + k * 32), idesc, 1u);
                  ptx::umma_bf16(d_tmem, make_desc3<TKH>(a_lo + k * 32), make_desc3<TKH>(b_hi + k * 32), idesc, 1u);
                }
              } else {
                for (int k = 0; k < prm.last_ksteps; ++k) {
                  ptx::umma_bf16(d_tmem, make_desc3<TKH>(a_hi + k * 32), make_desc3<TKH>(b_lo + k * 32), idesc, 1u);
                  ptx::umma_bf16(d_tmem, make_desc3<TKH>(a_lo + k * 32), make_desc3<TKH>(b_hi + k * 32), idesc, 1u);
                }
              }
              ptx::umma_commit(&free_[sh]);
              ptx::umma_commit(&free_[sl]);
            }
            __syncwarp();
          }
          if (issuer) ptx::umma_commit(&tfull[b]);             // accumulator complete
          __syncwarp();
        }
      }
    }
  } else if (warp >= 4) {
    // ================================ epilogue ==============================================
    const int quad = warp & 3;                // TMEM lanes [32*quad, +32)
    const int half = (warp - 4) >> 2;         // columns [128*half, +128) of the accumulator
    float* my_stg = stg + (warp - 4) * 32 * STG_LD;
    const EpiParams& P = prm.epi;
    uint32_t it = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      int qt, et0, et1, ec;
      work_range(w, qt, et0, et1, ec);
      const int64_t row = (int64_t)qt * TM + quad * 32 + lane;   // this thread's query row
      const bool row_ok = row < prm.nq;
      RowState<EPI> st;
      st.init();
      const float aux = row_ok ? epi_row_aux<EPI>(P, row) : 0.f;
      const float qs = row_ok ? __ldg(prm.q_scale + row) : 0.f;
      const int64_t csr_end = (P.csr_off && row_ok) ? __ldg(P.csr_off + row + 1) : 0;
      for (int et = et0; et < et1; ++et, ++it) {
        const int b = it & 1;
        ptx::mbar_wait_bounded(&tfull[b], (it >> 1) & 1);
        ptx::tc_fence_after();
        const int64_t tile_end = (int64_t)(et + 1) * prm.tn;
        tc::epilogue_tile<EPI, 4, true>(P, st, aux,
                                        tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(b * TN + half * 128),
                                        (int64_t)qt * TM + quad * 32, (int64_t)et * prm.tn + half * 128, prm.nq,
                                        tile_end < prm.m ? tile_end : prm.m, my_stg, lane, qs, prm.t_scale,
                                        csr_end);
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&tempty[b]);
      }
      if constexpr (EPI != EPI_STORE) {
        if (row_ok) epi_flush<EPI>(P, st, row, ec * 2 + half);
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
using tc::num_sms;

// same tiling policy as pairwise_tc.cu: tile width (multiple of 16 in [128, 256]) minimising the per-SM
// makespan in columns, entity tiles split into `echunks` ranges so that q_tiles * echunks ~ #SMs
void plan3(int64_t nq, int64_t m, int& q_tiles, int& e_tiles, int& echunks, int& tn) {
  q_tiles = (int)((nq + TM - 1) / TM);
  if (q_tiles < 1) q_tiles = 1;
  const int units = num_sms();
  int64_t best_cost = -1;
  tn = TN;
  for (int cand = TN; cand >= 128; cand -= 16) {
    const int64_t et = (m + cand - 1) / cand;
    int per = units / q_tiles; if (per < 1) per = 1; if (per > et) per = (int)et;
    const int64_t tiles_per_cta = (et + per - 1) / per;
    const int64_t waves = ((int64_t)q_tiles * per + units - 1) / units;
    const int64_t cost = waves * tiles_per_cta * cand + tiles_per_cta * 24;
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; tn = cand; }
  }
  e_tiles = (int)((m + tn - 1) / tn);
  int per = units / q_tiles;
  if (per < 1) per = 1;
  if (per > e_tiles) per = e_tiles;
  echunks = per;
}

template <int EPI, int TKH>
int launch_k3t(const CUtensorMap& qh, const CUtensorMap& ql, const CUtensorMap& th, const CUtensorMap& tl,
               const Tc3Params& prm, int grid, cudaStream_t st) {
  auto kern = pairwise_tc3_kernel<EPI, TKH>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg3<TKH>::SMEM_BYTES);
  if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(pairwise_tc3_kernel)");
  profile_begin(st);
  kern<<<grid, NTHREADS3, Cfg3<TKH>::SMEM_BYTES, st>>>(qh, ql, th, tl, prm);
  profile_end(st);
  B2K_LAUNCH_CHECK("pairwise_tc3_kernel");
  return 0;
}
// K chunk = 64 halfs (128-byte rows, 128-byte swizzle, 4 slots of 48 KB).  The 32-half variant (8 slots of 24 KB)
// measured slower on B200 (74.4 vs 69.3 us at the FB15k-237 shape, profiles/r2_summary.md) and was removed.
template <int EPI>
int launch_k3(int, const CUtensorMap& qh, const CUtensorMap& ql, const CUtensorMap& th, const CUtensorMap& tl,
              const Tc3Params& prm, int grid, cudaStream_t st) {
  return launch_k3t<EPI, 64>(qh, ql, th, tl, prm, grid, st);
}
int tk3_choice() { return 64; }

}  // namespace

int tc3_nchunks(int64_t nq, int64_t m) {
  int qt, et, ec, tn;
  plan3(nq, m, qt, et, ec, tn);
  return 2 * ec;
}

int launch_pairwise_tc3(int epi_kind, const SplitSet& Q, const SplitSet& T, const EpiParams& P, cudaStream_t st) {
  const int64_t nq = Q.rows, m = T.rows;
  if (nq == 0 || m == 0) return 0;
  if (Q.Kp != T.Kp || Q.Kp % 64 != 0) { set_error("operand planes disagree on the padded reduction length"); return B200KGE_ERR_INVALID; }
  const int tkh = tk3_choice();
  Tc3Params prm;
  prm.nq = nq; prm.m = m; prm.nk = Q.Kp / tkh;
  plan3(nq, m, prm.q_tiles, prm.e_tiles, prm.echunks, prm.tn);
  prm.ksplit = 1; prm.kseg = prm.nk;
  { const int rem = Q.K - (prm.nk - 1) * tkh;            // real reduction elements in the last chunk
    prm.last_ksteps = rem <= 0 ? 4 : (rem + 15) / 16; if (prm.last_ksteps > 4) prm.last_ksteps = 4; }
  if (P.accumulate_out) {
    // split-K GEMM: segments of 8 chunks (512 reduction elements) bound the tensor core's accumulator error, which
    // grows with the reduction length (2.4e-5 of rms at K=512, 2.8e-4 at K=14541: profiles/r2_summary.md); segment
    // results are added in fp32 by the epilogue (red.global.add).  One entity tile per work item.
    if (epi_kind != EPI_STORE) { set_error("split-K accumulation is a GEMM (store) mode"); return B200KGE_ERR_INVALID; }
    prm.kseg = 8;
    prm.ksplit = (prm.nk + prm.kseg - 1) / prm.kseg;
    prm.echunks = prm.e_tiles;
  }
  prm.q_scale = Q.inv_scale; prm.t_scale = T.inv_scale;
  CUtensorMap mQh, mQl, mTh, mTl;
  int rc;
  if ((rc = tc::make_map_f16(&mQh, Q.hi, nq, Q.Kp, Q.Kp, TM, tkh))) return rc;
  if ((rc = tc::make_map_f16(&mQl, Q.lo, nq, Q.Kp, Q.Kp, TM, tkh))) return rc;
  if ((rc = tc::make_map_f16(&mTh, T.hi, m, T.Kp, T.Kp, prm.tn, tkh))) return rc;
  if ((rc = tc::make_map_f16(&mTl, T.lo, m, T.Kp, T.Kp, prm.tn, tkh))) return rc;
  prm.epi = P;
  prm.epi.nchunks = 2 * prm.echunks;   // two epilogue warps (column halves) per row
  const int total = prm.q_tiles * prm.echunks * prm.ksplit;
  const int grid = total < num_sms() ? total : num_sms();
  switch (epi_kind) {
    case EPI_STORE: return launch_k3<EPI_STORE>(tkh, mQh, mQl, mTh, mTl, prm, grid, st);
    case EPI_BCE:   return launch_k3<EPI_BCE>(tkh, mQh, mQl, mTh, mTl, prm, grid, st);
    case EPI_KL:    return launch_k3<EPI_KL>(tkh, mQh, mQl, mTh, mTl, prm, grid, st);
    case EPI_RANK:  return launch_k3<EPI_RANK>(tkh, mQh, mQl, mTh, mTl, prm, grid, st);
  }
  set_error("bad epilogue kind %d", epi_kind);
  return B200KGE_ERR_INVALID;
}

}  // namespace b200kge
