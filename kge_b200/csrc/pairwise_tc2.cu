// pairwise_tc2.cu — CTA-pair (tcgen05 cta_group::2) version of the 3xTF32 dot-family scorer.
//
// Same math and epilogues as pairwise_tc.cu (see there for the precision scheme and reference
// citations), re-tiled for a cluster of two CTAs on the two SMs of a TPC:
//
//   cluster tile = 256 queries x 256 entities; UMMA M=256 (each CTA contributes its own 128 query
//   rows as A and receives their accumulator rows in its own TMEM), UMMA N=256 (each CTA stages
//   HALF of the entity tile — 128 rows — in its shared memory and the tensor cores of both SMs read
//   both halves).  Per CTA and K-chunk: TMA brings the RAW folded queries (128 x TK) and the raw
//   table half (128 x TK); the CTA's splitter warps derive Q_lo and T_lo on chip (raw = hi).
//
// Why: v1 (1 CTA, 96 KB stages x 2) was latency-bound on TMA -> split -> MMA (profiles/r1_tc_summary.md).
// Here a stage is 64 KB (TK=32) or 32 KB (TK=16), so 3 or 6 stages are in flight per SM, the
// split work per SM halves, and the B-operand shared-memory traffic of the MMAs halves.
//
// Cross-CTA protocol (leader = cluster rank 0 issues every MMA):
//   full[s]    local   : TMA bytes of this CTA's stage s landed             (count 1 + tx)
//   landed[s]  leader  : both CTAs' TMA data of stage s landed              (count 2; hi*hi MMAs may start)
//   split[s]   leader  : both CTAs' splitters finished stage s              (count 8: one per warp)
//   empty[s]   both    : MMAs reading stage s retired (multicast tcgen05.commit)
//   tfull[b]   both    : accumulator b complete      (multicast tcgen05.commit)
//   tempty[b]  leader  : both CTAs' epilogues drained accumulator b         (count 8: one per warp)
#include "tc_common.cuh"

namespace b200kge {

namespace {

constexpr int TM = 128;     // query rows per CTA (cluster: 256)
constexpr int TNH = 128;    // entity rows staged per CTA (cluster N: 256)
constexpr int TN = 256;
using tc::EPI_WARPS;
using tc::SPLIT_WARPS;
using tc::NTHREADS;
constexpr int STG_BYTES = EPI_WARPS * 32 * tc::STG_LD * 4;
constexpr int PIPE_BYTES = 192 * 1024;
constexpr int TMEM_COLS = 512;

template <int TK> struct Cfg {
  static constexpr int A_BYTES = TM * TK * 4;
  static constexpr int STAGE_BYTES = 4 * A_BYTES;            // Q_hi, Q_lo, T_half(raw = hi), T_lo
  static constexpr int STAGES = PIPE_BYTES / STAGE_BYTES;    // 3 (TK=32) or 6 (TK=16)
  static constexpr int NBARS = 4 * STAGES + 4;
  static constexpr int SMEM_BYTES = 1024 + PIPE_BYTES + STG_BYTES + NBARS * 8 + 64;
};

struct Tc2Params {
  int64_t nq, m;
  int K;
  int q_tiles, e_tiles, echunks;   // q tiles of 256 rows, e tiles of 256 rows
  int dbg;                         // experiments (B200KGE_DBG): 1 = skip TMA after first fills, 2 = skip split math, 4 = skip MMAs
  EpiParams epi;
};

template <int TK>
__device__ __forceinline__ uint64_t make_desc(uint32_t addr) {
  return (TK == 32) ? ptx::umma_desc_sw128(addr) : ptx::umma_desc_sw64(addr);
}

template <int EPI, int PASSES, int TK>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
pairwise_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmT,
                    const Tc2Params prm) {
  using C = Cfg<TK>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* stg = reinterpret_cast<float*>(smem + PIPE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + PIPE_BYTES + STG_BYTES);
  uint64_t* full = bars;
  uint64_t* split = bars + STAGES;
  uint64_t* empty = bars + 2 * STAGES;
  uint64_t* landed = bars + 3 * STAGES;   // leader: both CTAs' TMA data of stage s landed (count 2)
  uint64_t* tfull = bars + 4 * STAGES;
  uint64_t* tempty = bars + 4 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  const int nk = (prm.K + TK - 1) / TK;
  const int total_work = prm.q_tiles * prm.echunks;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmQ);
    ptx::prefetch_tensormap(&tmT);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&split[s], 2 * SPLIT_WARPS);      // one arrive per splitter warp
      ptx::mbar_init(&empty[s], 1);
      ptx::mbar_init(&landed[s], 2);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&tfull[b], 1);
      ptx::mbar_init(&tempty[b], 2 * EPI_WARPS);     // one arrive per epilogue warp
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc_2cta<TMEM_COLS>(tmem_slot);
  ptx::tc_fence_before();
  ptx::cluster_sync_all();          // barriers of BOTH CTAs initialised before any remote arrive
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto stage_ptr = [&](int s) { return smem + s * C::STAGE_BYTES; };
  auto work_range = [&](int w, int& qt, int& et0, int& et1, int& ec) {
    qt = w / prm.echunks;
    ec = w - qt * prm.echunks;
    const int base = prm.e_tiles / prm.echunks, rem = prm.e_tiles % prm.echunks;
    et0 = ec * base + (ec < rem ? ec : rem);
    et1 = et0 + base + (ec < rem ? 1 : 0);
  };

  if (warp == 0) {
    // ================================ TMA producer (both CTAs) ==============================
    if (lane == 0) {
      uint32_t c = 0;
      for (int w = cluster_id; w < total_work; w += nclusters) {
        int qt, et0, et1, ec;
        work_range(w, qt, et0, et1, ec);
        const int q_row = qt * 256 + (int)rank * TM;
        for (int et = et0; et < et1; ++et) {
          const int e_row = et * TN + (int)rank * TNH;
          for (int kc = 0; kc < nk; ++kc, ++c) {
            const int s = c % STAGES;
            const uint32_t ph = (c / STAGES) & 1;
            ptx::mbar_wait_cluster(&empty[s], ph ^ 1);
            uint8_t* sp = stage_ptr(s);
            if ((prm.dbg & 1) && c >= (uint32_t)STAGES) { ptx::mbar_arrive(&full[s]); continue; }
            ptx::mbar_arrive_expect_tx(&full[s], 2 * C::A_BYTES);
            ptx::tma_load_2d(sp, &tmQ, &full[s], kc * TK, q_row);                      // raw queries
            ptx::tma_load_2d(sp + 2 * C::A_BYTES, &tmT, &full[s], kc * TK, e_row);     // raw table half
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (leader CTA only) ===========================
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_tf32(256, TN);
      uint32_t c = 0, it = 0;
      for (int w = cluster_id; w < total_work; w += nclusters) {
        int qt, et0, et1, ec;
        work_range(w, qt, et0, et1, ec);
        for (int et = et0; et < et1; ++et, ++it) {
          const int b = it & 1;
          ptx::mbar_wait_cluster(&tempty[b], ((it >> 1) & 1) ^ 1);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(b * TN);
          for (int kc = 0; kc < nk; ++kc, ++c) {
            const int s = c % STAGES;
            const uint32_t ph = (c / STAGES) & 1;
            const uint32_t a_hi = ptx::smem_u32(stage_ptr(s));
            const uint32_t a_lo = a_hi + C::A_BYTES;
            const uint32_t b_hi = a_hi + 2 * C::A_BYTES;
            const uint32_t b_lo = a_hi + 3 * C::A_BYTES;
            // raw tiles are the hi operands: the hi*hi products start as soon as both CTAs' TMA
            // data has landed, overlapping the splitters' work on the same stage
            if ((prm.dbg & 8) && PASSES != 1) ptx::mbar_wait_cluster(&split[s], ph);   // experiment: no early issue
            ptx::mbar_wait_cluster(&landed[s], ph);
            ptx::tc_fence_after();
            if (!(prm.dbg & 4)) {
#pragma unroll
              for (int k8 = 0; k8 < TK / 8; ++k8)
                ptx::umma_tf32_2cta(d_tmem, make_desc<TK>(a_hi + k8 * 32), make_desc<TK>(b_hi + k8 * 32), idesc,
                                    (kc > 0 || k8 > 0) ? 1u : 0u);
            }
            if (PASSES == 3) {
              ptx::mbar_wait_cluster(&split[s], ph);
              ptx::tc_fence_after();
              if (!(prm.dbg & 4)) {
#pragma unroll
                for (int k8 = 0; k8 < TK / 8; ++k8) {
                  ptx::umma_tf32_2cta(d_tmem, make_desc<TK>(a_lo + k8 * 32), make_desc<TK>(b_hi + k8 * 32), idesc, 1u);
                  ptx::umma_tf32_2cta(d_tmem, make_desc<TK>(a_hi + k8 * 32), make_desc<TK>(b_lo + k8 * 32), idesc, 1u);
                }
              }
            } else if (PASSES == 2) {
              // mixed mode (TK == 32 only): cross terms in bf16 on 64-B swizzled tiles (hi16 | lo16) that
              // the splitters wrote where the fp32 lo tiles would be
              ptx::mbar_wait_cluster(&split[s], ph);
              ptx::tc_fence_after();
              constexpr uint32_t idesc16 = ptx::umma_idesc_bf16(256, TN);
              const uint32_t a16h = a_lo, a16l = a_lo + C::A_BYTES / 2;
              const uint32_t b16h = b_lo, b16l = b_lo + C::A_BYTES / 2;
              if (!(prm.dbg & 4)) {
#pragma unroll
                for (int k2 = 0; k2 < TK / 16; ++k2) {
                  ptx::umma_bf16_2cta(d_tmem, ptx::umma_desc_sw64(a16l + k2 * 32), ptx::umma_desc_sw64(b16h + k2 * 32), idesc16, 1u);
                  ptx::umma_bf16_2cta(d_tmem, ptx::umma_desc_sw64(a16h + k2 * 32), ptx::umma_desc_sw64(b16l + k2 * 32), idesc16, 1u);
                }
              }
            }
            ptx::umma_commit_2cta(&empty[s], 0b11);
          }
          ptx::umma_commit_2cta(&tfull[b], 0b11);
        }
      }
    }
  } else if (warp >= 12) {
    // ================================ splitters (both CTAs) ==================================
    const int t = threadIdx.x - 12 * 32;
    uint32_t c = 0;
    for (int w = cluster_id; w < total_work; w += nclusters) {
      int qt, et0, et1, ec;
      work_range(w, qt, et0, et1, ec);
      for (int et = et0; et < et1; ++et) {
        for (int kc = 0; kc < nk; ++kc, ++c) {
          const int s = c % STAGES;
          const uint32_t ph = (c / STAGES) & 1;
          ptx::mbar_wait(&full[s], ph);
          // tell the leader this CTA's stage has landed (its hi*hi MMAs may start)
          if (t == 0) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&landed[s]), 0));
          if (PASSES != 1) {
            if (!(prm.dbg & 2)) {
              // both operands are split here (raw tile = hi; write lo next to it): nothing derivable
              // on chip is fetched over the L2->SM fabric
              const uint32_t sp = ptx::smem_u32(stage_ptr(s));
              if (PASSES == 3) {
                tc::split_tile<C::A_BYTES, SPLIT_WARPS * 32>(sp, sp + C::A_BYTES, t);
                tc::split_tile<C::A_BYTES, SPLIT_WARPS * 32>(sp + 2 * C::A_BYTES, sp + 3 * C::A_BYTES, t);
              } else if (TK == 32) {
                tc::split_tile_bf16<TM, SPLIT_WARPS * 32>(sp, sp + C::A_BYTES, sp + C::A_BYTES + C::A_BYTES / 2, t);
                tc::split_tile_bf16<TNH, SPLIT_WARPS * 32>(sp + 2 * C::A_BYTES, sp + 3 * C::A_BYTES,
                                                           sp + 3 * C::A_BYTES + C::A_BYTES / 2, t);
              }
            }
            ptx::fence_proxy_async_smem();
          }
          // one remote arrive per WARP (128 per-thread DSMEM arrives per chunk serialise on the
          // leader's barrier and were the bottleneck): every lane fenced its own writes to the
          // async proxy, __syncwarp orders them before lane 0's cluster-scope release.
          if (PASSES != 1) {
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&split[s]), 0));
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ================================ epilogue (both CTAs) ===================================
    const int quad = warp & 3;                // TMEM lanes [32*quad, +32)
    const int half = (warp - 4) >> 2;         // columns [128*half, +128) of the accumulator
    float* my_stg = stg + (warp - 4) * 32 * tc::STG_LD;
    const EpiParams& P = prm.epi;
    uint32_t it = 0;
    for (int w = cluster_id; w < total_work; w += nclusters) {
      int qt, et0, et1, ec;
      work_range(w, qt, et0, et1, ec);
      const int64_t row0 = (int64_t)qt * 256 + (int64_t)rank * TM + quad * 32;
      const int64_t row = row0 + lane;
      const bool row_ok = row < prm.nq;
      RowState<EPI> st;
      st.init();
      const float aux = row_ok ? epi_row_aux<EPI>(P, row) : 0.f;
      for (int et = et0; et < et1; ++et, ++it) {
        const int b = it & 1;
        ptx::mbar_wait_cluster(&tfull[b], (it >> 1) & 1);
        ptx::tc_fence_after();
        tc::epilogue_tile<EPI, 4>(P, st, aux,
                                  tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(b * TN + half * 128),
                                  row0, (int64_t)et * TN + half * 128, prm.nq, prm.m, my_stg, lane);
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&tempty[b]), 0));
      }
      if constexpr (EPI != EPI_STORE) {
        if (row_ok) epi_flush<EPI>(P, st, row, ec * 2 + half);
      }
    }
  }

  ptx::tc_fence_before();
  ptx::cluster_sync_all();   // nobody exits (or frees TMEM) while the peer may still signal it
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2cta<TMEM_COLS>(tmem_base);
  }}

void plan2(int64_t nq, int64_t m, int& q_tiles, int& e_tiles, int& echunks) {
  q_tiles = (int)((nq + 255) / 256);
  e_tiles = (int)((m + TN - 1) / TN);
  const int nclusters = tc::num_sms() / 2;
  int per = nclusters / (q_tiles > 0 ? q_tiles : 1);
  if (per < 1) per = 1;
  if (per > e_tiles) per = e_tiles;
  echunks = per;
}

int tk_choice() {
  const char* e = getenv("B200KGE_TC2_TK");   // experiments: 16 (64 B swizzle, 6 stages) | 32 (128 B, 3 stages)
  return (e && atoi(e) == 16) ? 16 : 32;
}

template <int EPI, int PASSES, int TK>
int launch_k(const CUtensorMap& a, const CUtensorMap& c, const Tc2Params& prm, int grid, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(pairwise_tc2_kernel<EPI, PASSES, TK>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<TK>::SMEM_BYTES);
  if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(pairwise_tc2_kernel)");
  profile_begin(st);
  pairwise_tc2_kernel<EPI, PASSES, TK><<<grid, NTHREADS, Cfg<TK>::SMEM_BYTES, st>>>(a, c, prm);
  profile_end(st);
  B2K_LAUNCH_CHECK("pairwise_tc2_kernel");
  return 0;
}

template <int EPI>
int launch_e(int passes, int tk, const CUtensorMap& a, const CUtensorMap& c, const Tc2Params& prm, int grid,
             cudaStream_t st) {
  if (passes == 3) return tk == 32 ? launch_k<EPI, 3, 32>(a, c, prm, grid, st) : launch_k<EPI, 3, 16>(a, c, prm, grid, st);
  if (passes == 2) return launch_k<EPI, 2, 32>(a, c, prm, grid, st);   // mixed mode: 32-wide chunks only
  return tk == 32 ? launch_k<EPI, 1, 32>(a, c, prm, grid, st) : launch_k<EPI, 1, 16>(a, c, prm, grid, st);
}

}  // namespace

int tc2_nchunks(int64_t nq, int64_t m) {
  int qt, et, ec;
  plan2(nq, m, qt, et, ec);
  return 2 * ec;
}

int launch_pairwise_tc2(int epi_kind, int passes, const float* Q, int64_t ldq,
                        int64_t nq, const float* T, int64_t ldt, int64_t m, int K,
                        const EpiParams& P, cudaStream_t st) {
  if (nq == 0 || m == 0) return 0;
  const int tk = (passes == 2) ? 32 : tk_choice();
  CUtensorMap mQ, mT;
  int rc;
  if ((rc = tc::make_map(&mQ, Q, nq, K, ldq, tk, TM))) return rc;
  if ((rc = tc::make_map(&mT, T, m, K, ldt, tk, TNH))) return rc;
  Tc2Params prm;
  prm.nq = nq; prm.m = m; prm.K = K;
  plan2(nq, m, prm.q_tiles, prm.e_tiles, prm.echunks);
  prm.epi = P;
  prm.epi.nchunks = 2 * prm.echunks;   // two epilogue warps (column halves) per row
  { const char* e = getenv("B200KGE_DBG"); prm.dbg = e ? atoi(e) : 0; }
  const int total = prm.q_tiles * prm.echunks;
  const int nclusters = tc::num_sms() / 2;
  const int grid = 2 * (total < nclusters ? total : nclusters);
  switch (epi_kind) {
    case EPI_STORE: return launch_e<EPI_STORE>(passes, tk, mQ, mT, prm, grid, st);
    case EPI_BCE:   return launch_e<EPI_BCE>(passes, tk, mQ, mT, prm, grid, st);
    case EPI_KL:    return launch_e<EPI_KL>(passes, tk, mQ, mT, prm, grid, st);
    case EPI_RANK:  return launch_e<EPI_RANK>(passes, tk, mQ, mT, prm, grid, st);
  }
  set_error("bad epilogue kind %d", epi_kind);
  return B200KGE_ERR_INVALID;
}

}  // namespace b200kge
