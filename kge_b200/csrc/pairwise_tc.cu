// pairwise_tc.cu — tcgen05 tensor-core 1-vs-N scorer for the dot-product family
// (ComplEx / DistMult / SimplE / CP / RESCAL after folding), fp32-equivalent via 3xTF32.
//
//   S[q, e] = sum_k Q[q,k] * T[e,k]          Q: folded queries [nq, K]  (fold.cu, pre-split hi/lo)
//                                            T: entity table   [m,  K]  streamed RAW from HBM/L2
//
// replaces the reference's torch.mm over concatenated operands (complex.py:37,39,
// distmult.py:19,21, simple.py:25-29, cp.py:24,26, rescal.py:41,47) AND whatever consumes the
// scores next (BCE / KL loss, rank counting, or the plain [n,E] store) in ONE kernel.
//
// Precision: the reference is a true fp32 GEMM; single-pass TF32 misses the 1e-4 bar by 17x
// (hard part 1).  We split x = hi + lo with hi = x & 0xFFFFE000 (tf32-exact) and issue
//   D += Q_lo*T_hi + Q_hi*T_lo + Q_hi*T_hi          (kind::tf32, fp32 accumulate in TMEM)
// which is fp32-equivalent (dropped term lo*lo ~ 2^-22).  Both operands are split ON THE FLY (the
// L2->SM fabric is the scarce resource, so nothing derivable on chip is fetched): TMA lands the raw fp32 tile in shared memory; the raw tile IS the hi
// operand (kind::tf32 ignores the low 13 mantissa bits — truncation, measured on B200), and four
// "splitter" warps write lo = rn_tf32(x - trunc_tf32(x)) next to it (same swizzled layout,
// element-wise), fence to the async proxy, and only then may the MMA warp consume the stage.
//
// CTA = 16 warps, one CTA per SM, persistent over (query tile, range of entity tiles):
//   warp 0      TMA producer   (one elected lane)     full[s]   <- expect_tx
//   warp 1      MMA issuer     (one elected lane)     empty[s]  <- tcgen05.commit ; tmem_full[b]
//   warps 4-11  epilogue       tcgen05.ld -> regs -> {transposed coalesced store | BCE | KL | rank}
//                              (2 warps per TMEM lane quadrant, each takes 128 of the 256 columns)
//   warps 2-3 / 12-15 splitters (query tile / table tile)      split[s] <- one arrival per warp
// Tile = 128 queries (UMMA M, TMEM lanes) x 256 entities (UMMA N, TMEM columns), K in chunks of 32
// floats (one 128-byte swizzle atom), 2 smem stages of 96 KB, 2 TMEM accumulators of 256 columns
// so the epilogue of tile i overlaps the MMAs of tile i+1.  TMEM lane = query row, so every
// per-row reduction (loss terms, logsumexp, rank counters) is thread-local.
#include "tc_common.cuh"

namespace b200kge {

namespace {

constexpr int TM = 128;             // queries per tile  (UMMA M)
constexpr int TN = 256;             // entities per tile (UMMA N)
constexpr int TK = 32;              // floats per K chunk (128 B swizzle atom)
constexpr int STAGES = 2;
constexpr int A_BYTES = TM * TK * 4;   // 16 KB
constexpr int B_BYTES = TN * TK * 4;   // 32 KB
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;  // A_hi, A_lo, B_hi(raw), B_lo
using tc::EPI_WARPS;
using tc::SPLIT_WARPS;
using tc::NTHREADS;
using tc::STG_LD;
constexpr int STG_BYTES = EPI_WARPS * 32 * STG_LD * 4;
constexpr int SMEM_BYTES = 1024 /*align slack*/ + STAGES * STAGE_BYTES + STG_BYTES + 256 /*barriers*/;
constexpr int TMEM_COLS = 512;

struct TcParams {
  int64_t nq, m;
  int K;            // reduction length (floats)
  int q_tiles, e_tiles, echunks;
  int q_groups;     // work is (q group) x (e chunk): a group is one q tile, or a pair of q tiles when the
                    // table tile is multicast across a 2-CTA cluster (MC)
  int dbg;          // experiments (B200KGE_DBG bit-mask): 1 = skip TMA after the first fills, 2 = skip split
                    // math, 4 = skip MMAs — isolates the pipeline phases (results are garbage)
  int tn;           // entities per tile actually used (multiple of 16, <= TN): chosen per problem so that
                    // ceil(tiles / SMs) * tn — the makespan in columns — is minimal
  EpiParams epi;
};

// MC: launched as clusters of 2 CTAs that walk the SAME entity tiles with DIFFERENT query tiles; each CTA
// fetches half of every table tile and TMA-multicasts it into both CTAs' shared memory, so the table
// bytes cross the L2->SM fabric once per pair (-33 % TMA traffic per SM: the TMA phase is on the
// critical path of the 2-stage pipeline).  MMAs stay cta_group::1 and per-CTA; only the stage-free
// signal is shared (each CTA's commit arrives on both CTAs' empty barriers).
template <int EPI, int PASSES, bool MC>
__global__ void __launch_bounds__(NTHREADS, 1)
pairwise_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmT,
                   const TcParams prm) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* stg = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + STG_BYTES);
  uint64_t* full = bars;                 // [STAGES]
  uint64_t* split = bars + STAGES;       // [STAGES]
  uint64_t* empty = bars + 2 * STAGES;   // [STAGES]
  uint64_t* tfull = bars + 3 * STAGES;   // [2]
  uint64_t* tempty = bars + 3 * STAGES + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk = (prm.K + TK - 1) / TK;
  const int total_work = prm.q_groups * prm.echunks;
  const uint32_t rank = MC ? ptx::cluster_ctarank() : 0u;
  const int wstart = MC ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int wstep = MC ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmQ);
    ptx::prefetch_tensormap(&tmT);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&split[s], SPLIT_WARPS + 2);   // B splitters (4 warps) + A splitters (2 warps)
      ptx::mbar_init(&empty[s], MC ? 2 : 1);   // MC: both CTAs of the pair must have retired the stage
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&tfull[b], 1);
      ptx::mbar_init(&tempty[b], EPI_WARPS);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<TMEM_COLS>(tmem_slot);
  ptx::tc_fence_before();
  if (MC) ptx::cluster_sync_all(); else __syncthreads();   // MC: the peer's barriers must be initialised too
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto stage_ptr = [&](int s) { return smem + s * STAGE_BYTES; };
  // e-tile range of work item w
  auto work_range = [&](int w, int& qt, int& et0, int& et1, int& ec) {
    const int qg = w / prm.echunks;
    qt = MC ? 2 * qg + (int)rank : qg;
    ec = w - qg * prm.echunks;
    const int base = prm.e_tiles / prm.echunks, rem = prm.e_tiles % prm.echunks;
    et0 = ec * base + (ec < rem ? ec : rem);
    et1 = et0 + base + (ec < rem ? 1 : 0);
  };

  if (warp == 0) {
    // ================================ TMA producer =========================================
    if (lane == 0) {
      uint32_t c = 0;
      for (int w = wstart; w < total_work; w += wstep) {
        int qt, et0, et1, ec;
        work_range(w, qt, et0, et1, ec);
        for (int et = et0; et < et1; ++et) {
          for (int kc = 0; kc < nk; ++kc, ++c) {
            const int s = c % STAGES;
            const uint32_t ph = (c / STAGES) & 1;
            if (MC) ptx::mbar_wait_cluster(&empty[s], ph ^ 1); else ptx::mbar_wait(&empty[s], ph ^ 1);
            uint8_t* sp = stage_ptr(s);
            if ((prm.dbg & 1) && c >= (uint32_t)STAGES) { ptx::mbar_arrive(&full[s]); continue; }
            ptx::mbar_arrive_expect_tx(&full[s], A_BYTES + prm.tn * TK * 4);
            ptx::tma_load_2d(sp, &tmQ, &full[s], kc * TK, qt * TM);                 // raw queries
            if (MC) {
              // this CTA's half of the table tile, delivered to both CTAs (box = tn/2 rows)
              const int hrows = prm.tn >> 1;
              ptx::tma_load_2d_mc(sp + 2 * A_BYTES + (int)rank * hrows * TK * 4, &tmT, &full[s], kc * TK,
                                  et * prm.tn + (int)rank * hrows, (uint16_t)0b11);
            } else {
              ptx::tma_load_2d(sp + 2 * A_BYTES, &tmT, &full[s], kc * TK, et * prm.tn);   // raw table tile
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ============================================
    if (lane == 0) {
      const uint32_t idesc = ptx::umma_idesc_tf32(TM, prm.tn);
      uint32_t c = 0, it = 0;
      for (int w = wstart; w < total_work; w += wstep) {
        int qt, et0, et1, ec;
        work_range(w, qt, et0, et1, ec);
        for (int et = et0; et < et1; ++et, ++it) {
          const int b = it & 1;
          ptx::mbar_wait(&tempty[b], ((it >> 1) & 1) ^ 1);   // epilogue drained this accumulator
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(b * TN);
          for (int kc = 0; kc < nk; ++kc, ++c) {
            const int s = c % STAGES;
            const uint32_t ph = (c / STAGES) & 1;
            const uint32_t a_hi = ptx::smem_u32(stage_ptr(s));
            const uint32_t a_lo = a_hi + A_BYTES;
            const uint32_t b_hi = a_hi + 2 * A_BYTES;
            const uint32_t b_lo = b_hi + B_BYTES;
            // raw tiles are the hi operands: hi*hi starts when the TMA data lands, overlapping the split
            ptx::mbar_wait(&full[s], ph);
            ptx::tc_fence_after();
            const bool do_mma = !(prm.dbg & 4);
#pragma unroll
            for (int k4 = 0; k4 < TK / 8; ++k4)
              if (do_mma)
                ptx::umma_tf32(d_tmem, ptx::umma_desc_sw128(a_hi + k4 * 32), ptx::umma_desc_sw128(b_hi + k4 * 32), idesc,
                               (kc > 0 || k4 > 0) ? 1u : 0u);
            if (PASSES == 3) {
              ptx::mbar_wait(&split[s], ph);
              ptx::tc_fence_after();
#pragma unroll
              for (int k4 = 0; k4 < TK / 8; ++k4) {
                if (!do_mma) break;
                ptx::umma_tf32(d_tmem, ptx::umma_desc_sw128(a_lo + k4 * 32), ptx::umma_desc_sw128(b_hi + k4 * 32), idesc, 1u);
                ptx::umma_tf32(d_tmem, ptx::umma_desc_sw128(a_hi + k4 * 32), ptx::umma_desc_sw128(b_lo + k4 * 32), idesc, 1u);
              }
            } else if (PASSES == 2) {
              // cross terms in bf16 (K = 16 per MMA, 64-B swizzled tiles written by the splitters):
              //   Q_lo16 * T_hi16 + Q_hi16 * T_lo16        (operand error ~2^-20, below the accumulator's)
              ptx::mbar_wait(&split[s], ph);
              ptx::tc_fence_after();
              const uint32_t idesc16 = ptx::umma_idesc_bf16(TM, prm.tn);
              const uint32_t a16h = a_hi + A_BYTES, a16l = a16h + A_BYTES / 2;
              const uint32_t b16h = b_hi + B_BYTES, b16l = b16h + B_BYTES / 2;
#pragma unroll
              for (int k2 = 0; k2 < TK / 16; ++k2) {
                if (!do_mma) break;
                ptx::umma_bf16(d_tmem, ptx::umma_desc_sw64(a16l + k2 * 32), ptx::umma_desc_sw64(b16h + k2 * 32), idesc16, 1u);
                ptx::umma_bf16(d_tmem, ptx::umma_desc_sw64(a16h + k2 * 32), ptx::umma_desc_sw64(b16l + k2 * 32), idesc16, 1u);
              }
            }
            if (MC) ptx::umma_commit_mc(&empty[s], (uint16_t)0b11);   // frees the stage in BOTH CTAs
            else    ptx::umma_commit(&empty[s]);                      // smem stage free once these MMAs retire
          }
          ptx::umma_commit(&tfull[b]);             // accumulator complete
        }
      }
    }
  } else if (warp >= 12 || warp == 2 || warp == 3) {
    // ================================ splitters =============================================
    // warps 12-15 derive T_lo from the raw table tile, warps 2-3 derive Q_lo from the raw query tile
    // (16 float4 per thread each: the split phase is on the critical path with only two stages)
    if (PASSES != 1) {
      const bool is_b = warp >= 12;
      const int t = is_b ? threadIdx.x - 12 * 32 : threadIdx.x - 2 * 32;
      uint32_t c = 0;
      for (int w = wstart; w < total_work; w += wstep) {
        int qt, et0, et1, ec;
        work_range(w, qt, et0, et1, ec);
        for (int et = et0; et < et1; ++et) {
          for (int kc = 0; kc < nk; ++kc, ++c) {
            const int s = c % STAGES;
            const uint32_t ph = (c / STAGES) & 1;
            ptx::mbar_wait(&full[s], ph);
            // raw tile = hi operand; write lo next to it, for the table tile AND the query tile
            const uint32_t sp = ptx::smem_u32(stage_ptr(s));
            if (prm.dbg & 2) {
              // experiment: no split work
            } else if (PASSES == 3) {
              if (is_b) tc::split_tile<B_BYTES, SPLIT_WARPS * 32>(sp + 2 * A_BYTES, sp + 2 * A_BYTES + B_BYTES, t);
              else      tc::split_tile<A_BYTES, 2 * 32>(sp, sp + A_BYTES, t);
            } else {
              // mixed mode: the fp32 lo buffers hold two bf16 tiles (hi16 | lo16) instead
              if (is_b) tc::split_tile_bf16<TN, SPLIT_WARPS * 32>(sp + 2 * A_BYTES, sp + 2 * A_BYTES + B_BYTES,
                                                                  sp + 2 * A_BYTES + B_BYTES + B_BYTES / 2, t);
              else      tc::split_tile_bf16<TM, 2 * 32>(sp, sp + A_BYTES, sp + A_BYTES + A_BYTES / 2, t);
            }
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&split[s]);
          }
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
    for (int w = wstart; w < total_work; w += wstep) {
      int qt, et0, et1, ec;
      work_range(w, qt, et0, et1, ec);
      const int64_t row = (int64_t)qt * TM + quad * 32 + lane;   // this thread's query row
      const bool row_ok = row < prm.nq;
      RowState<EPI> st;
      st.init();
      const float aux = row_ok ? epi_row_aux<EPI>(P, row) : 0.f;
      for (int et = et0; et < et1; ++et, ++it) {
        const int b = it & 1;
        ptx::mbar_wait(&tfull[b], (it >> 1) & 1);
        ptx::tc_fence_after();
        // columns beyond this tile's tn entities were never computed: clip the valid range
        const int64_t tile_end = (int64_t)(et + 1) * prm.tn;
        tc::epilogue_tile<EPI, 4>(P, st, aux,
                                  tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(b * TN + half * 128),
                                  (int64_t)qt * TM + quad * 32, (int64_t)et * prm.tn + half * 128, prm.nq,
                                  tile_end < prm.m ? tile_end : prm.m, my_stg, lane);
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
  if (MC) ptx::cluster_sync_all(); else __syncthreads();   // MC: the peer may still signal this CTA's barriers
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<TMEM_COLS>(tmem_base);
  }}

// ---------------------------------------------------------------------------------------------
using tc::num_sms;

// mc: pairs of q tiles share table tiles (2-CTA clusters); groups = q pairs, CTAs available = sms/2 clusters
void plan(int64_t nq, int64_t m, bool mc, int& q_tiles, int& q_groups, int& e_tiles, int& echunks, int& tn) {
  q_tiles = (int)((nq + TM - 1) / TM);
  if (q_tiles < 1) q_tiles = 1;
  q_groups = mc ? (q_tiles + 1) / 2 : q_tiles;
  const int units = mc ? num_sms() / 2 : num_sms();
  // pick the tile width (multiple of 16 in [128, 256]) minimising the per-SM makespan in columns
  int64_t best_cost = -1;
  tn = TN;
  for (int cand = TN; cand >= 128; cand -= 16) {
    const int64_t et = (m + cand - 1) / cand;
    int per = units / q_groups; if (per < 1) per = 1; if (per > et) per = (int)et;
    const int64_t tiles_per_cta = (et + per - 1) / per;                   // largest e-range of a work item
    const int64_t waves = ((int64_t)q_groups * per + units - 1) / units;  // work items per CTA
    const int64_t cost = waves * tiles_per_cta * cand + tiles_per_cta * 24;   // + per-tile fixed overhead
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; tn = cand; }
  }
  e_tiles = (int)((m + tn - 1) / tn);
  int per = units / q_groups;
  if (per < 1) per = 1;
  if (per > e_tiles) per = e_tiles;
  echunks = per;
}

bool use_mc(int64_t nq) {
  // Table-tile multicast across 2-CTA clusters is parity-green and cuts L2->SM traffic by a third, but
  // measured no faster on B200 (the 1-CTA kernel is bound by shared-memory bandwidth: UMMA operand
  // reads 144 KB + TMA writes 46 KB + split 92 KB per K-chunk ~ 2200 of the 2750 cycles at 128 B/clk).
  // Off by default; B200KGE_TC_MC=1 enables it.
  const char* e = getenv("B200KGE_TC_MC");
  return e && atoi(e) != 0 && nq > TM;
}

template <int EPI, int PASSES, bool MC>
int launch_k(const CUtensorMap& a, const CUtensorMap& c, const TcParams& prm, int grid, cudaStream_t st) {
  auto kern = pairwise_tc_kernel<EPI, PASSES, MC>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(pairwise_tc_kernel)");
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = MC ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  profile_begin(st);
  e = cudaLaunchKernelEx(&cfg, kern, a, c, prm);
  profile_end(st);
  count_launch();
  if (e != cudaSuccess) return check_cuda(e, "cudaLaunchKernelEx(pairwise_tc_kernel)");
  return check_cuda(cudaGetLastError(), "pairwise_tc_kernel");
}

template <int EPI>
int launch_e(int passes, bool mc, const CUtensorMap& a, const CUtensorMap& c, const TcParams& prm, int grid,
             cudaStream_t st) {
  if (passes == 3) return mc ? launch_k<EPI, 3, true>(a, c, prm, grid, st) : launch_k<EPI, 3, false>(a, c, prm, grid, st);
  if (passes == 2) return mc ? launch_k<EPI, 2, true>(a, c, prm, grid, st) : launch_k<EPI, 2, false>(a, c, prm, grid, st);
  return mc ? launch_k<EPI, 1, true>(a, c, prm, grid, st) : launch_k<EPI, 1, false>(a, c, prm, grid, st);
}

}  // namespace

bool tc_supported(int pair_op, int K, const Rows& cand, int col_off) {
  if (pair_op != PAIR_DOT) return false;
  if (K < TK) return false;
  if (cand.ld % 4 != 0 || col_off % 4 != 0) return false;
  if ((reinterpret_cast<uintptr_t>(cand.base) & 15) != 0) return false;
  if (cand.rows >= (1ll << 31)) return false;
  return true;
}

int tc_nchunks(int64_t nq, int64_t m) {
  int qt, qg, et, ec, tn;
  plan(nq, m, use_mc(nq), qt, qg, et, ec, tn);
  return 2 * ec;
}

int launch_pairwise_tc(int epi_kind, int passes, const float* Q, int64_t ldq,
                       int64_t nq, const float* T, int64_t ldt, int64_t m, int K,
                       const EpiParams& P, cudaStream_t st) {
  if (nq == 0 || m == 0) return 0;
  const bool mc = use_mc(nq);
  CUtensorMap mQ, mT;
  int rc;
  if ((rc = tc::make_map(&mQ, Q, nq, K, ldq, TK, TM))) return rc;
  TcParams prm;
  prm.nq = nq; prm.m = m; prm.K = K;
  plan(nq, m, mc, prm.q_tiles, prm.q_groups, prm.e_tiles, prm.echunks, prm.tn);
  if ((rc = tc::make_map(&mT, T, m, K, ldt, TK, mc ? prm.tn / 2 : prm.tn))) return rc;
  prm.epi = P;
  prm.epi.nchunks = 2 * prm.echunks;   // two epilogue warps (column halves) per row
  { const char* e = getenv("B200KGE_DBG"); prm.dbg = e ? atoi(e) : 0; }
  const int total = prm.q_groups * prm.echunks;
  const int units = mc ? num_sms() / 2 : num_sms();
  const int grid = (mc ? 2 : 1) * (total < units ? total : units);
  switch (epi_kind) {
    case EPI_STORE: return launch_e<EPI_STORE>(passes, mc, mQ, mT, prm, grid, st);
    case EPI_BCE:   return launch_e<EPI_BCE>(passes, mc, mQ, mT, prm, grid, st);
    case EPI_KL:    return launch_e<EPI_KL>(passes, mc, mQ, mT, prm, grid, st);
    case EPI_RANK:  return launch_e<EPI_RANK>(passes, mc, mQ, mT, prm, grid, st);
  }
  set_error("bad epilogue kind %d", epi_kind);
  return B200KGE_ERR_INVALID;
}

}  // namespace b200kge
