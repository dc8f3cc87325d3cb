// pairwise_tc.cu — tcgen05 tensor-core 1-vs-N scorer for the dot-product family
// (ComplEx / DistMult / SimplE / CP / RESCAL after folding), fp32-equivalent via an on-chip operand split.
//
//   S[q, e] = sum_k Q[q,k] * T[e,k]          Q: folded queries [nq, K]  (fold.cu, RAW fp32)
//                                            T: entity table   [m,  K]  streamed RAW from HBM/L2
//
// replaces the reference's torch.mm over concatenated operands (complex.py:37,39,
// distmult.py:19,21, simple.py:25-29, cp.py:24,26, rescal.py:41,47) AND whatever consumes the
// scores next (BCE / KL loss, rank counting, or the plain [n,E] store) in ONE kernel.
//
// Precision (the reference is a true fp32 GEMM; single-pass TF32 misses the 1e-4 bar 40x):
//   mixed (default)  D += Q*T [tf32, raw tiles: kind::tf32 truncates the low 13 mantissa bits, measured]
//                         + Q_lo16*T_hi16 + Q_hi16*T_lo16   [bf16, kind::f16]        2.4e-5 of rms
//   3xTF32           D += Q*T + Q_lo*T + Q*T_lo              [tf32]                   3.0e-5 of rms
// Nothing derivable on chip crosses the L2->SM fabric: TMA lands RAW fp32 tiles, splitter warps derive
// the lo / bf16 operand tiles in shared memory (fence.proxy.async before the MMA warp may read them).
//
// Pipeline (measured: with coupled 2x96 KB stages the three phases TMA 0.034 / split 0.041 / MMA 0.046 ms
// overlapped poorly, 0.135 ms total).  Raw tiles and derived operand tiles are therefore DECOUPLED rings:
//   raw[2]  48 KB each: Q tile 16 KB | T tile 32 KB            filled by TMA
//   op[2]   48 KB each: derived Q operands 16 KB | derived T operands 32 KB   written by the splitters
//   raw_full[r]  TMA bytes landed                                  -> splitters, MMA (hi*hi)
//   raw_free[r]  hi*hi MMAs retired (tcgen05.commit) + 6 splitter warps done reading -> TMA producer
//   op_full[o]   6 splitter warps wrote + fenced the derived tiles -> MMA (cross terms)
//   op_free[o]   cross-term MMAs retired (tcgen05.commit)          -> splitters
// so the producer refills a raw buffer half a chunk earlier and the split of chunk c+1 overlaps the
// cross-term MMAs of chunk c.
//
// CTA = 16 warps, one CTA per SM, persistent over (query tile, range of entity tiles):
//   warp 0      TMA producer (one lane)          warp 1       MMA issuer (one lane), TMEM alloc
//   warps 2-3   query-tile splitters             warps 12-15  table-tile splitters
//   warps 4-11  epilogue: tcgen05.ld -> regs -> {transposed coalesced store | BCE | KL | rank}
//               (2 warps per TMEM lane quadrant, each takes 128 of the 256 accumulator columns)
// Tile = 128 queries (UMMA M, TMEM lanes) x <=256 entities (UMMA N, TMEM columns), K in chunks of 32
// floats (one 128-byte swizzle atom), 2 TMEM accumulators of 256 columns so the epilogue of tile i
// overlaps the MMAs of tile i+1.  TMEM lane = query row, so every per-row reduction is thread-local.
#include "tc_common.cuh"

namespace b200kge {

namespace {

constexpr int TM = 128;             // queries per tile  (UMMA M)
constexpr int TN = 256;             // max entities per tile (UMMA N)
constexpr int TK = 32;              // floats per K chunk (128 B swizzle atom)
constexpr int NR = 2, NO = 2;       // raw / operand ring depths
constexpr int A_BYTES = TM * TK * 4;   // 16 KB
constexpr int B_BYTES = TN * TK * 4;   // 32 KB
constexpr int RING_BYTES = A_BYTES + B_BYTES;   // one raw or one operand buffer: 48 KB
using tc::EPI_WARPS;
using tc::SPLIT_WARPS;
using tc::NTHREADS;
using tc::STG_LD;
constexpr int NSPLIT = SPLIT_WARPS + 2;   // 4 table-tile + 2 query-tile splitter warps
constexpr int STG_BYTES = EPI_WARPS * 32 * STG_LD * 4;
constexpr int SMEM_BYTES = 1024 /*align slack*/ + (NR + NO) * RING_BYTES + STG_BYTES + 256 /*barriers*/;
constexpr int TMEM_COLS = 512;

struct TcParams {
  int64_t nq, m;
  int K;            // reduction length (floats)
  int q_tiles, e_tiles, echunks;
  int dbg;          // experiments (B200KGE_DBG bit-mask): 1 = skip TMA after the first fills, 2 = skip split
                    // math, 4 = skip MMAs — isolates the pipeline phases (results are garbage)
  int tn;           // entities per tile actually used (multiple of 16, <= TN): chosen per problem so that
                    // ceil(tiles / SMs) * tn — the makespan in columns — is minimal
  EpiParams epi;
};

// PASSES: 1 = single-pass tf32 (experiments), 2 = mixed tf32 + bf16 cross terms, 3 = 3xTF32
template <int EPI, int PASSES>
__global__ void __launch_bounds__(NTHREADS, 1)
pairwise_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmT,
                   const TcParams prm) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* raw_base = smem;
  uint8_t* op_base = smem + NR * RING_BYTES;
  float* stg = reinterpret_cast<float*>(smem + (NR + NO) * RING_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (NR + NO) * RING_BYTES + STG_BYTES);
  uint64_t* raw_full = bars;                    // [NR]
  uint64_t* raw_free = bars + NR;               // [NR]
  uint64_t* op_full = bars + 2 * NR;            // [NO]
  uint64_t* op_free = bars + 2 * NR + NO;       // [NO]
  uint64_t* tfull = bars + 2 * NR + 2 * NO;     // [2]
  uint64_t* tempty = tfull + 2;                 // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk = (prm.K + TK - 1) / TK;
  const int total_work = prm.q_tiles * prm.echunks;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmQ);
    ptx::prefetch_tensormap(&tmT);
    for (int r = 0; r < NR; ++r) {
      ptx::mbar_init(&raw_full[r], 1);
      ptx::mbar_init(&raw_free[r], PASSES == 1 ? 1 : 1 + NSPLIT);   // MMA commit (+ splitter warps)
    }
    for (int o = 0; o < NO; ++o) {
      ptx::mbar_init(&op_full[o], NSPLIT);
      ptx::mbar_init(&op_free[o], 1);
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

  // e-tile range of work item w
  auto work_range = [&](int w, int& qt, int& et0, int& et1, int& ec) {
    qt = w / prm.echunks;
    ec = w - qt * prm.echunks;
    const int base = prm.e_tiles / prm.echunks, rem = prm.e_tiles % prm.echunks;
    et0 = ec * base + (ec < rem ? ec : rem);
    et1 = et0 + base + (ec < rem ? 1 : 0);
  };

  if (warp == 0) {
    // ================================ TMA producer =========================================
    if (lane == 0) {
      uint32_t c = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        int qt, et0, et1, ec;
        work_range(w, qt, et0, et1, ec);
        for (int et = et0; et < et1; ++et) {
          for (int kc = 0; kc < nk; ++kc, ++c) {
            const int r = c % NR;
            ptx::mbar_wait(&raw_free[r], ((c / NR) & 1) ^ 1);
            if ((prm.dbg & 1) && c >= (uint32_t)NR) { ptx::mbar_arrive(&raw_full[r]); continue; }
            uint8_t* rp = raw_base + r * RING_BYTES;
            ptx::mbar_arrive_expect_tx(&raw_full[r], A_BYTES + prm.tn * TK * 4);
            ptx::tma_load_2d(rp, &tmQ, &raw_full[r], kc * TK, qt * TM);                  // raw queries
            ptx::tma_load_2d(rp + A_BYTES, &tmT, &raw_full[r], kc * TK, et * prm.tn);    // raw table tile
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ============================================
    if (lane == 0) {
      const uint32_t idesc = ptx::umma_idesc_tf32(TM, prm.tn);
      const uint32_t idesc16 = ptx::umma_idesc_bf16(TM, prm.tn);
      const bool do_mma = !(prm.dbg & 4);
      uint32_t c = 0, it = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        int qt, et0, et1, ec;
        work_range(w, qt, et0, et1, ec);
        for (int et = et0; et < et1; ++et, ++it) {
          const int b = it & 1;
          ptx::mbar_wait(&tempty[b], ((it >> 1) & 1) ^ 1);   // epilogue drained this accumulator
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(b * TN);
          for (int kc = 0; kc < nk; ++kc, ++c) {
            const int r = c % NR, o = c % NO;
            const uint32_t a_raw = ptx::smem_u32(raw_base + r * RING_BYTES);
            const uint32_t b_raw = a_raw + A_BYTES;
            const uint32_t a_op = ptx::smem_u32(op_base + o * RING_BYTES);
            const uint32_t b_op = a_op + A_BYTES;
            // hi*hi on the raw tiles as soon as they land
            ptx::mbar_wait(&raw_full[r], (c / NR) & 1);
            ptx::tc_fence_after();
#pragma unroll
            for (int k4 = 0; k4 < TK / 8; ++k4)
              if (do_mma)
                ptx::umma_tf32(d_tmem, ptx::umma_desc_sw128(a_raw + k4 * 32), ptx::umma_desc_sw128(b_raw + k4 * 32),
                               idesc, (kc > 0 || k4 > 0) ? 1u : 0u);
            if (PASSES == 1) { ptx::umma_commit(&raw_free[r]); continue; }
            // cross terms on the derived tiles
            ptx::mbar_wait(&op_full[o], (c / NO) & 1);
            ptx::tc_fence_after();
            if (PASSES == 3) {
              // Q_lo * T_raw needs the raw table tile as well: it is issued before raw_free is committed
#pragma unroll
              for (int k4 = 0; k4 < TK / 8; ++k4)
                if (do_mma)
                  ptx::umma_tf32(d_tmem, ptx::umma_desc_sw128(a_op + k4 * 32), ptx::umma_desc_sw128(b_raw + k4 * 32), idesc, 1u);
#pragma unroll
              for (int k4 = 0; k4 < TK / 8; ++k4)
                if (do_mma)
                  ptx::umma_tf32(d_tmem, ptx::umma_desc_sw128(a_raw + k4 * 32), ptx::umma_desc_sw128(b_op + k4 * 32), idesc, 1u);
              ptx::umma_commit(&raw_free[r]);
            } else {
              ptx::umma_commit(&raw_free[r]);        // raw tiles are only read by the hi*hi MMAs above
              // op buffers hold bf16 tiles (64-B swizzle): [hi16 | lo16] for Q (8 KB each) and T (16 KB each)
              const uint32_t a16h = a_op, a16l = a_op + A_BYTES / 2;
              const uint32_t b16h = b_op, b16l = b_op + B_BYTES / 2;
#pragma unroll
              for (int k2 = 0; k2 < TK / 16; ++k2) {
                if (!do_mma) break;
                ptx::umma_bf16(d_tmem, ptx::umma_desc_sw64(a16l + k2 * 32), ptx::umma_desc_sw64(b16h + k2 * 32), idesc16, 1u);
                ptx::umma_bf16(d_tmem, ptx::umma_desc_sw64(a16h + k2 * 32), ptx::umma_desc_sw64(b16l + k2 * 32), idesc16, 1u);
              }
            }
            ptx::umma_commit(&op_free[o]);
          }
          ptx::umma_commit(&tfull[b]);             // accumulator complete
        }
      }
    }
  } else if (warp >= 12 || warp == 2 || warp == 3) {
    // ================================ splitters =============================================
    // warps 12-15 derive the table-tile operands, warps 2-3 the query-tile operands
    if (PASSES != 1) {
      const bool is_b = warp >= 12;
      const int t = is_b ? threadIdx.x - 12 * 32 : threadIdx.x - 2 * 32;
      uint32_t c = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        int qt, et0, et1, ec;
        work_range(w, qt, et0, et1, ec);
        for (int et = et0; et < et1; ++et) {
          for (int kc = 0; kc < nk; ++kc, ++c) {
            const int r = c % NR, o = c % NO;
            ptx::mbar_wait(&raw_full[r], (c / NR) & 1);
            ptx::mbar_wait(&op_free[o], ((c / NO) & 1) ^ 1);
            const uint32_t rp = ptx::smem_u32(raw_base + r * RING_BYTES);
            const uint32_t opp = ptx::smem_u32(op_base + o * RING_BYTES);
            if (prm.dbg & 2) {
              // experiment: no split work
            } else if (PASSES == 3) {
              if (is_b) tc::split_tile<B_BYTES, SPLIT_WARPS * 32>(rp + A_BYTES, opp + A_BYTES, t);
              else      tc::split_tile<A_BYTES, 2 * 32>(rp, opp, t);
            } else {
              if (is_b) tc::split_tile_bf16<TN, SPLIT_WARPS * 32>(rp + A_BYTES, opp + A_BYTES, opp + A_BYTES + B_BYTES / 2, t);
              else      tc::split_tile_bf16<TM, 2 * 32>(rp, opp, opp + A_BYTES / 2, t);
            }
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              ptx::mbar_arrive(&op_full[o]);
              ptx::mbar_arrive(&raw_free[r]);      // this warp is done reading the raw tile
            }
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
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
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
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
using tc::num_sms;

void plan(int64_t nq, int64_t m, int& q_tiles, int& e_tiles, int& echunks, int& tn) {
  q_tiles = (int)((nq + TM - 1) / TM);
  if (q_tiles < 1) q_tiles = 1;
  const int units = num_sms();
  // pick the tile width (multiple of 16 in [128, 256]) minimising the per-SM makespan in columns
  int64_t best_cost = -1;
  tn = TN;
  for (int cand = TN; cand >= 128; cand -= 16) {
    const int64_t et = (m + cand - 1) / cand;
    int per = units / q_tiles; if (per < 1) per = 1; if (per > et) per = (int)et;
    const int64_t tiles_per_cta = (et + per - 1) / per;                  // largest e-range of a work item
    const int64_t waves = ((int64_t)q_tiles * per + units - 1) / units;  // work items per CTA
    const int64_t cost = waves * tiles_per_cta * cand + tiles_per_cta * 24;   // + per-tile fixed overhead
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; tn = cand; }
  }
  e_tiles = (int)((m + tn - 1) / tn);
  int per = units / q_tiles;
  if (per < 1) per = 1;
  if (per > e_tiles) per = e_tiles;
  echunks = per;
}

template <int EPI, int PASSES>
int launch_k(const CUtensorMap& a, const CUtensorMap& c, const TcParams& prm, int grid, cudaStream_t st) {
  auto kern = pairwise_tc_kernel<EPI, PASSES>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(pairwise_tc_kernel)");
  profile_begin(st);
  kern<<<grid, NTHREADS, SMEM_BYTES, st>>>(a, c, prm);
  profile_end(st);
  B2K_LAUNCH_CHECK("pairwise_tc_kernel");
  return 0;
}

template <int EPI>
int launch_e(int passes, const CUtensorMap& a, const CUtensorMap& c, const TcParams& prm, int grid, cudaStream_t st) {
  if (passes == 3) return launch_k<EPI, 3>(a, c, prm, grid, st);
  if (passes == 2) return launch_k<EPI, 2>(a, c, prm, grid, st);
  return launch_k<EPI, 1>(a, c, prm, grid, st);
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
  int qt, et, ec, tn;
  plan(nq, m, qt, et, ec, tn);
  return 2 * ec;
}

int launch_pairwise_tc(int epi_kind, int passes, const float* Q, int64_t ldq,
                       int64_t nq, const float* T, int64_t ldt, int64_t m, int K,
                       const EpiParams& P, cudaStream_t st) {
  if (nq == 0 || m == 0) return 0;
  CUtensorMap mQ, mT;
  int rc;
  if ((rc = tc::make_map(&mQ, Q, nq, K, ldq, TK, TM))) return rc;
  TcParams prm;
  prm.nq = nq; prm.m = m; prm.K = K;
  plan(nq, m, prm.q_tiles, prm.e_tiles, prm.echunks, prm.tn);
  if ((rc = tc::make_map(&mT, T, m, K, ldt, TK, prm.tn))) return rc;
  prm.epi = P;
  prm.epi.nchunks = 2 * prm.echunks;   // two epilogue warps (column halves) per row
  { const char* e = getenv("B200KGE_DBG"); prm.dbg = e ? atoi(e) : 0; }
  const int total = prm.q_tiles * prm.echunks;
  const int grid = total < num_sms() ? total : num_sms();
  switch (epi_kind) {
    case EPI_STORE: return launch_e<EPI_STORE>(passes, mQ, mT, prm, grid, st);
    case EPI_BCE:   return launch_e<EPI_BCE>(passes, mQ, mT, prm, grid, st);
    case EPI_KL:    return launch_e<EPI_KL>(passes, mQ, mT, prm, grid, st);
    case EPI_RANK:  return launch_e<EPI_RANK>(passes, mQ, mT, prm, grid, st);
  }
  set_error("bad epilogue kind %d", epi_kind);
  return B200KGE_ERR_INVALID;
}

}  // namespace b200kge
