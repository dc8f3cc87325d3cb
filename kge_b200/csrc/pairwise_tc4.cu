// pairwise_tc4.cu — CTA-pair (tcgen05 cta_group::2) scorer on PRE-SPLIT fp16 operand planes, selected with
// B200KGE_TC_VERSION=4 (the default is the 1-CTA pairwise_tc3.cu); same scores, losses and rank counts; parity-green
// on B200 (tests/test_gpu_presplit.py), 2/3 of the L2->SM operand traffic, within +-3 % of the 1-CTA kernel's time.
//
// pairwise_tc3.cu (1 CTA, pre-split planes) has no shared-memory round trip left, which leaves two limits:
// the L2->SM feed (each 128 x 256 tile pulls 384 operand rows) and the MMAs' own operand reads (96 of the
// SM's 128 B/clk at M=128 x N=256).  A CTA pair on the two SMs of a TPC relaxes both: cluster tile = 256
// queries x 256 entities, UMMA M=256 / N=256; each CTA stages its own 128 query rows and HALF of the entity
// tile (128 rows), i.e. 256 operand rows per 128 x 256 outputs (2/3 of the feed), and each SM's tensor core
// reads 8 KB instead of 12 KB per instruction.  pairwise_tc2.cu is the same idea on raw fp32 tiles with an
// in-kernel split; it measured slower than the 1-CTA kernel because its splitters added a cross-CTA
// "split done" handshake per stage on top of the landing handshake (profiles/r1_notes.md).  Here nothing is
// derived on chip, so one handshake per slot remains and the ring is 6 slots deep (3 K chunks of 64).
//
//   slot = 32 KB = (query plane box 128 x 64 halfs | table-half plane box 128 x 64 halfs);  K chunk c uses
//   slots 2c%6 (hi planes) and 2c%6+1 (lo planes).
//   FORWARD signalling (template DIRECT=false; the mechanics validated in pairwise_tc2.cu):
//     full[s]    local  : this CTA's TMA bytes of slot s landed            (count 1 + tx)
//     landed[s]  leader : both CTAs' slot s landed (a forwarder lane in each CTA waits full[s] and arrives
//                         remotely)                                         (count 2)
//   DIRECT signalling (DIRECT=true, B200KGE_TC4_DIRECT=1): both CTAs' TMA loads complete on the LEADER's
//     full[s] (shared::cluster mbarrier operand of cp.async.bulk.tensor), the leader's producer expects the
//     bytes of both; no forwarding hop.
//   empty[s]   both   : MMAs reading slot s retired (multicast tcgen05.commit)
//   tfull[b]   both   : accumulator b complete      (multicast tcgen05.commit)
//   tempty[b]  leader : both CTAs' epilogues drained accumulator b          (count 16: one per warp)
// CTA = 12 warps: warp 0 TMA producer | warp 1 MMA issuer (leader only) + TMEM alloc | warp 2 forwarder |
// warp 3 idle | warps 4-11 epilogue (shared verbatim with the other tcgen05 kernels).
#include "tc_common.cuh"

namespace b200kge {

namespace {

constexpr int TM = 128;     // query rows per CTA (cluster: 256)
constexpr int TN = 256;
constexpr int TKH = 64;     // halfs per K chunk (128-byte swizzle atom)
constexpr int NSLOT = 6;
constexpr int A_BYTES = TM * TKH * 2;     // 16 KB
constexpr int SLOT_BYTES = 2 * A_BYTES;   // 32 KB
using tc::EPI_WARPS;
constexpr int NTHREADS4 = 12 * 32;
constexpr int STG_BYTES = EPI_WARPS * 32 * tc::STG_LD * 4;
constexpr int NBARS = 3 * NSLOT + 4;
constexpr int SMEM_BYTES = 1024 + NSLOT * SLOT_BYTES + STG_BYTES + NBARS * 8 + 64;
constexpr int TMEM_COLS = 512;

struct Tc4Params {
  int64_t nq, m;
  int nk;                          // K chunks of 64 halfs
  int last_ksteps;                 // 16-element MMA steps of the last chunk that hold real data (1..4)
  int q_tiles, echunks;            // q tiles of 256 rows; each q tile's columns are cut into `echunks` ranges
  int cols_per;                    // columns per range (multiple of 32): a range = full 256-column tiles + ONE narrower
                                   // last tile (width multiple of 32) — no quantisation to whole 256-column tiles
  const float* q_scale;            // [nq]
  const float* t_scale;            // [m + 32], zero beyond m
  int dbg;                         // B200KGE_DBG ablations (measurement only; results are garbage): 1 = epilogue releases
                                   // the accumulator without reading it, 2 = no MMAs are issued, 4 = no TMA loads
  EpiParams epi;
};

template <int EPI, bool DIRECT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS4, 1)
pairwise_tc4_kernel(const __grid_constant__ CUtensorMap tmQh, const __grid_constant__ CUtensorMap tmQl,
                    const __grid_constant__ CUtensorMap tmTh, const __grid_constant__ CUtensorMap tmTl,
                    const Tc4Params prm) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* stg = reinterpret_cast<float*>(smem + NSLOT * SLOT_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NSLOT * SLOT_BYTES + STG_BYTES);
  uint64_t* full = bars;                    // [NSLOT]
  uint64_t* landed = bars + NSLOT;          // [NSLOT] (leader's copy is the one used)
  uint64_t* empty = bars + 2 * NSLOT;       // [NSLOT]
  uint64_t* tfull = bars + 3 * NSLOT;       // [2]
  uint64_t* tempty = tfull + 2;             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  const int nk = prm.nk;
  const int total_work = prm.q_tiles * prm.echunks;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmQh);
    ptx::prefetch_tensormap(&tmQl);
    ptx::prefetch_tensormap(&tmTh);
    ptx::prefetch_tensormap(&tmTl);
    for (int s = 0; s < NSLOT; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&landed[s], 2);
      ptx::mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&tfull[b], 1);
      ptx::mbar_init(&tempty[b], 2 * EPI_WARPS);     // one arrive per epilogue warp of both CTAs
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc_2cta<TMEM_COLS>(tmem_slot);
  ptx::tc_fence_before();
  ptx::cluster_sync_all();          // barriers of BOTH CTAs initialised before any remote arrive
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // work item w -> query tile qt, column range [c_lo, c_hi) (ec-th range); tile t of the range starts at
  // c_lo + 256 t and is tile_n(c_lo, c_hi, t) columns wide
  auto work_range = [&](int w, int& qt, int64_t& c_lo, int64_t& c_hi, int& ec, int& ntiles) {
    qt = w / prm.echunks;
    ec = w - qt * prm.echunks;
    c_lo = (int64_t)ec * prm.cols_per;
    c_hi = c_lo + prm.cols_per < prm.m ? c_lo + prm.cols_per : prm.m;
    ntiles = c_hi > c_lo ? (int)((c_hi - c_lo + TN - 1) / TN) : 0;
  };
  auto tile_n = [](int64_t c_lo, int64_t c_hi, int t) {      // UMMA N of tile t: 256, or the remainder rounded up to 32
    const int64_t left = c_hi - (c_lo + (int64_t)t * TN);
    return (int)(left >= TN ? TN : ((left + 31) / 32) * 32);
  };
  // slot pair and phase of K chunk c: slots 2c%6, 2c%6+1; both on their (c/3)-th use
  auto slot_of = [](uint32_t c) { return (int)((2 * c) % NSLOT); };
  auto phase_of = [](uint32_t c) { return (c / (NSLOT / 2)) & 1u; };

  if (warp == 0) {
    // ================================ TMA producer (both CTAs) ==============================
    // The WHOLE warp walks the loops (uniform control flow, loop state in uniform registers) and one elected lane
    // issues: with the loops inside `if (lane == 0)` every operand of UTMALDG / UTCHMMA / UTCBAR lives in per-thread
    // registers and ptxas wraps each instruction in an ELECT / R2UR.BROADCAST x7 / BRA.U.ANY "waterfall" (~24 SASS
    // instructions per MMA): the issuing thread, not the tensor pipe, then paces the kernel.
    {
      const bool issuer = ptx::elect_one();
      uint32_t c = 0;
      for (int w = cluster_id; w < total_work; w += nclusters) {
        int qt, ec, ntiles;
        int64_t c_lo, c_hi;
        work_range(w, qt, c_lo, c_hi, ec, ntiles);
        const int q_row = qt * 256 + (int)rank * TM;
        for (int t = 0; t < ntiles; ++t) {
          // this CTA stages rows [rank * N/2, (rank + 1) * N/2) of the tile; the box always has TM rows (rows beyond
          // N/2 land in the slot unused, rows beyond the table are zero-filled)
          const int e_row = (int)(c_lo + (int64_t)t * TN) + (int)rank * (tile_n(c_lo, c_hi, t) >> 1);
          for (int kc = 0; kc < nk; ++kc, ++c) {
            const int s0 = slot_of(c);
            const uint32_t ph = phase_of(c);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int s = s0 + h;
              ptx::mbar_wait_cluster_bounded(&empty[s], ph ^ 1);
              uint8_t* sp = smem + s * SLOT_BYTES;
              // completion of BOTH CTAs' boxes is counted on the leader's barrier
              const uint32_t bar = ptx::mapa(ptx::smem_u32(&full[s]), 0);
              if (issuer) {
                if (prm.dbg & 4) {
                  if (rank == 0) ptx::mbar_arrive(&full[s]);
                } else {
                  if (rank == 0) ptx::mbar_arrive_expect_tx(&full[s], 2 * SLOT_BYTES);
                  ptx::tma_load_2d_cluster_bar(sp, h ? &tmQl : &tmQh, bar, kc * TKH, q_row);
                  ptx::tma_load_2d_cluster_bar(sp + A_BYTES, h ? &tmTl : &tmTh, bar, kc * TKH, e_row);
                }
              }
              __syncwarp();
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (leader CTA only) ===========================
    if (rank == 0) {
      const bool issuer = ptx::elect_one();
      uint32_t c = 0, it = 0;
      for (int w = cluster_id; w < total_work; w += nclusters) {
        int qt, ec, ntiles;
        int64_t c_lo, c_hi;
        work_range(w, qt, c_lo, c_hi, ec, ntiles);
        for (int t = 0; t < ntiles; ++t, ++it) {
          const uint32_t idesc = ptx::umma_idesc_f16(256, tile_n(c_lo, c_hi, t));
          const int b = it & 1;
          ptx::mbar_wait_cluster_bounded(&tempty[b], ((it >> 1) & 1) ^ 1);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(b * TN);
          for (int kc = 0; kc < nk; ++kc, ++c) {
            const int sh = slot_of(c), sl = sh + 1;
            const uint32_t ph = phase_of(c);
            const uint32_t a_hi = ptx::smem_u32(smem + sh * SLOT_BYTES), b_hi = a_hi + A_BYTES;
            const uint32_t a_lo = ptx::smem_u32(smem + sl * SLOT_BYTES), b_lo = a_lo + A_BYTES;
            ptx::mbar_wait_cluster_bounded(&full[sh], ph);
            ptx::tc_fence_after();
            const bool tail = (kc == nk - 1) && prm.last_ksteps < TKH / 16;      // zero-padded steps are not multiplied
            if (prm.dbg & 2) {
              ptx::mbar_wait_cluster_bounded(&full[sl], ph);
            } else if (!tail) {
              if (issuer) {
#pragma unroll
                for (int k = 0; k < TKH / 16; ++k)
                  ptx::umma_bf16_2cta(d_tmem, ptx::umma_desc_sw128(a_hi + k * 32), ptx::umma_desc_sw128(b_hi + k * 32), idesc,
                                      (kc > 0 || k > 0) ? 1u : 0u);
              }
              ptx::mbar_wait_cluster_bounded(&full[sl], ph);
              ptx::tc_fence_after();
              if (issuer) {
#pragma unroll
                for (int k = 0; k < TKH / 16; ++k) {
                  ptx::umma_bf16_2cta(d_tmem, ptx::umma_desc_sw128(a_hi + k * 32), ptx::umma_desc_sw128(b_lo + k * 32), idesc, 1u);
                  ptx::umma_bf16_2cta(d_tmem, ptx::umma_desc_sw128(a_lo + k * 32), ptx::umma_desc_sw128(b_hi + k * 32), idesc, 1u);
                }
              }
            } else {
              if (issuer) {
                for (int k = 0; k < prm.last_ksteps; ++k)
                  ptx::umma_bf16_2cta(d_tmem, ptx::umma_desc_sw128(a_hi + k * 32), ptx::umma_desc_sw128(b_hi + k * 32), idesc,
                                      (kc > 0 || k > 0) ? 1u : 0u);
              }
              ptx::mbar_wait_cluster_bounded(&full[sl], ph);
              ptx::tc_fence_after();
              if (issuer) {
                for (int k = 0; k < prm.last_ksteps; ++k) {
                  ptx::umma_bf16_2cta(d_tmem, ptx::umma_desc_sw128(a_hi + k * 32), ptx::umma_desc_sw128(b_lo + k * 32), idesc, 1u);
                  ptx::umma_bf16_2cta(d_tmem, ptx::umma_desc_sw128(a_lo + k * 32), ptx::umma_desc_sw128(b_hi + k * 32), idesc, 1u);
                }
              }
            }
            if (issuer) {
              ptx::umma_commit_2cta(&empty[sh], 0b11);
              ptx::umma_commit_2cta(&empty[sl], 0b11);
            }
            __syncwarp();
          }
          if (issuer) ptx::umma_commit_2cta(&tfull[b], 0b11);
          __syncwarp();
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
      int qt, ec, ntiles;
      int64_t c_lo, c_hi;
      work_range(w, qt, c_lo, c_hi, ec, ntiles);
      const int64_t row0 = (int64_t)qt * 256 + (int64_t)rank * TM + quad * 32;
      const int64_t row = row0 + lane;
      const bool row_ok = row < prm.nq;
      RowState<EPI> st;
      st.init();
      const float aux = row_ok ? epi_row_aux<EPI>(P, row) : 0.f;
      const float qs = row_ok ? __ldg(prm.q_scale + row) : 0.f;
      const int64_t csr_end = (P.csr_off && row_ok) ? __ldg(P.csr_off + row + 1) : 0;
      for (int t = 0; t < ntiles; ++t, ++it) {
        const int b = it & 1;
        ptx::mbar_wait_cluster_bounded(&tfull[b], (it >> 1) & 1);
        ptx::tc_fence_after();
        const int64_t tile_lo = c_lo + (int64_t)t * TN;
        const int64_t tile_end = tile_lo + TN < c_hi ? tile_lo + TN : c_hi;     // valid columns only (<= m)
        if (!(prm.dbg & 1))
        tc::epilogue_tile<EPI, 4, true>(P, st, aux,
                                        tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(b * TN + half * 128),
                                        row0, tile_lo + half * 128, prm.nq, tile_end, my_stg, lane, qs, prm.t_scale,
                                        csr_end);
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
  }
}

// CTA pairs as the scheduling unit: every query tile's columns are cut into `echunks` equal ranges (multiples of 32
// columns) so that q_tiles * echunks ~ #clusters; a range runs as full 256-column tiles plus one narrower last tile.
// (Whole-tile ranges lose 10 % at the FB15k-237 shape: 7 x 256 columns per cluster for 14 541 / 9 = 1616 needed.)
void plan4(int64_t nq, int64_t m, int& q_tiles, int& echunks, int& cols_per) {
  q_tiles = (int)((nq + 255) / 256);
  if (q_tiles < 1) q_tiles = 1;
  const int units = tc::num_sms() / 2;
  int per = units / q_tiles;
  if (per < 1) per = 1;
  const int64_t max_ranges = (m + 127) / 128;          // at least 128 columns per range
  if (per > max_ranges) per = (int)max_ranges;
  cols_per = (int)(((m + per - 1) / per + 31) / 32 * 32);
  echunks = (int)((m + cols_per - 1) / cols_per);
}

template <int EPI, bool DIRECT>
int launch_k4(const CUtensorMap& qh, const CUtensorMap& ql, const CUtensorMap& th, const CUtensorMap& tl,
              const Tc4Params& prm, int grid, cudaStream_t st) {
  auto kern = pairwise_tc4_kernel<EPI, DIRECT>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(pairwise_tc4_kernel)");
  profile_begin(st);
  kern<<<grid, NTHREADS4, SMEM_BYTES, st>>>(qh, ql, th, tl, prm);
  profile_end(st);
  B2K_LAUNCH_CHECK("pairwise_tc4_kernel");
  return 0;
}

// DIRECT signalling only: the forwarding variant (local barrier + forwarder lane) measured 85.9 us against 68.4 us
// at the FB15k-237 shape (profiles/r2_summary.md) and is no longer instantiated.
template <int EPI>
int launch_e4(bool, const CUtensorMap& qh, const CUtensorMap& ql, const CUtensorMap& th, const CUtensorMap& tl,
              const Tc4Params& prm, int grid, cudaStream_t st) {
  return launch_k4<EPI, true>(qh, ql, th, tl, prm, grid, st);
}

}  // namespace

int tc4_nchunks(int64_t nq, int64_t m) {
  int qt, ec, cp;
  plan4(nq, m, qt, ec, cp);
  return 2 * ec;
}

int launch_pairwise_tc4(int epi_kind, const SplitSet& Q, const SplitSet& T, const EpiParams& P, cudaStream_t st) {
  const int64_t nq = Q.rows, m = T.rows;
  if (nq == 0 || m == 0) return 0;
  if (Q.Kp != T.Kp || Q.Kp % TKH != 0) { set_error("operand planes disagree on the padded reduction length"); return B200KGE_ERR_INVALID; }
  Tc4Params prm;
  prm.nq = nq; prm.m = m; prm.nk = Q.Kp / TKH;
  plan4(nq, m, prm.q_tiles, prm.echunks, prm.cols_per);
  { const int rem = Q.K - (prm.nk - 1) * TKH;
    prm.last_ksteps = rem <= 0 ? 4 : (rem + 15) / 16; if (prm.last_ksteps > 4) prm.last_ksteps = 4; }
  prm.q_scale = Q.inv_scale; prm.t_scale = T.inv_scale;
  CUtensorMap mQh, mQl, mTh, mTl;
  int rc;
  if ((rc = tc::make_map_f16(&mQh, Q.hi, nq, Q.Kp, Q.Kp, TM))) return rc;
  if ((rc = tc::make_map_f16(&mQl, Q.lo, nq, Q.Kp, Q.Kp, TM))) return rc;
  if ((rc = tc::make_map_f16(&mTh, T.hi, m, T.Kp, T.Kp, TM))) return rc;      // box = 128 rows: half of a full tile
  if ((rc = tc::make_map_f16(&mTl, T.lo, m, T.Kp, T.Kp, TM))) return rc;
  prm.epi = P;
  prm.epi.nchunks = 2 * prm.echunks;   // two epilogue warps (column halves) per row
  { const char* d = getenv("B200KGE_DBG"); prm.dbg = d ? atoi(d) : 0; }
  const char* e = getenv("B200KGE_TC4_DIRECT");
  const bool direct = e && atoi(e) == 1;
  const int total = prm.q_tiles * prm.echunks;
  const int nclusters = tc::num_sms() / 2;
  const int grid = 2 * (total < nclusters ? total : nclusters);
  switch (epi_kind) {
    case EPI_STORE: return launch_e4<EPI_STORE>(direct, mQh, mQl, mTh, mTl, prm, grid, st);
    case EPI_BCE:   return launch_e4<EPI_BCE>(direct, mQh, mQl, mTh, mTl, prm, grid, st);
    case EPI_KL:    return launch_e4<EPI_KL>(direct, mQh, mQl, mTh, mTl, prm, grid, st);
    case EPI_RANK:  return launch_e4<EPI_RANK>(direct, mQh, mQl, mTh, mTl, prm, grid, st);
  }
  set_error("bad epilogue kind %d", epi_kind);
  return B200KGE_ERR_INVALID;
}

}  // namespace b200kge
