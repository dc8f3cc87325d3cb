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
// which is fp32-equivalent (dropped term lo*lo ~ 2^-22).  Q is split once by the prologue; the
// table is split ON THE FLY: TMA lands the raw fp32 tile in shared memory; the raw tile IS the hi
// operand (kind::tf32 ignores the low 13 mantissa bits — truncation, measured on B200), and four
// "splitter" warps write lo = rn_tf32(x - trunc_tf32(x)) next to it (same swizzled layout,
// element-wise), fence to the async proxy, and only then may the MMA warp consume the stage.
//
// CTA = 12 warps, one CTA per SM, persistent over (query tile, range of entity tiles):
//   warp 0      TMA producer   (one elected lane)     full[s]   <- expect_tx
//   warp 1      MMA issuer     (one elected lane)     empty[s]  <- tcgen05.commit ; tmem_full[b]
//   warps 4-7   epilogue       tcgen05.ld -> regs -> {transposed coalesced store | BCE | KL | rank}
//   warps 8-11  splitters      split[s] <- 128 arrivals
// Tile = 128 queries (UMMA M, TMEM lanes) x 256 entities (UMMA N, TMEM columns), K in chunks of 32
// floats (one 128-byte swizzle atom), 2 smem stages of 96 KB, 2 TMEM accumulators of 256 columns
// so the epilogue of tile i overlaps the MMAs of tile i+1.  TMEM lane = query row, so every
// per-row reduction (loss terms, logsumexp, rank counters) is thread-local.
#include <cuda.h>
#include <cstdlib>
#include "common.cuh"
#include "ptx.cuh"

namespace b200kge {

namespace {

constexpr int TM = 128;             // queries per tile  (UMMA M)
constexpr int TN = 256;             // entities per tile (UMMA N)
constexpr int TK = 32;              // floats per K chunk (128 B swizzle atom)
constexpr int STAGES = 2;
constexpr int A_BYTES = TM * TK * 4;   // 16 KB
constexpr int B_BYTES = TN * TK * 4;   // 32 KB
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;  // A_hi, A_lo, B_hi(raw), B_lo
constexpr int EPI_WARPS = 4, SPLIT_WARPS = 4;
constexpr int NTHREADS = 12 * 32;
constexpr int STG_LD = 33;
constexpr int STG_BYTES = EPI_WARPS * 32 * STG_LD * 4;
constexpr int SMEM_BYTES = 1024 /*align slack*/ + STAGES * STAGE_BYTES + STG_BYTES + 256 /*barriers*/;
constexpr int TMEM_COLS = 512;

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

struct TcParams {
  int64_t nq, m;
  int K;            // reduction length (floats)
  int q_tiles, e_tiles, echunks;
  EpiParams epi;
};

template <int EPI, int PASSES>
__global__ void __launch_bounds__(NTHREADS, 1)
pairwise_tc_kernel(const __grid_constant__ CUtensorMap tmQhi, const __grid_constant__ CUtensorMap tmQlo,
                   const __grid_constant__ CUtensorMap tmT, const TcParams prm) {
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
  const int total_work = prm.q_tiles * prm.echunks;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmQhi);
    if (PASSES == 3) ptx::prefetch_tensormap(&tmQlo);
    ptx::prefetch_tensormap(&tmT);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&split[s], SPLIT_WARPS * 32);
      ptx::mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&tfull[b], 1);
      ptx::mbar_init(&tempty[b], EPI_WARPS * 32);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<TMEM_COLS>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto stage_ptr = [&](int s) { return smem + s * STAGE_BYTES; };
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
            const int s = c % STAGES;
            const uint32_t ph = (c / STAGES) & 1;
            ptx::mbar_wait(&empty[s], ph ^ 1);
            uint8_t* sp = stage_ptr(s);
            ptx::mbar_arrive_expect_tx(&full[s], (PASSES == 3 ? 2 : 1) * A_BYTES + B_BYTES);
            ptx::tma_load_2d(sp, &tmQhi, &full[s], kc * TK, qt * TM);
            if (PASSES == 3) ptx::tma_load_2d(sp + A_BYTES, &tmQlo, &full[s], kc * TK, qt * TM);
            ptx::tma_load_2d(sp + 2 * A_BYTES, &tmT, &full[s], kc * TK, et * TN);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ============================================
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_tf32(TM, TN);
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
            const int s = c % STAGES;
            const uint32_t ph = (c / STAGES) & 1;
            if (PASSES == 3) ptx::mbar_wait(&split[s], ph); else ptx::mbar_wait(&full[s], ph);
            ptx::tc_fence_after();
            const uint32_t a_hi = ptx::smem_u32(stage_ptr(s));
            const uint32_t a_lo = a_hi + A_BYTES;
            const uint32_t b_hi = a_hi + 2 * A_BYTES;
            const uint32_t b_lo = b_hi + B_BYTES;
#pragma unroll
            for (int k4 = 0; k4 < TK / 8; ++k4) {
              const uint32_t ko = k4 * 32;  // 8 tf32 = 32 bytes along K inside the swizzle atom
              const uint32_t acc0 = (kc > 0 || k4 > 0) ? 1u : 0u;
              if (PASSES == 3) {
                ptx::umma_tf32(d_tmem, ptx::umma_desc_sw128(a_lo + ko), ptx::umma_desc_sw128(b_hi + ko), idesc, acc0);
                ptx::umma_tf32(d_tmem, ptx::umma_desc_sw128(a_hi + ko), ptx::umma_desc_sw128(b_lo + ko), idesc, 1u);
                ptx::umma_tf32(d_tmem, ptx::umma_desc_sw128(a_hi + ko), ptx::umma_desc_sw128(b_hi + ko), idesc, 1u);
              } else {
                ptx::umma_tf32(d_tmem, ptx::umma_desc_sw128(a_hi + ko), ptx::umma_desc_sw128(b_hi + ko), idesc, acc0);
              }
            }
            ptx::umma_commit(&empty[s]);           // smem stage free once these MMAs retire
          }
          ptx::umma_commit(&tfull[b]);             // accumulator complete
        }
      }
    }
  } else if (warp >= 8) {
    // ================================ splitters =============================================
    if (PASSES == 3) {
      const int t = threadIdx.x - 8 * 32;  // 0..127
      uint32_t c = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        int qt, et0, et1, ec;
        work_range(w, qt, et0, et1, ec);
        for (int et = et0; et < et1; ++et) {
          for (int kc = 0; kc < nk; ++kc, ++c) {
            const int s = c % STAGES;
            const uint32_t ph = (c / STAGES) & 1;
            ptx::mbar_wait(&full[s], ph);
            float4* bh = reinterpret_cast<float4*>(stage_ptr(s) + 2 * A_BYTES);
            float4* bl = reinterpret_cast<float4*>(stage_ptr(s) + 2 * A_BYTES + B_BYTES);
#pragma unroll 4
            for (int i = t; i < B_BYTES / 16; i += SPLIT_WARPS * 32) {
              // hi needs no write: the tensor core ignores the low 13 mantissa bits of the raw
              // fp32 tile (truncation, verified on B200); lo = rn_tf32(x - trunc_tf32(x)).
              const float4 v = bh[i];
              float4 l;
              l.x = tf32_rna(v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u));
              l.y = tf32_rna(v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u));
              l.z = tf32_rna(v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u));
              l.w = tf32_rna(v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u));
              bl[i] = l;
            }
            ptx::fence_proxy_async_smem();
            ptx::mbar_arrive(&split[s]);
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ================================ epilogue ==============================================
    const int ew = warp - 4;                  // == warp % 4: TMEM lanes [32*ew, 32*ew+32)
    float* my_stg = stg + ew * 32 * STG_LD;
    const EpiParams& P = prm.epi;
    uint32_t it = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      int qt, et0, et1, ec;
      work_range(w, qt, et0, et1, ec);
      const int64_t row = (int64_t)qt * TM + ew * 32 + lane;   // this thread's query row
      const bool row_ok = row < prm.nq;
      RowState<EPI> st;
      st.init();
      const float aux = row_ok ? epi_row_aux<EPI>(P, row) : 0.f;
      for (int et = et0; et < et1; ++et, ++it) {
        const int b = it & 1;
        ptx::mbar_wait(&tfull[b], (it >> 1) & 1);
        ptx::tc_fence_after();
        const int64_t e0 = (int64_t)et * TN;
#pragma unroll 1
        for (int j = 0; j < TN / 32; ++j) {
          uint32_t v[32];
          ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(b * TN + j * 32), v);
          ptx::tmem_ld_wait();
          const int64_t c0 = e0 + j * 32;
          if constexpr (EPI == EPI_STORE) {
            // transpose through smem so that each store instruction writes 32 consecutive
            // entities of ONE query row (coalesced 128 B), whatever the row stride is
#pragma unroll
            for (int c = 0; c < 32; ++c) my_stg[lane * STG_LD + c] = __uint_as_float(v[c]);
            __syncwarp();
            const int64_t col = c0 + lane;
            if (col < prm.m) {
#pragma unroll 4
              for (int rr = 0; rr < 32; ++rr) {
                int64_t r = (int64_t)qt * TM + ew * 32 + rr;
                if (r < prm.nq) {
                  int64_t cb = 0;
                  if (P.n_rows_out > 0 && r >= P.n_rows_out) { r -= P.n_rows_out; cb = P.col_block; }
                  P.out[r * P.ldo + cb + col] = my_stg[rr * STG_LD + lane];
                }
              }
            }
            __syncwarp();
          } else {
            if (row_ok) {
#pragma unroll
              for (int c = 0; c < 32; ++c)
                if (c0 + c < prm.m) epi_elem<EPI>(P, st, row, c0 + c, __uint_as_float(v[c]), aux);
            }
          }
        }
        ptx::tc_fence_before();
        ptx::mbar_arrive(&tempty[b]);
      }
      if constexpr (EPI != EPI_STORE) {
        if (row_ok) epi_flush<EPI>(P, st, row, ec);
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
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled not available from the driver"); return B200KGE_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)TK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld", (int)r, (long long)rows,
              (long long)cols, (long long)ld);
    return B200KGE_ERR_CUDA;
  }
  return 0;
}

int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

void plan(int64_t nq, int64_t m, int& q_tiles, int& e_tiles, int& echunks) {
  q_tiles = (int)((nq + TM - 1) / TM);
  e_tiles = (int)((m + TN - 1) / TN);
  int per = num_sms() / (q_tiles > 0 ? q_tiles : 1);
  if (per < 1) per = 1;
  if (per > e_tiles) per = e_tiles;
  echunks = per;
}

template <int EPI>
int launch_e(int passes, const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& c, const TcParams& prm,
             int grid, cudaStream_t st) {
  cudaError_t e;
  profile_begin(st);
  if (passes == 3) {
    e = cudaFuncSetAttribute(pairwise_tc_kernel<EPI, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(pairwise_tc_kernel)");
    pairwise_tc_kernel<EPI, 3><<<grid, NTHREADS, SMEM_BYTES, st>>>(a, b, c, prm);
  } else {
    e = cudaFuncSetAttribute(pairwise_tc_kernel<EPI, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(pairwise_tc_kernel)");
    pairwise_tc_kernel<EPI, 1><<<grid, NTHREADS, SMEM_BYTES, st>>>(a, b, c, prm);
  }
  profile_end(st);
  B2K_LAUNCH_CHECK("pairwise_tc_kernel");
  return 0;
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
  int qt, et, ec;
  plan(nq, m, qt, et, ec);
  return ec;
}

int launch_pairwise_tc(int epi_kind, int passes, const float* Qhi, const float* Qlo, int64_t ldq,
                       int64_t nq, const float* T, int64_t ldt, int64_t m, int K,
                       const EpiParams& P, cudaStream_t st) {
  if (nq == 0 || m == 0) return 0;
  CUtensorMap mQhi, mQlo, mT;
  int rc;
  if ((rc = make_map(&mQhi, Qhi, nq, K, ldq, TM))) return rc;
  if ((rc = make_map(&mQlo, passes == 3 ? Qlo : Qhi, nq, K, ldq, TM))) return rc;
  if ((rc = make_map(&mT, T, m, K, ldt, TN))) return rc;
  TcParams prm;
  prm.nq = nq; prm.m = m; prm.K = K;
  plan(nq, m, prm.q_tiles, prm.e_tiles, prm.echunks);
  prm.epi = P;
  prm.epi.nchunks = prm.echunks;
  const int total = prm.q_tiles * prm.echunks;
  const int grid = total < num_sms() ? total : num_sms();
  switch (epi_kind) {
    case EPI_STORE: return launch_e<EPI_STORE>(passes, mQhi, mQlo, mT, prm, grid, st);
    case EPI_BCE:   return launch_e<EPI_BCE>(passes, mQhi, mQlo, mT, prm, grid, st);
    case EPI_KL:    return launch_e<EPI_KL>(passes, mQhi, mQlo, mT, prm, grid, st);
    case EPI_RANK:  return launch_e<EPI_RANK>(passes, mQhi, mQlo, mT, prm, grid, st);
  }
  set_error("bad epilogue kind %d", epi_kind);
  return B200KGE_ERR_INVALID;
}

}  // namespace b200kge
