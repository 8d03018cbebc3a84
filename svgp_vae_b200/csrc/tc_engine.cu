// tcgen05 / TMEM / TMA implementation of the three O(N M^2 L) contractions of the SVGP step
// (sm_100a only).  fp32-accurate on TF32 tensor cores through the 3xTF32 split: every fp32 operand
// lives in memory as a TF32 pair x = hi + lo (written by K1 / svgp_split_tf32) and each k-step issues
//     D += A_lo * B_hi;   D += A_hi * B_lo;   D += A_hi * B_hi          (fp32 accumulators in TMEM)
//
//   MODE_SYRK    A_l[a,b]  += sum_n (w[n,l] Kt[a,n]) * Kt[b,n]   K2, SVGPVAE_model.py:328-330 and the
//                                                                adjoint of the row-wise quadratic forms
//   MODE_ROWQUAD q[i,l]     = sum_c (sum_a K[i,a] B_l[c,a]) * X   K4, :336-337, :284   (X = same product when
//                                                                B_l is a triangular factor, else K[i,c])
//   MODE_SCALED  out[i,c]   = sum_l sum_a (w[i,l] K[i,a]) G_l[c,a]   dObjective/dK_nm
//
// One persistent CTA per SM, 10 warps with fixed roles:
//   warp 0      TMA producer: four 2-D tensor maps (A_hi, A_lo, B_hi, B_lo), 128B-swizzled K-major boxes
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (128 x BN x 8 TF32 atoms)
//   warps 2-5   epilogue: tcgen05.ld the accumulator (one TMEM lane == one output row per thread)
//   warps 6-9   operand transform (SYRK / SCALED only): the per-channel diag(w) scaling cannot be
//               precomputed for L channels, so the TMA-landed tile is rescaled and re-split into a TF32
//               pair in place in shared memory, then handed to the MMA warp through a second mbarrier
// Pipelines: smem stages (full -> [ready] -> empty) and two TMEM accumulator buffers (tmem_full/empty)
// so that the epilogue of one tile overlaps the MMAs of the next.
#include <cuda.h>

#include "common.cuh"

namespace svgp {

constexpr int MODE_SYRK = 0, MODE_ROWQUAD = 1, MODE_SCALED = 2;
constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 32;                 // 32 TF32 = 128 bytes = one swizzle row
constexpr int UMMA_K = 8;
constexpr int NUM_THREADS = 320;
constexpr int TC_SMEM_LIMIT = 227 * 1024;

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart (SBO)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);        // start address, 16-byte units
  d |= (uint64_t)1 << 16;                         // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset
  d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                         // SWIZZLE_128B
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4)                 // D format: F32
         | (2u << 7)               // A format: TF32
         | (2u << 10)              // B format: TF32
         | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------
// kernel parameters
// ---------------------------------------------------------------------------------------------
struct TcParams {
  int64_t N, M, L;
  // SYRK
  const float* Wt;        // (L, ldwt) channel-major weights, contiguous in n
  int64_t ldwt;
  double* A;              // (L, M, M) accumulated
  int64_t chunk_rows, nchunk;
  int ntile;              // number of (ta, tb) tile pairs
  // ROWQUAD
  int tri;
  const float* K_hi;      // for the DOT epilogue
  const float* K_lo;
  int64_t ldk;
  float* q;
  int64_t ldq;
  int lgroup;             // channels scheduled together (L2 residency of their B planes)
  // SCALED
  const float* W;         // (N, ldw)
  int64_t ldw;
  float* out;
  int64_t ldo;
  int accumulate;
  int lflush;             // channels per TMEM accumulation chain
  int64_t n_items;
};

template <int BN, int STAGES>
struct SmemLayout {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 4;      // one plane
  static constexpr int B_BYTES = BN * BLOCK_K * 4;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;      // barriers + tmem ptr + alignment slack
};

// SYRK tile pairs: a-tiles of 128 rows, b-tiles of BN columns, kept when the tile touches the lower triangle
__host__ __device__ inline int syrk_tile_count(int64_t M, int BN) {
  int cnt = 0;
  for (int64_t ta = 0; ta * BLOCK_M < M; ++ta)
    for (int64_t tb = 0; tb * BN < M && tb * BN <= ta * BLOCK_M + BLOCK_M - 1; ++tb) ++cnt;
  return cnt;
}
__device__ inline void syrk_tile_decode(int idx, int64_t M, int BN, int& ta_out, int& tb_out) {
  int cnt = 0;
  for (int ta = 0; (int64_t)ta * BLOCK_M < M; ++ta)
    for (int tb = 0; (int64_t)tb * BN < M && (int64_t)tb * BN <= (int64_t)ta * BLOCK_M + BLOCK_M - 1; ++tb) {
      if (cnt == idx) { ta_out = ta; tb_out = tb; return; }
      ++cnt;
    }
  ta_out = tb_out = 0;
}

template <int MODE, int BN, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_kernel(const __grid_constant__ CUtensorMap mapA_hi, const __grid_constant__ CUtensorMap mapA_lo,
          const __grid_constant__ CUtensorMap mapB_hi, const __grid_constant__ CUtensorMap mapB_lo, const TcParams P) {
  using SL = SmemLayout<BN, STAGES>;
  constexpr bool HAS_XFORM = (MODE != MODE_ROWQUAD);
  constexpr int ACC_BUFS = (2 * BN <= 512) ? 2 : 1;
  constexpr uint32_t IDESC = make_idesc(BN);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + SL::BAR_OFFSET);
  uint64_t* full = bars;                         // [STAGES]
  uint64_t* ready = bars + STAGES;               // [STAGES]
  uint64_t* empty = bars + 2 * STAGES;           // [STAGES]
  uint64_t* tmem_full = bars + 3 * STAGES;       // [2]
  uint64_t* tmem_empty = bars + 3 * STAGES + 2;  // [2]
  uint32_t* tmem_ptr = (uint32_t*)(bars + 3 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&ready[s], 4);
      mbar_init(&empty[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&tmem_full[t], 1);
      mbar_init(&tmem_empty[t], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapA_hi); prefetch_tmap(&mapA_lo); prefetch_tmap(&mapB_hi); prefetch_tmap(&mapB_lo);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // ---- work decomposition -------------------------------------------------------------------
  // Every role walks the same sequence: item -> sub-tiles -> k-blocks.  One TMEM accumulator buffer
  // holds one sub-tile; its MMA chain is kept short on purpose: the tensor core accumulates with
  // truncation, so a chain of n MMAs carries a bias of up to ~n * 2^-24.  Long reductions (SYRK over
  // N datapoints, SCALED over L * M) are therefore cut into sub-tiles that the epilogue folds into a
  // CTA-owned global tile (double for SYRK, float for SCALED) with ordinary round-to-nearest adds.
  //   SYRK    item = (tile pair, channel)      sub = chunk of `chunk_rows` datapoints
  //   ROWQUAD item = (row tile, channel)       sub = column tile of B_l
  //   SCALED  item = (row tile, column tile)   sub = group of `lflush` channels
  const int64_t M = P.M;
  const int nct = (int)((M + BN - 1) / BN);
  const int kb_full = (int)((M + BLOCK_K - 1) / BLOCK_K);

  auto item_subtiles = [&]() -> int {
    if (MODE == MODE_SYRK) return (int)P.nchunk;
    if (MODE == MODE_ROWQUAD) return nct;
    return (int)((P.L + P.lflush - 1) / P.lflush);
  };
  auto subtile_kblocks = [&](int sub) -> int {
    if (MODE == MODE_SYRK) {
      int64_t n0 = (int64_t)sub * P.chunk_rows;
      int64_t n1 = n0 + P.chunk_rows < P.N ? n0 + P.chunk_rows : P.N;
      return (int)((n1 - n0 + BLOCK_K - 1) / BLOCK_K);
    } else if (MODE == MODE_ROWQUAD) {
      if (!P.tri) return kb_full;
      int64_t kend = (int64_t)(sub + 1) * BN < M ? (int64_t)(sub + 1) * BN : M;
      return (int)((kend + BLOCK_K - 1) / BLOCK_K);
    } else {
      int64_t l0 = (int64_t)sub * P.lflush;
      int64_t nl = l0 + P.lflush < P.L ? P.lflush : P.L - l0;
      return (int)(nl * kb_full);
    }
  };
  // item -> coordinates
  struct Item { int64_t l, itile; int a_row0, b_row0; };
  auto decode = [&](int64_t item) -> Item {
    Item it{0, 0, 0, 0};
    if (MODE == MODE_SYRK) {
      it.l = item % P.L;
      int ta, tb;
      syrk_tile_decode((int)(item / P.L), M, BN, ta, tb);
      it.a_row0 = ta * BLOCK_M; it.b_row0 = tb * BN;
    } else if (MODE == MODE_ROWQUAD) {
      int64_t ntile_r = (P.N + BLOCK_M - 1) / BLOCK_M;
      int64_t per_group = ntile_r * P.lgroup;
      int64_t g = item / per_group, rem = item % per_group;
      it.itile = rem / P.lgroup; it.l = g * P.lgroup + rem % P.lgroup;
      it.a_row0 = (int)(it.itile * BLOCK_M);
    } else {
      it.itile = item / nct;
      it.a_row0 = (int)(it.itile * BLOCK_M); it.b_row0 = (int)((item % nct) * BN);
    }
    return it;
  };

  if (warp == 0) {
    // =============================== TMA producer ===============================================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int64_t item = blockIdx.x; item < P.n_items; item += gridDim.x) {
        const Item it = decode(item);
        const int nsub = item_subtiles();
        for (int sub = 0; sub < nsub; ++sub) {
          const int nkb = subtile_kblocks(sub);
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* st = smem + stage * SL::STAGE_BYTES;
            int32_t ak, ar, bk, br;
            if (MODE == MODE_SYRK) {
              ak = (int32_t)((int64_t)sub * P.chunk_rows) + kb * BLOCK_K; ar = it.a_row0; bk = ak; br = it.b_row0;
            } else if (MODE == MODE_ROWQUAD) {
              ak = kb * BLOCK_K; ar = it.a_row0; bk = ak; br = (int32_t)(it.l * M + (int64_t)sub * BN);
            } else {
              int lc = sub * P.lflush + kb / kb_full, kk = kb % kb_full;
              ak = kk * BLOCK_K; ar = it.a_row0; bk = ak; br = (int32_t)((int64_t)lc * M + it.b_row0);
            }
            mbar_expect_tx(&full[stage], SL::STAGE_BYTES);
            tma_load_2d(st, &mapA_hi, &full[stage], ak, ar);
            tma_load_2d(st + SL::A_BYTES, &mapA_lo, &full[stage], ak, ar);
            tma_load_2d(st + 2 * SL::A_BYTES, &mapB_hi, &full[stage], bk, br);
            tma_load_2d(st + 2 * SL::A_BYTES + SL::B_BYTES, &mapB_lo, &full[stage], bk, br);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ==================================================
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int64_t item = blockIdx.x; item < P.n_items; item += gridDim.x) {
      const int nsub = item_subtiles();
      for (int sub = 0; sub < nsub; ++sub) {
        const int nkb = subtile_kblocks(sub);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full[stage], phase);
          if (HAS_XFORM) mbar_wait(&ready[stage], phase);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t sa = smem_u32(smem + stage * SL::STAGE_BYTES);
            const uint64_t a_hi = make_smem_desc(sa), a_lo = make_smem_desc(sa + SL::A_BYTES);
            const uint64_t b_hi = make_smem_desc(sa + 2 * SL::A_BYTES), b_lo = make_smem_desc(sa + 2 * SL::A_BYTES + SL::B_BYTES);
#pragma unroll
            for (int ks = 0; ks < BLOCK_K / UMMA_K; ++ks) {
              const uint64_t adv = (uint64_t)((ks * UMMA_K * 4) >> 4);
              umma_tf32(d_tmem, a_lo + adv, b_hi + adv, IDESC, (kb > 0 || ks > 0) ? 1u : 0u);
              umma_tf32(d_tmem, a_hi + adv, b_lo + adv, IDESC, 1u);
              umma_tf32(d_tmem, a_hi + adv, b_hi + adv, IDESC, 1u);
            }
            umma_commit(&empty[stage]);
            if (kb == nkb - 1) umma_commit(&tmem_full[acc]);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == ACC_BUFS) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp < 6) {
    // =============================== epilogue ====================================================
    const int qd = warp & 3;                         // TMEM lane quarter this warp may access
    const int row = qd * 32 + lane;                  // output row inside the tile
    int acc = 0; uint32_t acc_phase = 0;
    for (int64_t item = blockIdx.x; item < P.n_items; item += gridDim.x) {
      const Item it = decode(item);
      float qsum = 0.f;
      const int nsub = item_subtiles();
      for (int sub = 0; sub < nsub; ++sub) {
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          float v[32];
          tmem_ld32(taddr + c0, v);
          if (MODE == MODE_SYRK) {
            // CTA-owned tile of the double accumulator: plain read-modify-write, lower triangle only
            const int64_t a = it.a_row0 + row;
            if (a < M) {
              double* Arow = P.A + (it.l * M + a) * M + it.b_row0 + c0;
              const int64_t bmax = a - (it.b_row0 + c0);          // columns j <= bmax are on/below the diagonal
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j <= bmax) Arow[j] += (double)v[j];
            }
          } else if (MODE == MODE_ROWQUAD) {
            const int64_t i = it.itile * BLOCK_M + row;
            const int64_t cbase = (int64_t)sub * BN + c0;
            if (P.tri) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (cbase + j < M) qsum = fmaf(v[j], v[j], qsum);
            } else if (i < P.N) {
              const float* kh = P.K_hi + i * P.ldk + cbase;
              const float* kl = P.K_lo + i * P.ldk + cbase;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (cbase + j < M) qsum = fmaf(v[j], kh[j] + kl[j], qsum);
            }
          } else {
            const int64_t i = it.itile * BLOCK_M + row;
            if (i < P.N) {
              float* o = P.out + i * P.ldo + it.b_row0 + c0;
              const bool add = (sub > 0) || P.accumulate;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if ((int64_t)it.b_row0 + c0 + j < M) o[j] = add ? o[j] + v[j] : v[j];
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        if (++acc == ACC_BUFS) { acc = 0; acc_phase ^= 1; }
      }
      if (MODE == MODE_ROWQUAD) {
        const int64_t i = it.itile * BLOCK_M + row;
        if (i < P.N && it.l < P.L) P.q[i * P.ldq + it.l] = qsum;
      }
    }
  } else if (HAS_XFORM) {
    // =============================== operand transform ===========================================
    const int t = threadIdx.x - 6 * 32;              // 0..127
    int stage = 0; uint32_t phase = 0;
    // SW128: the 16-byte chunk index is XORed with (row & 7); this thread always sits on physical chunk t%8 of
    // rows t/8 + 16*it, so its logical chunk (= k offset / 4) is the same for every row it touches
    const int pchunk = t & 7, rbase = t >> 3;
    const int lchunk = pchunk ^ (rbase & 7);
    for (int64_t item = blockIdx.x; item < P.n_items; item += gridDim.x) {
      const Item it = decode(item);
      float rs[8];                                    // SCALED: per-row weights of the current channel
      int cur_l = -1;
      const int nsub = item_subtiles();
      for (int sub = 0; sub < nsub; ++sub) {
        const int nkb = subtile_kblocks(sub);
        for (int kb = 0; kb < nkb; ++kb) {
          float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (MODE == MODE_SYRK) {
            const int64_t n = (int64_t)sub * P.chunk_rows + (int64_t)kb * BLOCK_K + lchunk * 4;
            const float* wp = P.Wt + it.l * P.ldwt + n;
            if (n + 3 < P.N) wv = *reinterpret_cast<const float4*>(wp);
            else { if (n < P.N) wv.x = wp[0]; if (n + 1 < P.N) wv.y = wp[1]; if (n + 2 < P.N) wv.z = wp[2]; }
          } else {
            int lc = sub * P.lflush + kb / kb_full;
            if (lc != cur_l) {
              cur_l = lc;
#pragma unroll
              for (int r = 0; r < 8; ++r) {
                int64_t i = it.itile * BLOCK_M + rbase + 16 * r;
                rs[r] = (i < P.N) ? P.W[i * P.ldw + lc] : 0.f;
              }
            }
          }
          mbar_wait(&full[stage], phase);
          uint8_t* st = smem + stage * SL::STAGE_BYTES;
          // both modes rescale the 128-row A operand (for the SYRK the weight may sit on either factor of
          // k_a k_b; the A tile is half the size of the B tile, which halves the shared-memory traffic here)
          constexpr int ROWS = BLOCK_M;
          uint8_t* hi_p = st;
          uint8_t* lo_p = st + SL::A_BYTES;
#pragma unroll
          for (int r = 0; r < ROWS / 16; ++r) {
            const int off = (rbase + 16 * r) * 128 + pchunk * 16;
            float4 h = *reinterpret_cast<float4*>(hi_p + off);
            float4 lo4 = *reinterpret_cast<float4*>(lo_p + off);
            float4 s = (MODE == MODE_SYRK) ? wv : make_float4(rs[r & 7], rs[r & 7], rs[r & 7], rs[r & 7]);
            float y0 = (h.x + lo4.x) * s.x, y1 = (h.y + lo4.y) * s.y, y2 = (h.z + lo4.z) * s.z, y3 = (h.w + lo4.w) * s.w;
            float4 nh = make_float4(to_tf32(y0), to_tf32(y1), to_tf32(y2), to_tf32(y3));
            float4 nl = make_float4(to_tf32(y0 - nh.x), to_tf32(y1 - nh.y), to_tf32(y2 - nh.z), to_tf32(y3 - nh.w));
            *reinterpret_cast<float4*>(hi_p + off) = nh;
            *reinterpret_cast<float4*>(lo_p + off) = nl;
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&ready[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// copy the lower triangle of every A_l onto its upper triangle (the SYRK kernel only writes b <= a)
__global__ void mirror_lower_kernel(double* __restrict__ A, int64_t M, int64_t L) {
  const int64_t total = L * M * M;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t l = idx / (M * M), rem = idx - l * M * M, r = rem / M, c = rem - r * M;
    if (c > r) A[idx] = A[(l * M + c) * M + r];
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D fp32 tensor (rows x cols, leading dimension ld elements), box = box_rows x 32 columns, 128B swizzle
static int make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return SVGP_ERR_CUDA; }
  if (((uintptr_t)base & 15) || (ld * 4) % 16) { set_error("TMA operand needs 16-byte aligned base and row pitch"); return SVGP_ERR_ARG; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return SVGP_ERR_CUDA; }
  return SVGP_OK;
}

static int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// tile configuration: 128 x 256 x 32 with two smem stages (default) or 128 x 128 x 32 with three (SVGP_TC_BN=128)
static int tc_bn() {
  static int bn = 0;
  if (!bn) {
    const char* e = getenv("SVGP_TC_BN");
    bn = (e && atoi(e) == 128) ? 128 : 256;
  }
  return bn;
}

template <int MODE, int BN, int STAGES>
static int launch_tc(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi, const CUtensorMap& b_lo,
                     const TcParams& P, cudaStream_t st, const char* name) {
  using SL = SmemLayout<BN, STAGES>;
  static_assert(SL::TOTAL <= TC_SMEM_LIMIT, "shared memory budget");
  auto kern = tc_kernel<MODE, BN, STAGES>;
  static bool attr_done = false;
  if (!attr_done) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SL::TOTAL) != cudaSuccess) return check_launch(name);
    attr_done = true;
  }
  int64_t grid = P.n_items < num_sms() ? P.n_items : num_sms();
  if (grid <= 0) return SVGP_OK;
  kern<<<(unsigned)grid, NUM_THREADS, SL::TOTAL, st>>>(a_hi, a_lo, b_hi, b_lo, P);
  return check_launch(name);
}

template <int MODE>
static int dispatch_tc(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi, const CUtensorMap& b_lo,
                       const TcParams& P, cudaStream_t st, const char* name) {
  if (tc_bn() == 128) return launch_tc<MODE, 128, 3>(a_hi, a_lo, b_hi, b_lo, P, st, name);
  return launch_tc<MODE, 256, 2>(a_hi, a_lo, b_hi, b_lo, P, st, name);
}

bool tc_shape_ok(const svgp_kop* kop) {
  // worth it only when tiles are mostly full; TMA needs 16-byte pitches
  return kop->M >= 128 && kop->N >= 2048 && (kop->ldk % 4) == 0 && (kop->ldkt % 4) == 0;
}

int tc_syrk(const svgp_kop* kop, const float* Wt, int64_t ldwt, int64_t L, double* A, int64_t chunk_rows, cudaStream_t st) {
  if (!kop->Kt || !kop->Kt_lo) { set_error("tc_syrk: transposed TF32 planes missing"); return SVGP_ERR_ARG; }
  if (((uintptr_t)Wt & 15) || (ldwt % 4)) { set_error("tc_syrk: weights need 16-byte alignment"); return SVGP_ERR_ARG; }
  const int BN = tc_bn();
  CUtensorMap a_hi, a_lo, b_hi, b_lo;
  int rc;
  if ((rc = make_map(&a_hi, kop->Kt, kop->M, kop->N, kop->ldkt, BLOCK_M))) return rc;
  if ((rc = make_map(&a_lo, kop->Kt_lo, kop->M, kop->N, kop->ldkt, BLOCK_M))) return rc;
  if ((rc = make_map(&b_hi, kop->Kt, kop->M, kop->N, kop->ldkt, BN))) return rc;
  if ((rc = make_map(&b_lo, kop->Kt_lo, kop->M, kop->N, kop->ldkt, BN))) return rc;
  TcParams P{};
  P.N = kop->N; P.M = kop->M; P.L = L;
  P.Wt = Wt; P.ldwt = ldwt; P.A = A;
  int64_t chunk = chunk_rows > 0 ? chunk_rows : 1024;
  chunk = (chunk + BLOCK_K - 1) / BLOCK_K * BLOCK_K;
  P.chunk_rows = chunk; P.nchunk = ceil_div(kop->N, chunk);
  P.ntile = syrk_tile_count(kop->M, BN);
  P.n_items = (int64_t)P.ntile * L;
  rc = dispatch_tc<MODE_SYRK>(a_hi, a_lo, b_hi, b_lo, P, st, "svgp_syrk(tc)");
  if (rc) return rc;
  int64_t blocks = ceil_div(L * kop->M * kop->M, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  mirror_lower_kernel<<<(unsigned)blocks, 256, 0, st>>>(A, kop->M, L);
  return check_launch("svgp_syrk(mirror)");
}

int tc_rowquad(const svgp_kop* kop, const float* S_hi, const float* S_lo, int64_t L, int tri, float* q, int64_t ldq,
               cudaStream_t st) {
  const int BN = tc_bn();
  CUtensorMap a_hi, a_lo, b_hi, b_lo;
  int rc;
  if ((rc = make_map(&a_hi, kop->K, kop->N, kop->M, kop->ldk, BLOCK_M))) return rc;
  if ((rc = make_map(&a_lo, kop->K_lo, kop->N, kop->M, kop->ldk, BLOCK_M))) return rc;
  if ((rc = make_map(&b_hi, S_hi, L * kop->M, kop->M, kop->M, BN))) return rc;
  if ((rc = make_map(&b_lo, S_lo, L * kop->M, kop->M, kop->M, BN))) return rc;
  TcParams P{};
  P.N = kop->N; P.M = kop->M; P.L = L;
  P.tri = tri; P.K_hi = kop->K; P.K_lo = kop->K_lo; P.ldk = kop->ldk; P.q = q; P.ldq = ldq;
  // channels whose factor planes (2 * M*M*4 bytes each) share ~half of the 126 MB L2
  int64_t per = 2 * kop->M * kop->M * 4;
  int64_t g = (64LL << 20) / (per > 0 ? per : 1);
  if (g < 1) g = 1;
  if (g > L) g = L;
  while (L % g) --g;                       // keep groups uniform
  P.lgroup = (int)g;
  P.n_items = ceil_div(kop->N, BLOCK_M) * L;
  return dispatch_tc<MODE_ROWQUAD>(a_hi, a_lo, b_hi, b_lo, P, st, "svgp_rowquad(tc)");
}

int tc_scaled_gemm(const svgp_kop* kop, const float* W, int64_t ldw, const float* G_hi, const float* G_lo, int64_t L, float* out,
                   int64_t ldo, int accumulate, cudaStream_t st) {
  const int BN = tc_bn();
  CUtensorMap a_hi, a_lo, b_hi, b_lo;
  int rc;
  if ((rc = make_map(&a_hi, kop->K, kop->N, kop->M, kop->ldk, BLOCK_M))) return rc;
  if ((rc = make_map(&a_lo, kop->K_lo, kop->N, kop->M, kop->ldk, BLOCK_M))) return rc;
  if ((rc = make_map(&b_hi, G_hi, L * kop->M, kop->M, kop->M, BN))) return rc;
  if ((rc = make_map(&b_lo, G_lo, L * kop->M, kop->M, kop->M, BN))) return rc;
  TcParams P{};
  P.N = kop->N; P.M = kop->M; P.L = L;
  P.W = W; P.ldw = ldw; P.out = out; P.ldo = ldo; P.accumulate = accumulate;
  {
    // keep one accumulation chain to ~768 MMAs (measured truncation bias ~2.7e-8 per MMA -> ~2e-5):
    // lflush channels x (M/32) k-blocks x 12 MMAs
    int64_t kb = ceil_div(kop->M, BLOCK_K);
    int64_t lf = 64 / (kb > 0 ? kb : 1);
    const char* e = getenv("SVGP_TC_LFLUSH");
    if (e && atoi(e) > 0) lf = atoi(e);
    if (lf < 1) lf = 1;
    P.lflush = (int)lf;
  }
  P.n_items = ceil_div(kop->N, BLOCK_M) * ceil_div(kop->M, BN);
  return dispatch_tc<MODE_SCALED>(a_hi, a_lo, b_hi, b_lo, P, st, "svgp_scaled_gemm(tc)");
}

}  // namespace svgp
