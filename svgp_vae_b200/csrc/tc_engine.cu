// tcgen05 / TMEM / TMA implementation of the three O(N M^2 L) contractions of the SVGP step
// (sm_100a only).  fp32-accurate products on the FP16 tensor-core path through a 3-term split:
// every operand lives in memory as an fp16 pair  x * s = hi + lo  (s a power of two chosen per
// matrix so that max|x| s < 2^14: 22 significand bits down to 2^-17 of the maximum, see
// svgp_split_f16 / svgp_kernel_fwd) and each k-step issues
//     D += A_lo * B_hi;   D += A_hi * B_lo;   D += A_hi * B_hi          (fp32 accumulators in TMEM)
// at twice the TF32 rate and half the operand bytes of a 3xTF32 scheme.
//
//   MODE_SYRK    A_l[a,b] += sum_n (w[n,l] Kt[a,n]) * Kt[b,n]       K2, SVGPVAE_model.py:328-330, and the
//                                                                   adjoint of the row-wise quadratic forms
//   MODE_QUAD    q[i,l]    = sum_c T_l[i,c] * X[i,c],  T_l = K B_l^T   K4, :336-337, :284
//                                                                   (X = T_l when B_l is a triangular factor, else K)
//   MODE_SCALED  out[i,c]  = sum_s w[i,s] T_s[i,c],   dots[i,s] = sum_c T_s[i,c] K[i,c]
//                                                                   dObjective/dK_nm and dObjective/dp in one pass
//
// One persistent CTA per SM, 10 warps with fixed roles:
//   warp 0      TMA producer: four 2-D tensor maps (A_hi, A_lo, B_hi, B_lo), swizzled K-major boxes
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (128 x BN x 16 FP16 atoms)
//   warps 2-5   epilogue: tcgen05.ld the accumulator (one TMEM lane == one output row per thread)
//   warps 6-9   SYRK: operand transform -- the per-channel diag(w) scaling sits on the reduction index, so the
//               TMA-landed A tile is rescaled and re-split into an fp16 pair in place in shared memory and
//               handed to the MMA warp through a second mbarrier;
//               SCALED: four more epilogue warps (each warp owns half of the tile's columns: the running
//               sum over the sub-tiles lives in registers, the per-row weights are applied in the epilogue)
// Pipelines: smem stages (full -> [ready] -> empty) and two TMEM accumulator buffers (tmem_full/empty)
// so that the epilogue of one sub-tile overlaps the MMAs of the next.
#include "tc_ptx.cuh"

namespace svgp {

constexpr int MODE_SYRK = 0, MODE_QUAD = 1, MODE_SCALED = 2;
constexpr int BLOCK_M = 128;
constexpr int UMMA_K = 16;                  // fp16 elements per MMA k-step (32 bytes)
constexpr int TC_SMEM_LIMIT = 227 * 1024;
// threads per CTA: QUAD runs 10 warps; SYRK runs 4 warpgroups and SCALED 3 (register budgets are re-split between
// the warpgroups with setmaxnreg at the top of each role branch)
__host__ __device__ constexpr int tc_threads(int mode) { return mode == 0 ? 512 : (mode == 2 ? 384 : 320); }

// K-major swizzled operand tile: rows of RB (= 128 or 64) bytes, 8-row groups 8*RB bytes apart (SBO)
template <int RB>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);        // start address, 16-byte units
  d |= (uint64_t)1 << 16;                         // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)((8 * RB) >> 4) << 32;           // stride byte offset
  d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell)
  d |= (uint64_t)(RB == 128 ? 2 : 4) << 61;       // SWIZZLE_128B / SWIZZLE_64B
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4)                 // D format: F32
         | (0u << 7)               // A format: F16
         | (0u << 10)              // B format: F16
         | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
}

// scale an fp16 pair (hi + lo) by w and split the product into a new pair
__device__ __forceinline__ void rescale_pair(uint32_t& hi2, uint32_t& lo2, float w0, float w1) {
  const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&hi2));
  const float2 l = __half22float2(*reinterpret_cast<const __half2*>(&lo2));
  const float y0 = (h.x + l.x) * w0, y1 = (h.y + l.y) * w1;
  const __half2 nh = __floats2half2_rn(y0, y1);
  const float2 nhf = __half22float2(nh);
  const __half2 nl = __floats2half2_rn(y0 - nhf.x, y1 - nhf.y);
  hi2 = *reinterpret_cast<const uint32_t*>(&nh);
  lo2 = *reinterpret_cast<const uint32_t*>(&nl);
}
// the same in packed-half arithmetic, with the weight itself given as an fp16 pair w = wh + wl:
//   (hi + lo)(wh + wl) = hi wh + [lo wh + hi wl] + O(2^-22);   nh = rn(hi wh), and fma(hi, wh, -nh) is the EXACT
// rounding residual of that product (it has at most 11 significant bits), so nh + nl carries 22 bits again.
__device__ __forceinline__ void rescale_pair_h2(uint32_t& hi2, uint32_t& lo2, uint32_t wh2, uint32_t wl2) {
  const __half2 h = *reinterpret_cast<const __half2*>(&hi2), l = *reinterpret_cast<const __half2*>(&lo2);
  const __half2 wh = *reinterpret_cast<const __half2*>(&wh2), wl = *reinterpret_cast<const __half2*>(&wl2);
  const __half2 nh = __hmul2(h, wh);
  __half2 nl = __hfma2(h, wh, __hneg2(nh));
  nl = __hfma2(l, wh, nl);
  nl = __hfma2(h, wl, nl);
  hi2 = *reinterpret_cast<const uint32_t*>(&nh);
  lo2 = *reinterpret_cast<const uint32_t*>(&nl);
}
__device__ __forceinline__ void split_weights(float a, float b, uint32_t& wh2, uint32_t& wl2) {
  const __half2 wh = __floats2half2_rn(a, b);
  const float2 f = __half22float2(wh);
  const __half2 wl = __floats2half2_rn(a - f.x, b - f.y);
  wh2 = *reinterpret_cast<const uint32_t*>(&wh);
  wl2 = *reinterpret_cast<const uint32_t*>(&wl);
}
__device__ __forceinline__ float pair_dot2(uint32_t hi2, uint32_t lo2, float v0, float v1, float acc) {
  const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&hi2));
  const float2 l = __half22float2(*reinterpret_cast<const __half2*>(&lo2));
  acc = fmaf(v0, h.x + l.x, acc);
  return fmaf(v1, h.y + l.y, acc);
}

// ---------------------------------------------------------------------------------------------
// kernel parameters
// ---------------------------------------------------------------------------------------------
struct TcParams {
  int64_t N, M, L;            // L = number of channels (SYRK / QUAD) or of stacked matrices (SCALED)
  const float* kscale;        // {scale, 1/scale} of the K planes
  const float* binv;          // 1/scale per B matrix (QUAD / SCALED); per channel weight 1/scale (SYRK)
  // SYRK
  const float* Wt;            // (L, ldwt) channel-major weights, pre-scaled, zero padded
  int64_t ldwt;
  double* A;                  // (L, M, M) accumulated
  int64_t chunk_rows;         // datapoints per MMA chain (one TMEM sub-tile)
  int64_t sc_rows;            // datapoints per super-chunk: an item covers one super-chunk of one (tile pair, channel)
  int nsc;                    // number of super-chunks
  int ntile;                  // number of (ta, tb) tile pairs
  const float* wneg;          // per channel: != 0 if the channel has a negative weight
  float bias_coef;            // expected relative truncation loss of a one-signed chain per MMA, in units of 2^-24
  int* locks;                 // one word per (tile pair, channel, epilogue warp): guards the float64 read-modify-write
  int flush_every;            // chunks folded in fp32 registers between two float64 read-modify-writes of the tile
  int debug;                  // experiments (SVGP_TC_DEBUG): bit 0 = skip the operand transform arithmetic (wrong results),
                              // bit 1 = fp32 transform instead of the packed-half one
  // QUAD
  int tri;
  const __half* K_hi;         // for the DOT epilogues
  const __half* K_lo;
  int64_t ldkh;
  const __half* Kt_hi;        // datapoint-blocked transposed planes (SCALED k-dots), may be null
  const __half* Kt_lo;
  int64_t ldkt;
  float* q;
  int64_t ldq;
  int lgroup;                 // channels scheduled together (L2 residency of their B planes)
  // SCALED
  int64_t Mc;                 // output columns = rows of each stacked B matrix (M for the dK_nm product, L for K Wm^T)
  const float* W;             // (N, ldw) per-row weights, or null (all ones)
  int64_t ldw;
  float* out;
  int64_t ldo;
  int accumulate;
  float* dots;                // (N, lddots) or null
  int64_t lddots, ndot;
  int kcache;                 // SCALED with k-dots: keep the item's K tile in the second TMEM accumulator buffer while the
                              // matrices that carry a k-dot are processed (single-buffered on the first)
  int kseg, kseg2;            // k-blocks per MMA chain (0 = the whole reduction in one chain) for the matrices with a
                              // k-dot (< ndot) / without: the reduction over M is cut into segments that the epilogue
                              // folds with round-to-nearest FMAs (see "work decomposition")
  int64_t n_items;
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

// 32 columns [c, c + 32) of row i of the fp16 K planes dotted with v (columns >= M masked by the caller's zero v)
__device__ __forceinline__ float dot_k_planes(const __half* __restrict__ Kh, const __half* __restrict__ Kl, int64_t ldkh,
                                              int64_t i, int64_t c, const float (&v)[32], float acc) {
  const uint4* ph = reinterpret_cast<const uint4*>(Kh + i * ldkh + c);
  const uint4* pl = reinterpret_cast<const uint4*>(Kl + i * ldkh + c);
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    if (c + 8 * g + 8 <= ldkh) {
      const uint4 h = __ldg(ph + g), l = __ldg(pl + g);
      acc = pair_dot2(h.x, l.x, v[8 * g + 0], v[8 * g + 1], acc);
      acc = pair_dot2(h.y, l.y, v[8 * g + 2], v[8 * g + 3], acc);
      acc = pair_dot2(h.z, l.z, v[8 * g + 4], v[8 * g + 5], acc);
      acc = pair_dot2(h.w, l.w, v[8 * g + 6], v[8 * g + 7], acc);
    }
  }
  return acc;
}

// the same dot through the datapoint-blocked TRANSPOSED planes Kt[n / 64][m][n % 64]: the 32 lanes of a warp are 32
// consecutive datapoints, so one column of the block is one 64-byte run -- 1 L1 wavefront per (column, plane) instead
// of the 32 of the row-major planes (every lane on its own 2 KB-strided row).  The SCALED epilogue repeats this dot for
// every k-segment of every matrix that carries a k-dot, so its LSU cost decides how short the MMA chains can be.
__device__ __forceinline__ float dot_kt_planes(const __half* __restrict__ Kth, const __half* __restrict__ Ktl, int64_t ldkt,
                                               int64_t i, int64_t c, int64_t M, const float (&v)[32], float acc) {
  const int64_t o = (i >> 6) * ldkt + c * 64 + (i & 63);
  const __half* ph = Kth + o;
  const __half* pl = Ktl + o;
  if (c + 32 <= M) {
#pragma unroll
    for (int j = 0; j < 32; ++j) acc = fmaf(v[j], __half2float(__ldg(ph + j * 64)) + __half2float(__ldg(pl + j * 64)), acc);
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (c + j < M) acc = fmaf(v[j], __half2float(__ldg(ph + j * 64)) + __half2float(__ldg(pl + j * 64)), acc);
  }
  return acc;
}

// CL = CTAs per cluster (SYRK only: 2).  The two CTAs of a cluster work on the same (tile pair, super-chunk) for two
// adjacent channels: they need the SAME K^T boxes, so each loads half of the rows of every box and TMA-multicasts
// them into both shared memories -- half the L2 -> SM traffic per CTA.  mapH_* are the half-height (64-row) boxes.
template <int MODE, int BN, int BK, int STAGES, int CL>
__global__ void __launch_bounds__(tc_threads(MODE), 1)
tc_kernel(const __grid_constant__ CUtensorMap mapA_hi, const __grid_constant__ CUtensorMap mapA_lo,
          const __grid_constant__ CUtensorMap mapB_hi, const __grid_constant__ CUtensorMap mapB_lo,
          const __grid_constant__ CUtensorMap mapH_hi, const __grid_constant__ CUtensorMap mapH_lo, const TcParams P) {
  static_assert(CL == 1 || (CL == 2 && MODE == MODE_SYRK), "clusters are used by the SYRK only");
  constexpr int RB = BK * 2;                                   // bytes per operand row of one k-block
  constexpr int A_BYTES = BLOCK_M * RB, B_BYTES = BN * RB;     // one plane
  constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  constexpr bool HAS_XFORM = (MODE == MODE_SYRK);
  constexpr int EPI_WARPS = (MODE == MODE_QUAD) ? 4 : 8;
  constexpr int EPI_WARP0 = (MODE == MODE_SYRK) ? 8 : (MODE == MODE_SCALED ? 4 : 2);   // SYRK: warpgroups 2-3, SCALED: 1-2
  constexpr int XF_WARP0 = 4;                                  // SYRK: warpgroup 1 transforms the A operand
  constexpr int ACC_BUFS = (2 * BN <= 512) ? 2 : 1;
  constexpr uint32_t IDESC = make_idesc(BN);
  static_assert(RB == 128 || RB == 64, "one swizzle row per k-block row");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;      // shared-window address, 1024-byte aligned
  const uint32_t bars = smem + STAGES * STAGE_BYTES;
  auto full = [&](int s) { return bars + 8u * s; };
  auto ready = [&](int s) { return bars + 8u * (STAGES + s); };
  auto empty = [&](int s) { return bars + 8u * (2 * STAGES + s); };
  auto tmem_full = [&](int t) { return bars + 8u * (3 * STAGES + t); };
  auto tmem_empty = [&](int t) { return bars + 8u * (3 * STAGES + 2 + t); };
  const uint32_t tmem_ptr_addr = bars + 8u * (3 * STAGES + 4);
  auto fullA = [&](int s) { return bars + 8u * (3 * STAGES + 5 + s); };    // SYRK: the A planes land (and are transformed) first

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = (CL == 2) ? cluster_ctarank() : 0u;
  constexpr uint16_t MC_MASK = (uint16_t)((1u << CL) - 1u);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full(s), 1);
      mbar_init(fullA(s), 1);
      mbar_init(ready(s), 4);
      mbar_init(empty(s), CL);          // a stage is free when the MMAs of EVERY CTA of the cluster have read it
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(tmem_full(t), 1);
      mbar_init(tmem_empty(t), EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapA_hi); prefetch_tmap(&mapA_lo); prefetch_tmap(&mapB_hi); prefetch_tmap(&mapB_lo);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  if (CL == 2) cluster_sync_all();     // the peer's barriers exist before anything is multicast at them
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));

  // ---- work decomposition -------------------------------------------------------------------
  // Every role walks the same sequence: item -> sub-tiles -> k-blocks.  One TMEM accumulator buffer
  // holds one sub-tile; its MMA chain is kept short on purpose: the tensor core accumulates with
  // truncation, so a chain of n MMAs carries a bias of up to ~n * 2^-24.  Long reductions are cut
  // into sub-tiles that the epilogue folds with ordinary round-to-nearest adds.
  //   SYRK    item = (tile pair, channel)      sub = chunk of `chunk_rows` datapoints  -> double tile in L2
  //   QUAD    item = (row tile, channel)       sub = column tile of B_l                -> row sum in a register
  //   SCALED  item = (row tile, column tile)   sub = (stacked matrix, k-segment)       -> tile sum in registers
  const int64_t M = P.M;
  const int64_t Mc = (MODE == MODE_SCALED) ? P.Mc : M;          // output columns (rows of one B matrix)
  const int nct = (int)((Mc + BN - 1) / BN);
  const int kb_full = (int)((M + BK - 1) / BK);
  const int ksegA = (MODE == MODE_SCALED && P.kseg > 0 && P.kseg < kb_full) ? P.kseg : kb_full;
  const int ksegB = (MODE == MODE_SCALED && P.kseg2 > 0 && P.kseg2 < kb_full) ? P.kseg2 : kb_full;
  const int nsegA = (kb_full + ksegA - 1) / ksegA, nsegB = (kb_full + ksegB - 1) / ksegB;
  const int nsubA = (MODE == MODE_SCALED) ? (int)P.ndot * nsegA : 0;      // sub-tiles of the matrices that carry a k-dot
  struct SubInfo { int mat, k0, nkb; bool last; };
  // SCALED: sub -> (stacked matrix, first k-block and length of this segment of its reduction, last segment?)
  auto scaled_sub = [&](int sub) -> SubInfo {
    SubInfo r;
    int seg;
    if (sub < nsubA) {
      r.mat = sub / nsegA; seg = sub - r.mat * nsegA;
      r.k0 = seg * ksegA; r.nkb = kb_full - r.k0 < ksegA ? kb_full - r.k0 : ksegA; r.last = seg == nsegA - 1;
    } else {
      const int t = sub - nsubA, m = t / nsegB;
      r.mat = (int)P.ndot + m; seg = t - m * nsegB;
      r.k0 = seg * ksegB; r.nkb = kb_full - r.k0 < ksegB ? kb_full - r.k0 : ksegB; r.last = seg == nsegB - 1;
    }
    return r;
  };

  // TMEM accumulator buffer of sub-tile `sub`.  Normally the two buffers alternate (the epilogue of one sub-tile overlaps
  // the MMAs of the next).  SCALED with a K-tile cache: the matrices with a k-dot run single-buffered on buffer 0 while
  // buffer 1 holds the item's K tile (fp32, written once per item by the epilogue warps); the remaining matrices then
  // alternate again starting at buffer 0 -- whose release implies that every epilogue warp is done with the cache.
  const bool kcache = (MODE == MODE_SCALED) && P.kcache && nsubA > 0 && ACC_BUFS == 2;
  struct BufSel {
    int toggle = 0;
    uint32_t phase[2] = {0u, 0u};
  };
  auto pick_buf = [&](BufSel& b, int sub) -> int {
    if (kcache) return sub < nsubA ? 0 : ((sub - nsubA) & 1);
    const int r = b.toggle;
    b.toggle = (ACC_BUFS == 2) ? (b.toggle ^ 1) : 0;
    return r;
  };

  struct Item { int64_t l, itile, n0, n1; int a_row0, b_row0, tile; bool half_tile; };
  auto decode = [&](int64_t item) -> Item {
    Item it{0, 0, 0, 0, 0, 0, 0, false};
    if (MODE == MODE_SYRK) {
      // super-chunk major: all (tile pair, channel) items of one window of datapoints are scheduled together, so the
      // CTAs in flight stream the same <= ~48 MB slice of K^T out of L2 instead of thrashing it with M x N planes
      const int64_t per_sc = (int64_t)P.ntile * P.L;
      const int64_t sc = item / per_sc, rem = item - sc * per_sc;
      it.l = rem % P.L;
      it.tile = (int)(rem / P.L);
      int ta, tb;
      syrk_tile_decode(it.tile, M, BN, ta, tb);
      it.a_row0 = ta * BLOCK_M; it.b_row0 = tb * BN;
      // a tile whose right half lies entirely above the diagonal (128-row tile on the diagonal of a 256-column tile)
      // only needs its first BN / 2 columns: half the B rows are loaded and the MMAs run with N = BN / 2
      it.half_tile = (BN == 2 * BLOCK_M) && (it.b_row0 >= it.a_row0);
      it.n0 = sc * P.sc_rows;
      it.n1 = it.n0 + P.sc_rows < P.N ? it.n0 + P.sc_rows : P.N;
    } else if (MODE == MODE_QUAD) {
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
  auto item_subtiles = [&](const Item& it) -> int {
    if (MODE == MODE_SYRK) return (int)((it.n1 - it.n0 + P.chunk_rows - 1) / P.chunk_rows);
    if (MODE == MODE_QUAD) return nct;
    return nsubA + ((int)P.L - (int)P.ndot) * nsegB;      // SCALED: sub = (matrix, k-segment)
  };
  auto subtile_kblocks = [&](const Item& it, int sub) -> int {
    if (MODE == MODE_SYRK) {
      int64_t c0 = it.n0 + (int64_t)sub * P.chunk_rows;
      int64_t c1 = c0 + P.chunk_rows < it.n1 ? c0 + P.chunk_rows : it.n1;
      return (int)((c1 - c0 + BK - 1) / BK);
    } else if (MODE == MODE_QUAD) {
      if (!P.tri) return kb_full;
      int64_t kend = (int64_t)(sub + 1) * BN < M ? (int64_t)(sub + 1) * BN : M;
      return (int)((kend + BK - 1) / BK);
    }
    return scaled_sub(sub).nkb;
  };

  auto role_producer = [&]() {
    // =============================== TMA producer ===============================================
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int64_t item = blockIdx.x; item < P.n_items; item += gridDim.x) {
        const Item it = decode(item);
        const int nsub = item_subtiles(it);
        for (int sub = 0; sub < nsub; ++sub) {
          const int nkb = subtile_kblocks(it, sub);
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(empty(stage), phase ^ 1);
            const uint32_t st = smem + stage * STAGE_BYTES;
            int32_t ak, ar, bk, br;
            if (MODE == MODE_SYRK) {
              ak = (int32_t)(it.n0 + (int64_t)sub * P.chunk_rows) + kb * BK; ar = it.a_row0; bk = ak; br = it.b_row0;
            } else if (MODE == MODE_QUAD) {
              ak = kb * BK; ar = it.a_row0; bk = ak; br = (int32_t)(it.l * M + (int64_t)sub * BN);
            } else {
              const SubInfo si = scaled_sub(sub);
              ak = (si.k0 + kb) * BK; ar = it.a_row0; bk = ak; br = (int32_t)((int64_t)si.mat * Mc + it.b_row0);
            }
            if (MODE == MODE_SYRK) {
              // datapoint-blocked transposed planes [n / 64][m][n % 64]: a box is one contiguous run of rows.
              // The A planes get their own barrier and go first: the transform warps rescale them while the
              // (twice as large) B planes are still in flight.
              const int32_t nb = ak >> 6, ni = ak & 63;
              if (CL == 2) {
                // this CTA fetches rows [crank * h, (crank + 1) * h) of every box and multicasts them to both CTAs;
                // each CTA's barriers still expect the full box (its own half + the peer's)
                const int32_t ha = crank * (BLOCK_M / 2);
                const uint32_t oa = crank * (A_BYTES / 2);
                mbar_expect_tx(fullA(stage), 2 * A_BYTES);
                tma_load_3d_mc(st + oa, &mapH_hi, fullA(stage), ni, ar + ha, nb, MC_MASK);
                tma_load_3d_mc(st + A_BYTES + oa, &mapH_lo, fullA(stage), ni, ar + ha, nb, MC_MASK);
                if (it.half_tile) {
                  mbar_expect_tx(full(stage), 2 * A_BYTES);
                  tma_load_3d_mc(st + 2 * A_BYTES + oa, &mapH_hi, full(stage), ni, br + ha, nb, MC_MASK);
                  tma_load_3d_mc(st + 2 * A_BYTES + B_BYTES + oa, &mapH_lo, full(stage), ni, br + ha, nb, MC_MASK);
                } else {
                  const int32_t hb = crank * (BN / 2);
                  const uint32_t ob = crank * (B_BYTES / 2);
                  mbar_expect_tx(full(stage), 2 * B_BYTES);
                  tma_load_3d_mc(st + 2 * A_BYTES + ob, &mapA_hi, full(stage), ni, br + hb, nb, MC_MASK);    // BN / 2 == BLOCK_M rows
                  tma_load_3d_mc(st + 2 * A_BYTES + B_BYTES + ob, &mapA_lo, full(stage), ni, br + hb, nb, MC_MASK);
                }
              } else {
              mbar_expect_tx(fullA(stage), 2 * A_BYTES);
              tma_load_3d(st, &mapA_hi, fullA(stage), ni, ar, nb);
              tma_load_3d(st + A_BYTES, &mapA_lo, fullA(stage), ni, ar, nb);
              if (it.half_tile) {                       // BN / 2 == BLOCK_M rows: the A maps have the right box
                mbar_expect_tx(full(stage), 2 * A_BYTES);
                tma_load_3d(st + 2 * A_BYTES, &mapA_hi, full(stage), ni, br, nb);
                tma_load_3d(st + 2 * A_BYTES + B_BYTES, &mapA_lo, full(stage), ni, br, nb);
              } else {
                mbar_expect_tx(full(stage), 2 * B_BYTES);
                tma_load_3d(st + 2 * A_BYTES, &mapB_hi, full(stage), ni, br, nb);
                tma_load_3d(st + 2 * A_BYTES + B_BYTES, &mapB_lo, full(stage), ni, br, nb);
              }
              }
            } else {
              mbar_expect_tx(full(stage), STAGE_BYTES);
              tma_load_2d(st, &mapA_hi, full(stage), ak, ar);
              tma_load_2d(st + A_BYTES, &mapA_lo, full(stage), ak, ar);
              tma_load_2d(st + 2 * A_BYTES, &mapB_hi, full(stage), bk, br);
              tma_load_2d(st + 2 * A_BYTES + B_BYTES, &mapB_lo, full(stage), bk, br);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  };
  auto role_mma = [&]() {
    // =============================== MMA issuer ==================================================
    int stage = 0; uint32_t phase = 0;
    BufSel bs;
    for (int64_t item = blockIdx.x; item < P.n_items; item += gridDim.x) {
      const Item it = decode(item);
      const int nsub = item_subtiles(it);
      const uint32_t idesc = it.half_tile ? make_idesc(BN / 2) : IDESC;
      for (int sub = 0; sub < nsub; ++sub) {
        const int nkb = subtile_kblocks(it, sub);
        const int acc = pick_buf(bs, sub);
        mbar_wait(tmem_empty(acc), bs.phase[acc] ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(full(stage), phase);
          if (HAS_XFORM) mbar_wait(ready(stage), phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = smem + stage * STAGE_BYTES;
            const uint64_t a_hi = make_smem_desc<RB>(sa), a_lo = make_smem_desc<RB>(sa + A_BYTES);
            const uint64_t b_hi = make_smem_desc<RB>(sa + 2 * A_BYTES), b_lo = make_smem_desc<RB>(sa + 2 * A_BYTES + B_BYTES);
#pragma unroll
            for (int ks = 0; ks < BK / UMMA_K; ++ks) {
              const uint64_t adv = (uint64_t)((ks * UMMA_K * 2) >> 4);
              umma_f16(d_tmem, a_lo + adv, b_hi + adv, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
              umma_f16(d_tmem, a_hi + adv, b_lo + adv, idesc, 1u);
              umma_f16(d_tmem, a_hi + adv, b_hi + adv, idesc, 1u);
            }
            if (CL == 2) umma_commit_mc(empty(stage), MC_MASK); else umma_commit(empty(stage));
            if (kb == nkb - 1) umma_commit(tmem_full(acc));
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        bs.phase[acc] ^= 1;
      }
    }
  };
  auto role_epilogue = [&]() {
    // =============================== epilogue ====================================================
    const int qd = warp & 3;                         // TMEM lane quarter this warp may access
    const int row = qd * 32 + lane;                  // output row inside the tile
    const int half = (warp - EPI_WARP0) >> 2;        // SYRK / SCALED: which half of the tile's columns
    const float inv_ks = P.kscale[1];
    BufSel bsel;
    for (int64_t item = blockIdx.x; item < P.n_items; item += gridDim.x) {
      const Item it = decode(item);
      const int nsub = item_subtiles(it);
      if constexpr (MODE == MODE_SYRK) {
        // The tile's running sum lives in REGISTERS: warp (quarter qd, column half) owns rows [a0, a0 + 32) x 128
        // columns, one row per thread, 128 fp32 sums.  Every chunk (one short, truncating MMA chain in TMEM) is
        // folded with round-to-nearest adds; every `flush_every` chunks -- and at the end of the item -- the sums
        // are added into the float64 tile of A_l (owned by this CTA: plain read-modify-write, lower triangle only)
        // and cleared.  Per chunk the epilogue costs 4 tcgen05.ld + 128 FADD per thread; global traffic is
        // 1 / flush_every of a per-chunk write-back.
        const double sc = (double)inv_ks * (double)inv_ks * (double)P.binv[it.l];
        const int64_t a0 = (int64_t)it.a_row0 + qd * 32;                   // first row of this warp's block
        const int64_t r = a0 + lane;                                       // this thread's output row
        const int64_t cw0 = (int64_t)it.b_row0 + half * (BN / 2);          // first column of this warp
        int nlive = 0;                                                     // 32-column chunks touching the lower triangle
        if (a0 < M) {
#pragma unroll
          for (int ch = 0; ch < BN / 64; ++ch)
            if (cw0 + 32 * ch <= a0 + 31 && cw0 + 32 * ch < M) nlive = ch + 1;
        }
        float run[BN / 64][32];
#pragma unroll
        for (int ch = 0; ch < BN / 64; ++ch)
#pragma unroll
          for (int j = 0; j < 32; ++j) run[ch][j] = 0.f;
        int pending = 0;
        // Truncation-bias correction.  The tensor core truncates towards zero when it aligns and adds into its fp32
        // accumulator; when all terms of a chain have one sign (weights >= 0 and an element-wise non-negative kernel:
        // the forward A_l of the SE / periodic kernels) the partial sum grows monotonically from 0 to P and the loss
        // is a fixed fraction of P: one truncation of ulp(P k / n) / 2 per MMA would give
        //   sum_k ulp(P k / n) / 2 = n 2^-24 P (f - 2/3) / f^2 = (0.33 .. 0.375) n 2^-24 P   (f = significand of P),
        // and the hardware loses about twice that (products are truncated individually).  MEASURED on this part
        // (tests/probes/accum_probe.py, profiles/r01_syrk_shrink_calibration.jsonl): A_tc = (1 - beta) A with
        // beta / n = 3.86e-8, 3.97e-8, 4.00e-8, 4.04e-8 for n = 48, 96, 192, 384 MMAs, i.e. 0.666 n 2^-24, independent
        // of the entry.  The factor (1 + 0.666 n 2^-24) on every chunk's partial sum removes it (S_l = (K + c A_l + J)^-1
        // amplifies a relative error of A_l by the condition number).  Chains with mixed signs (adjoint SYRKs,
        // linear kernels) are left alone.
        const bool one_signed = (P.kscale[6] != 0.f) && (P.wneg[it.l] == 0.f);
        for (int sub = 0; sub < nsub; ++sub) {
          const float fix = one_signed ? 1.0f + P.bias_coef * 5.9604645e-8f * (float)(subtile_kblocks(it, sub) * (BK / UMMA_K) * 3) : 1.0f;
          const int acc = pick_buf(bsel, sub);
          mbar_wait(tmem_full(acc), bsel.phase[acc]);
          tc_fence_after();
          const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(acc * BN + half * (BN / 2));
#pragma unroll
          for (int ch = 0; ch < BN / 64; ++ch) {
            if (ch < nlive) {
              float v[32];
              tmem_ld32(taddr + ch * 32, v);
#pragma unroll
              for (int j = 0; j < 32; ++j) run[ch][j] = fmaf(v[j], fix, run[ch][j]);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty(acc));
          bsel.phase[acc] ^= 1;
          if (++pending == P.flush_every || sub == nsub - 1) {
            pending = 0;
            if (nlive > 0) {
              // Other CTAs add other super-chunks of the same tile: the 32 x 128 block of this warp is guarded by a
              // spin lock (contention is rare: same-tile items are ntile * L items apart) and accessed through L2 only.
              // the float64 tiles stream through L2 once per super-chunk: evict-first + no L1 allocation, so that they
              // do not push the K^T slice (re-read by every tile pair of the super-chunk) out of the cache
              uint64_t pol;
              asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
              int* lock = P.locks + ((int64_t)it.tile * P.L + it.l) * 8 + (warp - EPI_WARP0);
              if (lane == 0) {
                while (atomicCAS(lock, 0, 1) != 0) __nanosleep(100);
                __threadfence();
              }
              __syncwarp();
#pragma unroll
              for (int ch = 0; ch < BN / 64; ++ch) {
                if (ch < nlive) {
                  const int64_t c0 = cw0 + 32 * ch;
                  double* dst = P.A + (it.l * M + r) * M + c0;
                  const int64_t nv64 = (r < M) ? r - c0 + 1 : 0;           // columns c0 .. min(c0 + 31, r) of row r
                  const int nv = nv64 > 32 ? 32 : (int)nv64;
#pragma unroll
                  for (int j0 = 0; j0 < 32; j0 += 8) {
                    double t[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) t[j] = (j0 + j < nv) ? ld_stream_f64(dst + j0 + j, pol) : 0.0;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                      if (j0 + j < nv) st_stream_f64(dst + j0 + j, t[j] + (double)run[ch][j0 + j] * sc, pol);
                  }
#pragma unroll
                  for (int j = 0; j < 32; ++j) run[ch][j] = 0.f;
                }
              }
              __threadfence();
              __syncwarp();
              if (lane == 0) atomicExch(lock, 0);
            }
          }
        }
      } else if (MODE == MODE_QUAD) {
        const int64_t i = it.itile * BLOCK_M + row;
        const bool live = (i < P.N) && (it.l < P.L);
        const float bs = (it.l < P.L) ? P.binv[it.l] : 0.f;
        float qsum = 0.f;
        for (int sub = 0; sub < nsub; ++sub) {
          const int acc = pick_buf(bsel, sub);
          mbar_wait(tmem_full(acc), bsel.phase[acc]);
          tc_fence_after();
          const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
          for (int c0 = 0; c0 < BN; c0 += 32) {
            const int64_t cbase = (int64_t)sub * BN + c0;
            if (cbase >= M) break;
            float v[32];
            tmem_ld32(taddr + c0, v);
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (cbase + j >= M) v[j] = 0.f;
            if (P.tri) {
#pragma unroll
              for (int j = 0; j < 32; ++j) qsum = fmaf(v[j], v[j], qsum);
            } else if (live) {
              qsum = dot_k_planes(P.K_hi, P.K_lo, P.ldkh, i, cbase, v, qsum);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty(acc));
          bsel.phase[acc] ^= 1;
        }
        if (live) {
          const float s1 = inv_ks * bs;
          P.q[i * P.ldq + it.l] = P.tri ? qsum * s1 * s1 : qsum * s1 * inv_ks;
        }
      } else {
        constexpr int NCH = BN / 64;                   // 32-column chunks owned by this warp
        const int64_t i = it.itile * BLOCK_M + row;
        const bool live = i < P.N;
        float run[NCH][32];
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
          for (int j = 0; j < 32; ++j) run[ch][j] = 0.f;
        float dsum = 0.f;
        const uint32_t tcache = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(BN + half * (BN / 2));
        if (kcache) {
          // this thread's row of the item's K tile (its half of the columns), fp32 hi + lo in plane units, into buffer 1:
          // read back with tcgen05.ld for every k-segment of every matrix that carries a k-dot -- no L2 traffic
#pragma unroll
          for (int ch = 0; ch < NCH; ++ch) {
            const int64_t cbase = (int64_t)it.b_row0 + half * (BN / 2) + ch * 32;
            float kv[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) kv[j] = 0.f;
            if (live && cbase < M) {
              const int64_t o = (i >> 6) * P.ldkt + cbase * 64 + (i & 63);
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (cbase + j < M) kv[j] = __half2float(__ldg(P.Kt_hi + o + j * 64)) + __half2float(__ldg(P.Kt_lo + o + j * 64));
            }
            __syncwarp();                                          // reconverge: tcgen05.st is .sync.aligned
            tmem_st32(tcache + ch * 32, kv);
          }
        }
        for (int sub = 0; sub < nsub; ++sub) {
          const SubInfo si = scaled_sub(sub);
          const int mat = si.mat;                                  // stacked matrix; this sub-tile = one k-segment of its reduction
          // mixed-sign chains: the truncating accumulation shrinks K G_s uniformly by 0.263 .. 0.28 n 2^-24 on this workload
          // class (profiles/r01_scaled_shrink_calibration.jsonl).  Removing that uniform part (SVGP_SCALED_BIAS=0.27) was
          // measured to change NO parity figure (profiles/r01_parity_scaled_bias_sweep.jsonl): what hurts the
          // inducing-point gradient is the non-uniform part of the truncation error.  Default: no correction.
          const float fix = 1.0f + P.bias_coef * 5.9604645e-8f * (float)(si.nkb * (BK / UMMA_K) * 3);
          const float bs = P.binv[mat] * fix;
          const float wgt = live ? (P.W ? P.W[i * P.ldw + mat] : 1.f) * inv_ks * bs : 0.f;
          const bool want_dot = (P.dots != nullptr) && (mat < P.ndot) && live;
          if (si.k0 == 0) dsum = 0.f;
          const int acc = pick_buf(bsel, sub);
          mbar_wait(tmem_full(acc), bsel.phase[acc]);
          tc_fence_after();
          const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(acc * BN + half * (BN / 2));
#pragma unroll
          for (int ch = 0; ch < NCH; ++ch) {
            const int64_t cbase = (int64_t)it.b_row0 + half * (BN / 2) + ch * 32;
            if (cbase < Mc) {
              float v[32];
              tmem_ld32(taddr + ch * 32, v);
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                if (cbase + j >= Mc) v[j] = 0.f;
                run[ch][j] = fmaf(wgt, v[j], run[ch][j]);
              }
              if (kcache && P.dots != nullptr && mat < P.ndot) {      // warp-uniform: tcgen05.ld is .sync.aligned
                float kv[32];
                tmem_ld32(tcache + ch * 32, kv);
#pragma unroll
                for (int j = 0; j < 32; ++j) dsum = fmaf(v[j], kv[j], dsum);
              } else if (want_dot) {
                {
                  dsum = P.Kt_hi ? dot_kt_planes(P.Kt_hi, P.Kt_lo, P.ldkt, i, cbase, M, v, dsum)
                                 : dot_k_planes(P.K_hi, P.K_lo, P.ldkh, i, cbase, v, dsum);
                }
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty(acc));
          bsel.phase[acc] ^= 1;
          if (want_dot && si.last) atomicAdd(&P.dots[i * P.lddots + mat], dsum * inv_ks * inv_ks * bs);
        }
        if (live) {
#pragma unroll
          for (int ch = 0; ch < NCH; ++ch) {
            const int64_t cbase = (int64_t)it.b_row0 + half * (BN / 2) + ch * 32;
            float* o = P.out + i * P.ldo + cbase;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (cbase + j < Mc) o[j] = P.accumulate ? o[j] + run[ch][j] : run[ch][j];
          }
        }
      }
    }
  };
  auto role_transform = [&]() {
    // =============================== operand transform (SYRK) ====================================
    // The A tile (128 rows a x BK datapoints n) is rescaled by the channel's weights w[n] in place.  A thread
    // owns one logical 16-byte chunk (8 consecutive n) of CPR rows: its 8 weights are loaded once per k-block.
    // Swizzle: the physical chunk is the logical one XORed with (row & 7) [128-byte rows] or ((row >> 1) & 3)
    // [64-byte rows]; both are constant over the rows a thread visits, and a quarter-warp always covers
    // 128 contiguous bytes, so the 128-bit accesses are bank-conflict free.
    constexpr int CPR = RB / 16;                     // chunks per row
    constexpr int RSTEP = BLOCK_M / CPR;             // rows between two visits of a thread
    const int t = threadIdx.x - XF_WARP0 * 32;       // 0..127
    const int lchunk = t % CPR, rbase = t / CPR;
    const int pchunk = (RB == 128) ? (lchunk ^ (rbase & 7)) : (lchunk ^ ((rbase >> 1) & 3));
    int stage = 0; uint32_t phase = 0;
    for (int64_t item = blockIdx.x; item < P.n_items; item += gridDim.x) {
      const Item it = decode(item);
      const int nsub = item_subtiles(it);
      for (int sub = 0; sub < nsub; ++sub) {
        const int nkb = subtile_kblocks(it, sub);
        for (int kb = 0; kb < nkb; ++kb) {
          const int64_t n = it.n0 + (int64_t)sub * P.chunk_rows + (int64_t)kb * BK + lchunk * 8;
          float4 w0 = make_float4(0.f, 0.f, 0.f, 0.f), w1 = w0;
          if (n + 8 <= P.ldwt) {
            const float4* wp = reinterpret_cast<const float4*>(P.Wt + it.l * P.ldwt + n);
            w0 = __ldg(wp); w1 = __ldg(wp + 1);
          }
          // the thread's 8 weights as fp16 pairs (|w| <= 1 after the per-channel scaling)
          uint32_t wh[4], wl[4];
          split_weights(w0.x, w0.y, wh[0], wl[0]);
          split_weights(w0.z, w0.w, wh[1], wl[1]);
          split_weights(w1.x, w1.y, wh[2], wl[2]);
          split_weights(w1.z, w1.w, wh[3], wl[3]);
          mbar_wait(fullA(stage), phase);
          const uint32_t hi_p = smem + stage * STAGE_BYTES + rbase * RB + pchunk * 16;
          const uint32_t lo_p = hi_p + A_BYTES;
          if (P.debug & 2) {
            // reference variant: rescale in fp32 and re-split (3.5x the instructions of the packed-half path)
#pragma unroll
            for (int r = 0; r < CPR; ++r) {
              const uint32_t off = r * RSTEP * RB;
              uint4 h = lds128(hi_p + off), l = lds128(lo_p + off);
              rescale_pair(h.x, l.x, w0.x, w0.y);
              rescale_pair(h.y, l.y, w0.z, w0.w);
              rescale_pair(h.z, l.z, w1.x, w1.y);
              rescale_pair(h.w, l.w, w1.z, w1.w);
              sts128(hi_p + off, h);
              sts128(lo_p + off, l);
            }
          } else if (!(P.debug & 1)) {
#pragma unroll
            for (int r = 0; r < CPR; ++r) {
              const uint32_t off = r * RSTEP * RB;
              uint4 h = lds128(hi_p + off), l = lds128(lo_p + off);
              rescale_pair_h2(h.x, l.x, wh[0], wl[0]);
              rescale_pair_h2(h.y, l.y, wh[1], wl[1]);
              rescale_pair_h2(h.z, l.z, wh[2], wl[2]);
              rescale_pair_h2(h.w, l.w, wh[3], wl[3]);
              sts128(hi_p + off, h);
              sts128(lo_p + off, l);
            }
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(ready(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  };

  if constexpr (MODE == MODE_SYRK) {
    // 512 threads start with 128 registers each.  The TMA / MMA warpgroup and the transform warpgroup hand
    // registers to the two epilogue warpgroups, whose threads each keep 128 fp32 running sums of the output
    // tile.  setmaxnreg is executed by all four warps of a warpgroup at the top of that warpgroup's branch.
    const int wg = warp >> 2;
    if (wg == 0) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
      if (warp == 0) role_producer();
      else if (warp == 1) role_mma();
    } else if (wg == 1) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
      role_transform();
    } else {
      asm volatile("setmaxnreg.inc.sync.aligned.u32 192;");
      role_epilogue();
    }
  } else if constexpr (MODE == MODE_SCALED) {
    // 384 threads start with 168 registers each; the TMA / MMA warpgroup hands registers to the two epilogue
    // warpgroups (128 running sums + the k-dot operands per thread)
    if ((warp >> 2) == 0) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
      if (warp == 0) role_producer();
      else if (warp == 1) role_mma();
    } else {
      asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
      role_epilogue();
    }
  } else {
    if (warp == 0) role_producer();
    else if (warp == 1) role_mma();
    else if (warp >= EPI_WARP0 && warp < EPI_WARP0 + EPI_WARPS) role_epilogue();
  }

  tc_fence_before();
  __syncthreads();
  if (CL == 2) cluster_sync_all();     // nobody leaves while the peer may still multicast into / signal this CTA
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

int launch_mirror_lower(double* A, int64_t M, int64_t L, cudaStream_t st) {
  int64_t blocks = ceil_div(L * M * M, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks <= 0) return SVGP_OK;
  mirror_lower_kernel<<<(unsigned)blocks, 256, 0, st>>>(A, M, L);
  return check_launch("svgp_syrk(mirror)");
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// 2-D fp16 tensor (rows x cols, leading dimension ld elements), box = box_rows x bk columns, swizzle = row bytes
static int make_map(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int bk) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return SVGP_ERR_CUDA; }
  if (((uintptr_t)base & 15) || (ld * 2) % 16) { set_error("TMA operand needs 16-byte aligned base and row pitch"); return SVGP_ERR_ARG; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return SVGP_ERR_CUDA; }
  return SVGP_OK;
}

// datapoint-blocked transposed plane [nblk][rows][64] fp16 (block stride ldb elements), box = box_rows x bk datapoints
static int make_map_blocked(CUtensorMap* map, const void* base, int64_t rows, int64_t N, int64_t ldb, int box_rows, int bk) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return SVGP_ERR_CUDA; }
  if (((uintptr_t)base & 15) || (ldb * 2) % 16 || ldb < rows * 64) { set_error("TMA operand: bad blocked transposed plane"); return SVGP_ERR_ARG; }
  cuuint64_t dims[3] = {64, (cuuint64_t)rows, (cuuint64_t)((N + 63) / 64)};
  cuuint64_t strides[2] = {128, (cuuint64_t)ldb * 2};
  cuuint32_t box[3] = {(cuuint32_t)bk, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (blocked) failed (%d)", (int)r); return SVGP_ERR_CUDA; }
  return SVGP_OK;
}

// k-block depth: 64 fp16 (128-byte swizzle rows, 2 smem stages of 96 KB; default: measured faster in every mode,
// profiles/r01_tc_probe_fp16_v1.jsonl) or 32 (64-byte rows, 4 stages of 48 KB) with SVGP_TC_BK=32.
static int tc_bk() {
  static int bk = 0;
  if (!bk) {
    const char* e = getenv("SVGP_TC_BK");
    bk = (e && atoi(e) == 32) ? 32 : 64;
  }
  return bk;
}

// chunks folded in fp32 registers (round-to-nearest adds) between two float64 write-backs of a SYRK tile: one
// write-back per 32768 datapoints whatever the chain length; SVGP_SYRK_FLUSH overrides (in chunks).
static int syrk_flush_every(int64_t chunk) {
  const char* e = getenv("SVGP_SYRK_FLUSH");
  if (e && atoi(e) > 0) return atoi(e);
  int64_t f = 32768 / chunk;
  return f < 1 ? 1 : (int)f;
}

template <int MODE, int BN, int BK, int STAGES, int CL>
static int launch_tc(const CUtensorMap* maps, const TcParams& P, cudaStream_t st, const char* name) {
  constexpr int SMEM_TOTAL = STAGES * (2 * BLOCK_M * BK * 2 + 2 * BN * BK * 2) + 256 + 1024;
  static_assert(SMEM_TOTAL <= TC_SMEM_LIMIT, "shared memory budget");
  auto kern = tc_kernel<MODE, BN, BK, STAGES, CL>;
  static bool attr_done = false;
  static int max_ctas = 0;             // CTAs that can be co-resident (CL == 2: 2 x the number of active clusters)
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  cfg.blockDim = dim3(tc_threads(MODE), 1, 1);
  cfg.dynamicSmemBytes = SMEM_TOTAL;
  cfg.stream = st;
  if (CL > 1) {
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
  }
  if (!attr_done) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL) != cudaSuccess) return check_launch(name);
    max_ctas = num_sms();
    if (CL > 1) {
      // a persistent grid must be fully co-resident: ask how many clusters fit (GPCs with an odd number of usable
      // SMs leave one SM without a partner)
      int nclusters = 0;
      cfg.gridDim = dim3(num_sms() / CL * CL, 1, 1);
      if (cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg) != cudaSuccess || nclusters <= 0) {
        cudaGetLastError();
        set_error("%s: no co-resident clusters of %d CTAs available", name, CL);
        return SVGP_ERR_CUDA;
      }
      max_ctas = nclusters * CL;
    }
    attr_done = true;
  }
  int64_t grid = P.n_items < max_ctas ? P.n_items : max_ctas;
  grid = grid / CL * CL;
  if (grid <= 0) return SVGP_OK;
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  if (cudaLaunchKernelEx(&cfg, kern, maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], P) != cudaSuccess) return check_launch(name);
  return check_launch(name);
}

template <int MODE>
static int dispatch_tc(int bk, const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi, const CUtensorMap& b_lo,
                       const TcParams& P, cudaStream_t st, const char* name) {
  const CUtensorMap maps[6] = {a_hi, a_lo, b_hi, b_lo, a_hi, a_lo};
  if (bk == 32) return launch_tc<MODE, 256, 32, 4, 1>(maps, P, st, name);
  return launch_tc<MODE, 256, 64, 2, 1>(maps, P, st, name);
}

bool tc_shape_ok(const svgp_kop* kop) {
  // worth it only when tiles are mostly full; TMA needs 16-byte pitches
  return kop->M >= 128 && kop->N >= 2048 && (kop->ldkh % 8) == 0 && (kop->ldkt % 8) == 0;
}

// super-chunk length: the K^T slice of one super-chunk (M rows x sc datapoints x hi/lo fp16) is what the CTAs in flight
// stream out of L2, and every item ends with a float64 read-modify-write of its tile (55 GB of DRAM traffic per call
// at 12288-row super-chunks).  Measured at (1e6, 1024, 64) (profiles/r01_syrk_superchunk_sweep.jsonl): 212 / 201 / 199 /
// 203 / 215 ms at 16384 / 24576 / 28672 / 32768 / 40960 rows (204 at 20480) -- the speed optimum is a slice of about
// the L2 size.  The chunk partial sums of an item are folded in fp32 registers, though, and the forward A_l feeds an
// ill-conditioned inverse: p_m against the float64 oracle went 4.7e-5 -> 8.1e-5 between 24 and 56 partials per item
// (M = 1024).  80 MB (20480 rows, 40 partials) keeps most of the speed and the accuracy margin.
static int64_t syrk_superchunk_rows(int64_t N, int64_t M, int64_t chunk) {
  const char* e = getenv("SVGP_SYRK_SC");
  int64_t sc = (e && atoll(e) > 0) ? atoll(e) : (80LL << 20) / (M * 4);
  // small problems: at least 8 items per tile, i.e. fewer fp32 partial sums per float64 write-back -- the write-back
  // traffic is irrelevant there and the few items do not average the fp32 rounding of long register sums
  if (!(e && atoll(e) > 0) && sc > N / 8) sc = N / 8 > 4 * chunk ? N / 8 : 4 * chunk;
  sc = sc / chunk * chunk;
  return sc < chunk ? chunk : sc;
}

int64_t tc_syrk_lock_words(int64_t M, int64_t L) { return (int64_t)syrk_tile_count(M, 256) * L * 8; }

int tc_syrk(const svgp_kop* kop, const float* Wt, int64_t ldwt, const float* winv, const float* wneg, int64_t L, double* A,
            int64_t chunk_rows, int* locks, cudaStream_t st) {
  if (!kop->Kth || !kop->Ktl || !kop->kscale) { set_error("tc_syrk: transposed fp16 planes missing"); return SVGP_ERR_ARG; }
  if (((uintptr_t)Wt & 15) || (ldwt % 8)) { set_error("tc_syrk: weights need 16-byte alignment"); return SVGP_ERR_ARG; }
  const int BN = 256, bk = tc_bk();
  CUtensorMap a_hi, a_lo, b_hi, b_lo;
  int rc;
  if ((rc = make_map_blocked(&a_hi, kop->Kth, kop->M, kop->N, kop->ldkt, BLOCK_M, bk))) return rc;
  if ((rc = make_map_blocked(&a_lo, kop->Ktl, kop->M, kop->N, kop->ldkt, BLOCK_M, bk))) return rc;
  if ((rc = make_map_blocked(&b_hi, kop->Kth, kop->M, kop->N, kop->ldkt, BN, bk))) return rc;
  if ((rc = make_map_blocked(&b_lo, kop->Ktl, kop->M, kop->N, kop->ldkt, BN, bk))) return rc;
  TcParams P{};
  P.N = kop->N; P.M = kop->M; P.L = L; P.kscale = kop->kscale; P.binv = winv;
  P.Wt = Wt; P.ldwt = ldwt; P.A = A; P.locks = locks; P.wneg = wneg;
  // truncation-bias correction of one-signed chains (see the SYRK epilogue); SVGP_SYRK_BIAS overrides the coefficient
  { const char* e = getenv("SVGP_SYRK_BIAS"); P.bias_coef = e ? (float)atof(e) : 0.666f; }
  // one accumulation chain = chunk / 16 k-steps x 3 MMAs.  The tensor core accumulates with truncation: measured on
  // A_l (all terms positive) the bias is -1.8e-7 x chunk / 1024 of the largest entry (tests/probes/accum_probe.py), and S_l =
  // (K + c A_l + J)^-1 amplifies it by the condition number.  512 rows = 96 MMAs per chain costs ~3 % of SYRK time.
  int64_t chunk = chunk_rows > 0 ? chunk_rows : 512;
  chunk = (chunk + 63) / 64 * 64;
  P.chunk_rows = chunk;
  P.sc_rows = syrk_superchunk_rows(kop->N, kop->M, chunk);
  P.nsc = (int)ceil_div(kop->N, P.sc_rows);
  P.ntile = syrk_tile_count(kop->M, BN);
  P.flush_every = syrk_flush_every(chunk);
  { const char* e = getenv("SVGP_TC_DEBUG"); P.debug = e ? atoi(e) : 0; }
  P.n_items = (int64_t)P.nsc * P.ntile * L;
  if (cudaMemsetAsync(locks, 0, sizeof(int) * tc_syrk_lock_words(kop->M, L), st) != cudaSuccess) return check_launch("svgp_syrk(locks)");
  // even channel counts run as clusters of two CTAs (adjacent channels of one tile pair) sharing their loads by TMA
  // multicast; SVGP_SYRK_CLUSTER=1 forces the single-CTA kernel
  const char* ecl = getenv("SVGP_SYRK_CLUSTER");
  const bool pair = (L % 2 == 0) && !(ecl && atoi(ecl) == 1);
  if (pair) {
    CUtensorMap h_hi, h_lo;
    if ((rc = make_map_blocked(&h_hi, kop->Kth, kop->M, kop->N, kop->ldkt, BLOCK_M / 2, bk))) return rc;
    if ((rc = make_map_blocked(&h_lo, kop->Ktl, kop->M, kop->N, kop->ldkt, BLOCK_M / 2, bk))) return rc;
    const CUtensorMap maps[6] = {a_hi, a_lo, b_hi, b_lo, h_hi, h_lo};
    rc = (bk == 32) ? launch_tc<MODE_SYRK, 256, 32, 4, 2>(maps, P, st, "svgp_syrk(tc, cluster)")
                    : launch_tc<MODE_SYRK, 256, 64, 2, 2>(maps, P, st, "svgp_syrk(tc, cluster)");
  } else {
    rc = dispatch_tc<MODE_SYRK>(bk, a_hi, a_lo, b_hi, b_lo, P, st, "svgp_syrk(tc)");
  }
  if (rc) return rc;
  int64_t blocks = ceil_div(L * kop->M * kop->M, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  mirror_lower_kernel<<<(unsigned)blocks, 256, 0, st>>>(A, kop->M, L);
  return check_launch("svgp_syrk(mirror)");
}

int tc_rowquad(const svgp_kop* kop, const void* S_hi, const void* S_lo, const float* S_inv, int64_t L, int tri, float* q,
               int64_t ldq, cudaStream_t st) {
  if (!kop->Kh || !kop->Kl || !kop->kscale) { set_error("tc_rowquad: fp16 planes missing"); return SVGP_ERR_ARG; }
  const int BN = 256, bk = tc_bk();
  CUtensorMap a_hi, a_lo, b_hi, b_lo;
  int rc;
  if ((rc = make_map(&a_hi, kop->Kh, kop->N, kop->M, kop->ldkh, BLOCK_M, bk))) return rc;
  if ((rc = make_map(&a_lo, kop->Kl, kop->N, kop->M, kop->ldkh, BLOCK_M, bk))) return rc;
  if ((rc = make_map(&b_hi, S_hi, L * kop->M, kop->M, kop->M, BN, bk))) return rc;
  if ((rc = make_map(&b_lo, S_lo, L * kop->M, kop->M, kop->M, BN, bk))) return rc;
  TcParams P{};
  P.N = kop->N; P.M = kop->M; P.L = L; P.kscale = kop->kscale; P.binv = S_inv;
  P.tri = tri; P.K_hi = (const __half*)kop->Kh; P.K_lo = (const __half*)kop->Kl; P.ldkh = kop->ldkh; P.q = q; P.ldq = ldq;
  // channels whose factor planes (2 * M*M*2 bytes each) share ~half of the 126 MB L2
  int64_t per = 2 * kop->M * kop->M * 2;
  const char* eg = getenv("SVGP_QUAD_L2MB");
  int64_t g = ((eg && atoll(eg) > 0 ? atoll(eg) : 64LL) << 20) / (per > 0 ? per : 1);
  if (g < 1) g = 1;
  if (g > L) g = L;
  while (L % g) --g;                       // keep groups uniform
  P.lgroup = (int)g;
  P.n_items = ceil_div(kop->N, BLOCK_M) * L;
  return dispatch_tc<MODE_QUAD>(bk, a_hi, a_lo, b_hi, b_lo, P, st, "svgp_rowquad(tc)");
}

int tc_scaled_gemm(const svgp_kop* kop, const float* W, int64_t ldw, const void* G_hi, const void* G_lo, const float* G_inv,
                   int64_t L, int64_t Mc, float* out, int64_t ldo, int accumulate, float* dots, int64_t lddots, int64_t ndot,
                   cudaStream_t st) {
  if (!kop->Kh || !kop->Kl || !kop->kscale) { set_error("tc_scaled_gemm: fp16 planes missing"); return SVGP_ERR_ARG; }
  if (dots && Mc != kop->M) { set_error("tc_scaled_gemm: the k-dots need square M x M matrices"); return SVGP_ERR_ARG; }
  const int BN = 256, bk = tc_bk();
  CUtensorMap a_hi, a_lo, b_hi, b_lo;
  int rc;
  if ((rc = make_map(&a_hi, kop->Kh, kop->N, kop->M, kop->ldkh, BLOCK_M, bk))) return rc;
  if ((rc = make_map(&a_lo, kop->Kl, kop->N, kop->M, kop->ldkh, BLOCK_M, bk))) return rc;
  if ((rc = make_map(&b_hi, G_hi, L * Mc, kop->M, kop->M, BN, bk))) return rc;       // rows past L * Mc are zero-filled by TMA
  if ((rc = make_map(&b_lo, G_lo, L * Mc, kop->M, kop->M, BN, bk))) return rc;
  TcParams P{};
  P.N = kop->N; P.M = kop->M; P.L = L; P.Mc = Mc; P.kscale = kop->kscale; P.binv = G_inv;
  P.K_hi = (const __half*)kop->Kh; P.K_lo = (const __half*)kop->Kl; P.ldkh = kop->ldkh;
  P.W = W; P.ldw = ldw; P.out = out; P.ldo = ldo; P.accumulate = accumulate;
  if (kop->Kth && kop->Ktl && !getenv("SVGP_SCALED_ROWDOT")) { P.Kt_hi = (const __half*)kop->Kth; P.Kt_lo = (const __half*)kop->Ktl; P.ldkt = kop->ldkt; }
  P.dots = dots; P.lddots = lddots; P.ndot = dots ? ndot : 0;
  // Chain length.  The tensor core accumulates with truncation (an error of ~n 2^-24 after n MMAs, relative to the LARGEST
  // partial sum of the chain -- and K (dA_l + dA_l^T) cancels heavily).  The reduction over M is cut into segments of
  // `kseg` k-blocks that the epilogue folds in fp32 registers with round-to-nearest FMAs: SVGP_SCALED_KSEG (matrices with
  // a k-dot: dA_l + dA_l^T) / SVGP_SCALED_KSEG2 (the others: S_l - Kinv); 0 = one chain.  Measured at M = 1024
  // (profiles/r01_parity_chain_matrix.jsonl, r01_parity_kcache.jsonl): segments of 4 k-blocks (48 MMAs) on the k-dot
  // matrices take the inducing-point gradient from 1.9e-3 to 3.4e-4 of its maximum, shorter ones and segmenting the
  // other family change nothing.  Every segment repeats the k-dot, i.e. re-reads the item's K tile: from L2 that costs
  // +75 % kernel time (the kernel already runs at L2 -> SM bandwidth); with the tile cached in the second TMEM buffer
  // (kcache: those matrices run single-buffered on the first) it costs +10 % (610 -> 671 ms at N = 1e6, L = 64).
  { const char* e = getenv("SVGP_SCALED_KCACHE"); P.kcache = (P.Kt_hi != nullptr && P.ndot > 0) ? (e ? atoi(e) : 1) : 0; }
  { const char* e = getenv("SVGP_SCALED_KSEG"); P.kseg = e ? atoi(e) : (P.kcache ? 4 * 64 / bk : 0); }
  { const char* e = getenv("SVGP_SCALED_KSEG2"); P.kseg2 = e ? atoi(e) : 0; }
  { const char* e = getenv("SVGP_SCALED_BIAS"); P.bias_coef = e ? (float)atof(e) : 0.0f; }
  P.n_items = ceil_div(kop->N, BLOCK_M) * ceil_div(Mc, BN);
  return dispatch_tc<MODE_SCALED>(bk, a_hi, a_lo, b_hi, b_lo, P, st, "svgp_scaled_gemm(tc)");
}

}  // namespace svgp
