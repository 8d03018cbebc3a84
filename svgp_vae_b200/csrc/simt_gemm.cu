// Generic fp32 CUDA-core implementations of the GEMM-class entry points (any shape).
// They serve the reference's own small configurations (BALL m=15, MNIST m=32, SPRITES m=72..500,
// where a 128x256 tensor-core tile would be mostly padding) and are the on-device cross-check of
// the tcgen05 path (tests/test_tc_engine.py).  One 64x64x16 register-tiled kernel, parameterised
// by an operand functor, covers:
//   svgp_syrk        A_l  = sum_i W[i,l] k_i k_i^T                    (SVGPVAE_model.py:328-330)
//   svgp_gemm_tn     V_l  = sum_i X[i,l] k_i                          (:333-334)
//   svgp_gemm_nn     out  = K_nm Wm^T                                 (:332, :264-265)
//   svgp_scaled_gemm out  = sum_l diag(W[:,l]) K_nm G_l               (adjoint of the two above)
//   svgp_gemm_f32    plain C (+)= A B
// and a sibling kernel does the row-wise quadratic forms svgp_rowquad (:336-337, :284).
// Operands are fp32, every product and accumulation is fp64 (the S_l / Kinv operands have large
// cancelling entries), reductions over datapoints are folded across chunks with double atomics.
#include <cuda_fp16.h>

#include "common.cuh"

namespace svgp {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

template <class Op>
__global__ void __launch_bounds__(NT) tile_gemm_kernel(Op op) {
  using BT = typename Op::BT;                 // element type of the B operand (double for the M x M matrices)
  __shared__ float As[BK][BM + 4];
  __shared__ BT Bs[BK][BN + 4];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
  int64_t k0, k1;
  op.krange(blockIdx.z, k0, k1);
  // fp32 products; accumulation type per operator: double where the operand matrices have large cancelling entries
  // (S_l / Kinv, cond ~1e3..1e5), float for the short / chunked sums whose partials are folded in double afterwards
  using Acc = typename Op::Acc;
  Acc acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = (Acc)0;

  for (int64_t kb = k0; kb < k1; kb += BK) {
    // A tile: BK x BM, B tile: BK x BN; each thread loads 4 + 4 elements
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int idx = threadIdx.x + e * NT;              // 0..1023
      int kk, mm;
      if (Op::A_K_CONTIG) { kk = idx % BK; mm = idx / BK; } else { mm = idx % BM; kk = idx / BM; }
      int64_t gk = kb + kk, gm = m0 + mm;
      As[kk][mm] = (gk < k1 && gm < op.Mr) ? op.a(blockIdx.z, gm, gk) : 0.f;
      int nn;
      if (Op::B_K_CONTIG) { kk = idx % BK; nn = idx / BK; } else { nn = idx % BN; kk = idx / BN; }
      gk = kb + kk;
      int64_t gn = n0 + nn;
      Bs[kk][nn] = (gk < k1 && gn < op.Nc) ? op.b(blockIdx.z, gk, gn) : (BT)0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      Acc a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = (Acc)As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = (Acc)Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int64_t gm = m0 + ty * 4 + i, gn = n0 + tx * 4 + j;
      if (gm < op.Mr && gn < op.Nc) op.store(blockIdx.z, gm, gn, acc[i][j]);
    }
}

// ---- operand functors -------------------------------------------------------------------------
struct SyrkOp {     // z = l * nchunk + chunk
  using BT = float;
  using Acc = double;
  static constexpr bool A_K_CONTIG = false, B_K_CONTIG = false;
  const float* K; int64_t ldk; const float* W; int64_t ldw; double* A; int64_t N, M, L, chunk, nchunk;
  int64_t Mr, Nc;
  __device__ void krange(int z, int64_t& k0, int64_t& k1) const {
    int64_t c = z % nchunk; k0 = c * chunk; k1 = min(N, k0 + chunk);
  }
  __device__ float a(int z, int64_t m, int64_t k) const { return K[k * ldk + m]; }
  __device__ float b(int z, int64_t k, int64_t n) const { return W[k * ldw + z / nchunk] * K[k * ldk + n]; }
  __device__ void store(int z, int64_t m, int64_t n, double v) const {
    atomicAdd(&A[((int64_t)(z / nchunk) * M + m) * M + n], v);
  }
};
// K_nm element for the skinny products: plain fp32, or reassembled from the fp16 planes of a TC operand
struct KAccess {
  const float* K; const __half* Kh; const __half* Kl; int64_t ld; const float* kscale;
  __device__ float operator()(int64_t i, int64_t a) const {
    if (K) return K[i * ld + a];
    return (__half2float(Kh[i * ld + a]) + __half2float(Kl[i * ld + a])) * kscale[1];
  }
};
template <class AccT>
struct TnOpT {       // rows = channel l, cols = inducing index; z = chunk
  using BT = float;
  using Acc = AccT;
  static constexpr bool A_K_CONTIG = false, B_K_CONTIG = false;
  KAccess Ka; const float* X; int64_t ldx; double* V; int64_t N, M, chunk;
  int64_t Mr, Nc;
  __device__ void krange(int z, int64_t& k0, int64_t& k1) const { k0 = (int64_t)z * chunk; k1 = min(N, k0 + chunk); }
  __device__ float a(int z, int64_t m, int64_t k) const { return X[k * ldx + m]; }
  __device__ float b(int z, int64_t k, int64_t n) const { return Ka(k, n); }
  __device__ void store(int z, int64_t m, int64_t n, double v) const { atomicAdd(&V[m * M + n], v); }
};
struct NnOp {       // out[i,l] = sum_a K[i,a] Wm[l,a]
  using BT = float;
  using Acc = double;
  static constexpr bool A_K_CONTIG = true, B_K_CONTIG = true;
  KAccess Ka; int64_t row0; const float* Wm; int64_t ldwm; float* out; int64_t ldo; int64_t M;
  int64_t Mr, Nc;
  __device__ void krange(int z, int64_t& k0, int64_t& k1) const { k0 = 0; k1 = M; }
  __device__ float a(int z, int64_t m, int64_t k) const { return Ka(row0 + m, k); }
  __device__ float b(int z, int64_t k, int64_t n) const { return Wm[n * ldwm + k]; }
  __device__ void store(int z, int64_t m, int64_t n, double v) const { out[(row0 + m) * ldo + n] = (float)v; }
};
struct ScaledOp {   // out[i,c] = sum_{l,a} W[i,l] K[i,a] G[l,a,c]
  using BT = double;
  using Acc = double;
  static constexpr bool A_K_CONTIG = true, B_K_CONTIG = false;
  const float* K; int64_t ldk; const float* W; int64_t ldw; const double* G; float* out; int64_t ldo;
  int64_t M, L; int accumulate;
  int64_t Mr, Nc;
  __device__ void krange(int z, int64_t& k0, int64_t& k1) const { k0 = 0; k1 = L * M; }
  __device__ float a(int z, int64_t m, int64_t k) const {
    int64_t l = k / M, aa = k - l * M;
    return W[m * ldw + l] * K[m * ldk + aa];
  }
  __device__ double b(int z, int64_t k, int64_t n) const { return G[k * M + n]; }   // (l*M + a)*M + c
  __device__ void store(int z, int64_t m, int64_t n, double v) const {
    float* o = out + m * ldo + n;
    *o = accumulate ? (float)((double)*o + v) : (float)v;
  }
};
struct PlainOp {
  using BT = float;
  using Acc = float;
  static constexpr bool A_K_CONTIG = true, B_K_CONTIG = false;
  const float* A; int64_t lda; const float* B; int64_t ldb; float* C; int64_t ldc; int64_t Kd; int accumulate;
  int64_t Mr, Nc;
  __device__ void krange(int z, int64_t& k0, int64_t& k1) const { k0 = 0; k1 = Kd; }
  __device__ float a(int z, int64_t m, int64_t k) const { return A[m * lda + k]; }
  __device__ float b(int z, int64_t k, int64_t n) const { return B[k * ldb + n]; }
  __device__ void store(int z, int64_t m, int64_t n, double v) const {
    float* o = C + m * ldc + n;
    *o = accumulate ? (float)((double)*o + v) : (float)v;
  }
};

// ---- row-wise quadratic forms -------------------------------------------------------------------
// block = 64 datapoints x one channel; walks the column tiles of S_l, forms U = K_tile S_l[:, tile] in
// registers and folds it straight into q (DOT: sum_b U_ib K_ib; SQUARE/tri: sum_c U_ic^2).
__global__ void __launch_bounds__(NT) rowquad_simt_kernel(const float* __restrict__ K, int64_t ldk, int64_t N, int64_t M,
                                                          const double* __restrict__ S, int tri, float* __restrict__ q,
                                                          int64_t ldq) {
  __shared__ float As[BK][BM + 4];
  __shared__ double Bs[BK][BN + 4];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int l = blockIdx.y;
  const double* Sl = S + (int64_t)l * M * M;
  double qacc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int64_t n0 = 0; n0 < M; n0 += BN) {
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    const int64_t kend = tri ? min(M, n0 + BN) : M;     // Rinv[c][a] == 0 for a > c
    for (int64_t kb = 0; kb < kend; kb += BK) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        int idx = threadIdx.x + e * NT;
        int kk = idx % BK, mm = idx / BK;
        int64_t gk = kb + kk, gm = m0 + mm;
        As[kk][mm] = (gk < M && gm < N) ? K[gm * ldk + gk] : 0.f;
        int64_t gn;
        if (tri) { kk = idx % BK; int nn = idx / BK; gk = kb + kk; gn = n0 + nn;
                   Bs[kk][nn] = (gk < M && gn < M) ? Sl[gn * M + gk] : 0.0; }
        else     { int nn = idx % BN; kk = idx / BN; gk = kb + kk; gn = n0 + nn;
                   Bs[kk][nn] = (gk < M && gn < M) ? Sl[gk * M + gn] : 0.0; }
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        double a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = (double)As[kk][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = (double)Bs[kk][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int64_t gm = m0 + ty * 4 + i;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int64_t gn = n0 + tx * 4 + j;
        if (gm < N && gn < M) {
          double other = tri ? acc[i][j] : (double)K[gm * ldk + gn];
          qacc[i] = fma(acc[i][j], other, qacc[i]);
        }
      }
    }
  }
  // the 16 threads tx = 0..15 of one ty hold partial sums of the same 4 rows (lanes [0,16) / [16,32) of a warp)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double v = qacc[i];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    int64_t gm = m0 + ty * 4 + i;
    if (tx == 0 && gm < N) q[gm * ldq + l] = (float)v;
  }
}

// ---- fp16 operand planes ------------------------------------------------------------------------
// per-matrix max |x| (as float bits) -> scratch[b]
__global__ void absmax_f64_kernel(const double* __restrict__ x, int64_t count, float* __restrict__ scratch) {
  const double* xb = x + (int64_t)blockIdx.y * count;
  float m = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf((float)xb[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(reinterpret_cast<int*>(scratch + blockIdx.y), __float_as_int(m));
}
// scale 2^(target - e) for max = m 2^e (m in [0.5, 1)); 1 for an all-zero / non-finite matrix
__device__ __forceinline__ float pow2_scale(float mx, int target) {
  if (!(mx > 0.f) || !isfinite(mx)) return 1.0f;
  int e;
  frexpf(mx, &e);
  e = target - e;
  if (e > 100) e = 100;
  if (e < -100) e = -100;
  return ldexpf(1.0f, e);
}
__global__ void split_f16_kernel(const double* __restrict__ x, int64_t count, __half* __restrict__ hi, __half* __restrict__ lo,
                                 float* __restrict__ inv_scale, int64_t nb) {
  const int64_t b = blockIdx.y;
  const float s = pow2_scale(inv_scale[nb + b], 14);
  const double sd = (double)s;
  const double* xb = x + b * count;
  __half* hb = hi + b * count;
  __half* lb = lo + b * count;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = xb[i] * sd;
    const __half h = __float2half_rn((float)v);
    hb[i] = h;
    lb[i] = __float2half_rn((float)(v - (double)__half2float(h)));
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) inv_scale[b] = 1.0f / s;
}

// SYRK weights: column max of |W| -> per-channel power-of-two scale (|w s_l| <= 1), then the channel-major,
// zero-padded, pre-scaled copy Wt (L x ldwt) the operand transform reads with 128-bit loads
// (mx[L + l] becomes 1.0 if channel l has a negative weight: the SYRK's truncation-bias correction is only valid for
// chains whose terms all have one sign)
__global__ void colabsmax_kernel(const float* __restrict__ W, int64_t ldw, int64_t N, int64_t L, float* __restrict__ mx) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int64_t l0 = 0; l0 < L; l0 += 32) {
    const int64_t l = l0 + lane;
    float m = 0.f;
    bool neg = false;
    if (l < L)
      for (int64_t i = (int64_t)blockIdx.x * nwarp + warp; i < N; i += (int64_t)gridDim.x * nwarp) {
        const float w = W[i * ldw + l];
        m = fmaxf(m, fabsf(w));
        neg = neg || (w < 0.f);
      }
    if (l < L && m > 0.f) atomicMax(reinterpret_cast<int*>(mx + l), __float_as_int(m));
    if (l < L && neg) atomicMax(reinterpret_cast<int*>(mx + L + l), __float_as_int(1.0f));
  }
}
__global__ void transpose_scale_kernel(const float* __restrict__ W, int64_t ldw, int64_t N, int64_t L, const float* __restrict__ mx,
                                       float* __restrict__ Wt, int64_t ldwt, float* __restrict__ winv) {
  __shared__ float tile[32][33];
  const int64_t n0 = (int64_t)blockIdx.x * 32, l0 = (int64_t)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    int64_t n = n0 + r, l = l0 + tx;
    tile[r][tx] = (n < N && l < L) ? W[n * ldw + l] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    int64_t l = l0 + r, n = n0 + tx;
    if (l < L && n < ldwt) {
      const float s = pow2_scale(mx[l], 0);                         // max * s in [0.5, 1)
      Wt[l * ldwt + n] = tile[tx][r] * s;
      if (blockIdx.x == 0 && tx == 0) winv[l] = 1.0f / s;
    }
  }
}

template <class Op>
static int launch_tile(Op op, int64_t nz, cudaStream_t st, const char* name) {
  if (op.Mr == 0 || op.Nc == 0 || nz == 0) return SVGP_OK;
  dim3 grid((unsigned)ceil_div(op.Nc, BN), (unsigned)ceil_div(op.Mr, BM), (unsigned)nz);
  if (grid.y > 65535 || grid.z > 65535) { set_error("%s: grid too large for the SIMT path", name); return SVGP_ERR_UNSUPPORTED; }
  tile_gemm_kernel<Op><<<grid, NT, 0, st>>>(op);
  return check_launch(name);
}

// entry points of the tcgen05 implementation (tc_engine.cu)
int64_t tc_syrk_lock_words(int64_t M, int64_t L);
int tc_syrk(const svgp_kop* kop, const float* Wt, int64_t ldwt, const float* winv, const float* wneg, int64_t L, double* A, int64_t chunk_rows, int* locks,
            cudaStream_t st);
int tc_rowquad(const svgp_kop* kop, const void* S_hi, const void* S_lo, const float* S_inv, int64_t L, int tri, float* q,
               int64_t ldq, cudaStream_t st);
int tc_scaled_gemm(const svgp_kop* kop, const float* W, int64_t ldw, const void* G_hi, const void* G_lo, const float* G_inv,
                   int64_t L, int64_t Mc, float* out, int64_t ldo, int accumulate, float* dots, int64_t lddots, int64_t ndot,
                   cudaStream_t st);
bool tc_shape_ok(const svgp_kop* kop);
// entry points of the integer tcgen05 implementation (tc_i8_engine.cu)
int64_t i8_syrk_ws_floats(int64_t N, int64_t M, int64_t L);
int tc_syrk_i8_prep_run(const svgp_kop* kop, const float* W, int64_t ldw, int64_t L, double* A, float* ws, cudaStream_t st, int digits3);
int tc_scaled_gemm_i8(const svgp_kop* kop, const float* W, int64_t ldw, const void* Gp, int64_t ldg, const float* gscale, int64_t L,
                      int64_t Mc, float* out, int64_t ldo, int accumulate, float* dots, int64_t lddots, int64_t ndot, int64_t nfull,
                      const float* kcorr, const float* gcorr, cudaStream_t st);

static inline int64_t pad8(int64_t n) { return (n + 7) / 8 * 8; }
static inline KAccess kaccess(const svgp_kop* kop) {
  if (kop->K) return KAccess{kop->K, nullptr, nullptr, kop->ldk, nullptr};
  return KAccess{nullptr, (const __half*)kop->Kh, (const __half*)kop->Kl, kop->ldkh, kop->kscale};
}

}  // namespace svgp

using namespace svgp;

static bool use_tc(const svgp_kop* kop, int impl) {
  if (impl == SVGP_IMPL_SIMT) return false;
  if (impl == SVGP_IMPL_TC) return true;
  return kop->Kh && kop->Kl && kop->kscale && tc_shape_ok(kop);     // (the fp16 SYRK additionally needs the transposed planes)
}

extern "C" {

int64_t svgp_syrk_ws_floats(int64_t N, int64_t M, int64_t L) {
  const int64_t a = L * pad8(N) + 3 * L + tc_syrk_lock_words(M, L), b = i8_syrk_ws_floats(N, M, L);
  return a > b ? a : b;
}

int svgp_syrk(const svgp_kop* kop, const float* W, int64_t ldw, int64_t L, double* A, int impl, int64_t chunk_rows,
              float* ws, void* stream) {
  SVGP_REQUIRE(kop && W && A && L >= 1, "null argument");
  if (kop->N == 0 || kop->M == 0) return SVGP_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (impl == SVGP_IMPL_TC_I8 || impl == SVGP_IMPL_TC_I8_D3 || impl == SVGP_IMPL_TC_I8_O4 || (impl == SVGP_IMPL_AUTO && kop->Kc && kop->cscale && tc_shape_ok(kop))) {
    SVGP_REQUIRE(ws != nullptr, "TC path needs the workspace (svgp_syrk_ws_floats)");
    SVGP_REQUIRE(kop->Kc && kop->cscale && kop->M >= 128 && kop->N >= 128, "integer path needs the int8 planes (svgp_kernel_fwd_i8) and M, N >= 128");
    return tc_syrk_i8_prep_run(kop, W, ldw, L, A, ws, st, impl == SVGP_IMPL_TC_I8_D3 ? 1 : (impl == SVGP_IMPL_TC_I8_O4 ? 2 : 0));
  }
  if (use_tc(kop, impl)) {
    SVGP_REQUIRE(ws != nullptr, "TC path needs the workspace (svgp_syrk_ws_floats)");
    const int64_t ldwt = pad8(kop->N);
    float* Wt = ws;
    float* mx = ws + L * ldwt;                 // [max |w| per channel | any-negative flag per channel]
    float* wneg = mx + L;
    float* winv = wneg + L;
    if (cudaMemsetAsync(mx, 0, 2 * L * sizeof(float), st) != cudaSuccess) return check_launch("svgp_syrk(memset)");
    int64_t blocks = ceil_div(kop->N, 8);
    if (blocks > 148 * 8) blocks = 148 * 8;
    colabsmax_kernel<<<(unsigned)blocks, 256, 0, st>>>(W, ldw, kop->N, L, mx);
    dim3 grid((unsigned)ceil_div(ldwt, 32), (unsigned)ceil_div(L, 32));
    transpose_scale_kernel<<<grid, 256, 0, st>>>(W, ldw, kop->N, L, mx, Wt, ldwt, winv);
    int rc = check_launch("svgp_syrk(prep)");
    if (rc) return rc;
    return tc_syrk(kop, Wt, ldwt, winv, wneg, L, A, chunk_rows, reinterpret_cast<int*>(winv + L), st);
  }
  SVGP_REQUIRE(kop->K != nullptr, "SIMT path needs the fp32 K");
  int64_t chunk = chunk_rows > 0 ? chunk_rows : 2048;
  // keep the z-grid (L x chunks) within limits and the atomic traffic modest
  int64_t nchunk = ceil_div(kop->N, chunk);
  while (L * nchunk > 65535) { chunk *= 2; nchunk = ceil_div(kop->N, chunk); }
  SyrkOp op{kop->K, kop->ldk, W, ldw, A, kop->N, kop->M, L, chunk, nchunk, kop->M, kop->M};
  return launch_tile(op, L * nchunk, st, "svgp_syrk");
}

int svgp_gemm_tn(const svgp_kop* kop, const float* X, int64_t ldx, int64_t L, double* V, void* stream) {
  SVGP_REQUIRE(kop && (kop->K || (kop->Kh && kop->Kl && kop->kscale)) && X && V && L >= 1, "null argument");
  if (kop->N == 0 || kop->M == 0) return SVGP_OK;
  int64_t chunk = 2048, nchunk = ceil_div(kop->N, chunk);
  while (nchunk > 65535) { chunk *= 2; nchunk = ceil_div(kop->N, chunk); }
  // fp32 chunk partials are fine once many of them are folded in double (their rounding errors average out, same
  // accuracy class as the tensor-core SYRK); a short reduction keeps double accumulators throughout
  if (nchunk >= 32) {
    TnOpT<float> op{kaccess(kop), X, ldx, V, kop->N, kop->M, chunk, L, kop->M};
    return launch_tile(op, nchunk, (cudaStream_t)stream, "svgp_gemm_tn");
  }
  TnOpT<double> op{kaccess(kop), X, ldx, V, kop->N, kop->M, chunk, L, kop->M};
  return launch_tile(op, nchunk, (cudaStream_t)stream, "svgp_gemm_tn");
}

int svgp_gemm_nn(const svgp_kop* kop, const float* Wm, int64_t ldwm, int64_t L, float* out, int64_t ldo, void* stream) {
  SVGP_REQUIRE(kop && (kop->K || (kop->Kh && kop->Kl && kop->kscale)) && Wm && out && L >= 1, "null argument");
  if (kop->N == 0) return SVGP_OK;
  const int64_t slab = 65535LL * BM;   // walk row slabs (grid.y limit)
  for (int64_t r0 = 0; r0 < kop->N; r0 += slab) {
    int64_t rows = kop->N - r0 < slab ? kop->N - r0 : slab;
    NnOp op{kaccess(kop), r0, Wm, ldwm, out, ldo, kop->M, rows, L};
    int rc = launch_tile(op, 1, (cudaStream_t)stream, "svgp_gemm_nn");
    if (rc) return rc;
  }
  return SVGP_OK;
}

int svgp_gemm_nn_tc(const svgp_kop* kop, const void* Wm_hi, const void* Wm_lo, const float* Wm_inv, int64_t L, float* out,
                    int64_t ldo, void* stream) {
  SVGP_REQUIRE(kop && Wm_hi && Wm_lo && Wm_inv && out && L >= 1 && ldo >= L, "null argument");
  SVGP_REQUIRE(kop->Kh && kop->Kl && kop->kscale && tc_shape_ok(kop), "needs the fp16 planes of K_nm and a tensor-core sized problem");
  if (kop->N == 0) return SVGP_OK;
  // one "stacked matrix" of L rows: out[i, l] = sum_c K[i, c] Wm[l, c]
  return tc_scaled_gemm(kop, nullptr, 0, Wm_hi, Wm_lo, Wm_inv, 1, L, out, ldo, 0, nullptr, 0, 0, (cudaStream_t)stream);
}

int svgp_rowquad(const svgp_kop* kop, const void* S_hi, const void* S_lo, const float* S_inv, int64_t L, int tri, float* q,
                 int64_t ldq, int impl, void* stream) {
  SVGP_REQUIRE(kop && S_hi && q && L >= 1, "null argument");
  if (kop->N == 0) return SVGP_OK;
  if (use_tc(kop, impl)) {
    SVGP_REQUIRE(S_lo != nullptr && S_inv != nullptr, "TC path needs the fp16 planes and scales of S (svgp_split_f16)");
    return tc_rowquad(kop, S_hi, S_lo, S_inv, L, tri, q, ldq, (cudaStream_t)stream);
  }
  SVGP_REQUIRE(L <= 65535, "too many channels");
  SVGP_REQUIRE(kop->K != nullptr && S_lo == nullptr && S_inv == nullptr, "SIMT path takes fp32 K and the float64 matrices themselves (G_lo = G_inv = NULL)");
  dim3 grid((unsigned)ceil_div(kop->N, BM), (unsigned)L);
  rowquad_simt_kernel<<<grid, NT, 0, (cudaStream_t)stream>>>(kop->K, kop->ldk, kop->N, kop->M, (const double*)S_hi, tri, q, ldq);
  return check_launch("svgp_rowquad");
}

int svgp_scaled_gemm(const svgp_kop* kop, const float* W, int64_t ldw, const void* G_hi, const void* G_lo, const float* G_inv,
                     int64_t L, float* out, int64_t ldo, int accumulate, float* dots, int64_t lddots, int64_t ndot, int impl,
                     void* stream) {
  SVGP_REQUIRE(kop && W && G_hi && out && L >= 1, "null argument");
  SVGP_REQUIRE(!dots || (ndot >= 0 && ndot <= L && lddots >= ndot), "bad dots argument");
  if (kop->N == 0 || kop->M == 0) return SVGP_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (use_tc(kop, impl)) {
    SVGP_REQUIRE(G_lo != nullptr && G_inv != nullptr, "TC path needs the fp16 planes and scales of G (svgp_split_f16)");
    return tc_scaled_gemm(kop, W, ldw, G_hi, G_lo, G_inv, L, kop->M, out, ldo, accumulate, dots, lddots, ndot, st);
  }
  SVGP_REQUIRE(kop->K != nullptr && G_lo == nullptr && G_inv == nullptr, "SIMT path takes fp32 K and the float64 matrices themselves (G_lo = G_inv = NULL)");
  int64_t slab = 65535LL * BM;
  for (int64_t r0 = 0; r0 < kop->N; r0 += slab) {
    int64_t rows = kop->N - r0 < slab ? kop->N - r0 : slab;
    ScaledOp op{kop->K + r0 * kop->ldk, kop->ldk, W + r0 * ldw, ldw, (const double*)G_hi, out + r0 * ldo, ldo, kop->M, L, accumulate, rows, kop->M};
    int rc = launch_tile(op, 1, st, "svgp_scaled_gemm");
    if (rc) return rc;
  }
  if (dots && ndot > 0) {
    // no fused epilogue on the SIMT path: k^T G_l k comes from the row-quad kernel, which OVERWRITES dots
    // (equal to the documented accumulate-into-zeroed-buffer contract)
    SVGP_REQUIRE(ndot <= 65535, "too many channels");
    for (int64_t r0 = 0; r0 < kop->N; r0 += slab) {
      int64_t rows = kop->N - r0 < slab ? kop->N - r0 : slab;
      dim3 grid((unsigned)ceil_div(rows, BM), (unsigned)ndot);
      rowquad_simt_kernel<<<grid, NT, 0, st>>>(kop->K + r0 * kop->ldk, kop->ldk, rows, kop->M, (const double*)G_hi, 0,
                                               dots + r0 * lddots, lddots);
    }
    return check_launch("svgp_scaled_gemm(dots)");
  }
  return SVGP_OK;
}

int svgp_scaled_gemm_i8(const svgp_kop* kop, const float* W, int64_t ldw, const void* G_planes, int64_t ldg, const float* G_scale,
                        int64_t L, int64_t Mc, float* out, int64_t ldo, int accumulate, float* dots, int64_t lddots, int64_t ndot,
                        int64_t nfull, const float* K_bias, const float* G_bias, void* stream) {
  SVGP_REQUIRE(kop && G_planes && G_scale && out && L >= 1 && Mc >= 1 && nfull >= 0, "null argument");
  SVGP_REQUIRE((K_bias == nullptr) == (G_bias == nullptr), "K_bias and G_bias come together (svgp_i8_pair_bias of both operands)");
  SVGP_REQUIRE(!dots || (ndot >= 0 && ndot <= L && lddots >= ndot), "bad dots argument");
  SVGP_REQUIRE(kop->Kr && kop->rscale && kop->M >= 128 && kop->N >= 128, "integer path needs the int8 planes (svgp_kernel_fwd_i8) and M, N >= 128");
  SVGP_REQUIRE(ldg % 16 == 0 && ldg >= kop->M && ldo >= Mc, "bad pitch");
  if (kop->N == 0) return SVGP_OK;
  return tc_scaled_gemm_i8(kop, W, ldw, G_planes, ldg, G_scale, L, Mc, out, ldo, accumulate, dots, lddots, ndot, nfull < L ? nfull : L,
                           K_bias, G_bias, (cudaStream_t)stream);
}

int svgp_gemm_f32(int64_t Mr, int64_t Nc, int64_t Kd, const float* A, int64_t lda, const float* B, int64_t ldb, float* C,
                  int64_t ldc, int accumulate, void* stream) {
  SVGP_REQUIRE(A && B && C && Mr >= 0 && Nc >= 0 && Kd >= 0, "bad argument");
  int64_t slab = 65535LL * BM;
  for (int64_t r0 = 0; r0 < Mr; r0 += slab) {
    int64_t rows = Mr - r0 < slab ? Mr - r0 : slab;
    PlainOp op{A + r0 * lda, lda, B, ldb, C + r0 * ldc, ldc, Kd, accumulate, rows, Nc};
    int rc = launch_tile(op, 1, (cudaStream_t)stream, "svgp_gemm_f32");
    if (rc) return rc;
  }
  return SVGP_OK;
}

int svgp_split_f16(const double* x, int64_t nb, int64_t count, void* hi, void* lo, float* inv_scale, void* stream) {
  SVGP_REQUIRE(x && hi && lo && inv_scale && nb >= 0 && count >= 0 && nb <= 65535, "bad argument");
  if (nb == 0 || count == 0) return SVGP_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(inv_scale + nb, 0, nb * sizeof(float), st) != cudaSuccess) return check_launch("svgp_split_f16(memset)");
  int64_t bx = ceil_div(count, 256 * 8);
  if (bx > 148 * 4) bx = 148 * 4;
  if (bx < 1) bx = 1;
  dim3 grid((unsigned)bx, (unsigned)nb);
  absmax_f64_kernel<<<grid, 256, 0, st>>>(x, count, inv_scale + nb);
  split_f16_kernel<<<grid, 256, 0, st>>>(x, count, (__half*)hi, (__half*)lo, inv_scale, nb);
  return check_launch("svgp_split_f16");
}

}  // extern "C"
