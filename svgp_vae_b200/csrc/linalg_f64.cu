// K3: batched float64 factorisations over the latent channels (and over K_mm).
// Replaces tf.linalg.cholesky (SVGPVAE_model.py:270-274, :129-130) and the LU-based
// tf.linalg.inv (:239, :319, :331, :83, :154, :161) with a blocked right-looking Cholesky, a
// blocked triangular inverse and a batched DGEMM, all batched over L matrices of M x M.
//
// The O(L M^3) stage is kept in float64 on purpose: the M x M systems have condition
// numbers up to ~1e5-1e7 and the ELBO must match an fp64 reference to 1e-4.
//
// Panel factorisation: one warp per 32 x 32 diagonal block, lane r owns row r in registers and
// columns are exchanged with warp shuffles (no shared memory, no block barrier).  The same warp
// also inverts the block so that the panel solve below it becomes a multiply.
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace svgp {

constexpr int NB = 32;

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// factor one 32 x 32 SPD block held row-per-lane; returns the first failing column (+1) or 0
__device__ __forceinline__ int warp_potrf32(double (&a)[NB], int lane) {
  int fail = 0;
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    double akk = shfl_d(a[k], k);
    if (!(akk > 0.0) && fail == 0) fail = k + 1;
    double s = sqrt(akk);
    if (lane == k) a[k] = s;
    else if (lane > k) a[k] = a[k] / s;
#pragma unroll
    for (int j = k + 1; j < NB; ++j) {
      double ljk = shfl_d(a[k], j);
      if (lane >= j) a[j] = fma(-a[k], ljk, a[j]);
    }
  }
  return fail;
}

// X = inverse of the lower-triangular block whose row `lane` is a[]; on return x[] is COLUMN `lane` of X
__device__ __forceinline__ void warp_trtri32(const double (&a)[NB], double (&x)[NB], int lane) {
#pragma unroll
  for (int r = 0; r < NB; ++r) {
    double acc = (r == lane) ? 1.0 : 0.0;
#pragma unroll
    for (int k = 0; k < r; ++k) {
      double lrk = shfl_d(a[k], r);
      acc = fma(-lrk, x[k], acc);
    }
    double lrr = shfl_d(a[r], r);
    x[r] = (r >= lane) ? acc / lrr : 0.0;
  }
}

// diagonal block at (j, j): Cholesky in place (+ zero the strict upper triangle of the block), inverse -> dinv
__global__ void __launch_bounds__(32) potrf_diag_kernel(double* __restrict__ A, int64_t M, int64_t ld, int64_t stride,
                                                        int64_t j, int jb, int* __restrict__ status,
                                                        double* __restrict__ dinv /* [batch][NB][NB] */) {
  const int lane = threadIdx.x;
  double* Ab = A + (int64_t)blockIdx.x * stride + j * ld + j;
  double a[NB], x[NB];
#pragma unroll
  for (int c = 0; c < NB; ++c) {
    double v = (lane == c) ? 1.0 : 0.0;                 // identity padding for a partial last block
    if (lane < jb && c < jb && c <= lane) v = Ab[(int64_t)lane * ld + c];
    a[c] = v;
  }
  int fail = warp_potrf32(a, lane);
  unsigned anyfail = __ballot_sync(0xffffffffu, fail != 0);
  if (status && lane == 0 && anyfail && status[blockIdx.x] == 0) status[blockIdx.x] = (int)j + fail;
  warp_trtri32(a, x, lane);
#pragma unroll
  for (int c = 0; c < NB; ++c) {
    if (lane < jb && c < jb) Ab[(int64_t)lane * ld + c] = (c <= lane) ? a[c] : 0.0;
  }
  double* D = dinv + (int64_t)blockIdx.x * NB * NB;
#pragma unroll
  for (int r = 0; r < NB; ++r) D[r * NB + lane] = x[r];   // x[] is column `lane`
}

// inverse of every diagonal 32 x 32 block of a lower-triangular factor: dinv[b][blk][NB][NB]
__global__ void __launch_bounds__(32) trtri_diag_kernel(const double* __restrict__ Lf, int64_t M, int64_t ld, int64_t stride,
                                                        int64_t nblk, double* __restrict__ dinv) {
  const int lane = threadIdx.x;
  const int64_t b = blockIdx.y, blk = blockIdx.x, j = blk * NB;
  const int jb = (int)min((int64_t)NB, M - j);
  const double* Lb = Lf + b * stride + j * ld + j;
  double a[NB], x[NB];
#pragma unroll
  for (int c = 0; c < NB; ++c) {
    double v = (lane == c) ? 1.0 : 0.0;
    if (lane < jb && c < jb && c <= lane) v = Lb[(int64_t)lane * ld + c];
    a[c] = v;
  }
  warp_trtri32(a, x, lane);
  double* D = dinv + (b * nblk + blk) * NB * NB;
#pragma unroll
  for (int r = 0; r < NB; ++r) D[r * NB + lane] = x[r];
}

// panel below the diagonal block: rows [j+jb, M) of columns [j, j+jb)  <-  row * Linv_jj^T
__global__ void __launch_bounds__(128) trsm_panel_kernel(double* __restrict__ A, int64_t M, int64_t ld, int64_t stride,
                                                         int64_t j, int jb, const double* __restrict__ dinv) {
  __shared__ double Li[NB][NB + 1];
  const double* D = dinv + (int64_t)blockIdx.y * NB * NB;
  for (int idx = threadIdx.x; idx < NB * NB; idx += blockDim.x) Li[idx / NB][idx % NB] = D[idx];
  __syncthreads();
  int64_t r = j + jb + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= M) return;
  double* row = A + (int64_t)blockIdx.y * stride + r * ld + j;
  double v[NB];
#pragma unroll
  for (int k = 0; k < NB; ++k) v[k] = (k < jb) ? row[k] : 0.0;
#pragma unroll
  for (int c = 0; c < NB; ++c) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k <= c; ++k) s = fma(v[k], Li[c][k], s);     // (Linv^T)[k][c] = Linv[c][k], zero for k > c
    if (c < jb) row[c] = s;
  }
}

// Linv[blk, blk] = dinv[b][blk]  for every diagonal block
__global__ void copy_diag_blocks_kernel(const double* __restrict__ dinv, double* __restrict__ Linv, int64_t M, int64_t ld,
                                        int64_t stride, int64_t nblk) {
  const int64_t b = blockIdx.y, blk = blockIdx.x, j = blk * NB;
  const double* D = dinv + (b * nblk + blk) * NB * NB;
  double* X = Linv + b * stride + j * ld + j;
  for (int idx = threadIdx.x; idx < NB * NB; idx += blockDim.x) {
    int r = idx / NB, c = idx % NB;
    if (j + r < M && j + c < M) X[(int64_t)r * ld + c] = D[idx];
  }
}

__global__ void zero_upper_kernel(double* __restrict__ A, int64_t M, int64_t ld, int64_t stride) {
  double* Ab = A + (int64_t)blockIdx.y * stride;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < M * M; idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = idx / M, c = idx - r * M;
    if (c > r) Ab[r * ld + c] = 0.0;
  }
}

// ------------------------------------------------------------------------------------------------
// batched DGEMM, row-major, 64 x 64 x 16 tiles, 4 x 4 per thread.  `lower_only` skips output tiles
// strictly above the diagonal (SYRK-style trailing updates).
// ------------------------------------------------------------------------------------------------
constexpr int DM = 64, DN = 64, DK = 16;
__global__ void __launch_bounds__(256) gemm_f64_kernel(int transA, int transB, int64_t Mr, int64_t Nc, int64_t Kd, double alpha,
                                                       const double* __restrict__ A, int64_t lda, int64_t strideA,
                                                       const double* __restrict__ B, int64_t ldb, int64_t strideB, double beta,
                                                       double* __restrict__ C, int64_t ldc, int64_t strideC, int lower_only) {
  const int64_t m0 = (int64_t)blockIdx.y * DM, n0 = (int64_t)blockIdx.x * DN;
  if (lower_only && n0 > m0 + DM - 1) return;
  __shared__ double As[DK][DM + 2];
  __shared__ double Bs[DK][DN + 2];
  const double* Ab = A + (int64_t)blockIdx.z * strideA;
  const double* Bb = B + (int64_t)blockIdx.z * strideB;
  double* Cb = C + (int64_t)blockIdx.z * strideC;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) acc[i][jj] = 0.0;
  for (int64_t kb = 0; kb < Kd; kb += DK) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int idx = threadIdx.x + e * 256;
      int kk, mm, nn;
      if (!transA) { kk = idx % DK; mm = idx / DK; } else { mm = idx % DM; kk = idx / DM; }
      int64_t gk = kb + kk, gm = m0 + mm;
      As[kk][mm] = (gk < Kd && gm < Mr) ? (transA ? Ab[gk * lda + gm] : Ab[gm * lda + gk]) : 0.0;
      if (!transB) { nn = idx % DN; kk = idx / DN; } else { kk = idx % DK; nn = idx / DK; }
      gk = kb + kk;
      int64_t gn = n0 + nn;
      Bs[kk][nn] = (gk < Kd && gn < Nc) ? (transB ? Bb[gn * ldb + gk] : Bb[gk * ldb + gn]) : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < DK; ++kk) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) b[jj] = Bs[kk][tx * 4 + jj];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc[i][jj] = fma(a[i], b[jj], acc[i][jj]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      int64_t gm = m0 + ty * 4 + i, gn = n0 + tx * 4 + jj;
      if (gm < Mr && gn < Nc) {
        double* c = Cb + gm * ldc + gn;
        *c = (beta == 0.0) ? alpha * acc[i][jj] : fma(alpha, acc[i][jj], beta * (*c));
      }
    }
}

// Large-matrix variant: 128 x 128 x 16 tiles, 256 threads, 8 x 8 doubles per thread (one LDS.128 per 4 DFMA),
// next k-tile prefetched into registers while the current one is multiplied.  A thread's 8 rows are
// {ty*4 .. ty*4+3} and {64 + ty*4 ..}, its 8 columns likewise in tx: every shared-memory read is a 16-byte
// access whose quarter-warp covers consecutive banks.
constexpr int GM = 128, GN = 128, GK = 16;
template <bool transA, bool transB>
__global__ void __launch_bounds__(256) gemm_f64_big_kernel(int64_t Mr, int64_t Nc, int64_t Kd, double alpha,
                                                           const double* __restrict__ A, int64_t lda, int64_t strideA,
                                                           const double* __restrict__ B, int64_t ldb, int64_t strideB, double beta,
                                                           double* __restrict__ C, int64_t ldc, int64_t strideC, int lower_only) {
  const int64_t m0 = (int64_t)blockIdx.y * GM, n0 = (int64_t)blockIdx.x * GN;
  if (lower_only && n0 > m0 + GM - 1) return;
  // +2 doubles per row: the transposing stores (16 consecutive kk of one column) then land 16 bytes apart
  // (2-way instead of 16-way bank conflicts) and every row start stays 16-byte aligned for the LDS.128 reads
  __shared__ __align__(16) double As[GK][GM + 2];
  __shared__ __align__(16) double Bs[GK][GN + 2];
  const double* Ab = A + (int64_t)blockIdx.z * strideA;
  const double* Bb = B + (int64_t)blockIdx.z * strideB;
  double* Cb = C + (int64_t)blockIdx.z * strideC;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  double acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;
  double ra[8], rb[8];
  // element e of this thread inside a k-tile: idx = tid + 256 e  (0 .. 2047)
  auto fetch = [&](int64_t kb) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int idx = threadIdx.x + e * 256;
      int kk, mm, nn;
      if (!transA) { kk = idx % GK; mm = idx / GK; } else { mm = idx % GM; kk = idx / GM; }
      int64_t gk = kb + kk;
      const int64_t gm = m0 + mm;
      ra[e] = (gk < Kd && gm < Mr) ? (transA ? Ab[gk * lda + gm] : Ab[gm * lda + gk]) : 0.0;
      if (!transB) { nn = idx % GN; kk = idx / GN; } else { kk = idx % GK; nn = idx / GK; }
      gk = kb + kk;
      const int64_t gn = n0 + nn;
      rb[e] = (gk < Kd && gn < Nc) ? (transB ? Bb[gn * ldb + gk] : Bb[gk * ldb + gn]) : 0.0;
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int idx = threadIdx.x + e * 256;
      int kk, mm, nn;
      if (!transA) { kk = idx % GK; mm = idx / GK; } else { mm = idx % GM; kk = idx / GM; }
      As[kk][mm] = ra[e];
      if (!transB) { nn = idx % GN; kk = idx / GN; } else { kk = idx % GK; nn = idx / GK; }
      Bs[kk][nn] = rb[e];
    }
  };
  fetch(0);
  for (int64_t kb = 0; kb < Kd; kb += GK) {
    stash();
    __syncthreads();
    if (kb + GK < Kd) fetch(kb + GK);
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      double a[8], b[8];
      const double2 a0 = *reinterpret_cast<const double2*>(&As[kk][ty * 4]), a1 = *reinterpret_cast<const double2*>(&As[kk][ty * 4 + 2]);
      const double2 a2 = *reinterpret_cast<const double2*>(&As[kk][64 + ty * 4]), a3 = *reinterpret_cast<const double2*>(&As[kk][64 + ty * 4 + 2]);
      const double2 b0 = *reinterpret_cast<const double2*>(&Bs[kk][tx * 4]), b1 = *reinterpret_cast<const double2*>(&Bs[kk][tx * 4 + 2]);
      const double2 b2 = *reinterpret_cast<const double2*>(&Bs[kk][64 + tx * 4]), b3 = *reinterpret_cast<const double2*>(&Bs[kk][64 + tx * 4 + 2]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a1.x; a[3] = a1.y; a[4] = a2.x; a[5] = a2.y; a[6] = a3.x; a[7] = a3.y;
      b[0] = b0.x; b[1] = b0.y; b[2] = b1.x; b[3] = b1.y; b[4] = b2.x; b[5] = b2.y; b[6] = b3.x; b[7] = b3.y;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
    if (gm >= Mr) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int64_t gn = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + j - 4);
      if (gn < Nc) {
        double* c = Cb + gm * ldc + gn;
        *c = (beta == 0.0) ? alpha * acc[i][j] : fma(alpha, acc[i][j], beta * (*c));
      }
    }
  }
}

// Tensor-pipe variant (DMMA.8x8x4, the FP64 MMA shape native to sm_100: a probe of the FP64 pipes measures 37 TFLOP/s
// for both DFMA and DMMA on B200, but the MMA form needs 1/64 of the register-file operand traffic per FMA).
// CTA tile 128 x 128 x 16, 16 warps in a 4 x 4 grid, warp tile 32 x 32 = 4 x 4 fragments of 8 x 8 (32 doubles / thread).
// Each operand tile sits in shared memory in the orientation it has in global memory (no transposing stores); the
// row pitches (20 resp. 132 doubles, both = 4 mod 16) make every fragment load -- 4 k-values x 8 rows or columns per
// warp -- hit 16 distinct 8-byte banks per half-warp.
//   mma.m8n8k4.row.col.f64:  a = A[lane / 4][lane % 4]   b = B[lane % 4][lane / 4]   c{0,1} = C[lane / 4][2 (lane % 4) + {0,1}]
constexpr int PK = GK + 4;      // pitch of a k-contiguous tile   [128][PK]
constexpr int PM = GM + 4;      // pitch of an m/n-contiguous tile [GK][PM]
constexpr int TILE_DOUBLES = (GM * PK > GK * PM) ? GM * PK : GK * PM;

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <bool transA, bool transB>
__global__ void __launch_bounds__(512, 1) gemm_f64_mma_kernel(int64_t Mr, int64_t Nc, int64_t Kd, double alpha,
                                                              const double* __restrict__ A, int64_t lda, int64_t strideA,
                                                              const double* __restrict__ B, int64_t ldb, int64_t strideB, double beta,
                                                              double* __restrict__ C, int64_t ldc, int64_t strideC, int lower_only) {
  const int64_t m0 = (int64_t)blockIdx.y * GM, n0 = (int64_t)blockIdx.x * GN;
  if (lower_only && n0 > m0 + GM - 1) return;
  __shared__ __align__(16) double As[TILE_DOUBLES];
  __shared__ __align__(16) double Bs[TILE_DOUBLES];
  const double* Ab = A + (int64_t)blockIdx.z * strideA;
  const double* Bb = B + (int64_t)blockIdx.z * strideB;
  double* Cb = C + (int64_t)blockIdx.z * strideC;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = (warp >> 2) * 32, wn = (warp & 3) * 32;       // warp tile origin inside the CTA tile
  const int lr = lane >> 2, lk = lane & 3;
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  double ra[4], rb[4];
  // element e of this thread inside a k-tile: idx = tid + 512 e (0 .. 2047); k-contiguous operands walk k fastest
  auto fetch = [&](int64_t kb) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = threadIdx.x + e * 512;
      {
        const int kk = transA ? idx / GM : idx % GK, mm = transA ? idx % GM : idx / GK;
        const int64_t gk = kb + kk, gm = m0 + mm;
        ra[e] = (gk < Kd && gm < Mr) ? (transA ? Ab[gk * lda + gm] : Ab[gm * lda + gk]) : 0.0;
      }
      {
        const int kk = transB ? idx % GK : idx / GN, nn = transB ? idx / GK : idx % GN;
        const int64_t gk = kb + kk, gn = n0 + nn;
        rb[e] = (gk < Kd && gn < Nc) ? (transB ? Bb[gn * ldb + gk] : Bb[gk * ldb + gn]) : 0.0;
      }
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = threadIdx.x + e * 512;
      if (transA) As[(idx / GM) * PM + idx % GM] = ra[e]; else As[(idx / GK) * PK + idx % GK] = ra[e];
      if (transB) Bs[(idx / GK) * PK + idx % GK] = rb[e]; else Bs[(idx / GN) * PM + idx % GN] = rb[e];
    }
  };
  // lower_only == 2: C = T^T T with T lower-triangular (S = Linv^T Linv): entries of a lower tile only receive
  // contributions from k >= its first row
  const int64_t kb0 = (lower_only == 2) ? (m0 / GK) * GK : 0;
  fetch(kb0);
  for (int64_t kb = kb0; kb < Kd; kb += GK) {
    stash();
    __syncthreads();
    if (kb + GK < Kd) fetch(kb + GK);
#pragma unroll
    for (int k0 = 0; k0 < GK; k0 += 4) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        a[i] = transA ? As[(k0 + lk) * PM + wm + i * 8 + lr] : As[(wm + i * 8 + lr) * PK + k0 + lk];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        b[j] = transB ? Bs[(wn + j * 8 + lr) * PK + k0 + lk] : Bs[(k0 + lk) * PM + wn + j * 8 + lr];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t gm = m0 + wm + i * 8 + lr;
    if (gm >= Mr) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int64_t gn = n0 + wn + j * 8 + 2 * lk + t;
        if (gn < Nc) {
          double* c = Cb + gm * ldc + gn;
          *c = (beta == 0.0) ? alpha * acc[i][j][t] : fma(alpha, acc[i][j][t], beta * (*c));
        }
      }
    }
  }
}

static int gemm_f64(int transA, int transB, int64_t Mr, int64_t Nc, int64_t Kd, double alpha, const double* A, int64_t lda,
                    int64_t strideA, const double* B, int64_t ldb, int64_t strideB, double beta, double* C, int64_t ldc,
                    int64_t strideC, int64_t batch, int lower_only, cudaStream_t st) {
  if (Mr <= 0 || Nc <= 0 || batch <= 0) return SVGP_OK;
  const bool big = Mr >= 96 && Nc >= 96 && Kd >= 32;
  for (int64_t b0 = 0; b0 < batch; b0 += 65535) {
    int64_t nb = batch - b0 < 65535 ? batch - b0 : 65535;
    if (big) {
      dim3 gridb((unsigned)ceil_div(Nc, GN), (unsigned)ceil_div(Mr, GM), (unsigned)nb);
      static const bool use_simt = getenv("SVGP_DGEMM") && !strcmp(getenv("SVGP_DGEMM"), "simt");
#define SVGP_BIG(TA, TB)                                                                                                     \
  do {                                                                                                                       \
    if (use_simt)                                                                                                            \
      gemm_f64_big_kernel<TA, TB><<<gridb, 256, 0, st>>>(Mr, Nc, Kd, alpha, A + b0 * strideA, lda, strideA, B + b0 * strideB,  \
                                                          ldb, strideB, beta, C + b0 * strideC, ldc, strideC, lower_only);    \
    else                                                                                                                     \
      gemm_f64_mma_kernel<TA, TB><<<gridb, 512, 0, st>>>(Mr, Nc, Kd, alpha, A + b0 * strideA, lda, strideA, B + b0 * strideB,  \
                                                          ldb, strideB, beta, C + b0 * strideC, ldc, strideC, lower_only);    \
  } while (0)
      if (transA && transB) SVGP_BIG(true, true);
      else if (transA) SVGP_BIG(true, false);
      else if (transB) SVGP_BIG(false, true);
      else SVGP_BIG(false, false);
#undef SVGP_BIG
      int rcb = check_launch("svgp_gemm_f64");
      if (rcb) return rcb;
      continue;
    }
    dim3 grid((unsigned)ceil_div(Nc, DN), (unsigned)ceil_div(Mr, DM), (unsigned)nb);
    gemm_f64_kernel<<<grid, 256, 0, st>>>(transA, transB, Mr, Nc, Kd, alpha, A + b0 * strideA, lda, strideA, B + b0 * strideB,
                                          ldb, strideB, beta, C + b0 * strideC, ldc, strideC, lower_only);
    int rc = check_launch("svgp_gemm_f64");
    if (rc) return rc;
  }
  return SVGP_OK;
}

}  // namespace svgp

using namespace svgp;

extern "C" {

int svgp_gemm_f64(int transA, int transB, int64_t Mr, int64_t Nc, int64_t Kd, double alpha, const double* A, int64_t lda,
                  int64_t strideA, const double* B, int64_t ldb, int64_t strideB, double beta, double* C, int64_t ldc,
                  int64_t strideC, int64_t batch, void* stream) {
  SVGP_REQUIRE(A && B && C && Mr >= 0 && Nc >= 0 && Kd >= 0 && batch >= 0, "bad argument");
  return gemm_f64(transA, transB, Mr, Nc, Kd, alpha, A, lda, strideA, B, ldb, strideB, beta, C, ldc, strideC, batch, 0,
                  (cudaStream_t)stream);
}

// copy the lower triangle of every matrix onto its upper triangle
__global__ void mirror_lower_f64_kernel(double* __restrict__ A, int64_t M, int64_t ld, int64_t stride) {
  double* Ab = A + (int64_t)blockIdx.y * stride;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < M * M; idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = idx / M, c = idx - r * M;
    if (c > r) Ab[r * ld + c] = Ab[c * ld + r];
  }
}

int svgp_ltl_f64(const double* T, double* S, int64_t M, int64_t ld, int64_t stride, int64_t batch, void* stream) {
  SVGP_REQUIRE(T && S && M >= 1 && ld >= M && batch >= 1 && batch <= 65535, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = gemm_f64(1, 0, M, M, M, 1.0, T, ld, stride, T, ld, stride, 0.0, S, ld, stride, batch, 2, st);
  if (rc) return rc;
  dim3 g((unsigned)(ceil_div(M * M, 256) < 1024 ? ceil_div(M * M, 256) : 1024), (unsigned)batch);
  mirror_lower_f64_kernel<<<g, 256, 0, st>>>(S, M, ld, stride);
  return check_launch("svgp_ltl_f64");
}

int svgp_chol_f64(double* A, int64_t M, int64_t ld, int64_t stride, int64_t batch, int* status, double* ws, void* stream) {
  SVGP_REQUIRE(A && ws && M >= 1 && ld >= M && batch >= 1 && batch <= 65535, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (status) cudaMemsetAsync(status, 0, sizeof(int) * batch, st);
  for (int64_t j = 0; j < M; j += NB) {
    int jb = (int)(M - j < NB ? M - j : NB);
    potrf_diag_kernel<<<(unsigned)batch, 32, 0, st>>>(A, M, ld, stride, j, jb, status, ws);
    int rc = check_launch("svgp_chol_f64(diag)");
    if (rc) return rc;
    int64_t rem = M - j - jb;
    if (rem <= 0) break;
    dim3 g((unsigned)ceil_div(rem, 128), (unsigned)batch);
    trsm_panel_kernel<<<g, 128, 0, st>>>(A, M, ld, stride, j, jb, ws);
    rc = check_launch("svgp_chol_f64(panel)");
    if (rc) return rc;
    // trailing update (lower tiles only): A22 -= L21 L21^T
    double* L21 = A + (j + jb) * ld + j;
    double* A22 = A + (j + jb) * ld + (j + jb);
    rc = gemm_f64(0, 1, rem, rem, jb, -1.0, L21, ld, stride, L21, ld, stride, 1.0, A22, ld, stride, batch, 1, st);
    if (rc) return rc;
  }
  dim3 gz((unsigned)(ceil_div(M * M, 256) < 1024 ? ceil_div(M * M, 256) : 1024), (unsigned)batch);
  zero_upper_kernel<<<gz, 256, 0, st>>>(A, M, ld, stride);
  return check_launch("svgp_chol_f64(zero)");
}

int svgp_trinv_f64(const double* Lf, double* Linv, int64_t M, int64_t ld, int64_t stride, int64_t batch, double* ws,
                   void* stream) {
  SVGP_REQUIRE(Lf && Linv && ws && M >= 1 && ld >= M && batch >= 1 && batch <= 65535, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t nblk = ceil_div(M, NB);
  double* dinv = ws;                                  // [batch][nblk][NB][NB]
  double* T = ws + batch * nblk * NB * NB;            // [batch][NB][M] scratch
  dim3 gd((unsigned)nblk, (unsigned)batch);
  trtri_diag_kernel<<<gd, 32, 0, st>>>(Lf, M, ld, stride, nblk, dinv);
  int rc = check_launch("svgp_trinv_f64(diag)");
  if (rc) return rc;
  cudaMemsetAsync(Linv, 0, sizeof(double) * stride * (batch - 1) + sizeof(double) * ((M - 1) * ld + M), st);
  copy_diag_blocks_kernel<<<gd, 256, 0, st>>>(dinv, Linv, M, ld, stride, nblk);
  rc = check_launch("svgp_trinv_f64(copy)");
  if (rc) return rc;
  for (int64_t i = 0; i < nblk; ++i) {
    const int64_t r0 = i * NB;
    const int64_t ib = M - r0 < NB ? M - r0 : NB;
    if (i == 0) continue;
    // T = L[i, 0:i] * X[0:i, 0:i]        (ib x r0)
    rc = gemm_f64(0, 0, ib, r0, r0, 1.0, Lf + r0 * ld, ld, stride, Linv, ld, stride, 0.0, T, M, NB * M, batch, 0, st);
    if (rc) return rc;
    // X[i, 0:i] = -Linv_ii * T
    rc = gemm_f64(0, 0, ib, r0, ib, -1.0, dinv + i * NB * NB, NB, nblk * NB * NB, T, M, NB * M, 0.0, Linv + r0 * ld, ld, stride,
                  batch, 0, st);
    if (rc) return rc;
  }
  return check_launch("svgp_trinv_f64");
}

}  // extern "C"
