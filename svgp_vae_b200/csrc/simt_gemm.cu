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
#include "common.cuh"

namespace svgp {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

template <class Op>
__global__ void __launch_bounds__(NT) tile_gemm_kernel(Op op) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
  int64_t k0, k1;
  op.krange(blockIdx.z, k0, k1);
  double acc[4][4];      // fp32 products, fp64 accumulation: S_l / Kinv have large cancelling entries (cond ~1e3..1e5)
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

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
      Bs[kk][nn] = (gk < k1 && gn < op.Nc) ? op.b(blockIdx.z, gk, gn) : 0.f;
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
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int64_t gm = m0 + ty * 4 + i, gn = n0 + tx * 4 + j;
      if (gm < op.Mr && gn < op.Nc) op.store(blockIdx.z, gm, gn, acc[i][j]);
    }
}

// ---- operand functors -------------------------------------------------------------------------
struct SyrkOp {     // z = l * nchunk + chunk
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
struct TnOp {       // rows = channel l, cols = inducing index; z = chunk
  static constexpr bool A_K_CONTIG = false, B_K_CONTIG = false;
  const float* K; int64_t ldk; const float* X; int64_t ldx; double* V; int64_t N, M, chunk;
  int64_t Mr, Nc;
  __device__ void krange(int z, int64_t& k0, int64_t& k1) const { k0 = (int64_t)z * chunk; k1 = min(N, k0 + chunk); }
  __device__ float a(int z, int64_t m, int64_t k) const { return X[k * ldx + m]; }
  __device__ float b(int z, int64_t k, int64_t n) const { return K[k * ldk + n]; }
  __device__ void store(int z, int64_t m, int64_t n, double v) const { atomicAdd(&V[m * M + n], v); }
};
struct NnOp {       // out[i,l] = sum_a K[i,a] Wm[l,a]
  static constexpr bool A_K_CONTIG = true, B_K_CONTIG = true;
  const float* K; int64_t ldk; const float* Wm; int64_t ldwm; float* out; int64_t ldo; int64_t M;
  int64_t Mr, Nc;
  __device__ void krange(int z, int64_t& k0, int64_t& k1) const { k0 = 0; k1 = M; }
  __device__ float a(int z, int64_t m, int64_t k) const { return K[m * ldk + k]; }
  __device__ float b(int z, int64_t k, int64_t n) const { return Wm[n * ldwm + k]; }
  __device__ void store(int z, int64_t m, int64_t n, double v) const { out[m * ldo + n] = (float)v; }
};
struct ScaledOp {   // out[i,c] = sum_{l,a} W[i,l] K[i,a] G[l,a,c]
  static constexpr bool A_K_CONTIG = true, B_K_CONTIG = false;
  const float* K; int64_t ldk; const float* W; int64_t ldw; const float* G; float* out; int64_t ldo;
  int64_t M, L; int accumulate;
  int64_t Mr, Nc;
  __device__ void krange(int z, int64_t& k0, int64_t& k1) const { k0 = 0; k1 = L * M; }
  __device__ float a(int z, int64_t m, int64_t k) const {
    int64_t l = k / M, aa = k - l * M;
    return W[m * ldw + l] * K[m * ldk + aa];
  }
  __device__ float b(int z, int64_t k, int64_t n) const { return G[k * M + n]; }   // (l*M + a)*M + c
  __device__ void store(int z, int64_t m, int64_t n, double v) const {
    float* o = out + m * ldo + n;
    *o = accumulate ? (float)((double)*o + v) : (float)v;
  }
};
struct PlainOp {
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
                                                          const float* __restrict__ S, int tri, float* __restrict__ q,
                                                          int64_t ldq) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int l = blockIdx.y;
  const float* Sl = S + (int64_t)l * M * M;
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
                   Bs[kk][nn] = (gk < M && gn < M) ? Sl[gn * M + gk] : 0.f; }
        else     { int nn = idx % BN; kk = idx / BN; gk = kb + kk; gn = n0 + nn;
                   Bs[kk][nn] = (gk < M && gn < M) ? Sl[gk * M + gn] : 0.f; }
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

__global__ void split_tf32_kernel(const double* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double v = x[i];
    float h = to_tf32((float)v);
    hi[i] = h;
    if (lo) lo[i] = to_tf32((float)(v - (double)h));
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
int tc_syrk(const svgp_kop* kop, const float* Wt, int64_t ldwt, int64_t L, double* A, int64_t chunk_rows, cudaStream_t st);
int tc_rowquad(const svgp_kop* kop, const float* S_hi, const float* S_lo, int64_t L, int tri, float* q, int64_t ldq,
               cudaStream_t st);
int tc_scaled_gemm(const svgp_kop* kop, const float* W, int64_t ldw, const float* G_hi, const float* G_lo, int64_t L,
                   float* out, int64_t ldo, int accumulate, cudaStream_t st);
bool tc_shape_ok(const svgp_kop* kop);

}  // namespace svgp

using namespace svgp;

static bool use_tc(const svgp_kop* kop, int impl) {
  if (impl == SVGP_IMPL_SIMT) return false;
  if (impl == SVGP_IMPL_TC) return true;
  return kop->K_lo && kop->Kt && kop->Kt_lo && tc_shape_ok(kop);
}

extern "C" {

int svgp_syrk(const svgp_kop* kop, const float* W, int64_t ldw, const float* Wt, int64_t ldwt, int64_t L, double* A,
              int impl, int64_t chunk_rows, void* stream) {
  SVGP_REQUIRE(kop && kop->K && A && L >= 1, "null argument");
  if (kop->N == 0 || kop->M == 0) return SVGP_OK;
  if (use_tc(kop, impl)) {
    SVGP_REQUIRE(Wt != nullptr, "TC path needs the channel-major weights Wt");
    return tc_syrk(kop, Wt, ldwt, L, A, chunk_rows, (cudaStream_t)stream);
  }
  SVGP_REQUIRE(W != nullptr, "SIMT path needs W (N x L)");
  int64_t chunk = chunk_rows > 0 ? chunk_rows : 2048;
  // keep the z-grid (L x chunks) within limits and the atomic traffic modest
  int64_t nchunk = ceil_div(kop->N, chunk);
  while (L * nchunk > 65535) { chunk *= 2; nchunk = ceil_div(kop->N, chunk); }
  SyrkOp op{kop->K, kop->ldk, W, ldw, A, kop->N, kop->M, L, chunk, nchunk, kop->M, kop->M};
  return launch_tile(op, L * nchunk, (cudaStream_t)stream, "svgp_syrk");
}

int svgp_gemm_tn(const svgp_kop* kop, const float* X, int64_t ldx, int64_t L, double* V, void* stream) {
  SVGP_REQUIRE(kop && kop->K && X && V && L >= 1, "null argument");
  if (kop->N == 0 || kop->M == 0) return SVGP_OK;
  int64_t chunk = 2048, nchunk = ceil_div(kop->N, chunk);
  while (nchunk > 65535) { chunk *= 2; nchunk = ceil_div(kop->N, chunk); }
  TnOp op{kop->K, kop->ldk, X, ldx, V, kop->N, kop->M, chunk, L, kop->M};
  return launch_tile(op, nchunk, (cudaStream_t)stream, "svgp_gemm_tn");
}

int svgp_gemm_nn(const svgp_kop* kop, const float* Wm, int64_t ldwm, int64_t L, float* out, int64_t ldo, void* stream) {
  SVGP_REQUIRE(kop && kop->K && Wm && out && L >= 1, "null argument");
  if (kop->N == 0) return SVGP_OK;
  if (kop->N > 65535LL * BM) {   // walk row slabs
    int64_t slab = 65535LL * BM;
    for (int64_t r0 = 0; r0 < kop->N; r0 += slab) {
      int64_t rows = kop->N - r0 < slab ? kop->N - r0 : slab;
      NnOp op{kop->K + r0 * kop->ldk, kop->ldk, Wm, ldwm, out + r0 * ldo, ldo, kop->M, rows, L};
      int rc = launch_tile(op, 1, (cudaStream_t)stream, "svgp_gemm_nn");
      if (rc) return rc;
    }
    return SVGP_OK;
  }
  NnOp op{kop->K, kop->ldk, Wm, ldwm, out, ldo, kop->M, kop->N, L};
  return launch_tile(op, 1, (cudaStream_t)stream, "svgp_gemm_nn");
}

int svgp_rowquad(const svgp_kop* kop, const float* S_hi, const float* S_lo, int64_t L, int tri, float* q, int64_t ldq,
                 int impl, void* stream) {
  SVGP_REQUIRE(kop && kop->K && S_hi && q && L >= 1, "null argument");
  if (kop->N == 0) return SVGP_OK;
  if (use_tc(kop, impl)) {
    SVGP_REQUIRE(S_lo != nullptr, "TC path needs the lo plane of S");
    return tc_rowquad(kop, S_hi, S_lo, L, tri, q, ldq, (cudaStream_t)stream);
  }
  SVGP_REQUIRE(L <= 65535, "too many channels");
  // SIMT: planes are summed on the fly only if a lo plane is given -- keep it simple: hi plane must be plain fp32
  SVGP_REQUIRE(S_lo == nullptr, "SIMT path takes a single fp32 plane");
  dim3 grid((unsigned)ceil_div(kop->N, BM), (unsigned)L);
  rowquad_simt_kernel<<<grid, NT, 0, (cudaStream_t)stream>>>(kop->K, kop->ldk, kop->N, kop->M, S_hi, tri, q, ldq);
  return check_launch("svgp_rowquad");
}

int svgp_scaled_gemm(const svgp_kop* kop, const float* W, int64_t ldw, const float* G_hi, const float* G_lo, int64_t L,
                     float* out, int64_t ldo, int accumulate, int impl, void* stream) {
  SVGP_REQUIRE(kop && kop->K && W && G_hi && out && L >= 1, "null argument");
  if (kop->N == 0 || kop->M == 0) return SVGP_OK;
  if (use_tc(kop, impl)) {
    SVGP_REQUIRE(G_lo != nullptr, "TC path needs the lo plane of G");
    return tc_scaled_gemm(kop, W, ldw, G_hi, G_lo, L, out, ldo, accumulate, (cudaStream_t)stream);
  }
  SVGP_REQUIRE(G_lo == nullptr, "SIMT path takes a single fp32 plane");
  int64_t slab = 65535LL * BM;
  for (int64_t r0 = 0; r0 < kop->N; r0 += slab) {
    int64_t rows = kop->N - r0 < slab ? kop->N - r0 : slab;
    ScaledOp op{kop->K + r0 * kop->ldk, kop->ldk, W + r0 * ldw, ldw, G_hi, out + r0 * ldo, ldo, kop->M, L, accumulate, rows, kop->M};
    int rc = launch_tile(op, 1, (cudaStream_t)stream, "svgp_scaled_gemm");
    if (rc) return rc;
  }
  return SVGP_OK;
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

int svgp_split_tf32(const double* x, float* hi, float* lo, int64_t n, void* stream) {
  SVGP_REQUIRE(x && hi && n >= 0, "bad argument");
  if (n == 0) return SVGP_OK;
  int64_t blocks = ceil_div(n, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  split_tf32_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, hi, lo, n);
  return check_launch("svgp_split_tf32");
}

}  // extern "C"
