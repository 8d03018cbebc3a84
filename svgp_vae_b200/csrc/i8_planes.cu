// Operand planes of the exact-accumulation tensor-core products (tcgen05.mma.kind::i8, tc_i8_engine.cu).
//
// Digit format.  A real matrix entry x is stored as a fixed-point integer X = rint(x / scale) with |X| <= 127 * 256^(S-1),
// written in BALANCED base-256 digits  X = sum_s d_s 256^(S-1-s),  d_s in [-128, 127]  (S = 3: 24 bits, S = 4: 32 bits),
// one signed-int8 plane per digit (plane 0 = most significant).  `scale` is constant along the reduction index of the
// product the plane feeds (per row of K_nm for the products that reduce over inducing points, per column of K_nm for
// the SYRK that reduces over datapoints, per output column of a G matrix), so it factors out of the integer dot
// product: the MMAs are exact, the only rounding is this one quantisation of the operand.
// Balanced digits of a two's-complement integer V cost two instructions:  D = (V + 0x..808080) ^ 0x..808080  -- adding
// 128 to every lower byte turns the unsigned bytes into digit + 128 with the carries propagating, the XOR removes the
// offsets again; the bytes of D are the digits (top byte signed).
#include <cuda_fp16.h>

#include "common.cuh"

namespace svgp {

constexpr float XMAX3 = 8323072.0f;        // 127 * 2^16
constexpr double XMAX4 = 2130706432.0;     // 127 * 2^24

__device__ __forceinline__ uint32_t digits3(int v) { return ((uint32_t)v + 0x00008080u) ^ 0x00008080u; }   // bytes [d2, d1, d0, sign]
__device__ __forceinline__ uint32_t digits4(int v) { return ((uint32_t)v + 0x00808080u) ^ 0x00808080u; }   // bytes [d3, d2, d1, d0]

// 4 x 4 byte transpose: element words e0..e3 (byte b = digit index from the least significant) -> one word per digit
// plane holding that digit of the four elements (byte j = element j)
__device__ __forceinline__ void transpose4(uint32_t e0, uint32_t e1, uint32_t e2, uint32_t e3, uint32_t (&p)[4]) {
  const uint32_t x01 = __byte_perm(e0, e1, 0x7362), y01 = __byte_perm(e0, e1, 0x5140);
  const uint32_t x23 = __byte_perm(e2, e3, 0x7362), y23 = __byte_perm(e2, e3, 0x5140);
  p[3] = __byte_perm(x01, x23, 0x7632);    // byte 3 of every element
  p[2] = __byte_perm(x01, x23, 0x5410);
  p[1] = __byte_perm(y01, y23, 0x7632);
  p[0] = __byte_perm(y01, y23, 0x5410);    // byte 0 of every element
}

__device__ __forceinline__ void atomic_max_pos(float* addr, float v) {          // v >= 0: float order == int order
  if (v > 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
}

// ---- K_nm: row / column maxima of |hi + lo| (plane units) ---------------------------------------------------------------
__global__ void rowabsmax_kernel(const __half* __restrict__ Kh, const __half* __restrict__ Kl, int64_t ldkh, int64_t N, int64_t M,
                                 float* __restrict__ rmax) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < N; i += nwarp) {
    const uint4* ph = reinterpret_cast<const uint4*>(Kh + i * ldkh);
    const uint4* pl = reinterpret_cast<const uint4*>(Kl + i * ldkh);
    float m = 0.f;
    for (int64_t c8 = lane; c8 * 8 < M; c8 += 32) {
      const uint4 h = __ldg(ph + c8), l = __ldg(pl + c8);
      const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&hw[k]));
        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&lw[k]));
        if (c8 * 8 + 2 * k < M) m = fmaxf(m, fabsf(a.x + b.x));
        if (c8 * 8 + 2 * k + 1 < M) m = fmaxf(m, fabsf(a.y + b.y));
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) rmax[i] = m;
  }
}
// column maxima from the datapoint-blocked transposed planes [nb64][M][64]: one warp per inducing point and block range
__global__ void colabsmax_t_kernel(const __half* __restrict__ Kth, const __half* __restrict__ Ktl, int64_t ldkt, int64_t nb64,
                                   int64_t M, float* __restrict__ cmax) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  if (m >= M) return;
  const int64_t per = (nb64 + gridDim.y - 1) / gridDim.y;
  const int64_t b0 = (int64_t)blockIdx.y * per, b1 = b0 + per < nb64 ? b0 + per : nb64;
  float mx = 0.f;
#pragma unroll 4
  for (int64_t b = b0; b < b1; ++b) {
    const __half2 h = __ldg(reinterpret_cast<const __half2*>(Kth + b * ldkt + m * 64) + lane);
    const __half2 l = __ldg(reinterpret_cast<const __half2*>(Ktl + b * ldkt + m * 64) + lane);
    const float2 a = __half22float2(h), c = __half22float2(l);
    mx = fmaxf(mx, fmaxf(fabsf(a.x + c.x), fabsf(a.y + c.y)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) atomic_max_pos(cmax + m, mx);
}

// 16 fp16 (hi, lo) pairs -> 16 fixed-point integers -> three 16-byte digit runs
// (hi + lo is exact in fp32; the product with q is taken in double: an fp32 product would add up to half a unit of its own)
__device__ __forceinline__ void quant16(const uint4 (&h)[2], const uint4 (&l)[2], double q, uint4 (&out)[3]) {
  const uint32_t hw[8] = {h[0].x, h[0].y, h[0].z, h[0].w, h[1].x, h[1].y, h[1].z, h[1].w};
  const uint32_t lw[8] = {l[0].x, l[0].y, l[0].z, l[0].w, l[1].x, l[1].y, l[1].z, l[1].w};
  uint32_t e[16];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&hw[k]));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&lw[k]));
    e[2 * k] = digits3(__double2int_rn((double)(a.x + b.x) * q));
    e[2 * k + 1] = digits3(__double2int_rn((double)(a.y + b.y) * q));
  }
  uint32_t w[3][4];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint32_t p[4];
    transpose4(e[4 * g], e[4 * g + 1], e[4 * g + 2], e[4 * g + 3], p);
    w[0][g] = p[2]; w[1][g] = p[1]; w[2][g] = p[0];          // plane 0 = most significant digit = byte 2
  }
#pragma unroll
  for (int s = 0; s < 3; ++s) out[s] = make_uint4(w[s][0], w[s][1], w[s][2], w[s][3]);
}

// Kr[s][i][c]: row-scaled digits of K_nm (reduction over the inducing points)
__global__ void quant_rows_kernel(const __half* __restrict__ Kh, const __half* __restrict__ Kl, int64_t ldkh, int64_t N, int64_t M,
                                  const float* __restrict__ rmax, const float* __restrict__ kscale,
                                  int8_t* __restrict__ Kr, int64_t ldkr, float* __restrict__ rscale) {
  const int64_t cpr = ldkr / 16;                                   // 16-column chunks per row
  const int64_t total = N * cpr;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx / cpr, c0 = (idx - i * cpr) * 16;
    const float mx = rmax[i];
    const double q = mx > 0.f ? (double)XMAX3 / (double)mx : 0.0;
    if (c0 == 0) rscale[i] = mx > 0.f ? (float)((double)mx / (double)XMAX3 * (double)kscale[1]) : 0.f;
    uint4 h[2], l[2];
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      if (c0 + 8 * g + 8 <= ldkh) {
        h[g] = __ldg(reinterpret_cast<const uint4*>(Kh + i * ldkh + c0 + 8 * g));
        l[g] = __ldg(reinterpret_cast<const uint4*>(Kl + i * ldkh + c0 + 8 * g));
      } else {
        h[g] = make_uint4(0, 0, 0, 0); l[g] = h[g];
      }
    }
    uint4 out[3];
    quant16(h, l, q, out);
    if (c0 + 16 > M) {                                             // columns >= M must be exact zeros (ldkh padding is, but mask anyway)
      uint32_t* w;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        w = reinterpret_cast<uint32_t*>(&out[s]);
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int b = 0; b < 4; ++b)
            if (c0 + 4 * g + b >= M) w[g] &= ~(0xFFu << (8 * b));
      }
    }
#pragma unroll
    for (int s = 0; s < 3; ++s) *reinterpret_cast<uint4*>(Kr + (int64_t)s * N * ldkr + i * ldkr + c0) = out[s];
  }
}

// Kc[s][n / 128][m][n % 128]: column-scaled digits of K_nm^T (reduction over the datapoints), from the transposed fp16
// planes [n / 64][m][n % 64]
__global__ void quant_cols_kernel(const __half* __restrict__ Kth, const __half* __restrict__ Ktl, int64_t ldkt, int64_t nb64, int64_t M,
                                  const float* __restrict__ cmax, const float* __restrict__ kscale, int8_t* __restrict__ Kc,
                                  int64_t nb128, float* __restrict__ cscale) {
  const int64_t total = nb128 * M * 8;
  const int64_t plane = nb128 * M * 128;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t ch = idx & 7, m = (idx >> 3) % M, b = (idx >> 3) / M;
    const float mx = cmax[m];
    const double q = mx > 0.f ? (double)XMAX3 / (double)mx : 0.0;
    if (b == 0 && ch == 0) cscale[m] = mx > 0.f ? (float)((double)mx / (double)XMAX3 * (double)kscale[1]) : 0.f;
    const int64_t sb = 2 * b + (ch >> 2);
    uint4 h[2], l[2];
    if (sb < nb64) {
      const uint4* ph = reinterpret_cast<const uint4*>(Kth + sb * ldkt + m * 64 + (ch & 3) * 16);
      const uint4* pl = reinterpret_cast<const uint4*>(Ktl + sb * ldkt + m * 64 + (ch & 3) * 16);
      h[0] = __ldg(ph); h[1] = __ldg(ph + 1); l[0] = __ldg(pl); l[1] = __ldg(pl + 1);
    } else {
      h[0] = h[1] = l[0] = l[1] = make_uint4(0, 0, 0, 0);
    }
    uint4 out[3];
    quant16(h, l, q, out);
#pragma unroll
    for (int s = 0; s < 3; ++s) *reinterpret_cast<uint4*>(Kc + (int64_t)s * plane + (b * M + m) * 128 + ch * 16) = out[s];
  }
}

// ---- float64 matrices: per-row scale, S = 3 or 4 digit planes ------------------------------------------------------------
// x (nrows x cols, row pitch ldx doubles) -> planes[s][r][c] (row pitch ldp bytes, columns >= cols zeroed) and scale[r];
// one warp per row
template <int S>
__global__ void split_i8_kernel(const double* __restrict__ x, int64_t nrows, int64_t cols, int64_t ldx, int8_t* __restrict__ planes,
                                int64_t ldp, float* __restrict__ scale) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const double top = (S == 4) ? XMAX4 : (double)XMAX3;
  const int64_t plane = nrows * ldp;
  for (int64_t r = warp; r < nrows; r += nwarp) {
    const double* xr = x + r * ldx;
    double mx = 0.0;
    for (int64_t c = lane; c < cols; c += 32) mx = fmax(mx, fabs(xr[c]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const bool ok = mx > 0.0 && isfinite(mx);
    if (lane == 0) scale[r] = ok ? (float)(mx / top) : 0.f;
    // the scale the consumers multiply back is the float written above: quantise against exactly that value
    // (the digit format has 2^-8 of headroom above `top`, so the half-ulp by which the float may be smaller is harmless)
    const double quse = ok ? 1.0 / (double)(float)(mx / top) : 0.0;
    for (int64_t c0 = lane * 16; c0 < ldp; c0 += 32 * 16) {
      uint32_t e[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const double v = (c0 + k < cols) ? xr[c0 + k] * quse : 0.0;
        const int vi = (int)llrint(v);
        e[k] = (S == 4) ? digits4(vi) : digits3(vi);
      }
      uint32_t w[4][4];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint32_t p[4];
        transpose4(e[4 * g], e[4 * g + 1], e[4 * g + 2], e[4 * g + 3], p);
#pragma unroll
        for (int s = 0; s < S; ++s) w[s][g] = p[S - 1 - s];
      }
#pragma unroll
      for (int s = 0; s < S; ++s)
        *reinterpret_cast<uint4*>(planes + (int64_t)s * plane + r * ldp + c0) = make_uint4(w[s][0], w[s][1], w[s][2], w[s][3]);
    }
  }
}

}  // namespace svgp

using namespace svgp;

extern "C" {

int64_t svgp_i8_ldkr(int64_t M) { return (M + 15) / 16 * 16; }
int64_t svgp_i8_nblk(int64_t N) { return (N + 127) / 128; }

int svgp_kplanes_i8(const svgp_kop* kop, void* Kr, int64_t ldkr, float* rscale, void* Kc, float* cscale, float* scratch,
                    void* stream) {
  SVGP_REQUIRE(kop && Kr && rscale && Kc && cscale && scratch, "null argument");
  SVGP_REQUIRE(kop->Kh && kop->Kl && kop->Kth && kop->Ktl && kop->kscale, "needs the fp16 planes of K_nm (svgp_kernel_fwd)");
  SVGP_REQUIRE(ldkr % 16 == 0 && ldkr >= kop->M, "ldkr must be a multiple of 16 and >= M");
  SVGP_REQUIRE(((uintptr_t)Kr & 15) == 0 && ((uintptr_t)Kc & 15) == 0 && kop->ldkh % 8 == 0, "planes need 16-byte alignment");
  const int64_t N = kop->N, M = kop->M;
  if (N == 0 || M == 0) return SVGP_OK;
  cudaStream_t st = (cudaStream_t)stream;
  float* rmax = scratch;
  float* cmax = scratch + N;
  if (cudaMemsetAsync(cmax, 0, M * sizeof(float), st) != cudaSuccess) return check_launch("svgp_kplanes_i8(memset)");
  const int64_t nb64 = (N + 63) / 64, nb128 = (N + 127) / 128;
  int64_t blocks = ceil_div(N, 8);
  if (blocks > 148 * 16) blocks = 148 * 16;
  rowabsmax_kernel<<<(unsigned)blocks, 256, 0, st>>>((const __half*)kop->Kh, (const __half*)kop->Kl, kop->ldkh, N, M, rmax);
  int64_t split = nb64 / 64;
  if (split < 1) split = 1;
  if (split > 64) split = 64;
  dim3 gc((unsigned)ceil_div(M, 8), (unsigned)split);
  colabsmax_t_kernel<<<gc, 256, 0, st>>>((const __half*)kop->Kth, (const __half*)kop->Ktl, kop->ldkt, nb64, M, cmax);
  blocks = ceil_div(N * (ldkr / 16), 256);
  if (blocks > 148 * 32) blocks = 148 * 32;
  quant_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>((const __half*)kop->Kh, (const __half*)kop->Kl, kop->ldkh, N, M, rmax,
                                                      kop->kscale, (int8_t*)Kr, ldkr, rscale);
  blocks = ceil_div(nb128 * M * 8, 256);
  if (blocks > 148 * 32) blocks = 148 * 32;
  quant_cols_kernel<<<(unsigned)blocks, 256, 0, st>>>((const __half*)kop->Kth, (const __half*)kop->Ktl, kop->ldkt, nb64, M, cmax,
                                                      kop->kscale, (int8_t*)Kc, nb128, cscale);
  return check_launch("svgp_kplanes_i8");
}

int svgp_split_i8(const double* x, int64_t nrows, int64_t cols, int64_t ldx, int nslices, void* planes, int64_t ldp, float* scale,
                  void* stream) {
  SVGP_REQUIRE(x && planes && scale && nrows >= 0 && cols >= 0 && ldx >= cols, "bad argument");
  SVGP_REQUIRE(nslices == 3 || nslices == 4, "3 or 4 digit planes");
  SVGP_REQUIRE(ldp % 16 == 0 && ldp >= cols && ((uintptr_t)planes & 15) == 0, "plane pitch must be a multiple of 16 bytes");
  if (nrows == 0 || cols == 0) return SVGP_OK;
  int64_t blocks = ceil_div(nrows, 8);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (nslices == 4)
    split_i8_kernel<4><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, nrows, cols, ldx, (int8_t*)planes, ldp, scale);
  else
    split_i8_kernel<3><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, nrows, cols, ldx, (int8_t*)planes, ldp, scale);
  return check_launch("svgp_split_i8");
}

}  // extern "C"
