// Operand planes of the exact-accumulation tensor-core products (tcgen05.mma.kind::i8, tc_i8_engine.cu).
//
// Digit format.  A real matrix entry x is stored as a fixed-point integer X = rint(x / scale) with |X| <= 127 * 256^(S-1),
// written in BALANCED base-256 digits  X = sum_s d_s 256^(S-1-s),  d_s in [-128, 127]  (S = 3: 24 bits, S = 4: 32 bits),
// one signed-int8 plane per digit (plane 0 = most significant).  `scale` is constant along the reduction index of the
// product the plane feeds (per row of K_nm for the products that reduce over inducing points, per column of K_nm for
// the SYRK that reduces over datapoints, per output column of a G matrix), so it factors out of the integer dot
// product: the MMAs are exact, the only rounding is this one quantisation of the operand.  The planes of K_nm itself are
// written by K1 (svgp_kernel_fwd_i8, kernel_matrix.cu); this file cuts the float64 M x M matrices.
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

// ---- expectation of the digit-plane pairs a product does NOT multiply -----------------------------------------------------
// The lower digits of the format lie in [-128, 127] and are uniformly distributed: mean -1/2, not 0.  A dropped pair
// sum_k a_t[k] b_u[k] (t + u >= 4, and (0, 3) / (3, 0) with three leading digits) therefore has the expectation
// -(sum_k a_t + sum_k b_u) / 2 - n / 4 (both lower digits) resp. -sum_k a_0 / 2 (a leading digit against a lower one), far below one
// fp32 rounding of the entry but of one sign for every entry of the product.  This kernel forms one operand's share from its
// digit sums, in units of the order-3 accumulator:  bias[0][r] = -(S_1 + S_2 + S_3) / 512  (the three order-4 pairs, ten pairs
// kept),  bias[1][r] = bias[0][r] - S_0 / 2  (eight pairs kept).  The constant -3 n / 1024 is the consumer's.
// One warp per row; the digits of four columns are summed by one dp4a.
__global__ void pair_bias_kernel(const int8_t* __restrict__ planes, int64_t nrows, int64_t ld, float* __restrict__ bias) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t plane = nrows * ld;
  for (int64_t r = warp; r < nrows; r += nwarp) {
    int s0 = 0, s123 = 0;
    for (int64_t c = lane * 16; c < ld; c += 32 * 16) {
      const int8_t* bp = planes + r * ld + c;
      const uint4 d0 = __ldg(reinterpret_cast<const uint4*>(bp));
      s0 = __dp4a((int)d0.x, 0x01010101, s0); s0 = __dp4a((int)d0.y, 0x01010101, s0);
      s0 = __dp4a((int)d0.z, 0x01010101, s0); s0 = __dp4a((int)d0.w, 0x01010101, s0);
#pragma unroll
      for (int t = 1; t < 4; ++t) {
        const uint4 d = __ldg(reinterpret_cast<const uint4*>(bp + t * plane));
        s123 = __dp4a((int)d.x, 0x01010101, s123); s123 = __dp4a((int)d.y, 0x01010101, s123);
        s123 = __dp4a((int)d.z, 0x01010101, s123); s123 = __dp4a((int)d.w, 0x01010101, s123);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s123 += __shfl_xor_sync(0xffffffffu, s123, o); }
    if (lane == 0) {
      const float b = -(float)s123 / 512.f;
      bias[r] = b;
      bias[nrows + r] = b - 0.5f * (float)s0;
    }
  }
}

}  // namespace svgp

using namespace svgp;

extern "C" {

int64_t svgp_i8_ldkr(int64_t M) { return (M + 15) / 16 * 16; }
int64_t svgp_i8_nblk(int64_t N) { return (N + 127) / 128; }

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

int svgp_i8_pair_bias(const void* planes, int64_t nrows, int64_t ld, float* bias, void* stream) {
  SVGP_REQUIRE(planes && bias && nrows >= 0, "null argument");
  SVGP_REQUIRE(ld % 16 == 0 && ld > 0 && ((uintptr_t)planes & 15) == 0, "plane pitch must be a multiple of 16 bytes");
  if (nrows == 0) return SVGP_OK;
  int64_t blocks = ceil_div(nrows, 8);
  if (blocks > 148 * 16) blocks = 148 * 16;
  pair_bias_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const int8_t*)planes, nrows, ld, bias);
  return check_launch("svgp_i8_pair_bias");
}

}  // extern "C"
