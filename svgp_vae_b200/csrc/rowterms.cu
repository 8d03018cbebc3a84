// K4 (row part): fused per-datapoint terms of the ELBO with warp-level reductions.
//   svgp_rowstats_fwd   p = reciprocal_no_nan(noise) (SVGPVAE_model.py:282,330), p*y, and the three
//                       per-channel row sums that L3 (:297-299) and the cross entropy
//                       (utils.py:498-502) need once their quadratic forms are collapsed onto A_l:
//                       sum_i p kappa_i, sum_i p y^2, sum_i log noise  -- accumulated in double.
//   svgp_predictive_fwd p_v = kappa - h + q1 (:336-337), optional clip to [1e-4, 100] (:891-892) and the
//                       clip correction of the collapsed cross-entropy sum.
// Layout: y / noise / p / py / q1 are (N, L) row-major fp32 with unit channel stride; a warp owns 32
// consecutive datapoints of one 32-channel slab, so loads are coalesced along L and the reduction over
// datapoints runs down registers first, then across blocks with one double atomic per (block, channel).
#include "common.cuh"

namespace svgp {

constexpr int RT_THREADS = 256;

__global__ void __launch_bounds__(RT_THREADS) rowstats_kernel(const float* __restrict__ y, const float* __restrict__ noise,
                                                              const float* __restrict__ kappa, int64_t N, int64_t L,
                                                              float* __restrict__ p, float* __restrict__ py,
                                                              double* __restrict__ sums) {
  // thread -> channel (threadIdx.x % 32 within a 32-channel slab), rows strided by warps
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = RT_THREADS / 32;
  __shared__ double red[3][RT_THREADS / 32][32];
  for (int64_t l0 = (int64_t)blockIdx.y * 32; l0 < L; l0 += (int64_t)gridDim.y * 32) {
    const int64_t l = l0 + lane;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    float f0 = 0.f, f1 = 0.f, f2 = 0.f;
    int cnt = 0;
    for (int64_t i = (int64_t)blockIdx.x * nwarp + warp; i < N; i += (int64_t)gridDim.x * nwarp) {
      if (l < L) {
        float nv = noise[i * L + l], yv = y[i * L + l];
        float pv = (nv == 0.f) ? 0.f : 1.0f / nv;
        p[i * L + l] = pv;
        py[i * L + l] = pv * yv;
        f0 = fmaf(pv, kappa[i], f0);
        f1 = fmaf(pv * yv, yv, f1);
        f2 += logf(nv);
      }
      if (++cnt == 64) {          // flush fp32 partials into double every 64 rows
        s0 += f0; s1 += f1; s2 += f2; f0 = f1 = f2 = 0.f; cnt = 0;
      }
    }
    s0 += f0; s1 += f1; s2 += f2;
    red[0][warp][lane] = s0; red[1][warp][lane] = s1; red[2][warp][lane] = s2;
    __syncthreads();
    if (warp < 3 && l < L) {
      double t = 0.0;
      for (int w = 0; w < nwarp; ++w) t += red[warp][w][lane];
      atomicAdd(&sums[warp * L + l], t);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(RT_THREADS) predictive_kernel(const float* __restrict__ kappa, const float* __restrict__ h,
                                                                float* __restrict__ q1_pv, const float* __restrict__ p,
                                                                int64_t N, int64_t L, int clip, float clip_lo, float clip_hi,
                                                                double* __restrict__ clipsum,
                                                                unsigned char* __restrict__ clipmask) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = RT_THREADS / 32;
  __shared__ double red[RT_THREADS / 32][32];
  for (int64_t l0 = (int64_t)blockIdx.y * 32; l0 < L; l0 += (int64_t)gridDim.y * 32) {
    const int64_t l = l0 + lane;
    double corr = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * nwarp + warp; i < N; i += (int64_t)gridDim.x * nwarp) {
      if (l < L) {
        float raw = kappa[i] - h[i] + q1_pv[i * L + l];
        float v = raw;
        if (clip) {
          v = fminf(fmaxf(raw, clip_lo), clip_hi);
          bool active = (v != raw);
          if (clipmask) clipmask[i * L + l] = active ? 1 : 0;
          if (active) corr += (double)p[i * L + l] * ((double)v - (double)raw);
        }
        q1_pv[i * L + l] = v;
      }
    }
    if (clip) {
      red[warp][lane] = corr;
      __syncthreads();
      if (warp == 0 && l < L) {
        double t = 0.0;
        for (int w = 0; w < nwarp; ++w) t += red[w][lane];
        if (t != 0.0) atomicAdd(&clipsum[l], t);
      }
      __syncthreads();
    }
  }
}

// ---- backward of the row terms (pass C / pass D of step.py), one warp per datapoint, lanes over the channels ----------
// pre: adjoints of the predictive moments -> the SYRK weights G_q1, the stacked row weights [p | 2 G_q1] and [p y | g_pm]
// of pass D (written side by side: no concatenation afterwards), the clip correction of d/dp, and sum_l G_q1 (= d/d kappa).
__global__ void __launch_bounds__(RT_THREADS) rowterms_bwd_pre_kernel(
    const float* __restrict__ g_pv, const float* __restrict__ g_pm, const float* __restrict__ p, const float* __restrict__ y,
    const unsigned char* __restrict__ mask, const float* __restrict__ pv, const float* __restrict__ kappa, const float* __restrict__ h,
    const float* __restrict__ q1raw, const float* __restrict__ gce, int64_t N, int64_t L, float* __restrict__ G_q1,
    float* __restrict__ Wst, float* __restrict__ PYst, float* __restrict__ G_p_clip, float* __restrict__ G_kappa) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < N; i += nwarp) {
    float ks = 0.f;
    const float kh = mask ? kappa[i] - h[i] : 0.f;
    for (int64_t l = lane; l < L; l += 32) {
      const int64_t o = i * L + l;
      const float pp = p[o];
      float gq = g_pv ? g_pv[o] : 0.f;
      if (mask) {
        // d/d pv_raw: unclipped entries pass g_pv; clipped ones only see the clip correction of the collapsed CE sum
        const float gc = gce[l];
        if (mask[o]) gq = 0.5f * gc * pp;
        G_p_clip[o] = -0.5f * gc * (pv[o] - (kh + q1raw[o]));
      }
      G_q1[o] = gq;
      Wst[i * 2 * L + l] = pp;
      Wst[i * 2 * L + L + l] = 2.0f * gq;
      PYst[i * 2 * L + l] = pp * y[o];
      PYst[i * 2 * L + L + l] = g_pm[o];
      ks += gq;
    }
    ks = warp_sum(ks);
    if (lane == 0) G_kappa[i] = ks;
  }
}
// post: dObjective/dy, dObjective/dnoise and the rest of d/d kappa from the products of pass D
//   G_p     = kGk / 2 + y G_py + kappa gs0 + y^2 gs1 (+ clip correction)
//   G_y     = p G_py + 2 p y gs1
//   G_noise = -p^2 G_p (0 where noise == 0: reciprocal_no_nan) + gs2 / noise
//   G_kappa += sum_l p gs0
__global__ void __launch_bounds__(RT_THREADS) rowterms_bwd_post_kernel(
    const float* __restrict__ y, const float* __restrict__ noise, const float* __restrict__ p, const float* __restrict__ kappa,
    const float* __restrict__ kGk, const float* __restrict__ G_py, const double* __restrict__ gs, const float* __restrict__ G_p_clip,
    int64_t N, int64_t L, float* __restrict__ G_y, float* __restrict__ G_noise, float* __restrict__ G_kappa) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < N; i += nwarp) {
    const float kap = kappa[i];
    float ks = 0.f;
    for (int64_t l = lane; l < L; l += 32) {
      const int64_t o = i * L + l;
      const float g0 = (float)gs[l], g1 = (float)gs[L + l], g2 = (float)gs[2 * L + l];
      const float pp = p[o], yy = y[o], nn = noise[o], gpy = G_py[o];
      float gp = 0.5f * kGk[o] + yy * gpy + kap * g0 + yy * yy * g1;
      if (G_p_clip) gp += G_p_clip[o];
      G_y[o] = pp * gpy + 2.0f * pp * yy * g1;
      G_noise[o] = (nn != 0.f) ? (-pp * pp * gp + g2 / nn) : g2;
      ks += pp * g0;
    }
    ks = warp_sum(ks);
    if (lane == 0) G_kappa[i] += ks;
  }
}

}  // namespace svgp

using namespace svgp;

extern "C" {

int svgp_rowstats_fwd(const float* y, const float* noise, const float* kappa, int64_t N, int64_t L, float* p, float* py,
                      double* sums, void* stream) {
  SVGP_REQUIRE(y && noise && kappa && p && py && sums && N >= 0 && L >= 1, "bad argument");
  if (N == 0) return SVGP_OK;
  int64_t gy = ceil_div(L, 32);
  int64_t gx = ceil_div(N, RT_THREADS / 32 * 16);
  int64_t cap = (148 * 8 + gy - 1) / gy;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, (unsigned)(gy > 65535 ? 65535 : gy));
  rowstats_kernel<<<grid, RT_THREADS, 0, (cudaStream_t)stream>>>(y, noise, kappa, N, L, p, py, sums);
  return check_launch("svgp_rowstats_fwd");
}

int svgp_predictive_fwd(const float* kappa, const float* h, float* q1_pv, const float* p, int64_t N, int64_t L, int clip,
                        float clip_lo, float clip_hi, double* clipsum, unsigned char* clipmask, void* stream) {
  SVGP_REQUIRE(kappa && h && q1_pv && N >= 0 && L >= 1, "bad argument");
  SVGP_REQUIRE(!clip || (p && clipsum), "clip needs p and clipsum");
  if (N == 0) return SVGP_OK;
  int64_t gy = ceil_div(L, 32);
  int64_t gx = ceil_div(N, RT_THREADS / 32 * 16);
  int64_t cap = (148 * 8 + gy - 1) / gy;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, (unsigned)(gy > 65535 ? 65535 : gy));
  predictive_kernel<<<grid, RT_THREADS, 0, (cudaStream_t)stream>>>(kappa, h, q1_pv, p, N, L, clip, clip_lo, clip_hi, clipsum,
                                                                   clipmask);
  return check_launch("svgp_predictive_fwd");
}

static unsigned rows_grid(int64_t N) {
  int64_t g = ceil_div(N, RT_THREADS / 32);
  if (g > 148 * 16) g = 148 * 16;
  return (unsigned)(g < 1 ? 1 : g);
}

int svgp_rowterms_bwd_pre(const float* g_pv, const float* g_pm, const float* p, const float* y, const unsigned char* clipmask,
                          const float* pv, const float* kappa, const float* h, const float* q1raw, const float* gce, int64_t N,
                          int64_t L, float* G_q1, float* Wst, float* PYst, float* G_p_clip, float* G_kappa, void* stream) {
  SVGP_REQUIRE(g_pm && p && y && G_q1 && Wst && PYst && G_kappa && N >= 0 && L >= 1, "bad argument");
  SVGP_REQUIRE(!clipmask || (pv && kappa && h && q1raw && gce && G_p_clip), "the clip branch needs pv, kappa, h, q1raw, gce and G_p_clip");
  if (N == 0) return SVGP_OK;
  rowterms_bwd_pre_kernel<<<rows_grid(N), RT_THREADS, 0, (cudaStream_t)stream>>>(g_pv, g_pm, p, y, clipmask, pv, kappa, h, q1raw, gce, N, L,
                                                                                 G_q1, Wst, PYst, G_p_clip, G_kappa);
  return check_launch("svgp_rowterms_bwd_pre");
}

int svgp_rowterms_bwd_post(const float* y, const float* noise, const float* p, const float* kappa, const float* kGk,
                           const float* G_py, const double* gsums, const float* G_p_clip, int64_t N, int64_t L, float* G_y,
                           float* G_noise, float* G_kappa, void* stream) {
  SVGP_REQUIRE(y && noise && p && kappa && kGk && G_py && gsums && G_y && G_noise && G_kappa && N >= 0 && L >= 1, "bad argument");
  if (N == 0) return SVGP_OK;
  rowterms_bwd_post_kernel<<<rows_grid(N), RT_THREADS, 0, (cudaStream_t)stream>>>(y, noise, p, kappa, kGk, G_py, gsums, G_p_clip, N, L,
                                                                                  G_y, G_noise, G_kappa);
  return check_launch("svgp_rowterms_bwd_post");
}

}  // extern "C"
