// Micro-probe: tcgen05.mma.kind::i8 on B200 (sm_100a) -- exactness of the int32 accumulation, operand signedness
// flags, and issue-bound MMA throughput against kind::f16 for the tile shapes the SVGP engine uses; single-CTA
// (cta_group::1) and CTA-pair (cta_group::2) forms.  Operands sit in shared memory (no loads in the timed loop), so the
// numbers are the tensor pipe + shared-memory operand fetch only.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o i8_mma_probe i8_mma_probe.cu && ./i8_mma_probe
// Output: JSON lines.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e_ = (x);                                                                      \
    if (e_ != cudaSuccess) {                                                                   \
      printf("{\"error\": \"%s at %s:%d\"}\n", cudaGetErrorString(e_), __FILE__, __LINE__);   \
      exit(1);                                                                                 \
    }                                                                                          \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
// bounded wait: a probe must never hang the box
__device__ __forceinline__ bool mbar_wait_bounded(uint32_t bar, uint32_t parity) {
  for (long long it = 0; it < 400000000LL; ++it) {
    uint32_t done;
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return true;
  }
  return false;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {   // K-major, 128-byte swizzle rows, 8-row groups 1024 B apart
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
template <int KIND, int CG>
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (KIND == 0 && CG == 1)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  if (KIND == 1 && CG == 1)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  if (KIND == 0 && CG == 2)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  if (KIND == 1 && CG == 2)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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
}

struct Params {
  int n;            // MMA N
  int a_fmt, b_fmt; // i8: 0 = unsigned, 1 = signed; f16: 0
  int reps;         // timed repetitions of the 4-k-step chain
  int* out;         // [grid][128][n] accumulator of the verification pass (CTA 0 / cluster 0 only)
  long long* cyc;   // [grid] cycles of the timed chain
  int* fail;        // set when a bounded wait expired
};

__host__ __device__ inline int a_val(int r, int k, int fmt) {      // deterministic operand bytes
  int v = (r * 7 + k * 13 + (r * k) % 5) % 251;
  return fmt ? v - 125 : v;                                        // signed: [-125, 125], unsigned: [0, 250]
}
__host__ __device__ inline int b_val(int c, int k, int fmt) {
  int v = (c * 11 + k * 3 + (c + 2 * k) % 7) % 241;
  return fmt ? v - 120 : v;
}

// KIND 0: fp16 operands (values a_val / 64, exact), K = 16 per MMA; KIND 1: int8 operands, K = 32 per MMA.
// Every operand row is 128 bytes (one swizzle row): 4 MMA k-steps per chain either way.
template <int KIND, int CG>
__global__ void __launch_bounds__(128, 1) probe_kernel(const Params P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (smem - smem_u32(smem_raw));
  const int n_local = P.n / CG;                        // rows of B held by this CTA
  const uint32_t a_off = 0, b_off = 128 * 128, bar_off = b_off + 256 * 128, tptr_off = bar_off + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t crank = 0;
  if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const int cluster_id = blockIdx.x / CG;

  // operands, written in the layout TMA SWIZZLE_128B would produce: 16-byte chunk index XOR (row & 7)
  for (int idx = threadIdx.x; idx < 128 * 128; idx += blockDim.x) {
    const int r = idx >> 7, byte = idx & 127;
    const int grow = r + 128 * (int)crank;             // CTA pair: this CTA holds rows [128 crank, +128) of A
    const uint32_t dst = a_off + (r >> 3) * 1024 + (r & 7) * 128 + (((byte >> 4) ^ (r & 7)) << 4) + (byte & 15);
    if (KIND == 1) {
      sm[dst] = (uint8_t)(int8_t)a_val(grow, byte, P.a_fmt);
    } else if ((byte & 1) == 0) {
      const int k = byte >> 1;
      __half h = __float2half((float)a_val(grow, k, 1) / 64.f);
      *reinterpret_cast<__half*>(sm + dst) = h;
    }
  }
  for (int idx = threadIdx.x; idx < n_local * 128; idx += blockDim.x) {
    const int r = idx >> 7, byte = idx & 127;
    const int gcol = r + n_local * (int)crank;         // CTA pair: this CTA holds columns [n/2 crank, +n/2) of B
    const uint32_t dst = b_off + (r >> 3) * 1024 + (r & 7) * 128 + (((byte >> 4) ^ (r & 7)) << 4) + (byte & 15);
    if (KIND == 1) {
      sm[dst] = (uint8_t)(int8_t)b_val(gcol, byte, P.b_fmt);
    } else if ((byte & 1) == 0) {
      const int k = byte >> 1;
      *reinterpret_cast<__half*>(sm + dst) = __float2half((float)b_val(gcol, k, 1) / 64.f);
    }
  }
  const uint32_t bar = smem + bar_off, tptr = smem + tptr_off;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tptr), "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tptr), "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tptr));

  const uint32_t mdim = (CG == 2) ? 256 : 128;
  const uint32_t idesc = (KIND == 1 ? (2u << 4) : (1u << 4)) | ((uint32_t)P.a_fmt << 7) | ((uint32_t)P.b_fmt << 10) |
                         ((uint32_t)(P.n >> 3) << 17) | ((mdim >> 4) << 24);
  const uint64_t ad = make_desc(smem + a_off), bd = make_desc(smem + b_off);
  uint32_t parity = 0;
  bool ok = true;
  auto commit = [&]() {
    if (CG == 1)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    else
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3)
                   : "memory");
  };
  // ---- verification pass: one chain of 4 k-steps into columns [0, n) ------------------------------------------
  if (warp == 1 && lane == 0 && crank == 0) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) umma<KIND, CG>(tmem_base, ad + 2 * ks, bd + 2 * ks, idesc, ks > 0);
    commit();
  }
  ok = mbar_wait_bounded(bar, parity) && ok;
  parity ^= 1;
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (cluster_id == 0) {
    for (int c0 = 0; c0 < P.n; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
      const int grow = warp * 32 + lane + 128 * (int)crank;
      for (int j = 0; j < 32; ++j) P.out[(size_t)grow * P.n + c0 + j] = (int)v[j];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  // ---- timed chain: reps x 4 MMAs, alternating between two accumulator regions --------------------------------
  long long t0 = 0, t1 = 0;
  if (warp == 1 && lane == 0 && crank == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    t0 = clock64();
    for (int r = 0; r < P.reps; ++r) {
      const uint32_t d = tmem_base + ((r & 1) ? 256u : 0u);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) umma<KIND, CG>(d, ad + 2 * ks, bd + 2 * ks, idesc, 1u);
    }
    commit();
  }
  ok = mbar_wait_bounded(bar, parity) && ok;
  parity ^= 1;
  if (warp == 1 && lane == 0 && crank == 0) {
    t1 = clock64();
    P.cyc[cluster_id] = t1 - t0;
  }
  if (!ok && threadIdx.x == 0) atomicExch(P.fail, 1);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (CG == 1)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    else
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

template <int KIND, int CG>
static void run(const char* name, int n, int a_fmt, int b_fmt, int reps, int grid) {
  const int rows = 128 * CG;
  int *d_out, *d_fail;
  long long* d_cyc;
  CK(cudaMalloc(&d_out, sizeof(int) * rows * n));
  CK(cudaMalloc(&d_cyc, sizeof(long long) * grid));
  CK(cudaMalloc(&d_fail, sizeof(int)));
  CK(cudaMemset(d_fail, 0, sizeof(int)));
  CK(cudaMemset(d_cyc, 0, sizeof(long long) * grid));
  Params P{n, a_fmt, b_fmt, reps, d_out, d_cyc, d_fail};
  const int smem = 128 * 128 + 256 * 128 + 64 + 1024;
  auto kern = probe_kernel<KIND, CG>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  cfg.gridDim = dim3(grid / CG * CG);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  if (CG == 2) {
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
  }
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaLaunchKernelEx(&cfg, kern, P));       // warm-up (also the verified launch)
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  CK(cudaLaunchKernelEx(&cfg, kern, P));
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<int> out((size_t)rows * n);
  std::vector<long long> cyc(grid);
  int fail = 0;
  CK(cudaMemcpy(out.data(), d_out, sizeof(int) * out.size(), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(cyc.data(), d_cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&fail, d_fail, sizeof(int), cudaMemcpyDeviceToHost));
  // expected values
  long long bad = 0, first_bad = -1;
  double worst = 0;
  const int K = (KIND == 1) ? 128 : 64;
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < n; ++c) {
      long long e = 0;
      for (int k = 0; k < K; ++k) e += (long long)a_val(r, k, KIND == 1 ? a_fmt : 1) * b_val(c, k, KIND == 1 ? b_fmt : 1);
      int got = out[(size_t)r * n + c];
      if (KIND == 1) {
        if ((long long)got != e) { if (!bad) first_bad = (long long)r * n + c; ++bad; }
      } else {
        float gf; memcpy(&gf, &got, 4);
        double d = fabs((double)gf - (double)e / 4096.0);
        if (d > worst) worst = d;
        if (d > 1e-3) { if (!bad) first_bad = (long long)r * n + c; ++bad; }
      }
    }
  long long cmin = 1LL << 60, cmax = 0;
  const int nclusters = grid / CG;
  for (int i = 0; i < nclusters; ++i) { if (cyc[i] < cmin) cmin = cyc[i]; if (cyc[i] > cmax) cmax = cyc[i]; }
  const double mmas = 4.0 * reps;
  const double macs_per_mma = (double)rows * n * ((KIND == 1) ? 32 : 16);
  const double total_mac = mmas * macs_per_mma * nclusters;
  printf("{\"probe\": \"%s\", \"kind\": \"%s\", \"cta_group\": %d, \"m\": %d, \"n\": %d, \"a_fmt\": %d, \"b_fmt\": %d, \"mismatches\": %lld, "
         "\"first_bad\": %lld, \"f16_worst_abs\": %.3g, \"wait_expired\": %d, \"cycles_per_mma_min\": %.1f, \"cycles_per_mma_max\": %.1f, "
         "\"mac_per_clk_per_sm\": %.0f, \"kernel_ms\": %.3f, \"chip_tmacs\": %.1f, \"grid\": %d}\n",
         name, KIND == 1 ? "i8" : "f16", CG, rows, n, a_fmt, b_fmt, bad, first_bad, worst, fail, cmin / mmas, cmax / mmas,
         macs_per_mma / (cmax / mmas) / CG, ms, total_mac / (ms * 1e-3) / 1e12, grid);
  fflush(stdout);
  cudaFree(d_out); cudaFree(d_cyc); cudaFree(d_fail);
}

template <int KIND>
static void sustained(const char* name, int grid) {
  int *d_out, *d_fail;
  long long* d_cyc;
  const int n = 256, reps = 100000;                  // 400k MMAs per CTA per launch: ~30 ms
  CK(cudaMalloc(&d_out, sizeof(int) * 128 * n));
  CK(cudaMalloc(&d_cyc, sizeof(long long) * grid));
  CK(cudaMalloc(&d_fail, sizeof(int)));
  CK(cudaMemset(d_fail, 0, sizeof(int)));
  Params P{n, KIND == 1 ? 1 : 0, KIND == 1 ? 1 : 0, reps, d_out, d_cyc, d_fail};
  const int smem = 128 * 128 + 256 * 128 + 64 + 1024;
  auto kern = probe_kernel<KIND, 1>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  kern<<<grid, 128, smem>>>(P);
  CK(cudaDeviceSynchronize());
  int launches = 0;
  float ms = 0;
  CK(cudaEventRecord(e0));
  do {
    for (int i = 0; i < 10; ++i) kern<<<grid, 128, smem>>>(P);
    launches += 10;
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
  } while (ms < 3000.f);
  const double macs = 4.0 * reps * 128.0 * n * (KIND == 1 ? 32 : 16) * grid * launches;
  printf("{\"probe\": \"%s\", \"kind\": \"%s\", \"seconds\": %.2f, \"chip_tmacs_sustained\": %.1f, \"tops_sustained\": %.1f}\n", name,
         KIND == 1 ? "i8" : "f16", ms * 1e-3, macs / (ms * 1e-3) / 1e12, 2.0 * macs / (ms * 1e-3) / 1e12);
  fflush(stdout);
  cudaFree(d_out); cudaFree(d_cyc); cudaFree(d_fail);
}

int main() {
  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int reps = 20000;                            // 80k MMAs per CTA: ~10 M cycles
  run<0, 1>("f16_n256", 256, 0, 0, reps, sms);
  run<1, 1>("i8_ss_n256", 256, 1, 1, reps, sms);
  run<1, 1>("i8_us_n256", 256, 0, 1, reps, sms);
  run<1, 1>("i8_su_n256", 256, 1, 0, reps, sms);
  run<1, 1>("i8_uu_n256", 256, 0, 0, reps, sms);
  run<1, 1>("i8_ss_n128", 128, 1, 1, reps, sms);
  run<1, 1>("i8_ss_n64", 64, 1, 1, reps, sms);
  run<0, 1>("f16_n128", 128, 0, 0, reps, sms);
  run<1, 2>("i8_ss_pair_n256", 256, 1, 1, reps, sms);
  run<1, 2>("i8_ss_pair_n128", 128, 1, 1, reps, sms);
  run<0, 2>("f16_pair_n256", 256, 0, 0, reps, sms);
  // sustained rates under the power cap: the same kernels back to back for ~3 s each (the roofline denominator of a
  // kernel timed inside a long step; MEASURED_PEAKS.json has the bf16 cuBLAS figure, this is the raw MMA-issue figure)
  sustained<0>("f16_n256_sustained_3s", sms);
  sustained<1>("i8_ss_n256_sustained_3s", sms);
  return 0;
}
