// Micro-benchmark: peak FP64 rate of the SIMT pipe (DFMA) and of the tensor pipe (mma.sync f64) on this GPU.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu && ./fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters) {
  double a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3 + i;
  const double b = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fma(a[i], b, c);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int SHAPE>
__global__ void dmma_kernel(double* out, int iters) {
  // SHAPE 0: m8n8k4 (a 1, b 1, c 2 doubles / thread)   SHAPE 1: m16n8k16 (a 8, b 4, c 4)
  double c[4][4];
#pragma unroll
  for (int t = 0; t < 4; ++t)
#pragma unroll
    for (int i = 0; i < 4; ++i) c[t][i] = 0.0;
  double a[8], b[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = 1.0 + threadIdx.x * 1e-6 + i;
#pragma unroll
  for (int i = 0; i < 4; ++i) b[i] = 1e-3 * (i + 1);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      if (SHAPE == 0) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[t][0]), "+d"(c[t][1]) : "d"(a[t]), "d"(b[t]));
      } else {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                     : "+d"(c[t][0]), "+d"(c[t][1]), "+d"(c[t][2]), "+d"(c[t][3])
                     : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]),
                       "d"(b[2]), "d"(b[3]));
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int t = 0; t < 4; ++t)
#pragma unroll
    for (int i = 0; i < 4; ++i) s += c[t][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
static float time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out; cudaMalloc(&out, sizeof(double) * sms * 4 * 1024);
  const int iters = 20000;
  for (int threads : {256, 512, 1024}) {
    float ms = time_ms([&] { dfma_kernel<<<sms * 2, threads>>>(out, iters); });
    double fl = 2.0 * 16 * iters * (double)threads * sms * 2;
    printf("{\"op\": \"dfma\", \"threads\": %d, \"ctas_per_sm\": 2, \"ms\": %.3f, \"TFLOPs\": %.2f}\n", threads, ms, fl / ms / 1e9);
  }
  for (int threads : {128, 256, 512}) {
    float ms = time_ms([&] { dmma_kernel<0><<<sms * 2, threads>>>(out, iters); });
    double fl = 2.0 * 8 * 8 * 4 * 4 * iters * (double)(threads / 32) * sms * 2;
    printf("{\"op\": \"dmma_m8n8k4\", \"threads\": %d, \"ms\": %.3f, \"TFLOPs\": %.2f}\n", threads, ms, fl / ms / 1e9);
    ms = time_ms([&] { dmma_kernel<1><<<sms * 2, threads>>>(out, iters); });
    fl = 2.0 * 16 * 8 * 16 * 4 * iters * (double)(threads / 32) * sms * 2;
    printf("{\"op\": \"dmma_m16n8k16\", \"threads\": %d, \"ms\": %.3f, \"TFLOPs\": %.2f}\n", threads, ms, fl / ms / 1e9);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e));
  return 0;
}
