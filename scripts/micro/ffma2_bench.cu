// Microbenchmark: scalar FFMA/FADD vs packed f32x2 (FFMA2/FADD2) throughput on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(float2 *out, int iters, float2 c1, float2 c2) {
  float2 a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { a[i].x = fmaf(a[i].x, c1.x, c2.x); a[i].y = fmaf(a[i].y, c1.y, c2.y); }
      if (MODE == 1) a[i] = __ffma2_rn(a[i], c1, c2);
      if (MODE == 2) { a[i].x = a[i].x + c2.x; a[i].y = a[i].y + c2.y; }
      if (MODE == 3) a[i] = __fadd2_rn(a[i], c2);
    }
  }
  float2 s = make_float2(0, 0);
#pragma unroll
  for (int i = 0; i < 8; ++i) { s.x += a[i].x; s.y += a[i].y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float2 *out; cudaMalloc(&out, 148 * 16 * 256 * sizeof(float2));
  const int iters = 20000; const dim3 g(148 * 8), b(256);
  const char *names[4] = {"FFMA x2 scalar", "FFMA2 packed", "FADD x2 scalar", "FADD2 packed"};
  for (int m = 0; m < 4; ++m) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (m == 0) k<0><<<g, b>>>(out, iters, make_float2(0.999f, 1.001f), make_float2(1e-3f, -1e-3f));
      if (m == 1) k<1><<<g, b>>>(out, iters, make_float2(0.999f, 1.001f), make_float2(1e-3f, -1e-3f));
      if (m == 2) k<2><<<g, b>>>(out, iters, make_float2(0.999f, 1.001f), make_float2(1e-3f, -1e-3f));
      if (m == 3) k<3><<<g, b>>>(out, iters, make_float2(0.999f, 1.001f), make_float2(1e-3f, -1e-3f));
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double lane_ops = (double)g.x * b.x * iters * 16.0;  // scalar-equivalent ops
    printf("%-16s %8.3f ms  %7.2f T lane-ops/s  (%.1f per clk per SM @1.965GHz)\n", names[m], ms, lane_ops / ms / 1e9,
           lane_ops / (ms * 1e-3) / 148 / 1.965e9);
  }
  return 0;
}
