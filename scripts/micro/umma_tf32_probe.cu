// Groundwork for round 2: a minimal hand-written tcgen05 GEMM on sm_100a —
//   D[128 x N] (fp32, TMEM) = A[128 x K] * B[N x K]^T, TF32 inputs from shared memory, K-major, no swizzle,
// once as plain TF32 and once as the 3xTF32 split (A_hi*B_hi + A_lo*B_hi + A_hi*B_lo) that an fp32-accurate
// DFT stage needs.  Checks both against a float64 CPU product and reports the error levels.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_tf32_probe umma_tf32_probe.cu && ./umma_tf32_probe
//
// Layout facts used (cute/arch/mma_sm100_desc.hpp, cute/atom/mma_traits_sm100.hpp of CUTLASS 3.9):
//  * K-major SWIZZLE_NONE canonical layout: 8-row x 16-byte core matrices stored as 128 contiguous bytes;
//    core (i = row/8, j = k/4 for tf32) lives at i*SBO + j*LBO; descriptor = addr>>4 | LBO>>4 <<16 | SBO>>4 <<32 | 1<<46.
//  * instruction descriptor (kind::tf32): c_format F32 (1)<<4 | a_format TF32 (2)<<7 | b_format TF32 (2)<<10 |
//    (N>>3)<<17 | (M>>4)<<24; K per instruction = 8.
//  * accumulator in TMEM: lane = row of D (0..127), column = n; warp w of the CTA reads lanes 32w..32w+31.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

constexpr int M = 128, N = 32, K = 32;
constexpr int SBO = 128 * (K / 4);  // bytes between 8-row groups: (K/4) core matrices of 128 B
constexpr int LBO = 128;            // bytes between adjacent core matrices along K

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((LBO >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((SBO >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version for sm_100
  return d;               // layout_type = SWIZZLE_NONE (0), base_offset 0
}

// element (row, k) of a K-major no-swizzle operand tile
__device__ __forceinline__ int canon_index(int row, int k) {  // in floats
  return (row / 8) * (SBO / 4) + (k / 4) * (LBO / 4) + (row % 8) * 4 + (k % 4);
}

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

__global__ void __launch_bounds__(128) probe(const float *A, const float *B, float *D1, float *D3) {
  __shared__ __align__(128) float sAhi[M * K], sAlo[M * K], sBhi[N * K], sBlo[N * K];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;

  for (int i = tid; i < M * K; i += 128) {
    const int r = i / K, k = i % K;
    const float a = A[i], h = tf32_hi(a);
    sAhi[canon_index(r, k)] = h;
    sAlo[canon_index(r, k)] = a - h;
  }
  for (int i = tid; i < N * K; i += 128) {
    const int r = i / K, k = i % K;
    const float b = B[i], h = tf32_hi(b);
    sBhi[canon_index(r, k)] = h;
    sBlo[canon_index(r, k)] = b - h;
  }
  if (warp == 0) {  // TMEM: 64 columns (two 32-column accumulators), allocated by one warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the MMA's async proxy
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm = tmem_base;

  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    auto mma = [&](uint32_t d_tmem, const float *a, const float *b, int kk, bool accumulate) {
      const uint64_t ad = make_desc(smem_u32(a) + kk * 2 * LBO), bd = make_desc(smem_u32(b) + kk * 2 * LBO);
      const uint32_t acc = accumulate ? 1u : 0u;
      asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0;\n"
                   "  tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }\n"
                   ::"r"(d_tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
    };
    for (int kk = 0; kk < K / 8; ++kk) mma(tm, sAhi, sBhi, kk, kk > 0);            // plain TF32 -> columns [0, 32)
    for (int kk = 0; kk < K / 8; ++kk) mma(tm + 32, sAlo, sBhi, kk, kk > 0);       // 3xTF32   -> columns [32, 64)
    for (int kk = 0; kk < K / 8; ++kk) mma(tm + 32, sAhi, sBlo, kk, true);
    for (int kk = 0; kk < K / 8; ++kk) mma(tm + 32, sAhi, sBhi, kk, true);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  {
    unsigned done = 0;
    while (!done)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  // each warp reads its 32 TMEM lanes (rows 32w..32w+31): 32 columns of each accumulator
  uint32_t r[32];
  for (int acc = 0; acc < 2; ++acc) {
    const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16) + acc * 32;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                   "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                   "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    float *D = acc ? D3 : D1;
    for (int n = 0; n < 32; ++n) D[tid * N + n] = __uint_as_float(r[n]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tm));
}

int main() {
  std::vector<float> A(M * K), B(N * K), D1(M * N), D3(M * N);
  srand(1);
  for (auto &v : A) v = (float)rand() / RAND_MAX - 0.5f;
  for (auto &v : B) v = (float)rand() / RAND_MAX - 0.5f;
  float *dA, *dB, *dD1, *dD3;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD1, D1.size() * 4); cudaMalloc(&dD3, D3.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD1, 0, D1.size() * 4); cudaMemset(dD3, 0, D3.size() * 4);
  probe<<<1, 128>>>(dA, dB, dD1, dD3);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel status: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  cudaMemcpy(D1.data(), dD1, D1.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(D3.data(), dD3, D3.size() * 4, cudaMemcpyDeviceToHost);
  double e1 = 0, e3 = 0, ref_max = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * (double)B[n * K + k];
      e1 = fmax(e1, fabs(D1[m * N + n] - ref));
      e3 = fmax(e3, fabs(D3[m * N + n] - ref));
      ref_max = fmax(ref_max, fabs(ref));
    }
  printf("max|ref| %.4f  max err plain TF32 %.3e  max err 3xTF32 %.3e\n", ref_max, e1, e3);
  printf("D1[0][0..3] = %f %f %f %f\n", D1[0], D1[1], D1[2], D1[3]);
  printf("%s\n", (e1 < 5e-3 && e3 < 5e-6) ? "UMMA PROBE OK" : "UMMA PROBE MISMATCH");
  return 0;
}
