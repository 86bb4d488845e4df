// Tensor-core kernel for fft_length N = 512 (kernel = "tc"): the 512-point real DFT of every frame is a two-stage
// Cooley-Tukey factorisation 512 = 32 x 16 whose two stages are GEMMs on the 5th-generation tensor cores
// (tcgen05.mma.kind::tf32, accumulators in TMEM), made fp32-accurate by the 3xTF32 split
//     A*B ~= A_lo*B_hi + A_hi*B_lo + A_hi*B_hi,   x_hi = x & 0xffffe000,  x_lo = x - x_hi.
//
//   sample n = 16*n1 + n2 of the pre-processed frame v (n1 = 0..31, zero for n >= L; n2 = 0..15), bin k = k1 + 32*k2:
//     stage 1   Y[n2][k1]  = sum_n1 v[16 n1 + n2] * W32^(n1 k1)          k1 = 0..16 (real input: the rest is the conjugate)
//     twiddle   Y'[n2][k1] = Y[n2][k1] * W512^(n2 k1)                     (CUDA cores, between the two GEMMs)
//     stage 2   X[k1 + 32 k2] = sum_n2 Y'[n2][k1] * W16^(n2 k2)           k2 = 0..15; k2 >= 8 is conj X[512 - k]
//
//   GEMM 1:  D1[(frame, n2)][32] = A1[(frame, n2)][n1 = 0..31] * B1[n1][32]
//            A1 is the frame itself: 16 consecutive samples per K-row, i.e. an MN-major operand in the 64-byte-swizzle
//            canonical layout (8 K-rows x 64 B atoms), so the pre-processing threads store float4 chunks in sample order.
//            B1 columns: {Re Y0, Y16, Re Y1, Im Y1, ..., Re Y15, Im Y15} (Im Y0 = Im Y16 = 0).
//   GEMM 2:  D2[(frame, k1)][32] = A2[(frame, k1)][(n2, re/im)] * B2[(n2, re/im)][(k2, re/im)]   (K-major, 128-byte swizzle)
//            rows k1 = 0..15 of 8 frames fill one 128-row tile; the k1 = 16 rows of the 16 frames go to a third tile.
//
// A tile is 16 consecutive frames of one cut; a CTA (256 threads, 2 CTAs per SM so that one CTA's CUDA-core phases
// overlap the other's tensor phases) runs, per tile:
//   PRE    global -> DC removal, pre-emphasis, window (layers.py:151-186) -> hi/lo -> A1            (2 frames per warp)
//   MMA1   one thread issues 24 tcgen05.mma (2 row tiles x 3 products x 4 K-steps), commit -> mbarrier
//   INTER  tcgen05.ld D1 (thread = (frame, n2)) -> twiddle -> hi/lo -> A2 (STS.64, swizzled)
//   MMA2   36 tcgen05.mma (3 row tiles), commit -> mbarrier
//   POWER  tcgen05.ld D2 (thread = (frame, k1)) -> |X|^2 -> P[frame][bin] in shared memory (layers.py:38-42)
//   MEL    lane = frame: every half-warp owns a set of filters, weights are broadcast operands (layers.py:565-578)
//   OUT    log-mel tile -> coalesced rows (or DCT + lifter for MFCC, layers.py:708-724)
// HBM traffic: 4*S bytes in (overlap served by L1/L2), 4*F bytes out per frame; no intermediate leaves the SM.
#pragma once
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

#define TC_NF 16                    // frames per tile
#define TC_THREADS 256
#define TC_PP 260                   // floats per P row (= 4 mod 32: the 16 frame-lanes of a 128-bit P load hit disjoint banks)
#define TC_A1_LO (32 * 1024)        // byte offset of the lo half of A1 (hi at 0): 16 frames x 2 KB each
#define TC_A2_LO (34 * 1024)        // lo half of A2 (hi at 0): tiles T0 [0,16K) T1 [16K,32K) T2 [32K,34K)
#define TC_OFF_ETILE (20 * 1024)    // log-mel tile [16][Mpad] (aliases A like P does)
#define TC_OFF_CONST (68 * 1024)    // constant blob: B1 hi/lo, B2 hi/lo (4 KB each, 1024-aligned), mel tables
#define TC_TMEM_COLS 256
// Accumulators: every row tile owns 64 TMEM columns, [0, 32) = A_hi * B_hi and [32, 64) = A_hi * B_lo + A_lo * B_hi (the two
// products that share A_hi are ONE N = 64 instruction against the adjacent [B_hi | B_lo] images).  D2 reuses D1's columns.
#define TC_DCOLS 64

struct Tc512Tables {
  const void *cblob;      // [B1hi 4K][B1lo 4K][B2hi 4K][B2lo 4K][mel descriptors][mel weights]
  int cblob_bytes;
  int off_md, off_mw;     // byte offsets of the mel descriptors / weights inside the blob
  int fpu;                // filters per unit = ceil(M / 16)
  const float *win4;      // [512] window, zero beyond L
  const float2 *tw;       // [16][16]: W512^(n2 * k1) for k1 = 1..16 at [n2][k1 - 1]
};

__device__ __forceinline__ uint32_t tc_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float tc_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp: SmemDescriptor): start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 |
// version 1 << 46 | layout << 61 (2 = SWIZZLE_128B, 4 = SWIZZLE_64B)
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
         (1ull << 46) | ((uint64_t)layout << 61);
}
// instruction descriptor, kind::tf32, fp32 accumulate (InstrDescriptor): M = 128, N = 32; bit 15 = A is MN-major
#define TC_IDESC(a_mn, n) ((1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a_mn) << 15) | (((uint32_t)(n) >> 3) << 17) | ((128u >> 4) << 24))

__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0;\n"
               "  tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }\n"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_wait(uint32_t bar, uint32_t parity) {
  unsigned done = 0;
  while (!done)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
// 32 consecutive TMEM columns of this thread's lane (tcgen05.ld.32x32b.x32): warp w reads lanes 32 (w % 4) ...
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
               "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                 "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                 "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// accumulator row of this thread: columns [0, 32) + [32, 64) of a row tile (hi*hi plus the two cross terms)
__device__ __forceinline__ void tc_ld_acc(uint32_t taddr, float (&v)[32]) {
  uint32_t a[32], c[32];
  tc_ld32(taddr, a);
  tc_ld32(taddr + 32, c);
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(a[i]) + __uint_as_float(c[i]);
}

template <int DT>
__device__ __forceinline__ float4 tc_ld_chunk(const void *base, int64_t i) {  // 4 consecutive samples, i % 4 == 0, aligned
  if (DT == B200FEAT_I16) {
    const short4 q = __ldg(reinterpret_cast<const short4 *>(reinterpret_cast<const int16_t *>(base) + i));
    const float k = 1.0f / 32768.0f;
    return make_float4((float)q.x * k, (float)q.y * k, (float)q.z * k, (float)q.w * k);
  } else {
    return __ldg(reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(base) + i));
  }
}

// DBG != 0: the raw accumulators of the FIRST tile of block 0 go to `dbg` ([2][128][32] D1 | [3][128][32] D2 | [16][TC_PP] P)
// A1M: shared-memory layout of the stage-1 A operand.
//   0  K-major, 128-byte swizzle (row = (frame, n2), 32 fp32 of K = n1 per row): the pre-processing threads transpose
//      their float4 chunk into four scalar stores
//   1  MN-major, SWIZZLE_128B_BASE32B (the only MN-major layout tcgen05 accepts for tf32, cutlass sm100_common.inl:92):
//      a K-row holds the 16 samples of n1 for TWO frames (32 fp32 = 128 B), 4 K-rows per 512-byte atom, 32-byte granules
//      XOR-swizzled with the K-row index: chunks are stored as float4 in sample order
template <int DT, int DBG, int A1M>
__global__ void __launch_bounds__(TC_THREADS, 2)
b200feat_tc512_kernel(const DevPlan p, const Tc512Tables tt, const DevBatch b, float *dbg) {
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned char *sA = tc_smem;
  unsigned char *sC = tc_smem + TC_OFF_CONST;
  float *Pbuf = reinterpret_cast<float *>(sA);
  float *Etile = reinterpret_cast<float *>(sA + TC_OFF_ETILE);
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(sC + tt.cblob_bytes);  // [0] tables, [1] GEMM 1, [2] GEMM 2
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 3);
  const int4 *s_md = reinterpret_cast<const int4 *>(sC + tt.off_md);      // [unit][slot] {first bin, float4 groups, weight index, filter}
  const float4 *s_mw4 = reinterpret_cast<const float4 *>(sC + tt.off_mw);
  const uint32_t a_base = tc_smem_u32(sA), c_base = tc_smem_u32(sC);
  const uint32_t bar0 = tc_smem_u32(bars), bar1 = bar0 + 8, bar2 = bar0 + 16;

  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 2;" ::"r"(bar1));  // two issuing threads (one per row tile)
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 3;" ::"r"(bar2));  // three
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {  // constant tables: one TMA bulk copy
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0), "r"(tt.cblob_bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(c_base), "l"(tt.cblob), "r"(tt.cblob_bytes), "r"(bar0) : "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(tmem_slot)), "n"(TC_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }

  // ---- per-thread constants
  const int L = p.L;
  const int NCH = (L + 3) >> 2;                 // 16-byte chunks per frame that carry samples
  float4 wreg[4];                               // window of this lane's chunks c = lane + 32 j
  uint32_t aoff[4];                             // byte offset of chunk c inside its frame's part of A1 (frame-independent part)
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = lane + 32 * j;
    wreg[j] = __ldg(reinterpret_cast<const float4 *>(tt.win4) + c);
    const int n1 = c >> 2, q = c & 3;
    if (A1M == 1) aoff[j] = (uint32_t)((n1 >> 2) * 512 + (n1 & 3) * 128 + (q & 1) * 16);  // + granule ((2 h + q / 2) ^ (n1 & 3)) * 32
    else aoff[j] = (uint32_t)((q >> 1) * 1024 + (q & 1) * 512 + (n1 & 3) * 4);           // rows 4 q + e of the frame: + e * 128 + chunk
  }
  float2 twr[16];                               // W512^(n2 k1), k1 = 1..16, n2 = lane % 16
#pragma unroll
  for (int k = 0; k < 16; ++k) twr[k] = __ldg(tt.tw + (lane & 15) * 16 + k);
  const float inv_L = 1.0f / (float)L;
  const float pre = p.preemph;

  tc_wait(bar0, 0);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = *tmem_slot;
  const uint32_t tlane = tm + ((uint32_t)((warp & 3) * 32) << 16);  // this warp's TMEM lane quadrant
  const int mt = warp >> 2;                                          // row tile served by this warp (frames 8 mt ...)

  // DBG == 2: thread 0 accumulates the cycles between phase boundaries (PRE, MMA1, INTER, MMA2, POWER, MEL, OUT) into dbg
  long long tacc[7] = {0, 0, 0, 0, 0, 0, 0}, tlast = 0;
#define TC_TICK(i) do { if (DBG == 2 && tid == 0) { const long long c_ = clock64(); tacc[i] += c_ - tlast; tlast = c_; } } while (0)
  uint32_t it = 0;
  for (int64_t tg = blockIdx.x; tg < b.num_tiles; tg += gridDim.x) {
    const int64_t tile = b.tile_base + tg;
    const int cut = __ldg(b.tile_cut + tile) - b.batch_first;
    const int64_t t0 = (tile - __ldg(b.tile_off + cut)) * TC_NF;
    const int64_t T = __ldg(b.row_off + cut + 1) - __ldg(b.row_off + cut);
    const int64_t rows_here = b.out_mode == B200FEAT_OUT_PADDED ? b.max_frames : T;
    const int nv = (int)max((int64_t)0, min((int64_t)TC_NF, T - t0));            // frames of this tile that exist
    const int nrows = (int)max((int64_t)0, min((int64_t)TC_NF, rows_here - t0));  // rows of this tile in the output
    const int64_t row0 = b.out_mode == B200FEAT_OUT_PADDED ? (int64_t)(b.batch_first + cut) * b.max_frames + t0
                                                            : __ldg(b.row_off + cut) + t0;
    float *out = b.out + row0 * p.F;
    if (nv == 0) {  // padded mode: a tile past the end of the cut
      for (int i = tid; i < nrows * p.F; i += TC_THREADS) out[i] = b.pad_value;
      continue;
    }
    const int64_t n = __ldg(b.nsamp + cut);
    const int64_t xoff = __ldg(b.samp_off + cut);
    const uint32_t par = it & 1;
    ++it;
    if (DBG == 2 && tid == 0) tlast = clock64();

    // ================================ PRE: frames 2 warp, 2 warp + 1 -> A1 (hi / lo)
#pragma unroll 1
    for (int ff = 0; ff < 2; ++ff) {
      const int f = 2 * warp + ff;
      if (f >= nv) break;
      const int64_t sb = (t0 + f) * p.S - (p.snip_edges ? 0 : p.pad_left);  // first sample of the frame, relative to the cut
      float4 x[4];
      const bool interior = sb >= 0 && sb + 4 * NCH <= n && (((xoff + sb) & 3) == 0);
      if (interior) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = lane + 32 * j;
          x[j] = c < NCH ? tc_ld_chunk<DT>(b.samples, xoff + sb + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      } else {  // cut edge (or an unaligned cut): per-sample reflection (layers.py:753-772)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = lane + 32 * j;
          float e[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int i = 4 * c + q;
            e[q] = 0.f;
            if (i < L) {
              int64_t s = sb + i;
              if (!p.snip_edges) s = reflect_index(s, n, p.pad_mode);
              e[q] = ld_sample<DT>(b.samples, xoff + s);
            }
          }
          x[j] = make_float4(e[0], e[1], e[2], e[3]);
        }
      }
      if (L & 3) {  // taps >= L inside the last chunk are not part of the frame
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int i0 = 4 * (lane + 32 * j);
          if (i0 + 1 >= L) x[j].y = 0.f;
          if (i0 + 2 >= L) x[j].z = 0.f;
          if (i0 + 3 >= L) x[j].w = 0.f;
        }
      }
      // the tap before chunk c is the last tap of chunk c - 1: the neighbour lane's .w (lane 0: lane 31 of the round before)
      float pv[4];
      {
        float last = x[0].x;  // lane 0, chunk 0: replicate-left (layers.py:166)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float up = __shfl_up_sync(0xffffffffu, x[j].w, 1);
          pv[j] = lane == 0 ? last : up;
          last = __shfl_sync(0xffffffffu, x[j].w, 31);
        }
      }
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) s += (x[j].x + x[j].y) + (x[j].z + x[j].w);
      const float mu = p.remove_dc ? warp_sum(s) * inv_L : 0.f;
      // A1M 1: frame pair f / 2 owns 4 KB, frame f % 2 the granules 2 h, 2 h + 1 of every K-row; A1M 0: 2 KB of rows per frame
      unsigned char *fr = A1M == 1 ? sA + (f >> 1) * 4096 : sA + f * 2048;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float d0 = x[j].x - mu, d1 = x[j].y - mu, d2 = x[j].z - mu, d3 = x[j].w - mu, dp = pv[j] - mu;
        float4 v;
        v.x = fmaf(-pre, dp, d0) * wreg[j].x;
        v.y = fmaf(-pre, d0, d1) * wreg[j].y;
        v.z = fmaf(-pre, d1, d2) * wreg[j].z;
        v.w = fmaf(-pre, d2, d3) * wreg[j].w;
        if (lane + 32 * j >= NCH) v = make_float4(0.f, 0.f, 0.f, 0.f);  // K-rows beyond the frame stay exact zeros
        const float4 h = make_float4(tc_hi(v.x), tc_hi(v.y), tc_hi(v.z), tc_hi(v.w));
        const float4 lo = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
        const int c = lane + 32 * j, n1 = c >> 2, q = c & 3;
        if (A1M == 1) {
          unsigned char *dst = fr + aoff[j] + (((((f & 1) << 1) | (q >> 1)) ^ (n1 & 3)) << 5);
          *reinterpret_cast<float4 *>(dst) = h;
          *reinterpret_cast<float4 *>(dst + TC_A1_LO) = lo;
        } else {  // rows r = 4 q + e (mod 8: 4 (q & 1) + e), element n1: chunk (n1 / 4) ^ (r % 8)
          const int r0 = (q & 1) << 2, kc = n1 >> 2;
          unsigned char *dst = fr + aoff[j];
          *reinterpret_cast<float *>(dst + 0 * 128 + ((kc ^ (r0 + 0)) << 4)) = h.x;
          *reinterpret_cast<float *>(dst + 1 * 128 + ((kc ^ (r0 + 1)) << 4)) = h.y;
          *reinterpret_cast<float *>(dst + 2 * 128 + ((kc ^ (r0 + 2)) << 4)) = h.z;
          *reinterpret_cast<float *>(dst + 3 * 128 + ((kc ^ (r0 + 3)) << 4)) = h.w;
          dst += TC_A1_LO;
          *reinterpret_cast<float *>(dst + 0 * 128 + ((kc ^ (r0 + 0)) << 4)) = lo.x;
          *reinterpret_cast<float *>(dst + 1 * 128 + ((kc ^ (r0 + 1)) << 4)) = lo.y;
          *reinterpret_cast<float *>(dst + 2 * 128 + ((kc ^ (r0 + 2)) << 4)) = lo.z;
          *reinterpret_cast<float *>(dst + 3 * 128 + ((kc ^ (r0 + 3)) << 4)) = lo.w;
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core's async proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    TC_TICK(0);

    // ================================ MMA1: D1[m] = A1[m] * B1 (3xTF32), one issuing thread per row tile
    if ((tid & 127) == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int m = tid >> 7;
      const uint32_t ahi = a_base + m * 16384, alo = ahi + TC_A1_LO, d = tm + TC_DCOLS * m;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        // MN-major: 32 rows (a frame pair) per 4 KB (LBO), 4 K-rows per 512-byte atom (SBO), K = 8 = two atoms per step
        const uint64_t dh = A1M == 1 ? tc_desc(ahi + ks * 1024, 4096, 512, 1) : tc_desc(ahi + ks * 32, 16, 1024, 2);
        const uint64_t dl = A1M == 1 ? tc_desc(alo + ks * 1024, 4096, 512, 1) : tc_desc(alo + ks * 32, 16, 1024, 2);
        const uint64_t db = tc_desc(c_base + ks * 32, 16, 1024, 2);            // rows 0..31 B1_hi, 32..63 B1_lo
        tc_mma(d, dh, db, TC_IDESC(A1M == 1 ? 1 : 0, 64), ks ? 1u : 0u);       // [hi*hi | hi*lo]
        tc_mma(d + 32, dl, db, TC_IDESC(A1M == 1 ? 1 : 0, 32), 1u);            //          + lo*hi
      }
      tc_commit(bar1);
    }
    tc_wait(bar1, par);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    TC_TICK(1);

    // ================================ INTER: thread = (frame 8 mt + row / 16, n2 = row % 16): twiddle, split, store A2
    {
      float r[32];
      tc_ld_acc(tlane + TC_DCOLS * mt, r);
      if (DBG == 1 && blockIdx.x == 0 && it == 1) {
        float *d = dbg + (mt * 128 + (warp & 3) * 32 + lane) * 32;
#pragma unroll
        for (int c = 0; c < 32; ++c) d[c] = r[c];
      }
      const int row = (warp & 3) * 32 + lane;
      const int fl = row >> 4, n2 = row & 15;
      // T0 / T1 rows r = 16 fl + k1: atom (r / 8) = 2 fl + (k1 >> 3), row in atom = k1 & 7; chunk (n2 / 2) ^ (k1 & 7)
      unsigned char *base = sA + mt * 16384 + fl * 2048 + (n2 & 1) * 8;
      const int ch = n2 >> 1;
#pragma unroll
      for (int k1 = 0; k1 < 16; ++k1) {
        float yr, yi;
        if (k1 == 0) { yr = r[0]; yi = 0.f; }
        else {
          const float a = r[2 * k1], bq = r[2 * k1 + 1];
          const float2 w = twr[k1 - 1];
          yr = fmaf(a, w.x, -bq * w.y);
          yi = fmaf(a, w.y, bq * w.x);
        }
        const float hr = tc_hi(yr), hi_ = tc_hi(yi);
        unsigned char *dst = base + (k1 >> 3) * 1024 + (k1 & 7) * 128 + ((ch ^ (k1 & 7)) << 4);
        *reinterpret_cast<float2 *>(dst) = make_float2(hr, hi_);
        *reinterpret_cast<float2 *>(dst + TC_A2_LO) = make_float2(yr - hr, yi - hi_);
      }
      {  // k1 = 16: Y16 real, times W32^n2; tile T2 row = frame = 8 mt + fl
        const float a = r[1];
        const float2 w = twr[15];
        const float yr = a * w.x, yi = a * w.y;
        const float hr = tc_hi(yr), hi_ = tc_hi(yi);
        unsigned char *dst = sA + 32768 + mt * 1024 + fl * 128 + ((ch ^ fl) << 4) + (n2 & 1) * 8;
        *reinterpret_cast<float2 *>(dst) = make_float2(hr, hi_);
        *reinterpret_cast<float2 *>(dst + TC_A2_LO) = make_float2(yr - hr, yi - hi_);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    TC_TICK(2);

    // ================================ MMA2: D2[t] = A2[t] * B2, 3 row tiles, one issuing thread each (warps 0, 2, 4)
    if ((tid & 63) == 0 && tid < 192) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int t = tid >> 6;
      const uint32_t ahi = a_base + t * 16384, alo = ahi + TC_A2_LO, d = tm + TC_DCOLS * t;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t db = tc_desc(c_base + 8192 + ks * 32, 16, 1024, 2);     // rows 0..31 B2_hi, 32..63 B2_lo
        tc_mma(d, tc_desc(ahi + ks * 32, 16, 1024, 2), db, TC_IDESC(0, 64), ks ? 1u : 0u);
        tc_mma(d + 32, tc_desc(alo + ks * 32, 16, 1024, 2), db, TC_IDESC(0, 32), 1u);
      }
      tc_commit(bar2);
    }
    tc_wait(bar2, par);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    TC_TICK(3);

    // ================================ POWER: thread = (frame, k1) -> P[frame][bin]
    {
      float r[32];
      tc_ld_acc(tlane + TC_DCOLS * mt, r);
      const int row = (warp & 3) * 32 + lane;
      const int fl = row >> 4, k1 = row & 15;
      if (DBG == 1 && blockIdx.x == 0 && it == 1) {
        float *d = dbg + 2 * 128 * 32 + (mt * 128 + row) * 32;
#pragma unroll
        for (int c = 0; c < 32; ++c) d[c] = r[c];
      }
      float *Pf = Pbuf + (8 * mt + fl) * TC_PP;
#pragma unroll
      for (int k2 = 0; k2 < 16; ++k2) {
        const float re = r[2 * k2], im = r[2 * k2 + 1];
        float pw = fmaf(re, re, im * im);
        if (p.use_mag) pw = sqrtf(pw);
        const int bin = k2 < 8 ? k1 + 32 * k2 : 512 - k1 - 32 * k2;
        if (k1 != 0 || k2 <= 8) Pf[k1 == 0 ? 32 * k2 : bin] = pw;
      }
      if (k1 == 0) Pf[257] = Pf[258] = Pf[259] = 0.f;  // row padding a 128-bit mel load may touch (weight 0): never stale NaNs
      if ((warp & 3) == 0 && mt == 0) {  // tile T2: rows 0..15 = the k1 = 16 row of every frame -> bins 16 + 32 k2
        tc_ld_acc(tlane + TC_DCOLS * 2, r);
        if (DBG == 1 && blockIdx.x == 0 && it == 1) {
          float *d = dbg + 2 * 128 * 32 + (2 * 128 + lane) * 32;
#pragma unroll
          for (int c = 0; c < 32; ++c) d[c] = r[c];
        }
        if (lane < 16) {
          float *Pg = Pbuf + lane * TC_PP;
#pragma unroll
          for (int k2 = 0; k2 < 8; ++k2) {
            const float re = r[2 * k2], im = r[2 * k2 + 1];
            float pw = fmaf(re, re, im * im);
            if (p.use_mag) pw = sqrtf(pw);
            Pg[16 + 32 * k2] = pw;
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    TC_TICK(4);
    if (DBG == 1 && blockIdx.x == 0 && it == 1)
      for (int i = tid; i < TC_NF * TC_PP; i += TC_THREADS) dbg[5 * 128 * 32 + i] = Pbuf[i];

    // ================================ MEL: lane = frame (lane % 16), unit = 2 warp + lane / 16 owns filters unit + 16 j
    const int Mpad = (p.M + 3) & ~3;
    {
      const int f = lane & 15, unit = 2 * warp + (lane >> 4);
      const float *Pf = Pbuf + f * TC_PP;
      const float lgk = p.log10_mel ? 0.30102999566398119521f : 0.69314718055994530942f;
      for (int j = 0; j < tt.fpu; ++j) {
        const int4 md = s_md[unit * tt.fpu + j];  // {first bin (multiple of 4), float4 groups, float4 weight index, filter or -1}
        if (md.w < 0) continue;
        const float4 *pp = reinterpret_cast<const float4 *>(Pf + md.x);
        const float4 *wp = s_mw4 + md.z;
        float acc = 0.f;
        for (int i = 0; i < md.y; ++i) {
          const float4 w = wp[i], q = pp[i];
          acc = fmaf(q.w, w.w, fmaf(q.z, w.z, fmaf(q.y, w.y, fmaf(q.x, w.x, acc))));
        }
        Etile[f * Mpad + md.w] = fast_lg2_normal(nanmax(acc, p.mel_floor)) * lgk;
      }
    }
    __syncthreads();
    TC_TICK(5);

    // ================================ OUT
    if (p.feature == B200FEAT_MFCC) {
      for (int idx = tid; idx < nv * p.C; idx += TC_THREADS) {
        const int f = idx / p.C, c = idx - f * p.C;
        float acc = 0.f;
        for (int m = 0; m < p.M; ++m) acc = fmaf(Etile[f * Mpad + m], __ldg(p.dct + m * p.C + c), acc);
        if (p.use_lifter) acc *= __ldg(p.lifter + c);
        out[(int64_t)f * p.F + c] = acc;
      }
    } else {
      for (int idx = tid; idx < nv * p.M; idx += TC_THREADS) {
        const int f = idx / p.M, m = idx - f * p.M;
        out[idx] = Etile[f * Mpad + m];
      }
    }
    for (int i = nv * p.F + tid; i < nrows * p.F; i += TC_THREADS) out[i] = b.pad_value;
    __syncthreads();  // P / E tile (aliasing A1) are free again
    TC_TICK(6);
  }
  if (DBG == 2 && tid == 0) {
    unsigned long long *g = reinterpret_cast<unsigned long long *>(dbg);
    for (int i = 0; i < 7; ++i) atomicAdd(g + i, (unsigned long long)tacc[i]);
    atomicAdd(g + 7, (unsigned long long)it);
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(TC_TMEM_COLS));
}

// ---------------------------------------------------------------------------------------------- host
struct Tc512Host {
  Tc512Tables t;
  size_t smem;
};

static inline bool tc512_supported(const DevPlan &p) {
  return p.N == 512 && p.L >= 16 && p.L <= 512 && (p.feature == B200FEAT_FBANK || p.feature == B200FEAT_MFCC) && !p.use_energy &&
         p.pad_mode == B200FEAT_PAD_KALDI && p.M >= 1 && p.M <= 128 && p.C <= 128;
}

// element (row, k) of a K-major 128-byte-swizzle operand with 32 fp32 per row (one 128-byte row per matrix row), in floats
static inline int tc_k128_index(int row, int k) { return (row >> 3) * 256 + (row & 7) * 32 + ((((k >> 2) ^ (row & 7)) & 7) << 2) + (k & 3); }

static inline float tc_hi_host(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u &= 0xFFFFE000u;
  float r;
  memcpy(&r, &u, 4);
  return r;
}

template <typename T>
static int tc_upload(const std::vector<T> &h, std::vector<void *> &allocs, const T **out) {
  void *d = nullptr;
  if (cudaMalloc(&d, h.size() * sizeof(T)) != cudaSuccess) return B200FEAT_ECUDA;
  allocs.push_back(d);
  if (cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) return B200FEAT_ECUDA;
  *out = reinterpret_cast<const T *>(d);
  return 0;
}

// host images of the constant tables (no CUDA calls: the probe and the tests use it too)
struct Tc512Image {
  std::vector<unsigned char> blob;
  std::vector<float> win4;
  std::vector<float2> tw;
  int off_md = 0, off_mw = 0, fpu = 0;
};

static inline int tc512_build_image(const DevPlan &p, const std::vector<float> &bank, const std::vector<float> &window, Tc512Image *img) {
  std::vector<float> b1(32 * 32), b2(32 * 32);  // logical [col][k]
  for (int n1 = 0; n1 < 32; ++n1) {
    b1[0 * 32 + n1] = 1.f;
    b1[1 * 32 + n1] = (n1 & 1) ? -1.f : 1.f;
    for (int k1 = 1; k1 < 16; ++k1) {
      const double a = 2.0 * M_PI * (double)((n1 * k1) % 32) / 32.0;
      b1[(2 * k1) * 32 + n1] = (float)cos(a);
      b1[(2 * k1 + 1) * 32 + n1] = (float)(-sin(a));
    }
  }
  for (int n2 = 0; n2 < 16; ++n2)
    for (int k2 = 0; k2 < 16; ++k2) {
      const double a = 2.0 * M_PI * (double)((n2 * k2) % 16) / 16.0;
      const float c = (float)cos(a), s = (float)sin(a);
      b2[(2 * k2) * 32 + 2 * n2] = c;       // Re X += Re Y' * cos
      b2[(2 * k2) * 32 + 2 * n2 + 1] = s;   //        + Im Y' * sin
      b2[(2 * k2 + 1) * 32 + 2 * n2] = -s;  // Im X += -Re Y' * sin
      b2[(2 * k2 + 1) * 32 + 2 * n2 + 1] = c;
    }
  std::vector<float> img4(4 * 1024, 0.f);  // B1hi | B1lo | B2hi | B2lo, each 32 rows x 128 B swizzled
  for (int c = 0; c < 32; ++c)
    for (int k = 0; k < 32; ++k) {
      const int idx = tc_k128_index(c, k);
      const float v1 = b1[c * 32 + k], h1 = tc_hi_host(v1);
      img4[idx] = h1; img4[1024 + idx] = v1 - h1;
      const float v2 = b2[c * 32 + k], h2 = tc_hi_host(v2);
      img4[2048 + idx] = h2; img4[3072 + idx] = v2 - h2;
    }
  // mel: unit u (16 of them) owns filters u, u + 16, ...; per filter a 4-aligned window of float4 weight groups inside [0, 260)
  const int fpu = std::max(1, (p.M + 15) / 16);
  std::vector<int> md((size_t)16 * fpu * 4, 0);
  std::vector<float> mw;
  for (int u = 0; u < 16; ++u)
    for (int j = 0; j < fpu; ++j) {
      int *d = &md[((size_t)u * fpu + j) * 4];
      const int m = u + 16 * j;
      d[3] = -1;
      if (m >= p.M) continue;
      int f0 = -1, f1 = -1;
      for (int k = 0; k < p.K; ++k)
        if (bank[(size_t)k * p.M + m] != 0.f) { if (f0 < 0) f0 = k; f1 = k; }
      d[3] = m;
      d[2] = (int)(mw.size() / 4);
      if (f0 < 0) { d[0] = 0; d[1] = 0; continue; }
      const int s0 = f0 & ~3;
      const int g = (f1 - s0) / 4 + 1;
      d[0] = s0; d[1] = g;
      for (int i = 0; i < 4 * g; ++i) {
        const int k = s0 + i;
        mw.push_back(k < p.K ? bank[(size_t)k * p.M + m] : 0.f);  // bins 257..259 of a P row are never written: weight 0 * stale
      }
    }
  if (mw.empty()) mw.assign(4, 0.f);
  img->blob.clear();
  auto append = [&](const void *src, size_t bytes) -> int {
    const size_t off = img->blob.size();
    img->blob.resize(off + ((bytes + 15) & ~(size_t)15), 0);
    memcpy(img->blob.data() + off, src, bytes);
    return (int)off;
  };
  append(img4.data(), img4.size() * 4);
  img->off_md = append(md.data(), md.size() * 4);
  img->off_mw = append(mw.data(), mw.size() * 4);
  img->fpu = fpu;
  img->win4.assign(512, 0.f);
  for (int i = 0; i < p.L; ++i) img->win4[i] = window[i];
  img->tw.resize(256);
  for (int n2 = 0; n2 < 16; ++n2)
    for (int k1 = 1; k1 <= 16; ++k1) {
      const double a = -2.0 * M_PI * (double)((n2 * k1) % 512) / 512.0;
      img->tw[n2 * 16 + k1 - 1] = make_float2((float)cos(a), (float)sin(a));
    }
  return 0;
}

static inline size_t tc512_smem_bytes(const Tc512Tables &t) { return (size_t)TC_OFF_CONST + (size_t)t.cblob_bytes + 64; }

#ifndef TC_A1_MODE
#define TC_A1_MODE 0
#endif
template <int DT, int DBG, int A1M = TC_A1_MODE>
static int tc512_go(bool launch, size_t smem, const DevPlan &p, const Tc512Tables &t, const DevBatch &b, dim3 grid, cudaStream_t stream,
                    float *dbg) {
  auto kern = b200feat_tc512_kernel<DT, DBG, A1M>;
  if (!launch)
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess ? 0 : B200FEAT_ECUDA;
  kern<<<grid, dim3(TC_THREADS), smem, stream>>>(p, t, b, dbg);
  return 0;
}

static inline int tc512_prepare(DevPlan &p, const std::vector<float> &bank, std::vector<void *> &allocs, int *frames_per_tile,
                                const std::vector<float> &window, Tc512Host *out) {
  Tc512Image img;
  tc512_build_image(p, bank, window, &img);
  Tc512Host hst;
  int rc;
  const unsigned char *d = nullptr;
  if ((rc = tc_upload(img.blob, allocs, &d))) return rc;
  hst.t.cblob = d;
  hst.t.cblob_bytes = (int)img.blob.size();
  hst.t.off_md = img.off_md; hst.t.off_mw = img.off_mw; hst.t.fpu = img.fpu;
  if ((rc = tc_upload(img.win4, allocs, &hst.t.win4))) return rc;
  if ((rc = tc_upload(img.tw, allocs, &hst.t.tw))) return rc;
  hst.smem = tc512_smem_bytes(hst.t);
  // the k1 = 16 row tile is read as 128 rows (16 KB) from 66 KB on: everything up to 82 KB must be mapped
  if (hst.smem < 82 * 1024 + 64) hst.smem = 82 * 1024 + 64;
  if (hst.smem > (size_t)(227 * 1024 / 2) - 1024) return B200FEAT_EUNSUPPORTED;  // two CTAs per SM
  DevBatch none{};
  if (tc512_go<B200FEAT_F32, 0>(false, hst.smem, p, hst.t, none, dim3(1), nullptr, nullptr)) return B200FEAT_ECUDA;
  if (tc512_go<B200FEAT_I16, 0>(false, hst.smem, p, hst.t, none, dim3(1), nullptr, nullptr)) return B200FEAT_ECUDA;
  *out = hst;
  *frames_per_tile = TC_NF;
  return 0;
}

static inline int tc512_launch(const DevPlan &p, const Tc512Host &hst, const DevBatch &b, int dt, int sm_count, cudaStream_t stream) {
  int64_t blocks = b.num_tiles;
  const int64_t cap = (int64_t)sm_count * 2;
  if (blocks > cap) blocks = cap;
  if (dt == B200FEAT_I16) tc512_go<B200FEAT_I16, 0>(true, hst.smem, p, hst.t, b, dim3((unsigned)blocks), stream, nullptr);
  else tc512_go<B200FEAT_F32, 0>(true, hst.smem, p, hst.t, b, dim3((unsigned)blocks), stream, nullptr);
  return (int)cudaGetLastError();
}
