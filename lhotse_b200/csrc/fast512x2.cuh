// Fast fused kernel for fft_length N = 512, Blackwell packed-FP32 edition.
//
// EXPERIMENTAL variant (kernel = "fast_x2"), not what AUTO selects: measured 12 % slower than fast512.cuh
// on the headline workload (profiles/README.md) because halving the instruction count also halves the
// independent instructions per warp that hide the FP-pipe latency at 4 warps/scheduler.
// Same algorithm as the scalar kernel (fast512.cuh) — one HALF-WARP per
// frame, 512-point real FFT as a packed 256-point complex FFT factored 16 x 16 in registers, one
// shared-memory transpose, real-FFT split against the mirrored lane, sparse mel bank, log — but every
// half-warp now carries TWO consecutive frames (A, B) side by side in 64-bit register pairs and all
// floating-point work is issued as sm_100 packed-FP32 instructions (FADD2 / FMUL2 / FFMA2, PTX
// add/mul/fma.rn.f32x2): one issue slot does the arithmetic of both frames, and per-lane constants
// (twiddles, window taps, mel weights) enter as broadcast scalar operands (SASS `Rn.F32`).
// It issues 976 instead of 1250 instructions per lane-frame, but ncu shows 50 % issue-active
// (wait / short-scoreboard / math-pipe-throttle stalls): FFMA2 occupies the FP32 pipe for two cycles.
//
// Replaces, per frame, lhotse/features/kaldi/layers.py:151-186 (Wav2Win._forward_strided),
// :32-42 (_rfft/_pow_spectrogram), :565-578 (mel+log), :708-724 (DCT/lifter), with the framing of
// :727-772 folded into the load addresses.  HBM traffic: 4*S bytes in, 4*F bytes out per frame.
#pragma once
#include <algorithm>
#include <vector>

#include "common.cuh"

#define FX_WARPS 8                    // warps per CTA
#define FX_HW (2 * FX_WARPS)          // half-warps per CTA
#define FX_PAIRS 2                    // frame pairs per half-warp per tile
#define FX_TILE (FX_HW * 2 * FX_PAIRS)  // frames per tile (64)
#define FX_XROW 17                    // float4 per transpose row (16 + 1 pad: 272 B keeps LDS.128 conflict-free)
#define FX_XBUF (16 * FX_XROW)        // float4 per half-warp transpose tile
#define FX_PBINS 264                  // float2 (A,B) power bins per half-warp: 257 + zeroed slack for mel over-reads
#define FX_FULL 0xffffffffu
#ifndef FX_PREFETCH
#define FX_PREFETCH 1
#endif

typedef float2 f2;  // (frame A, frame B)
struct C2 { f2 r, i; };  // packed complex: real parts of A,B and imaginary parts of A,B

__device__ __forceinline__ f2 bc(float s) { return make_float2(s, s); }
__device__ __forceinline__ f2 pneg(f2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ f2 padd(f2 a, f2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ f2 psub(f2 a, f2 b) { return __fadd2_rn(a, pneg(b)); }  // FADD2 a, -b
__device__ __forceinline__ f2 pmul(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 pfma(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
// z * (wr + i*wi) with scalar (broadcast) factors: 4 packed instructions for two frames
__device__ __forceinline__ C2 cmul_s(C2 z, float wr, float wi) {
  C2 o;
  o.r = pfma(z.r, bc(wr), pmul(z.i, bc(-wi)));
  o.i = pfma(z.r, bc(wi), pmul(z.i, bc(wr)));
  return o;
}

// forward 4-point DFT, in place, natural order (16 packed adds)
__device__ __forceinline__ void dft4p(C2 &a0, C2 &a1, C2 &a2, C2 &a3) {
  const f2 s02r = padd(a0.r, a2.r), s02i = padd(a0.i, a2.i);
  const f2 d02r = psub(a0.r, a2.r), d02i = psub(a0.i, a2.i);
  const f2 s13r = padd(a1.r, a3.r), s13i = padd(a1.i, a3.i);
  const f2 d13r = psub(a1.r, a3.r), d13i = psub(a1.i, a3.i);
  a0.r = padd(s02r, s13r); a0.i = padd(s02i, s13i);
  a2.r = psub(s02r, s13r); a2.i = psub(s02i, s13i);
  a1.r = padd(d02r, d13i); a1.i = psub(d02i, d13r);  // d02 - i*d13
  a3.r = psub(d02r, d13i); a3.i = padd(d02i, d13r);  // d02 + i*d13
}

#define FX_C1 0.92387953251128674f  // cos(pi/8)
#define FX_S1 0.38268343236508977f  // sin(pi/8)
#define FX_R2 0.70710678118654752f  // sqrt(1/2)

// forward 16-point DFT in registers (radix 4x4).  Input v[n]; output X[k] lands in v[4*(k&3) + (k>>2)].
__device__ __forceinline__ void dft16p(C2 (&v)[16]) {
#pragma unroll
  for (int b = 0; b < 4; ++b) dft4p(v[b], v[4 + b], v[8 + b], v[12 + b]);
  // v[4c + b] = y[b][c]; twiddle by W16^(b*c), W16^m = (cos(pi m/8), -sin(pi m/8))
  v[5] = cmul_s(v[5], FX_C1, -FX_S1);    // W^1
  v[6] = cmul_s(v[6], FX_R2, -FX_R2);    // W^2
  v[7] = cmul_s(v[7], FX_S1, -FX_C1);    // W^3
  v[9] = cmul_s(v[9], FX_R2, -FX_R2);    // W^2
  { const C2 t = v[10]; v[10].r = t.i; v[10].i = pneg(t.r); }  // W^4 = -i (sign folds into the next adds)
  v[11] = cmul_s(v[11], -FX_R2, -FX_R2); // W^6
  v[13] = cmul_s(v[13], FX_S1, -FX_C1);  // W^3
  v[14] = cmul_s(v[14], -FX_R2, -FX_R2); // W^6
  v[15] = cmul_s(v[15], -FX_C1, FX_S1);  // W^9
#pragma unroll
  for (int c = 0; c < 4; ++c) dft4p(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
}
#define FX_OUT(k) (4 * ((k) & 3) + ((k) >> 2))

// W32^m = exp(-2*pi*i*m/32) as (re, im), m = 0..15
__device__ __forceinline__ float2 fx_w32(int m) {
  const float c[16] = {1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                       0.70710678118654752f, 0.55557023301960218f, 0.38268343236508977f, 0.19509032201612825f,
                       0.0f, -0.19509032201612825f, -0.38268343236508977f, -0.55557023301960218f,
                       -0.70710678118654752f, -0.83146961230254524f, -0.92387953251128674f, -0.98078528040323043f};
  const float s[16] = {0.0f, 0.19509032201612825f, 0.38268343236508977f, 0.55557023301960218f,
                       0.70710678118654752f, 0.83146961230254524f, 0.92387953251128674f, 0.98078528040323043f,
                       1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                       0.70710678118654752f, 0.55557023301960218f, 0.38268343236508977f, 0.19509032201612825f};
  return make_float2(c[m], -s[m]);
}

__device__ __forceinline__ f2 fx_hw_sum(f2 v) {  // per-frame sums over the 16 lanes of each half-warp
#pragma unroll
  for (int o = 8; o > 0; o >>= 1)
    v = padd(v, make_float2(__shfl_xor_sync(FX_FULL, v.x, o, 16), __shfl_xor_sync(FX_FULL, v.y, o, 16)));
  return v;
}

struct FastX2Tables {   // device pointers, derived once per handle
  const float2 *win2;   // [16][16] window taps (w[32*n1+2l], w[32*n1+2l+1]), zero beyond L
  const float2 *tw1;    // [16][16] W256^(l*k1) indexed [k1][l]
  const float2 *w512;   // [16]     W512^l
  const int *rstart;    // [rounds][16] first FFT bin of filter m = lane + 16*round (0 if m >= M)
  const int *rlen;      // [rounds]     trip count of the round = its longest filter
  const int *rrow;      // [rounds]     first row of the round in wdense
  const float *wdense;  // [rows][16]   weights (x 1/4 or 1/2, see below), zero-padded to the round's trip count
  int mel_rounds;       // ceil(M / 16)
  int mel_wrows;        // sum of rlen
};

static inline size_t fastx2_smem_bytes(const DevPlan &p, const FastX2Tables &t) {
  size_t b = (size_t)FX_HW * (FX_XBUF * 16 + FX_PBINS * 8);
  b += 256 * 8 * 2;                                   // window taps + stage-1 twiddles
  b += (size_t)t.mel_rounds * 16 * 4 + (size_t)t.mel_rounds * 8;
  b += (size_t)t.mel_wrows * 16 * 4;
  (void)p;
  return (b + 15) & ~(size_t)15;
}

template <int DT> struct SampleT { typedef float type; };
template <> struct SampleT<B200FEAT_I16> { typedef int16_t type; };
__device__ __forceinline__ float fx_cvt(float v) { return v; }
__device__ __forceinline__ float fx_cvt(int16_t v) { return (float)v * (1.0f / 32768.0f); }

template <int DT>
__device__ __forceinline__ float fx_ld(const void *base, int64_t i) {
  if (DT == B200FEAT_I16) return (float)__ldg(reinterpret_cast<const int16_t *>(base) + i) * (1.0f / 32768.0f);
  return __ldg(reinterpret_cast<const float *>(base) + i);
}

template <int DT, int LCT>
__global__ void __launch_bounds__(FX_WARPS * 32, 2)
b200feat_fast512x2_kernel(const DevPlan p, const FastX2Tables ft, const DevBatch b) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int l = tid & 15;           // lane within the half-warp
  const int hw = tid >> 4;          // half-warp within the CTA
  const int L = LCT ? LCT : p.L;
  constexpr int NP = LCT ? (LCT + 31) / 32 : 16;  // n1 rows that carry samples

  // ---- shared memory carve-up
  float4 *xall = reinterpret_cast<float4 *>(smem_raw);
  float2 *pall = reinterpret_cast<float2 *>(xall + (size_t)FX_HW * FX_XBUF);
  float2 *s_win = pall + (size_t)FX_HW * FX_PBINS;
  float2 *s_tw1 = s_win + 256;
  int *s_rstart = reinterpret_cast<int *>(s_tw1 + 256);
  int *s_rlen = s_rstart + ft.mel_rounds * 16;
  int *s_rrow = s_rlen + ft.mel_rounds;
  float *s_mw = reinterpret_cast<float *>(s_rrow + ft.mel_rounds);
  float4 *X = xall + (size_t)hw * FX_XBUF;
  float2 *P = pall + (size_t)hw * FX_PBINS;  // P[k] = (|2X_A[k]|^2, |2X_B[k]|^2)

  for (int i = tid; i < 256; i += blockDim.x) { s_win[i] = __ldg(ft.win2 + i); s_tw1[i] = __ldg(ft.tw1 + i); }
  for (int i = tid; i < ft.mel_rounds * 16; i += blockDim.x) s_rstart[i] = __ldg(ft.rstart + i);
  for (int i = tid; i < ft.mel_rounds; i += blockDim.x) { s_rlen[i] = __ldg(ft.rlen + i); s_rrow[i] = __ldg(ft.rrow + i); }
  for (int i = tid; i < ft.mel_wrows * 16; i += blockDim.x) s_mw[i] = __ldg(ft.wdense + i);
  // P is read past a filter's support with zero weights: it must never hold NaN patterns
  for (int i = tid; i < FX_HW * FX_PBINS; i += blockDim.x) pall[i] = make_float2(0.f, 0.f);

  const float2 w512l = __ldg(ft.w512 + l);
  const int partner = (16 - l) & 15;
  const float inv_L = 1.0f / (float)L;
  const size_t esz = DT == B200FEAT_I16 ? 2 : 4;
  __syncthreads();

  for (int64_t tg = blockIdx.x; tg < b.num_tiles; tg += gridDim.x) {
    const int64_t tile = b.tile_base + tg;
    const int cut = __ldg(b.tile_cut + tile) - b.batch_first;  // host-built tile->cut table
    const int64_t tq = (tile - __ldg(b.tile_off + cut)) * FX_TILE + (int64_t)hw * (2 * FX_PAIRS);
    const int64_t T = __ldg(b.row_off + cut + 1) - __ldg(b.row_off + cut);
    const int64_t rows_here = b.out_mode == B200FEAT_OUT_PADDED ? b.max_frames : T;
    if (!__any_sync(FX_FULL, tq < rows_here)) continue;  // both halves idle for this tile
    const int64_t n = __ldg(b.nsamp + cut);
    const int64_t xoff = __ldg(b.samp_off + cut);
    const int64_t rowq = b.out_mode == B200FEAT_OUT_PADDED ? (int64_t)(b.batch_first + cut) * b.max_frames + tq
                                                           : __ldg(b.row_off + cut) + tq;

#pragma unroll 1
    for (int pr = 0; pr < FX_PAIRS; ++pr) {
      const int64_t t0 = tq + 2 * pr;
      if (!__any_sync(FX_FULL, t0 < rows_here)) continue;
      const bool compute = __any_sync(FX_FULL, t0 < T);
      // frames beyond the cut are clamped to its last frame: computed, never stored
      const int64_t tA = min(t0, T - 1), tB = min(t0 + 1, T - 1);
      const int64_t baseA = tA * p.S - (p.snip_edges ? 0 : p.pad_left);
      const int64_t baseB = tB * p.S - (p.snip_edges ? 0 : p.pad_left);
      f2 le2 = make_float2(0.f, 0.f);
      if (compute) {
        C2 v[16];
        if (FX_PREFETCH && l < 11) {  // the 2*S new samples of this half-warp's next pair
          const int64_t nx = baseA + L + p.S + 32 * l;
          if (nx >= 0 && nx + 32 <= n) {
            const char *pp = reinterpret_cast<const char *>(b.samples) + (xoff + nx) * esz;
            asm volatile("prefetch.global.L1 [%0];" ::"l"(pp));
          }
        }
        // ---- gather (layers.py:753-772) + DC removal (:155-157)
        f2 pv[NP];
        const bool interior = baseA >= 0 && baseB + L <= n;
        if (__all_sync(FX_FULL, interior)) {
          typedef typename SampleT<DT>::type ST;
          const ST *xa = reinterpret_cast<const ST *>(b.samples) + (xoff + baseA + 2 * l);
          const ST *xb = reinterpret_cast<const ST *>(b.samples) + (xoff + baseB + 2 * l);
#pragma unroll
          for (int n1 = 0; n1 < NP; ++n1) {
            const int j0 = 32 * n1 + 2 * l;
            v[n1].r = v[n1].i = pv[n1] = make_float2(0.f, 0.f);
            if (j0 < L) {
              v[n1].r = make_float2(fx_cvt(__ldg(xa + 32 * n1)), fx_cvt(__ldg(xb + 32 * n1)));
              const int back = j0 > 0 ? 1 : 0;
              pv[n1] = make_float2(fx_cvt(__ldg(xa + 32 * n1 - back)), fx_cvt(__ldg(xb + 32 * n1 - back)));
            }
            if (j0 + 1 < L) v[n1].i = make_float2(fx_cvt(__ldg(xa + 32 * n1 + 1)), fx_cvt(__ldg(xb + 32 * n1 + 1)));
          }
        } else {  // a cut edge in this warp: per-tap reflection; ~3 frames per cut
#pragma unroll
          for (int n1 = 0; n1 < NP; ++n1) {
            const int j0 = 32 * n1 + 2 * l;
            v[n1].r = v[n1].i = pv[n1] = make_float2(0.f, 0.f);
            if (j0 < L) {
              int64_t a = baseA + j0, c = baseB + j0, ap = baseA + (j0 > 0 ? j0 - 1 : 0), cp = baseB + (j0 > 0 ? j0 - 1 : 0);
              if (!p.snip_edges) { a = reflect_index(a, n); c = reflect_index(c, n); ap = reflect_index(ap, n); cp = reflect_index(cp, n); }
              v[n1].r = make_float2(fx_ld<DT>(b.samples, xoff + a), fx_ld<DT>(b.samples, xoff + c));
              pv[n1] = make_float2(fx_ld<DT>(b.samples, xoff + ap), fx_ld<DT>(b.samples, xoff + cp));
            }
            if (j0 + 1 < L) {
              int64_t a = baseA + j0 + 1, c = baseB + j0 + 1;
              if (!p.snip_edges) { a = reflect_index(a, n); c = reflect_index(c, n); }
              v[n1].i = make_float2(fx_ld<DT>(b.samples, xoff + a), fx_ld<DT>(b.samples, xoff + c));
            }
          }
        }
        f2 s = make_float2(0.f, 0.f);
#pragma unroll
        for (int n1 = 0; n1 < NP; ++n1) s = padd(s, padd(v[n1].r, v[n1].i));  // taps beyond L are exact zeros
        const f2 nmu = p.remove_dc ? pmul(fx_hw_sum(s), bc(-inv_L)) : make_float2(0.f, 0.f);  // -mean
        // ---- energy, pre-emphasis, window (layers.py:159-170); zero padding is implicit
        f2 e = make_float2(0.f, 0.f);
        const float npre = -p.preemph;
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) {
          if (n1 < NP) {
            const int j0 = 32 * n1 + 2 * l;
            const float2 w = s_win[n1 * 16 + l];  // zero beyond L
            f2 da = padd(v[n1].r, nmu), dc = padd(v[n1].i, nmu);
            const f2 dp = padd(pv[n1], nmu);
            if (p.use_energy) {
              if (j0 >= L) da = make_float2(0.f, 0.f);
              if (j0 + 1 >= L) dc = make_float2(0.f, 0.f);
              if (p.raw_energy) e = pfma(da, da, pfma(dc, dc, e));
            }
            const f2 ya = pmul(pfma(dp, bc(npre), da), bc(w.x));
            const f2 yc = pmul(pfma(da, bc(npre), dc), bc(w.y));
            if (p.use_energy && !p.raw_energy) e = pfma(ya, ya, pfma(yc, yc, e));
            v[n1].r = ya; v[n1].i = yc;
          } else {
            v[n1].r = v[n1].i = make_float2(0.f, 0.f);
          }
        }
        if (p.use_energy) {
          const f2 es = fx_hw_sum(e);
          le2 = make_float2(log_energy_value(p, es.x), log_energy_value(p, es.y));
        }

        // ---- stage 1: radix-16 over n1, twiddle W256^(l*k1), transpose through shared memory
        dft16p(v);
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) {
          C2 y = v[FX_OUT(k1)];
          if (k1 > 0) { const float2 tw = s_tw1[k1 * 16 + l]; y = cmul_s(y, tw.x, tw.y); }
          X[k1 * FX_XROW + l] = make_float4(y.r.x, y.r.y, y.i.x, y.i.y);
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const float4 r = X[l * FX_XROW + q];
          v[q].r = make_float2(r.x, r.y);
          v[q].i = make_float2(r.z, r.w);
        }
        __syncwarp();
        // ---- stage 2: radix-16 over n2 -> Z[l + 16*k2]
        dft16p(v);
        // ---- real-FFT split + power (layers.py:38-42).  E = Z[k] + conj(Z[256-k]), O = Z[k] - conj(Z[256-k]),
        // T = W512^k * O:  2*X[k] = E - i*T and 2*conj(X[256-k]) = E + i*T.  Each lane handles its EVEN k2 and gets
        // the mirror lane's ODD slots; lane 0 mirrors itself shifted by one slot and walks
        // (0,0) (2,14) (4,12) (6,10) (8,8) (1,15) (3,13) (5,11) here and (7,9) below.  The 1/4 of |2X|^2 lives in
        // the mel weights / spectrogram epilogue.
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          constexpr int kOwn0[8] = {0, 2, 4, 6, 8, 1, 3, 5};
          constexpr int kSend0[8] = {0, 14, 12, 10, 8, 15, 13, 11};
          C2 zk = v[FX_OUT(2 * i)];
          C2 zs = v[FX_OUT(15 - 2 * i)];
          if (l == 0) { zs = v[FX_OUT(kSend0[i])]; if (i >= 5) zk = v[FX_OUT(kOwn0[i])]; }
          C2 zc;
          zc.r = make_float2(__shfl_sync(FX_FULL, zs.r.x, partner, 16), __shfl_sync(FX_FULL, zs.r.y, partner, 16));
          zc.i = make_float2(__shfl_sync(FX_FULL, zs.i.x, partner, 16), __shfl_sync(FX_FULL, zs.i.y, partner, 16));
          C2 E, O;
          E.r = padd(zk.r, zc.r); E.i = psub(zk.i, zc.i);
          O.r = psub(zk.r, zc.r); O.i = padd(zk.i, zc.i);
          float2 wc = fx_w32(2 * i);  // W16^i; lane 0 needs W32^(own slot)
          if (i >= 5) { const float2 w0 = fx_w32(kOwn0[i]); wc = l == 0 ? w0 : wc; }
          const C2 tt = cmul_s(cmul_s(O, wc.x, wc.y), w512l.x, w512l.y);
          const f2 ar = padd(E.r, tt.i), ai = psub(E.i, tt.r);  // 2*X[k]
          const f2 br = psub(E.r, tt.i), bi = padd(E.i, tt.r);  // 2*conj(X[256-k])
          f2 pa = pfma(ar, ar, pmul(ai, ai)), pb = pfma(br, br, pmul(bi, bi));
          if (p.use_mag) { pa = make_float2(sqrtf(pa.x), sqrtf(pa.y)); pb = make_float2(sqrtf(pb.x), sqrtf(pb.y)); }
          const int k = l == 0 ? 16 * kOwn0[i] : l + 32 * i;
          P[k] = pa;
          P[256 - k] = pb;
        }
        if (l == 0) {  // lane 0's last pair: slots (7, 9) -> bins 112 and 144
          const C2 zk = v[FX_OUT(7)], zc = v[FX_OUT(9)];
          C2 E, O;
          E.r = padd(zk.r, zc.r); E.i = psub(zk.i, zc.i);
          O.r = psub(zk.r, zc.r); O.i = padd(zk.i, zc.i);
          const float2 wc = fx_w32(7);
          const C2 tt = cmul_s(O, wc.x, wc.y);
          const f2 ar = padd(E.r, tt.i), ai = psub(E.i, tt.r), br = psub(E.r, tt.i), bi = padd(E.i, tt.r);
          f2 pa = pfma(ar, ar, pmul(ai, ai)), pb = pfma(br, br, pmul(bi, bi));
          if (p.use_mag) { pa = make_float2(sqrtf(pa.x), sqrtf(pa.y)); pb = make_float2(sqrtf(pb.x), sqrtf(pb.y)); }
          P[112] = pa;
          P[144] = pb;
        }
        __syncwarp();
      }

      // ---- epilogue for frames t0 (A) and t0+1 (B) of this half-warp
      const bool vA = t0 < T, vB = t0 + 1 < T;                     // real frames
      const bool rA = t0 < rows_here, rB = t0 + 1 < rows_here;     // rows that exist (padded mode)
      float *outA = b.out + (rowq + 2 * pr) * p.F, *outB = outA + p.F;
      if (p.feature == B200FEAT_SPECTROGRAM || p.feature == B200FEAT_LOG_SPECTROGRAM) {
        const float sc = p.use_mag ? 0.5f : 0.25f;  // P holds |2X|^2 (or |2X|)
        for (int k = l; k < p.K; k += 16) {
          const float2 pw = P[k];
          float xa = pw.x * sc, xb = pw.y * sc;
          if (p.feature == B200FEAT_LOG_SPECTROGRAM) { xa = log_spec_value(p, xa); xb = log_spec_value(p, xb); }
          if (k == 0 && p.use_energy) { xa = le2.x; xb = le2.y; }
          if (vA) outA[k] = xa; else if (rA) outA[k] = b.pad_value;
          if (vB) outB[k] = xb; else if (rB) outB[k] = b.pad_value;
        }
      } else {
        const int shift = (p.feature == B200FEAT_FBANK && p.use_energy) ? 1 : 0;
        const int Mpad = (p.M + 3) & ~3;
        float *mlog = reinterpret_cast<float *>(X);  // the transpose tile is idle during the epilogue
        for (int j = 0; j < ft.mel_rounds; ++j) {
          const int m = l + 16 * j;
          const float2 *Pj = P + s_rstart[j * 16 + l];
          const float *wj = s_mw + s_rrow[j] * 16 + l;
          const int len = s_rlen[j];  // uniform: shorter filters continue on zero weights
          f2 acc = make_float2(0.f, 0.f);
#pragma unroll 4
          for (int i = 0; i < len; ++i) acc = pfma(Pj[i], bc(wj[i * 16]), acc);
          if (m < p.M) {
            const float ra = fast_log_normal(nanmax(acc.x, p.mel_floor)), rb = fast_log_normal(nanmax(acc.y, p.mel_floor));
            if (p.feature == B200FEAT_FBANK) {
              if (vA) outA[m + shift] = ra;
              if (vB) outB[m + shift] = rb;
            } else {
              mlog[m] = ra;
              mlog[Mpad + m] = rb;
            }
          }
        }
        if (p.feature == B200FEAT_FBANK) {
          if (shift && l == 0) { if (vA) outA[0] = le2.x; if (vB) outB[0] = le2.y; }
        } else {
          __syncwarp();
          for (int idx = l; idx < 2 * p.C; idx += 16) {
            const int fsel = idx >= p.C, c = idx - fsel * p.C;
            float acc = 0.f;
            for (int m = 0; m < p.M; ++m) acc = fmaf(mlog[fsel * Mpad + m], __ldg(p.dct + m * p.C + c), acc);
            if (p.use_lifter) acc *= __ldg(p.lifter + c);
            if (p.use_energy && c == 0) acc = fsel ? le2.y : le2.x;
            if (fsel ? vB : vA) (fsel ? outB : outA)[c] = acc;
          }
          __syncwarp();
        }
        if (!vA && rA) for (int k = l; k < p.F; k += 16) outA[k] = b.pad_value;
        if (!vB && rB) for (int k = l; k < p.F; k += 16) outB[k] = b.pad_value;
      }
      __syncwarp();
    }
  }
}

// ---------------------------------------------------------------------------------------------- host
struct FastX2Host {
  FastX2Tables t;
  size_t smem;
};

static inline bool fastx2_supported(const DevPlan &p) {
  return p.N == 512 && p.packed && p.L >= 2 && p.L <= 512 && p.C <= 128 && 2 * ((p.M + 3) & ~3) * 4 <= FX_XBUF * 16;
}

template <typename T>
static int fx_upload(const std::vector<T> &h, std::vector<void *> &allocs, const T **out) {
  void *d = nullptr;
  if (cudaMalloc(&d, h.size() * sizeof(T)) != cudaSuccess) return B200FEAT_ECUDA;
  allocs.push_back(d);
  if (cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) return B200FEAT_ECUDA;
  *out = reinterpret_cast<const T *>(d);
  return 0;
}

template <int DT, int LCT>
static int fx_set_attr(size_t smem) {
  return cudaFuncSetAttribute(b200feat_fast512x2_kernel<DT, LCT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess
             ? 0 : B200FEAT_ECUDA;
}

static inline int fastx2_prepare(DevPlan &p, const std::vector<float> &bank, std::vector<void *> &allocs,
                                 int *frames_per_tile, const std::vector<float> &window, FastX2Host *out) {
  FastX2Host hst;
  std::vector<float2> win2(256), tw1(256), w512(16);
  for (int n1 = 0; n1 < 16; ++n1)
    for (int l = 0; l < 16; ++l) {
      const int j0 = 32 * n1 + 2 * l;
      win2[n1 * 16 + l] = make_float2(j0 < p.L ? window[j0] : 0.f, j0 + 1 < p.L ? window[j0 + 1] : 0.f);
    }
  for (int k1 = 0; k1 < 16; ++k1)
    for (int l = 0; l < 16; ++l) {
      const double a = -2.0 * M_PI * (double)((l * k1) % 256) / 256.0;
      tw1[k1 * 16 + l] = make_float2((float)cos(a), (float)sin(a));
    }
  for (int l = 0; l < 16; ++l) {
    const double a = -2.0 * M_PI * (double)l / 512.0;
    w512[l] = make_float2((float)cos(a), (float)sin(a));
  }
  int rc;
  if ((rc = fx_upload(win2, allocs, &hst.t.win2))) return rc;
  if ((rc = fx_upload(tw1, allocs, &hst.t.tw1))) return rc;
  if ((rc = fx_upload(w512, allocs, &hst.t.w512))) return rc;
  // mel bank re-packed for the epilogue (pack_mel_rounds, common.cuh); 1/4 (1/2) of |2X|^2 (|2X|) folded into the weights
  const MelRounds mr = pack_mel_rounds(bank, p.K, p.M, p.use_mag ? 0.5f : 0.25f);
  if (mr.max_reach > FX_PBINS) return B200FEAT_EUNSUPPORTED;  // over-reads must stay inside the (zero-padded) P tile
  const int rounds = mr.rounds;
  const std::vector<int> &rstart = mr.rstart, &rlen = mr.rlen, &rrow = mr.rrow;
  const std::vector<float> &wdense = mr.wdense;
  if ((rc = fx_upload(rstart, allocs, &hst.t.rstart))) return rc;
  if ((rc = fx_upload(rlen, allocs, &hst.t.rlen))) return rc;
  if ((rc = fx_upload(rrow, allocs, &hst.t.rrow))) return rc;
  if ((rc = fx_upload(wdense, allocs, &hst.t.wdense))) return rc;
  hst.t.mel_rounds = rounds;
  hst.t.mel_wrows = rounds ? (int)(wdense.size() / 16) : 0;
  hst.smem = fastx2_smem_bytes(p, hst.t);
  if (hst.smem > 113 * 1024) return B200FEAT_EUNSUPPORTED;  // keep 2 CTAs per SM
  if (fx_set_attr<B200FEAT_F32, 400>(hst.smem) || fx_set_attr<B200FEAT_I16, 400>(hst.smem) ||
      fx_set_attr<B200FEAT_F32, 0>(hst.smem) || fx_set_attr<B200FEAT_I16, 0>(hst.smem))
    return B200FEAT_ECUDA;
  *out = hst;
  *frames_per_tile = FX_TILE;
  return 0;
}

static inline int fastx2_launch(const DevPlan &p, const FastX2Host &hst, const DevBatch &b, int dt, int sm_count,
                                cudaStream_t stream) {
  int64_t blocks = b.num_tiles;
  const int64_t cap = (int64_t)sm_count * 2;
  if (blocks > cap) blocks = cap;
  const dim3 grid((unsigned)blocks), block(FX_WARPS * 32);
  if (p.L == 400) {
    if (dt == B200FEAT_I16) b200feat_fast512x2_kernel<B200FEAT_I16, 400><<<grid, block, hst.smem, stream>>>(p, hst.t, b);
    else b200feat_fast512x2_kernel<B200FEAT_F32, 400><<<grid, block, hst.smem, stream>>>(p, hst.t, b);
  } else {
    if (dt == B200FEAT_I16) b200feat_fast512x2_kernel<B200FEAT_I16, 0><<<grid, block, hst.smem, stream>>>(p, hst.t, b);
    else b200feat_fast512x2_kernel<B200FEAT_F32, 0><<<grid, block, hst.smem, stream>>>(p, hst.t, b);
  }
  return (int)cudaGetLastError();
}
