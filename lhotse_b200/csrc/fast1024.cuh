// Fast fused kernel for fft_length N = 1024 (22.05 / 24 kHz with 25 ms frames: L = 551 / 600, and any plan with
// 512 < L <= 1024): one WARP per frame, built from the N = 512 machinery (fast512.cuh) plus one decimation-in-time step.
//
//   The frame's even samples e[n] = y[2n] go to half-warp 0 and its odd samples o[n] = y[2n+1] to half-warp 1; each
//   half-warp runs the 512-point real FFT of fast512.cuh on its sub-sequence (packed 256-point complex FFT, 16 x 16 in
//   registers, one shared-memory transpose, paired split) and ends up, for the same (lane, item), holding E[k], E[256-k]
//   resp. O[k], O[256-k] for the SAME k.  The 1024-point spectrum follows from
//       X[k] = E[k] + W1024^k O[k],      X[512-k] = conj(E[k] - W1024^k O[k]),       k = 0..256:
//   half-warp 1 multiplies its values by the twiddle, the halves swap values with one shuffle-xor-16 each, and each half
//   produces two of the four power bins {k, 256-k} / {512-k, 256+k}.
//   Power spectra of 4 consecutive frames are staged as P[frame][bin] (513 bins) and the mel bank runs in rounds of 32.
//
// Replaces the same reference code as fast512.cuh (lhotse/features/kaldi/layers.py:151-186, :32-42, :565-578, :708-724,
// framing :727-772).
#pragma once
#include "fast512.cuh"

#define F1K_PBINS 516                      // floats per P row (513 bins + pad)
#define F1K_PTAIL 64

// W64^s = exp(-2*pi*i*s/64), s = 0..8 (lane 0's combine twiddles W1024^(16 s))
__device__ __forceinline__ float2 w64_const(int s) {
  const float c[9] = {1.0f, 0.99518472667219693f, 0.98078528040323043f, 0.95694033573220882f, 0.92387953251128674f,
                      0.88192126434835505f, 0.83146961230254524f, 0.77301045336273699f, 0.70710678118654752f};
  const float sn[9] = {0.0f, 0.09801714032956060f, 0.19509032201612825f, 0.29028467725446233f, 0.38268343236508977f,
                       0.47139673682599764f, 0.55557023301960218f, 0.63439328416364549f, 0.70710678118654752f};
  return make_float2(c[s], -sn[s]);
}

struct Fast1024Tables {
  // one 16-byte-aligned blob (TMA bulk copy):
  //   [win2: 2*16*16 float2 (w[64 n1 + 4 l + h], w[64 n1 + 4 l + h + 2]) indexed [h][n1][l], zero beyond L]
  //   [rstart: rounds*32 int | rlen: rounds | rrow: rounds | wdense: rows*32 float]
  const void *cblob;
  int cblob_bytes;
  int off_rstart, off_rlen, off_rrow, off_mw;
  const float2 *tw1;     // [16][16] W256^(l*k1)
  const float2 *w512;    // [16]     W512^l      (split of the 512-point sub-FFTs)
  const float2 *w1024;   // [16]     W1024^l     (radix-2 combination)
  int mel_rounds, mel_wrows;
};

static inline size_t fast1024_smem_bytes(const Fast1024Tables &t, int warps, int slots) {
  size_t b = (size_t)(2 * warps) * F512_XBUF * 8 + (size_t)warps * F1K_PBINS * slots * 4 + F1K_PTAIL * 4;
  b += (size_t)t.cblob_bytes + 16;
  return (b + 15) & ~(size_t)15;
}

template <int DT, int LCT, int WARPS, int SLOTS>
__global__ void __launch_bounds__(WARPS * 32, 2)
b200feat_fast1024_kernel(const DevPlan p, const Fast1024Tables ft, const DevBatch b) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int l = tid & 15;            // lane within the half-warp
  const int h = (tid >> 4) & 1;      // 0: even samples, 1: odd samples
  const int w = tid >> 5;            // warp = frame owner
  const int L = LCT ? LCT : p.L;
  constexpr int NP = LCT ? (LCT + 63) / 64 : 16;  // rows of 64 samples that carry data

  float2 *xall = reinterpret_cast<float2 *>(smem_raw);
  float *pall = reinterpret_cast<float *>(xall + (size_t)(2 * WARPS) * F512_XBUF);
  unsigned char *s_const = reinterpret_cast<unsigned char *>(pall + (size_t)WARPS * (F1K_PBINS * SLOTS) + F1K_PTAIL);
  const float2 *s_win = reinterpret_cast<const float2 *>(s_const);
  const int *s_rstart = reinterpret_cast<const int *>(s_const + ft.off_rstart);
  const int *s_rlen = reinterpret_cast<const int *>(s_const + ft.off_rlen);
  const int *s_rrow = reinterpret_cast<const int *>(s_const + ft.off_rrow);
  const float *s_mw = reinterpret_cast<const float *>(s_const + ft.off_mw);
  unsigned long long *s_bar = reinterpret_cast<unsigned long long *>(s_const + ft.cblob_bytes);
  float2 *X = xall + (size_t)(tid >> 4) * F512_XBUF;   // per half-warp transpose tile
  float *P = pall + (size_t)w * (F1K_PBINS * SLOTS);              // per warp: [slot][F1K_PBINS]

  const unsigned bar = f512_smem_u32(s_bar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {  // constant tables: one TMA bulk copy
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(ft.cblob_bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(f512_smem_u32(s_const)), "l"(ft.cblob), "r"(ft.cblob_bytes), "r"(bar) : "memory");
  }
  for (int i = tid; i < WARPS * (F1K_PBINS * SLOTS) + F1K_PTAIL; i += blockDim.x) pall[i] = 0.f;

  float2 tw1[16];
#pragma unroll
  for (int k1 = 1; k1 < 16; ++k1) tw1[k1] = __ldg(ft.tw1 + k1 * 16 + l);
  const float2 w512l = __ldg(ft.w512 + l);
  const float2 w1024l = __ldg(ft.w1024 + l);
  const float sgn = h ? -1.0f : 1.0f;
  const int partner = (16 - l) & 15;
  const float inv_L = 1.0f / (float)L;
  {
    unsigned done = 0;
    while (!done)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(done) : "r"(bar), "r"(0u) : "memory");
  }
  __syncthreads();

  // radix-2 combination of one (k, 256-k) pair: xa = 2*Xh[k], xb = 2*Xh[256-k] of this half's sub-FFT, wk = W1024^k
  auto combine = [&](float2 xa, float2 xb, float2 wk, unsigned mask, int k, float *Pf) {
    const float2 wk2 = f2mi(f2conj(wk));                     // W1024^(256-k) = -i * conj(W1024^k)
    const float2 m1 = h ? f2mul(xa, wk) : xa;                // half 1 contributes W^k * O[k]
    const float2 m2 = h ? f2mul(xb, wk2) : xb;
    const float2 t1 = make_float2(__shfl_xor_sync(mask, m1.x, 16), __shfl_xor_sync(mask, m1.y, 16));
    const float2 t2 = make_float2(__shfl_xor_sync(mask, m2.x, 16), __shfl_xor_sync(mask, m2.y, 16));
    // half 0: E + W*O (bins k, 256-k); half 1: W*O - E, same modulus as E - W*O (bins 512-k, 256+k)
    const float r1x = fmaf(sgn, t1.x, m1.x), r1y = fmaf(sgn, t1.y, m1.y);
    const float r2x = fmaf(sgn, t2.x, m2.x), r2y = fmaf(sgn, t2.y, m2.y);
    float p1 = fmaf(r1x, r1x, r1y * r1y), p2 = fmaf(r2x, r2x, r2y * r2y);
    if (p.use_mag) { p1 = sqrtf(p1); p2 = sqrtf(p2); }
    Pf[h ? 512 - k : k] = p1;
    Pf[h ? 256 + k : 256 - k] = p2;
  };

  for (int64_t tg = blockIdx.x; tg < b.num_tiles; tg += gridDim.x) {
    const int64_t tile = b.tile_base + tg;
    const int cut = __ldg(b.tile_cut + tile) - b.batch_first;
    const int64_t t0 = (tile - __ldg(b.tile_off + cut)) * (WARPS * SLOTS) + (int64_t)w * SLOTS;
    const int64_t T = __ldg(b.row_off + cut + 1) - __ldg(b.row_off + cut);
    const int64_t rows_here = b.out_mode == B200FEAT_OUT_PADDED ? b.max_frames : T;
    if (t0 >= rows_here) continue;  // warp-uniform: the whole warp owns these frames
    const int64_t n = __ldg(b.nsamp + cut);
    const int64_t xoff = __ldg(b.samp_off + cut);
    const int64_t row0 = b.out_mode == B200FEAT_OUT_PADDED ? (int64_t)(b.batch_first + cut) * b.max_frames + t0
                                                           : __ldg(b.row_off + cut) + t0;
    float le[SLOTS];
#pragma unroll
    for (int k = 0; k < SLOTS; ++k) le[k] = 0.f;

#pragma unroll 1
    for (int f = 0; f < SLOTS; ++f) {
      const int64_t t = t0 + f;
      if (t >= T) continue;  // warp-uniform
      const int64_t base = t * p.S - (p.snip_edges ? 0 : p.pad_left);
      float2 v[16];
      float2 pv[NP];  // (sample before .x, sample before .y)
      const bool interior = base >= 0 && base + L <= n;
      if (interior) {
        const int64_t x0 = xoff + base + 4 * l + h;
#pragma unroll
        for (int n1 = 0; n1 < NP; ++n1) {
          const int ja = 64 * n1 + 4 * l + h;
          v[n1] = pv[n1] = make_float2(0.f, 0.f);
          if (ja < L) {
            v[n1].x = ld_sample<DT>(b.samples, x0 + 64 * n1);
          }
          if (ja + 2 < L) {
            v[n1].y = ld_sample<DT>(b.samples, x0 + 64 * n1 + 2);
          }
        }
      } else {  // cut edge: per-tap reflection (layers.py:753-772)
#pragma unroll
        for (int n1 = 0; n1 < NP; ++n1) {
          const int ja = 64 * n1 + 4 * l + h;
          v[n1] = pv[n1] = make_float2(0.f, 0.f);
          if (ja < L) {
            int64_t i = base + ja;
            if (!p.snip_edges) i = reflect_index(i, n, p.pad_mode);
            v[n1].x = ld_sample<DT>(b.samples, xoff + i);
          }
          if (ja + 2 < L) {
            int64_t i = base + ja + 2;
            if (!p.snip_edges) i = reflect_index(i, n, p.pad_mode);
            v[n1].y = ld_sample<DT>(b.samples, xoff + i);
          }
        }
      }
      {  // the taps before ja and ja + 2 live in neighbouring lanes: two shuffles instead of two more loads
        //   y[ja - 1]: h = 1 -> (l, 0).x ; h = 0 -> (l - 1, 1).y, and for l = 0 the previous row's (15, 1).y
        //   y[ja + 1]: h = 0 -> (l, 1).x ; h = 1 -> (l, 0).y
        float carry = v[0].x;  // lane (0, 0), row 0: replicate-left (layers.py:166)
        const int src1 = h ? lane - 16 : 16 + ((l + 15) & 15);
#pragma unroll
        for (int n1 = 0; n1 < NP; ++n1) {
          const float r1 = __shfl_sync(F512_FULL, h ? v[n1].y : v[n1].x, src1);
          const float r2 = __shfl_xor_sync(F512_FULL, h ? v[n1].x : v[n1].y, 16);
          pv[n1] = make_float2(lane == 0 ? carry : r1, r2);
          carry = r1;
        }
      }
      // ---- DC removal over the whole frame (both halves), energy, pre-emphasis, window (layers.py:155-170)
      float s = 0.f;
#pragma unroll
      for (int n1 = 0; n1 < NP; ++n1) s += v[n1].x + v[n1].y;
      const float mu = p.remove_dc ? warp_sum(s) * inv_L : 0.f;
      float e = 0.f;
#pragma unroll
      for (int n1 = 0; n1 < 16; ++n1) {
        if (n1 < NP) {
          const int ja = 64 * n1 + 4 * l + h;
          const float2 wv = s_win[(h * 16 + n1) * 16 + l];  // zero beyond L
          float2 d = f2add(v[n1], make_float2(-mu, -mu));
          const float2 dp = f2add(pv[n1], make_float2(-mu, -mu));
          if (ja >= L) d.x = 0.f;
          if (ja + 2 >= L) d.y = 0.f;
          if (p.raw_energy) e = fmaf(d.x, d.x, fmaf(d.y, d.y, e));
          const float2 y = __fmul2_rn(__ffma2_rn(dp, make_float2(-p.preemph, -p.preemph), d), wv);
          if (!p.raw_energy) e = fmaf(y.x, y.x, fmaf(y.y, y.y, e));
          v[n1] = y;
        } else {
          v[n1] = make_float2(0.f, 0.f);
        }
      }
      if (p.use_energy) {  // le[] stays in registers: no dynamic indexing
        const float lev = log_energy_value(p, warp_sum(e));
#pragma unroll
        for (int k = 0; k < SLOTS; ++k) le[k] = (f == k) ? lev : le[k];
      }

      // ---- 512-point real FFT of this half's sub-sequence: stage 1, transpose, stage 2 (as in fast512.cuh)
      dft16(v);
#pragma unroll
      for (int k1 = 0; k1 < 16; ++k1) {
        float2 y = v[F512_OUT(k1)];
        if (k1 > 0) y = f2mul(y, tw1[k1]);
        X[k1 * F512_XROW + l] = y;
      }
      __syncwarp();
      {
        const float4 *row = reinterpret_cast<const float4 *>(X + l * F512_XROW);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 r = row[q];
          v[2 * q] = make_float2(r.x, r.y);
          v[2 * q + 1] = make_float2(r.z, r.w);
        }
      }
      __syncwarp();
      dft16(v);
      // ---- paired split of the sub-FFT (see fast512.cuh), then the radix-2 combination across the two half-warps
      float *Pf = P + f * F1K_PBINS;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        constexpr int kOwn0[8] = {0, 2, 4, 6, 8, 1, 3, 5};
        constexpr int kSend0[8] = {0, 14, 12, 10, 8, 15, 13, 11};
        const float2 zo = v[F512_OUT(2 * i)], zo0 = v[F512_OUT(kOwn0[i])];
        const float2 zs = v[F512_OUT(15 - 2 * i)], zs0 = v[F512_OUT(kSend0[i])];
        const float2 zk = (i >= 5 && l == 0) ? zo0 : zo;
        const float sx = l == 0 ? zs0.x : zs.x, sy = l == 0 ? zs0.y : zs.y;
        const float2 cc = f2conj(make_float2(__shfl_sync(F512_FULL, sx, partner, 16), __shfl_sync(F512_FULL, sy, partner, 16)));
        const float2 E = f2add(zk, cc), O = f2sub(zk, cc);
        float2 wc = w32_const(2 * i);
        if (i >= 5) { const float2 w0 = w32_const(kOwn0[i]); wc = l == 0 ? w0 : wc; }
        const float2 mit = f2mi(f2mul(f2mul(O, wc), w512l));
        const float2 xa = f2add(E, mit);            // 2*Xh[k]
        const float2 xb = f2conj(f2sub(E, mit));    // 2*Xh[256-k]
        // W1024^k: k = l + 32 i -> W1024^l * W32^i;  lane 0: k = 16 * own slot -> W64^(own slot)
        const float2 wu = f2mul(w1024l, w32_const(i)), w0 = w64_const(kOwn0[i]);
        const float2 wk = l == 0 ? w0 : wu;
        const int k = l == 0 ? 16 * kOwn0[i] : l + 32 * i;
        combine(xa, xb, wk, F512_FULL, k, Pf);
      }
      if (l == 0) {  // lane 0's ninth pair: slots (7, 9) -> k = 112
        const float2 zk = v[F512_OUT(7)], cc = f2conj(v[F512_OUT(9)]);
        const float2 E = f2add(zk, cc), O = f2sub(zk, cc);
        const float2 mit = f2mi(f2mul(O, w32_const(7)));
        combine(f2add(E, mit), f2conj(f2sub(E, mit)), w64_const(7), 0x00010001u, 112, Pf);
      }
    }
    __syncwarp();

    // ---- epilogue: the warp's (up to) 4 frames
    const int nvalid = (int)max((int64_t)0, min((int64_t)SLOTS, T - t0));
    const int nrows = (int)max((int64_t)0, min((int64_t)SLOTS, rows_here - t0));
    float *out = b.out + row0 * p.F;
    if (p.feature == B200FEAT_SPECTROGRAM || p.feature == B200FEAT_LOG_SPECTROGRAM) {
      for (int f = 0; f < nrows; ++f) {
        float *o = out + (int64_t)f * p.F;
        if (f >= nvalid) { for (int k = lane; k < p.F; k += 32) o[k] = post_affine(p, k, b.pad_value); continue; }
        for (int k = lane; k < p.K; k += 32) {
          float x = P[f * F1K_PBINS + k] * (p.use_mag ? 0.5f : 0.25f);
          if (p.feature == B200FEAT_LOG_SPECTROGRAM) x = log_spec_value(p, x);
          if (k == 0 && p.use_energy) {
#pragma unroll
            for (int g = 0; g < SLOTS; ++g) x = (f == g) ? le[g] : x;
          }
          o[k] = post_affine(p, k, x);
        }
      }
    } else {
      const int shift = (p.feature == B200FEAT_FBANK && p.use_energy) ? 1 : 0;
      const float lgk = p.log10_mel ? 0.30102999566398119521f : 0.69314718055994530942f;  // log10 (librosa_fbank.py:126) or ln
      const int Mpad = (p.M + 3) & ~3;
      float *mlog = reinterpret_cast<float *>(xall + (size_t)(2 * w) * F512_XBUF);  // both transpose tiles of the warp
      for (int j = 0; j < ft.mel_rounds; ++j) {
        const int m = lane + 32 * j;
        const float *Pj = P + s_rstart[j * 32 + lane];
        const float *wj = s_mw + s_rrow[j] * 32 + lane;
        const int len = s_rlen[j];
        float acc[SLOTS];
#pragma unroll
        for (int f = 0; f < SLOTS; ++f) acc[f] = 0.f;
        const float4 *wp = reinterpret_cast<const float4 *>(wj - lane) + lane;  // [row / 4][lane][4]
        const float4 *pp = reinterpret_cast<const float4 *>(Pj);
#pragma unroll 1
        for (int i = len; i > 0; i -= 4, ++pp, wp += 32) {
          const float4 wi = *wp;
#pragma unroll
          for (int f = 0; f < SLOTS; ++f) {
            const float4 pv = pp[f * (F1K_PBINS / 4)];
            acc[f] = fmaf(pv.w, wi.w, fmaf(pv.z, wi.z, fmaf(pv.y, wi.y, fmaf(pv.x, wi.x, acc[f]))));
          }
        }
        if (m < p.M) {
          float r[SLOTS];
#pragma unroll
          for (int f = 0; f < SLOTS; ++f) r[f] = fast_lg2_normal(nanmax(acc[f], p.mel_floor)) * lgk;
          if (p.feature != B200FEAT_MFCC) {
            float *orow = out + m + shift;
#pragma unroll
            for (int f = 0; f < SLOTS; ++f)
              if (f < nvalid) orow[(int64_t)f * p.F] = post_affine(p, m + shift, r[f]);
          } else {
#pragma unroll
            for (int f = 0; f < SLOTS; ++f) mlog[f * Mpad + m] = r[f];
          }
        }
      }
      if (p.feature == B200FEAT_FBANK) {
        if (shift && lane < nvalid) {
          float v0 = 0.f;
#pragma unroll
          for (int f = 0; f < SLOTS; ++f) v0 = (lane == f) ? le[f] : v0;
          out[(int64_t)lane * p.F] = post_affine(p, 0, v0);
        }
      } else if (p.feature == B200FEAT_MFCC) {
        __syncwarp();
        for (int idx = lane; idx < nvalid * p.C; idx += 32) {
          const int f = idx / p.C, c = idx - f * p.C;
          float acc = 0.f;
          for (int m = 0; m < p.M; ++m) acc = fmaf(mlog[f * Mpad + m], __ldg(p.dct + m * p.C + c), acc);
          if (p.use_lifter) acc *= __ldg(p.lifter + c);
          if (p.use_energy && c == 0) {
#pragma unroll
            for (int g = 0; g < SLOTS; ++g) acc = (f == g) ? le[g] : acc;
          }
          out[(int64_t)f * p.F + c] = post_affine(p, c, acc);
        }
      }
      for (int f = nvalid; f < nrows; ++f)
        for (int k = lane; k < p.F; k += 32) out[(int64_t)f * p.F + k] = post_affine(p, k, b.pad_value);
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------- host
struct Fast1024Host {
  Fast1024Tables t;
  size_t smem;
  int variant;
};

// launch shapes {warps per CTA, frames per warp}; both keep 2 CTAs per SM.  B200FEAT_FAST1024_VARIANT selects.
struct F1kVariant { int warps, slots; };
static const F1kVariant kF1kVariants[2] = {{8, 3}, {6, 4}};

static inline bool fast1024_supported(const DevPlan &p) {
  return p.N == 1024 && p.packed && p.L > 2 && p.L <= 1024 && p.C <= 128 && 4 * ((p.M + 3) & ~3) <= 4 * F512_XBUF;
}

template <int DT, int LCT, int WARPS, int SLOTS>
static int f1k_go(bool launch, size_t smem, const DevPlan &p, const Fast1024Tables &t, const DevBatch &b, dim3 grid, cudaStream_t stream) {
  auto kern = b200feat_fast1024_kernel<DT, LCT, WARPS, SLOTS>;
  if (!launch)
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess ? 0 : B200FEAT_ECUDA;
  kern<<<grid, dim3(WARPS * 32), smem, stream>>>(p, t, b);
  return 0;
}

template <int DT, int LCT>
static int f1k_shape(int variant, bool launch, size_t smem, const DevPlan &p, const Fast1024Tables &t, const DevBatch &b, dim3 grid,
                     cudaStream_t stream) {
  return variant == 1 ? f1k_go<DT, LCT, 6, 4>(launch, smem, p, t, b, grid, stream)
                      : f1k_go<DT, LCT, 8, 3>(launch, smem, p, t, b, grid, stream);
}

static int f1k_dispatch(int dt, int L, int variant, bool launch, size_t smem, const DevPlan &p, const Fast1024Tables &t,
                        const DevBatch &b, dim3 grid, cudaStream_t stream) {
  if (L == 600) return dt == B200FEAT_I16 ? f1k_shape<B200FEAT_I16, 600>(variant, launch, smem, p, t, b, grid, stream)
                                          : f1k_shape<B200FEAT_F32, 600>(variant, launch, smem, p, t, b, grid, stream);
  if (L == 551) return dt == B200FEAT_I16 ? f1k_shape<B200FEAT_I16, 551>(variant, launch, smem, p, t, b, grid, stream)
                                          : f1k_shape<B200FEAT_F32, 551>(variant, launch, smem, p, t, b, grid, stream);
  return dt == B200FEAT_I16 ? f1k_shape<B200FEAT_I16, 0>(variant, launch, smem, p, t, b, grid, stream)
                            : f1k_shape<B200FEAT_F32, 0>(variant, launch, smem, p, t, b, grid, stream);
}

static inline int f1k_ct_length(int L) { return (L == 600 || L == 551) ? L : 0; }

static inline int fast1024_prepare(DevPlan &p, const std::vector<float> &bank, std::vector<void *> &allocs,
                                   int *frames_per_tile, const std::vector<float> &window, Fast1024Host *out) {
  Fast1024Host hst;
  hst.variant = 0;
  if (const char *e = getenv("B200FEAT_FAST1024_VARIANT")) hst.variant = atoi(e) == 1 ? 1 : 0;
  const F1kVariant shape = kF1kVariants[hst.variant];
  std::vector<float2> win2(2 * 16 * 16), tw1(256), w512(16), w1024(16);
  for (int h = 0; h < 2; ++h)
    for (int n1 = 0; n1 < 16; ++n1)
      for (int l = 0; l < 16; ++l) {
        const int ja = 64 * n1 + 4 * l + h;
        win2[(h * 16 + n1) * 16 + l] = make_float2(ja < p.L ? window[ja] : 0.f, ja + 2 < p.L ? window[ja + 2] : 0.f);
      }
  for (int k1 = 0; k1 < 16; ++k1)
    for (int l = 0; l < 16; ++l) {
      const double a = -2.0 * M_PI * (double)((l * k1) % 256) / 256.0;
      tw1[k1 * 16 + l] = make_float2((float)cos(a), (float)sin(a));
    }
  for (int l = 0; l < 16; ++l) {
    const double a = -2.0 * M_PI * (double)l / 512.0, c = -2.0 * M_PI * (double)l / 1024.0;
    w512[l] = make_float2((float)cos(a), (float)sin(a));
    w1024[l] = make_float2((float)cos(c), (float)sin(c));
  }
  int rc;
  if ((rc = f512_upload(tw1, allocs, &hst.t.tw1))) return rc;
  if ((rc = f512_upload(w512, allocs, &hst.t.w512))) return rc;
  if ((rc = f512_upload(w1024, allocs, &hst.t.w1024))) return rc;
  const MelRounds mr = pack_mel_rounds(bank, p.K, p.M, p.use_mag ? 0.5f : 0.25f, 32, 4);  // 128-bit mel loads
  if (mr.max_reach > F1K_PBINS) return B200FEAT_EUNSUPPORTED;
  hst.t.mel_rounds = mr.rounds;
  hst.t.mel_wrows = mr.rows;
  {
    std::vector<unsigned char> blob;
    auto append = [&](const void *src, size_t bytes) -> int {
      const size_t off = blob.size();
      blob.resize(off + ((bytes + 15) & ~(size_t)15), 0);
      if (bytes) memcpy(blob.data() + off, src, bytes);
      return (int)off;
    };
    append(win2.data(), win2.size() * sizeof(float2));
    hst.t.off_rstart = append(mr.rstart.data(), mr.rstart.size() * sizeof(int));
    hst.t.off_rlen = append(mr.rlen.data(), mr.rlen.size() * sizeof(int));
    hst.t.off_rrow = append(mr.rrow.data(), mr.rrow.size() * sizeof(int));
    hst.t.off_mw = append(mr.wdense.data(), mr.wdense.size() * sizeof(float));
    const unsigned char *d = nullptr;
    if ((rc = f512_upload(blob, allocs, &d))) return rc;
    hst.t.cblob = d;
    hst.t.cblob_bytes = (int)blob.size();
  }
  hst.smem = fast1024_smem_bytes(hst.t, shape.warps, shape.slots);
  if (hst.smem > 113 * 1024) return B200FEAT_EUNSUPPORTED;  // keep 2 CTAs per SM
  DevBatch none{};
  for (int dt = 0; dt < 2; ++dt)
    if (f1k_dispatch(dt, f1k_ct_length(p.L), hst.variant, false, hst.smem, p, hst.t, none, dim3(1), nullptr)) return B200FEAT_ECUDA;
  *out = hst;
  *frames_per_tile = shape.warps * shape.slots;
  return 0;
}

static inline int fast1024_launch(const DevPlan &p, const Fast1024Host &hst, const DevBatch &b, int dt, int sm_count, cudaStream_t stream) {
  int64_t blocks = b.num_tiles;
  const int64_t cap = (int64_t)sm_count * 2;
  if (blocks > cap) blocks = cap;
  f1k_dispatch(dt, f1k_ct_length(p.L), hst.variant, true, hst.smem, p, hst.t, b, dim3((unsigned)blocks), stream);
  return (int)cudaGetLastError();
}
