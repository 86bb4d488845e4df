// Fast fused kernel for fft_length N = 512 (the 16 kHz / 25 ms headline geometry and every other
// plan with 256 < L <= 512): one HALF-WARP owns one frame; the 512-point real FFT is a packed
// 256-point complex FFT factored 16 x 16 and kept in registers:
//
//   lane l holds z[16*n1 + l], n1 = 0..15         (z[n] = y[2n] + i*y[2n+1], y = windowed frame)
//   radix-16 DFT over n1 in registers  -> Y[k1][l], times W256^(l*k1)
//   16x16 transpose through a padded shared-memory tile (the only FFT traffic that leaves registers)
//   lane k1 holds Y'[k1][0..15], radix-16 DFT over n2 -> Z[k1 + 16*k2]
//   real-FFT split against the mirrored lane (16 - k1) with half-warp shuffles -> |X[k]|^2
//   power spectrum of 4 consecutive frames staged as P[k][4] -> sparse mel bank with the weights
//   shared by the 4 frames (LDS.128) -> log -> coalesced 64-byte row segments.
//
// Replaces, for one frame, lhotse/features/kaldi/layers.py:151-186 (Wav2Win._forward_strided),
// :32-42 (_rfft/_pow_spectrogram), :565-578 (mel+log), :708-724 (DCT/lifter), with the framing of
// :727-772 folded into the load addresses.  HBM traffic: 4*S bytes in, 4*F bytes out per frame.
#pragma once
#include <vector>

#include "common.cuh"

#define F512_WARPS 8                        // warps per CTA
#define F512_HW (2 * F512_WARPS)            // half-warps per CTA
#define F512_SLOTS 4                        // frames per half-warp per round (mel register blocking)
#define F512_TILE (F512_HW * F512_SLOTS)    // frames per tile (64)
#define F512_XROW 18                        // float2 per transpose row (16 + 2 pad: 144 B, LDS.128 conflict-free)
#define F512_XBUF (16 * F512_XROW)          // float2 per half-warp transpose tile
#define F512_PBINS 260                      // bins per P tile (257 rounded up)
#define F512_PBUF (F512_PBINS * F512_SLOTS) // floats per half-warp P tile

struct Fast512Tables {  // device pointers, derived once per handle
  const float2 *win2;   // [16][16] window pairs (w[32*n1+2l], w[32*n1+2l+1]), zero beyond L
  const float2 *tw1;    // [16][16] W256^(l*k1) indexed [k1][l]
  const float2 *w512;   // [16]     W512^l
  int mel_nnz;          // floats in mel_w
  int mel_rounds;       // ceil(M / 16)
};

__device__ __forceinline__ float2 f2add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 f2sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 f2mul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

// forward 4-point DFT, in place, natural order
__device__ __forceinline__ void dft4(float2 &a0, float2 &a1, float2 &a2, float2 &a3) {
  const float2 s02 = f2add(a0, a2), d02 = f2sub(a0, a2);
  const float2 s13 = f2add(a1, a3), d13 = f2sub(a1, a3);
  a0 = f2add(s02, s13);
  a2 = f2sub(s02, s13);
  a1 = make_float2(d02.x + d13.y, d02.y - d13.x);  // d02 - i*d13
  a3 = make_float2(d02.x - d13.y, d02.y + d13.x);  // d02 + i*d13
}

#define F512_C1 0.92387953251128674f  // cos(pi/8)
#define F512_S1 0.38268343236508977f  // sin(pi/8)
#define F512_R2 0.70710678118654752f  // sqrt(1/2)

// forward 16-point DFT in registers (radix 4x4).  Input v[n]; output X[k] lands in v[4*(k&3) + (k>>2)].
__device__ __forceinline__ void dft16(float2 (&v)[16]) {
#pragma unroll
  for (int b = 0; b < 4; ++b) dft4(v[b], v[4 + b], v[8 + b], v[12 + b]);
  // v[4c + b] = y[b][c]; twiddle by W16^(b*c)
  {
    float2 t;
    // c = 1: b = 1,2,3 -> W^1, W^2, W^3
    t = v[5]; v[5] = make_float2(fmaf(t.x, F512_C1, t.y * F512_S1), fmaf(t.y, F512_C1, -t.x * F512_S1));
    t = v[6]; v[6] = make_float2((t.x + t.y) * F512_R2, (t.y - t.x) * F512_R2);
    t = v[7]; v[7] = make_float2(fmaf(t.x, F512_S1, t.y * F512_C1), fmaf(t.y, F512_S1, -t.x * F512_C1));
    // c = 2: b = 1,2,3 -> W^2, W^4, W^6
    t = v[9]; v[9] = make_float2((t.x + t.y) * F512_R2, (t.y - t.x) * F512_R2);
    t = v[10]; v[10] = make_float2(t.y, -t.x);
    t = v[11]; v[11] = make_float2((t.y - t.x) * F512_R2, -(t.x + t.y) * F512_R2);
    // c = 3: b = 1,2,3 -> W^3, W^6, W^9
    t = v[13]; v[13] = make_float2(fmaf(t.x, F512_S1, t.y * F512_C1), fmaf(t.y, F512_S1, -t.x * F512_C1));
    t = v[14]; v[14] = make_float2((t.y - t.x) * F512_R2, -(t.x + t.y) * F512_R2);
    t = v[15]; v[15] = make_float2(fmaf(-t.x, F512_C1, -t.y * F512_S1), fmaf(t.x, F512_S1, -t.y * F512_C1));
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) dft4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
}
// register holding output bin k of dft16
#define F512_OUT(k) (4 * ((k) & 3) + ((k) >> 2))

// W32^k2 = exp(-2*pi*i*k2/32), k2 = 0..15 (compile-time immediates after unrolling)
__device__ __forceinline__ float2 w32_const(int k2) {
  const float c[16] = {1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                       0.70710678118654752f, 0.55557023301960218f, 0.38268343236508977f, 0.19509032201612825f,
                       0.0f, -0.19509032201612825f, -0.38268343236508977f, -0.55557023301960218f,
                       -0.70710678118654752f, -0.83146961230254524f, -0.92387953251128674f, -0.98078528040323043f};
  const float s[16] = {0.0f, 0.19509032201612825f, 0.38268343236508977f, 0.55557023301960218f,
                       0.70710678118654752f, 0.83146961230254524f, 0.92387953251128674f, 0.98078528040323043f,
                       1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                       0.70710678118654752f, 0.55557023301960218f, 0.38268343236508977f, 0.19509032201612825f};
  return make_float2(c[k2], -s[k2]);
}

template <int DT>
__device__ __forceinline__ float2 ld_pair(const void *base, int64_t i) {  // i even, element index
  if (DT == B200FEAT_I16) {
    const short2 v = __ldg(reinterpret_cast<const short2 *>(reinterpret_cast<const int16_t *>(base) + i));
    return make_float2((float)v.x * (1.0f / 32768.0f), (float)v.y * (1.0f / 32768.0f));
  } else {
    return __ldg(reinterpret_cast<const float2 *>(reinterpret_cast<const float *>(base) + i));
  }
}

// The two half-warps of a warp run independent frames (and may diverge at cut edges), so every
// warp-level primitive below is scoped to the calling half with `hmask`.
__device__ __forceinline__ float hw_sum(float v, unsigned hmask) {  // sum over the 16 lanes of a half-warp
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(hmask, v, o);
  return v;
}

static inline size_t fast512_smem_bytes(const DevPlan &p, const Fast512Tables &t) {
  size_t b = (size_t)F512_HW * (F512_XBUF * 8 + F512_PBUF * 4);
  b += 16 * 16 * 8;                          // window pairs
  b += (size_t)p.M * 3 * 4 + (size_t)t.mel_nnz * 4;  // mel start/len/woff + weights
  return (b + 15) & ~(size_t)15;
}

template <int DT, int LCT>
__global__ void __launch_bounds__(F512_WARPS * 32, 2)
b200feat_fast512_kernel(const DevPlan p, const Fast512Tables ft, const DevBatch b) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int l = tid & 15;           // lane within the half-warp
  const int hw = tid >> 4;          // half-warp within the CTA
  const unsigned hmask = (tid & 16) ? 0xffff0000u : 0x0000ffffu;
  const int L = LCT ? LCT : p.L;
  constexpr int NP = LCT ? (LCT + 31) / 32 : 16;  // sample-pair registers actually needed

  // ---- shared memory carve-up
  float2 *xall = reinterpret_cast<float2 *>(smem_raw);
  float *pall = reinterpret_cast<float *>(xall + (size_t)F512_HW * F512_XBUF);
  float2 *s_win = reinterpret_cast<float2 *>(pall + (size_t)F512_HW * F512_PBUF);
  int *s_mstart = reinterpret_cast<int *>(s_win + 256);
  int *s_mlen = s_mstart + p.M;
  int *s_mwoff = s_mlen + p.M;
  float *s_mw = reinterpret_cast<float *>(s_mwoff + p.M);
  float2 *X = xall + (size_t)hw * F512_XBUF;
  float *P = pall + (size_t)hw * F512_PBUF;

  for (int i = tid; i < 256; i += blockDim.x) s_win[i] = __ldg(ft.win2 + i);
  for (int i = tid; i < p.M; i += blockDim.x) {
    s_mstart[i] = __ldg(p.mel_start + i);
    s_mlen[i] = __ldg(p.mel_len + i);
    s_mwoff[i] = __ldg(p.mel_woff + i);
  }
  for (int i = tid; i < ft.mel_nnz; i += blockDim.x) s_mw[i] = __ldg(p.mel_w + i);

  // per-lane constants kept in registers for the whole kernel
  float2 tw1[16];
#pragma unroll
  for (int k1 = 1; k1 < 16; ++k1) tw1[k1] = __ldg(ft.tw1 + k1 * 16 + l);
  const float2 w512l = __ldg(ft.w512 + l);
  const int partner = (16 - l) & 15;
  const float inv_L = 1.0f / (float)L;
  __syncthreads();

  for (int64_t tg = blockIdx.x; tg < b.num_tiles; tg += gridDim.x) {
    const int64_t tile = b.tile_base + tg;
    const int cut = find_segment(b.tile_off, b.B, tile);
    const int64_t t0 = (tile - __ldg(b.tile_off + cut)) * F512_TILE + (int64_t)hw * F512_SLOTS;
    const int64_t T = __ldg(b.row_off + cut + 1) - __ldg(b.row_off + cut);
    const int64_t rows_here = b.out_mode == B200FEAT_OUT_PADDED ? b.max_frames : T;
    if (t0 >= rows_here) continue;  // whole half-warp idle for this tile (warp-divergent at most by halves)
    const int64_t n = __ldg(b.nsamp + cut);
    const int64_t xoff = __ldg(b.samp_off + cut);
    const int64_t row0 = b.out_mode == B200FEAT_OUT_PADDED
                             ? (int64_t)(b.batch_first + cut) * b.max_frames + t0
                             : __ldg(b.row_off + cut) + t0;
    float le[F512_SLOTS];

#pragma unroll 1
    for (int f = 0; f < F512_SLOTS; ++f) {
      const int64_t t = t0 + f;
      le[f] = 0.f;
      if (t >= T) {  // beyond the cut: zero the P column so the mel phase stays finite
        for (int k = l; k < F512_PBINS; k += 16) P[k * F512_SLOTS + f] = 0.f;
        continue;
      }
      const int64_t base = t * p.S - (p.snip_edges ? 0 : p.pad_left);
      float2 v[16];
      float prev[NP];
      const bool interior = base >= 0 && base + L <= n && (((xoff + base) & 1) == 0);
      if (interior) {
#pragma unroll
        for (int n1 = 0; n1 < NP; ++n1) {
          const int j0 = 32 * n1 + 2 * l;
          if (j0 + 1 < L) {
            v[n1] = ld_pair<DT>(b.samples, xoff + base + j0);
            prev[n1] = ld_sample<DT>(b.samples, xoff + base + (j0 > 0 ? j0 - 1 : 0));
          } else if (j0 < L) {  // odd L: last tap alone
            v[n1] = make_float2(ld_sample<DT>(b.samples, xoff + base + j0), 0.f);
            prev[n1] = ld_sample<DT>(b.samples, xoff + base + (j0 > 0 ? j0 - 1 : 0));
          } else {
            v[n1] = make_float2(0.f, 0.f);
            prev[n1] = 0.f;
          }
        }
      } else {  // edge frame: per-tap reflection (layers.py:753-772); ~3 frames per cut
#pragma unroll
        for (int n1 = 0; n1 < NP; ++n1) {
          const int j0 = 32 * n1 + 2 * l;
          float a = 0.f, c = 0.f, pr = 0.f;
          if (j0 < L) {
            int64_t i = base + j0;
            if (!p.snip_edges) i = reflect_index(i, n);
            a = ld_sample<DT>(b.samples, xoff + i);
            int64_t ip = base + (j0 > 0 ? j0 - 1 : 0);
            if (!p.snip_edges) ip = reflect_index(ip, n);
            pr = ld_sample<DT>(b.samples, xoff + ip);
          }
          if (j0 + 1 < L) {
            int64_t i = base + j0 + 1;
            if (!p.snip_edges) i = reflect_index(i, n);
            c = ld_sample<DT>(b.samples, xoff + i);
          }
          v[n1] = make_float2(a, c);
          prev[n1] = pr;
        }
      }
      // ---- DC removal (layers.py:155-157)
      float s = 0.f;
#pragma unroll
      for (int n1 = 0; n1 < NP; ++n1) s += v[n1].x + v[n1].y;  // taps beyond L are exact zeros
      const float mu = p.remove_dc ? hw_sum(s, hmask) * inv_L : 0.f;
      // ---- energy, pre-emphasis, window (layers.py:159-170); zero padding is implicit
      float e = 0.f;
#pragma unroll
      for (int n1 = 0; n1 < 16; ++n1) {
        if (n1 < NP) {
          const int j0 = 32 * n1 + 2 * l;
          const float2 w = s_win[n1 * 16 + l];  // zero beyond L
          float da = v[n1].x - mu, dc = v[n1].y - mu, dp = prev[n1] - mu;
          if (j0 >= L) da = 0.f;
          if (j0 + 1 >= L) dc = 0.f;
          if (p.raw_energy) e = fmaf(da, da, fmaf(dc, dc, e));
          const float ya = fmaf(-p.preemph, dp, da) * w.x;
          const float yc = fmaf(-p.preemph, da, dc) * w.y;
          if (!p.raw_energy) e = fmaf(ya, ya, fmaf(yc, yc, e));
          v[n1] = make_float2(ya, yc);
        } else {
          v[n1] = make_float2(0.f, 0.f);
        }
      }
      if (p.use_energy) le[f] = log_energy_value(p, hw_sum(e, hmask));

      // ---- stage 1: radix-16 over n1, twiddle, transpose
      dft16(v);
#pragma unroll
      for (int k1 = 0; k1 < 16; ++k1) {
        float2 y = v[F512_OUT(k1)];
        if (k1 > 0) y = f2mul(y, tw1[k1]);
        X[k1 * F512_XROW + l] = y;
      }
      __syncwarp(hmask);
      {
        const float4 *row = reinterpret_cast<const float4 *>(X + l * F512_XROW);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 r = row[q];
          v[2 * q] = make_float2(r.x, r.y);
          v[2 * q + 1] = make_float2(r.z, r.w);
        }
      }
      __syncwarp(hmask);
      // ---- stage 2: radix-16 over n2 -> Z[l + 16*k2]
      dft16(v);
      // ---- real-FFT split + power (layers.py:38-42): X[k] = 0.5*(E - i*W512^k*O)
#pragma unroll
      for (int k2 = 0; k2 < 16; ++k2) {
        const float2 zk = v[F512_OUT(k2)];
        // what my mirror lane needs from me at this step: Z[.. 15-k2] (lane 0 mirrors itself, shifted by one)
        const float2 za = v[F512_OUT(15 - k2)];
        const float2 zb = v[F512_OUT((16 - k2) & 15)];
        const float sx = l == 0 ? zb.x : za.x;
        const float sy = l == 0 ? zb.y : za.y;
        const float cx = __shfl_sync(hmask, sx, partner, 16);
        const float cy = __shfl_sync(hmask, sy, partner, 16);
        const float er = zk.x + cx, ei = zk.y - cy;   // E = Zk + conj(Zc)
        const float orr = zk.x - cx, oi = zk.y + cy;  // O = Zk - conj(Zc)
        const float2 u = f2mul(make_float2(orr, oi), w32_const(k2));
        const float2 tt = f2mul(u, w512l);
        const float xr = er + tt.y, xi = ei - tt.x;   // 2*X[k]
        float pw = 0.25f * fmaf(xr, xr, xi * xi);
        if (p.use_mag) pw = sqrtf(pw);
        P[(l + 16 * k2) * F512_SLOTS + f] = pw;
        if (k2 == 0 && l == 0) {  // Nyquist bin: X[256] = Re Z0 - Im Z0
          const float xn = zk.x - zk.y;
          P[256 * F512_SLOTS + f] = p.use_mag ? fabsf(xn) : xn * xn;
        }
      }
    }
    __syncwarp(hmask);

    // ---- epilogue over the (up to) 4 frames of this half-warp
    const int nvalid = (int)min((int64_t)F512_SLOTS, T - t0) < 0 ? 0 : (int)min((int64_t)F512_SLOTS, T - t0);
    const int nrows = (int)min((int64_t)F512_SLOTS, rows_here - t0);
    float *out = b.out + row0 * p.F;
    if (p.feature == B200FEAT_SPECTROGRAM || p.feature == B200FEAT_LOG_SPECTROGRAM) {
      for (int f = 0; f < nrows; ++f) {
        float *o = out + (int64_t)f * p.F;
        if (f >= nvalid) { for (int k = l; k < p.F; k += 16) o[k] = b.pad_value; continue; }
        for (int k = l; k < p.K; k += 16) {
          float x = P[k * F512_SLOTS + f];
          if (p.feature == B200FEAT_LOG_SPECTROGRAM) x = logf(x + p.log_spec_eps);
          if (k == 0 && p.use_energy) x = le[f];
          o[k] = x;
        }
      }
    } else {
      const int shift = (p.feature == B200FEAT_FBANK && p.use_energy) ? 1 : 0;
      const int Mpad = (p.M + 3) & ~3;
      float *mlog = reinterpret_cast<float *>(X);  // the transpose tile is idle during the epilogue
      for (int j = 0; j < ft.mel_rounds; ++j) {
        const int m = l + 16 * j;
        const bool mv = m < p.M;
        const int st = mv ? s_mstart[m] : 0, len = mv ? s_mlen[m] : 0;
        const float *w = s_mw + (mv ? s_mwoff[m] : 0);
        const float4 *P4 = reinterpret_cast<const float4 *>(P) + st;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = 0; i < len; ++i) {
          const float wi = w[i];
          const float4 pv = P4[i];
          acc.x = fmaf(pv.x, wi, acc.x); acc.y = fmaf(pv.y, wi, acc.y);
          acc.z = fmaf(pv.z, wi, acc.z); acc.w = fmaf(pv.w, wi, acc.w);
        }
        if (mv) {
          const float r[4] = {logf(fmaxf(acc.x, p.mel_floor)), logf(fmaxf(acc.y, p.mel_floor)),
                              logf(fmaxf(acc.z, p.mel_floor)), logf(fmaxf(acc.w, p.mel_floor))};
          if (p.feature == B200FEAT_FBANK) {
#pragma unroll
            for (int f = 0; f < F512_SLOTS; ++f)
              if (f < nvalid) out[(int64_t)f * p.F + m + shift] = r[f];
          } else {
#pragma unroll
            for (int f = 0; f < F512_SLOTS; ++f) mlog[f * Mpad + m] = r[f];
          }
        }
      }
      if (p.feature == B200FEAT_FBANK) {
        if (shift && l < nvalid) out[(int64_t)l * p.F] = le[0] * (l == 0) + le[1] * (l == 1) + le[2] * (l == 2) + le[3] * (l == 3);
      } else {
        __syncwarp(hmask);
        for (int idx = l; idx < nvalid * p.C; idx += 16) {
          const int f = idx / p.C, c = idx - f * p.C;
          float acc = 0.f;
          for (int m = 0; m < p.M; ++m) acc = fmaf(mlog[f * Mpad + m], __ldg(p.dct + m * p.C + c), acc);
          if (p.use_lifter) acc *= __ldg(p.lifter + c);
          if (p.use_energy && c == 0) acc = le[0] * (f == 0) + le[1] * (f == 1) + le[2] * (f == 2) + le[3] * (f == 3);
          out[(int64_t)f * p.F + c] = acc;
        }
      }
      // padded tail rows
      for (int f = nvalid; f < nrows; ++f)
        for (int k = l; k < p.F; k += 16) out[(int64_t)f * p.F + k] = b.pad_value;
    }
    __syncwarp(hmask);
  }
}

// ---------------------------------------------------------------------------------------------- host
struct Fast512Host {
  Fast512Tables t;
  size_t smem;
};

static inline bool fast512_supported(const DevPlan &p) {
  return p.N == 512 && p.packed && p.L >= 2 && p.L <= 512 && p.C <= 128 &&
         F512_SLOTS * ((p.M + 3) & ~3) <= 2 * F512_XBUF;  // log-mel staging reuses the transpose tile
}

template <typename T>
static int f512_upload(const std::vector<T> &h, std::vector<void *> &allocs, const T **out) {
  void *d = nullptr;
  if (cudaMalloc(&d, h.size() * sizeof(T)) != cudaSuccess) return B200FEAT_ECUDA;
  allocs.push_back(d);
  if (cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) return B200FEAT_ECUDA;
  *out = reinterpret_cast<const T *>(d);
  return 0;
}

template <int DT, int LCT>
static int f512_set_attr(size_t smem) {
  return cudaFuncSetAttribute(b200feat_fast512_kernel<DT, LCT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess
             ? 0 : B200FEAT_ECUDA;
}

static inline int fast512_prepare(DevPlan &p, const std::vector<float> &bank, std::vector<void *> &allocs,
                                  int *frames_per_tile, const std::vector<float> &window, Fast512Host *out) {
  Fast512Host hst;
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
  if ((rc = f512_upload(win2, allocs, &hst.t.win2))) return rc;
  if ((rc = f512_upload(tw1, allocs, &hst.t.tw1))) return rc;
  if ((rc = f512_upload(w512, allocs, &hst.t.w512))) return rc;
  int nnz = 0;  // must match the packing of mel_w in b200feat_create (0 when there is no bank)
  for (int m = 0; m < p.M; ++m) {
    int first = -1, last = -1;
    for (int k = 0; k < p.K; ++k)
      if (bank[(size_t)k * p.M + m] != 0.f) { if (first < 0) first = k; last = k; }
    if (first >= 0) nnz += last - first + 1;
  }
  hst.t.mel_nnz = nnz;
  hst.t.mel_rounds = (p.M + 15) / 16;
  hst.smem = fast512_smem_bytes(p, hst.t);
  if (hst.smem > 113 * 1024) return B200FEAT_EUNSUPPORTED;  // keep 2 CTAs per SM
  if (f512_set_attr<B200FEAT_F32, 400>(hst.smem) || f512_set_attr<B200FEAT_I16, 400>(hst.smem) ||
      f512_set_attr<B200FEAT_F32, 0>(hst.smem) || f512_set_attr<B200FEAT_I16, 0>(hst.smem))
    return B200FEAT_ECUDA;
  *out = hst;
  *frames_per_tile = F512_TILE;
  return 0;
}

static inline int fast512_launch(const DevPlan &p, const Fast512Host &hst, const DevBatch &b, int dt, int sm_count,
                                 cudaStream_t stream) {
  int64_t blocks = b.num_tiles;
  const int64_t cap = (int64_t)sm_count * 2;
  if (blocks > cap) blocks = cap;
  const dim3 grid((unsigned)blocks), block(F512_WARPS * 32);
  if (p.L == 400) {
    if (dt == B200FEAT_I16) b200feat_fast512_kernel<B200FEAT_I16, 400><<<grid, block, hst.smem, stream>>>(p, hst.t, b);
    else b200feat_fast512_kernel<B200FEAT_F32, 400><<<grid, block, hst.smem, stream>>>(p, hst.t, b);
  } else {
    if (dt == B200FEAT_I16) b200feat_fast512_kernel<B200FEAT_I16, 0><<<grid, block, hst.smem, stream>>>(p, hst.t, b);
    else b200feat_fast512_kernel<B200FEAT_F32, 0><<<grid, block, hst.smem, stream>>>(p, hst.t, b);
  }
  return (int)cudaGetLastError();
}
