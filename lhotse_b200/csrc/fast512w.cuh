// Second register-FFT kernel for fft_length N = 512 (`B200FEAT_FAST_VARIANT=3`, A/B partner of fast512.cuh): one WARP per frame
// instead of a half-warp, in the style of fast2048.cuh.  The 512-point real FFT is a packed 256-point complex FFT factored
// 8 x 8 x 4 with only 8 complex points per lane, so the kernel needs about half the registers of fast512.cuh (16 points + 15
// twiddles per lane, 128 registers, 16 warps per SM) and can keep 24-32 warps per SM resident:
//
//   z[n] = y[2n] + i*y[2n+1], n = 32*n1 + 4*n2 + n3            k = k1 + 8*k2 + 64*k3
//   stage 1  lane owns column c = lane = 4*n2 + n3: radix-8 DFT over n1, times W64^(n2*k1) -> tile A[k1][c]   (STS.64)
//   stage 2  lane (k1 = lane >> 2, n3 = lane & 3) reads A[k1][n2][n3], radix-8 DFT over n2, times W256^(n3*(k1 + 8*k2))
//            -> tile B[n3 >> 1][k = k1 + 8*k2][n3 & 1]                                                       (STS.64)
//   stage 3  lane owns k in {lane, 64 - lane}: radix-4 over n3 gives Z[k + 64*k3]; the mirror Z[256 - kappa] of every
//            kappa = k + 64*k3 is (64 - k) + 64*(3 - k3), in the SAME lane: the real-FFT split needs no shuffle.  The four
//            pair twiddles W512^kappa are per-lane register constants; lane 0 (residues 0 and 32, both self-mirrored) walks
//            the pairs (0,0) (64,192) (128,128) (32,224) in the same four slots and (96,160) in a fifth.
//   power spectra of SLOTS consecutive frames as P[slot][bin]; mel bank as balanced 12-tap work items (MelItems, common.cuh).
//
// The stage functions are __host__ __device__ (scripts/micro/f2k_host_check.cu, tests/test_fast2048_host.py).
// Replaces the same reference code as fast512.cuh (lhotse/features/kaldi/layers.py:151-186, :32-42, :565-578, :708-724,
// framing :727-772).
#pragma once
#include "fast2048.cuh"

#define F5W_PBINS 272                      // floats per P row: 257 bins + zero pad (a 12-tap mel piece may start at bin 256)
#define F5W_PTAIL 64
#define F5W_XROW 36                        // float2 per k1-row of tile A (32 + 4 pad: the strided LDS.64 of stage 2 is conflict-free)
#define F5W_XBUF (8 * F5W_XROW)            // float2 per warp (2304 B); tile B aliases it
#define F5W_PLANE2 136                     // float2 per n3-pair plane of tile B (128 + 8 pad)
#define F5W_TW2ROW 68                      // float2 per n3-row of the stage-2 twiddle table (64 + 4 pad)
#define F5W_PIECE F2K_PIECE

// forward 8-point DFT, natural order in and out (the fast256.cuh butterfly, host + device)
F512_HD void f5w_dft8(float2 (&x)[8]) {
  float2 e0 = x[0], e1 = x[2], e2 = x[4], e3 = x[6], o0 = x[1], o1 = x[3], o2 = x[5], o3 = x[7];
  dft4(e0, e1, e2, e3);
  dft4(o0, o1, o2, o3);
  o1 = f2mul_w8_1(o1);
  o2 = f2mi(o2);
  o3 = f2mul_w8_3(o3);
  x[0] = f2add(e0, o0); x[4] = f2sub(e0, o0);
  x[1] = f2add(e1, o1); x[5] = f2sub(e1, o1);
  x[2] = f2add(e2, o2); x[6] = f2sub(e2, o2);
  x[3] = f2add(e3, o3); x[7] = f2sub(e3, o3);
}

// ---- stage 1: v[n1] = z[32*n1 + lane]; tw1[k1*8 + n2] = W64^(n2*k1)
F512_HD void f5w_stage1(int lane, float2 (&v)[8], const float2 *tw1, float2 *xa) {
  f5w_dft8(v);
  const int n2 = lane >> 2;
#pragma unroll
  for (int k1 = 0; k1 < 8; ++k1) {
    float2 a = v[k1];
    if (k1 > 0) a = f2mul(a, tw1[k1 * 8 + n2]);
    xa[k1 * F5W_XROW + lane] = a;
  }
}

F512_HD void f5w_stage2_load(int lane, const float2 *xa, float2 (&u)[8]) {
  const float2 *row = xa + (lane >> 2) * F5W_XROW + (lane & 3);  // A[k1][4*n2 + n3]
#pragma unroll
  for (int n2 = 0; n2 < 8; ++n2) u[n2] = row[4 * n2];
}

// tw[k2] = W256^(n3*(k1 + 8*k2)) of this lane (register constants); tile B float2 index (n3 >> 1)*PLANE2 + 2*k + (n3 & 1)
F512_HD void f5w_stage2_store(int lane, float2 (&u)[8], const float2 (&tw)[8], float2 *xb) {
  f5w_dft8(u);
  const int k1 = lane >> 2, n3 = lane & 3;
  float2 *dst = xb + (n3 >> 1) * F5W_PLANE2 + 2 * k1 + (n3 & 1);
#pragma unroll
  for (int k2 = 0; k2 < 8; ++k2) dst[16 * k2] = f2mul(u[k2], tw[k2]);
}

F512_HD void f5w_pair(float2 zk, float2 zm, float2 w, bool use_mag, float &pa, float &pb) {
  const float2 cc = f2conj(zm);
  const float2 E = f2add(zk, cc), O = f2sub(zk, cc);
  const float2 mit = f2mi(f2mul(O, w));  // -i * W512^kappa * O
  const float2 a = f2add(E, mit), q = f2sub(E, mit);
  pa = fmaf(a.x, a.x, a.y * a.y);
  pb = fmaf(q.x, q.x, q.y * q.y);
  if (use_mag) { pa = sqrtf(pa); pb = sqrtf(pb); }
}

// kappa of pair slot s for this lane (host + device; the twiddle table is built from it)
F512_HD int f5w_kappa(int lane, int s) { return lane ? lane + 64 * s : (s == 3 ? 32 : 64 * s); }

// ---- stage 3 + split + power: tk[s] = W512^kappa(lane, s); Pf[0..256]
F512_HD void f5w_stage3(int lane, const float4 *xb4, const float2 (&tk)[4], float *Pf, bool use_mag) {
  const int ka = lane, kb = lane ? 64 - lane : 32;
  float2 a[4], q[4];
  {
    const float4 r0 = xb4[ka], r1 = xb4[F5W_PLANE2 / 2 + ka];
    a[0] = make_float2(r0.x, r0.y); a[1] = make_float2(r0.z, r0.w);
    a[2] = make_float2(r1.x, r1.y); a[3] = make_float2(r1.z, r1.w);
    dft4(a[0], a[1], a[2], a[3]);  // a[k3] = Z[ka + 64*k3]
    const float4 s0 = xb4[kb], s1 = xb4[F5W_PLANE2 / 2 + kb];
    q[0] = make_float2(s0.x, s0.y); q[1] = make_float2(s0.z, s0.w);
    q[2] = make_float2(s1.x, s1.y); q[3] = make_float2(s1.z, s1.w);
    dft4(q[0], q[1], q[2], q[3]);  // q[k3] = Z[kb + 64*k3]
  }
  const bool z = lane == 0;
  const float2 zk[4] = {a[0], a[1], a[2], z ? q[0] : a[3]};
  const float2 zm[4] = {z ? a[0] : q[3], z ? a[3] : q[2], z ? a[2] : q[1], z ? q[3] : q[0]};
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    float pa, pb;
    f5w_pair(zk[s], zm[s], tk[s], use_mag, pa, pb);
    const int kappa = f5w_kappa(lane, s);
    Pf[kappa] = pa;
    Pf[256 - kappa] = pb;
  }
  if (z) {  // lane 0's fifth pair: kappa = 96 <-> 160, W512^96 = W16^3
    float pa, pb;
    f5w_pair(q[1], q[2], make_float2(F512_S1, -F512_C1), use_mag, pa, pb);
    Pf[96] = pa;
    Pf[160] = pb;
  }
}

struct Fast512wTables {
  // one 16-byte-aligned blob (TMA bulk copy):
  //   [win2: 8*32 float2 (w[64 n1 + 2 lane], w[.. + 1]), zero beyond L] [tw1: 8*8 float2 W64^(n2*k1) at [k1][n2]]
  //   [rstart: rounds*32 int | fdesc: M int2 + ceil(M/32) int2 | wdense: rounds*3*32 float4]
  const void *cblob;
  int cblob_bytes;
  int off_tw1, off_rstart, off_fdesc, off_mw;
  const float2 *tw2;   // [32 lanes][8]  W256^(n3*(k1 + 8*k2)), lane = 4*k1 + n3 (global: loaded into registers once)
  const float2 *tk;    // [32 lanes][4]  W512^kappa(lane, s)
  int mel_rounds;
  int xfloats;         // floats of per-warp scratch: the exchange tile, reused by the epilogue for SLOTS x (work-item sums + log-mel row)
};

static inline size_t fast512w_smem_bytes(const Fast512wTables &t, int warps, int slots) {
  size_t b = (size_t)warps * t.xfloats * 4 + (size_t)warps * F5W_PBINS * slots * 4 + F5W_PTAIL * 4;
  b += (size_t)t.cblob_bytes + 16;
  return (b + 15) & ~(size_t)15;
}

template <int DT, int LCT, int WARPS, int SLOTS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
b200feat_fast512w_kernel(const DevPlan p, const Fast512wTables ft, const DevBatch b) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int w = tid >> 5;            // warp = frame owner
  const int L = LCT ? LCT : p.L;
  constexpr int NP = LCT ? (LCT + 63) / 64 : 8;  // rows of 64 samples that carry data

  float *xall = reinterpret_cast<float *>(smem_raw);
  float *pall = xall + (size_t)WARPS * ft.xfloats;
  unsigned char *s_const = reinterpret_cast<unsigned char *>(pall + (size_t)WARPS * (F5W_PBINS * SLOTS) + F5W_PTAIL);
  const float2 *s_win = reinterpret_cast<const float2 *>(s_const);
  const float2 *s_tw1 = reinterpret_cast<const float2 *>(s_const + ft.off_tw1);
  const int *s_rstart = reinterpret_cast<const int *>(s_const + ft.off_rstart);
  const int2 *s_fdesc = reinterpret_cast<const int2 *>(s_const + ft.off_fdesc);
  const float4 *s_mw4 = reinterpret_cast<const float4 *>(s_const + ft.off_mw);
  unsigned long long *s_bar = reinterpret_cast<unsigned long long *>(s_const + ft.cblob_bytes);
  float2 *X = reinterpret_cast<float2 *>(xall + (size_t)w * ft.xfloats);  // per warp exchange tile (xfloats is a multiple of 4)
  float *P = pall + (size_t)w * (F5W_PBINS * SLOTS);             // per warp: [slot][F5W_PBINS]

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
  for (int i = tid; i < WARPS * (F5W_PBINS * SLOTS) + F5W_PTAIL; i += blockDim.x) pall[i] = 0.f;
  // per-lane twiddles kept in registers for the whole kernel
  float2 tw2[8], tk[4];
#pragma unroll
  for (int k2 = 0; k2 < 8; ++k2) tw2[k2] = __ldg(ft.tw2 + lane * 8 + k2);
#pragma unroll
  for (int s = 0; s < 4; ++s) tk[s] = __ldg(ft.tk + lane * 4 + s);
  const float inv_L = 1.0f / (float)L;
  const int up_lane = (lane + 31) & 31;
  {
    unsigned done = 0;
    while (!done)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(done) : "r"(bar), "r"(0u) : "memory");
  }
  __syncthreads();

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
      if (t >= T) break;  // warp-uniform; the frames of a warp are consecutive
      const int64_t base = t * p.S - (p.snip_edges ? 0 : p.pad_left);
      float2 v[8];
      const bool interior = base >= 0 && base + L <= n;
      if (F512_PREFETCH && f + 1 < SLOTS && t + 1 < T) {  // one L1 prefetch per 32-byte sector of the next frame's new samples
        constexpr int PER = DT == B200FEAT_I16 ? 16 : 8;
        const int64_t q = base + L + PER * lane;
        if (PER * lane < p.S + PER && q >= 0 && q < n)
          asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const char *>(b.samples) + (xoff + q) * (DT == B200FEAT_I16 ? 2 : 4)));
      }
      if (interior && (((xoff + base) & 1) == 0)) {  // aligned 8-byte (4-byte for PCM16) pairs, coalesced
        if (DT == B200FEAT_I16) {
          const int16_t *xp = reinterpret_cast<const int16_t *>(b.samples) + (xoff + base + 2 * lane);
#pragma unroll
          for (int n1 = 0; n1 < NP; ++n1) {
            const int j = 64 * n1 + 2 * lane;
            float2 r = make_float2(0.f, 0.f);
            if (j + 1 < L) {
              const short2 q = __ldg(reinterpret_cast<const short2 *>(xp + 64 * n1));
              r = make_float2((float)q.x * (1.0f / 32768.0f), (float)q.y * (1.0f / 32768.0f));
            } else if (j < L) {
              r.x = (float)__ldg(xp + 64 * n1) * (1.0f / 32768.0f);
            }
            v[n1] = r;
          }
        } else {
          const float *xp = reinterpret_cast<const float *>(b.samples) + (xoff + base + 2 * lane);
#pragma unroll
          for (int n1 = 0; n1 < NP; ++n1) {
            const int j = 64 * n1 + 2 * lane;
            float2 r = make_float2(0.f, 0.f);
            if (j + 1 < L) r = __ldg(reinterpret_cast<const float2 *>(xp + 64 * n1));
            else if (j < L) r.x = __ldg(xp + 64 * n1);  // odd L: last tap alone
            v[n1] = r;
          }
        }
      } else if (interior) {  // odd element offset: two 4-byte loads per pair
        const int64_t x0 = xoff + base + 2 * lane;
#pragma unroll
        for (int n1 = 0; n1 < NP; ++n1) {
          const int j = 64 * n1 + 2 * lane;
          float2 r = make_float2(0.f, 0.f);
          if (j < L) r.x = ld_sample<DT>(b.samples, x0 + 64 * n1);
          if (j + 1 < L) r.y = ld_sample<DT>(b.samples, x0 + 64 * n1 + 1);
          v[n1] = r;
        }
      } else {  // a cut edge: per-tap reflection (layers.py:753-772)
#pragma unroll
        for (int n1 = 0; n1 < NP; ++n1) {
          const int j = 64 * n1 + 2 * lane;
          float2 r = make_float2(0.f, 0.f);
          if (j < L) {
            int64_t i = base + j;
            if (!p.snip_edges) i = reflect_index(i, n, p.pad_mode);
            r.x = ld_sample<DT>(b.samples, xoff + i);
          }
          if (j + 1 < L) {
            int64_t i = base + j + 1;
            if (!p.snip_edges) i = reflect_index(i, n, p.pad_mode);
            r.y = ld_sample<DT>(b.samples, xoff + i);
          }
          v[n1] = r;
        }
      }
      // ---- DC removal (layers.py:155-157)
      float s = 0.f, s2 = 0.f;
#pragma unroll
      for (int n1 = 0; n1 < NP; ++n1) {  // taps beyond L are exact zeros
        if (n1 & 1) s2 += v[n1].x + v[n1].y; else s += v[n1].x + v[n1].y;
      }
      const float mu = p.remove_dc ? warp_sum(s + s2) * inv_L : 0.f;
      // ---- energy, pre-emphasis, window (layers.py:159-170).  The tap before y[64 n1 + 2 lane] is the neighbour lane's odd tap;
      // lane 0 takes lane 31's of the previous row
      float e = 0.f;
      float carry = v[0].x;  // lane 0, row 0: replicate-left (layers.py:166)
#pragma unroll
      for (int n1 = 0; n1 < 8; ++n1) {
        if (n1 < NP) {
          const float up = __shfl_sync(F512_FULL, v[n1].y, up_lane);
          const float pr = lane == 0 ? carry : up;
          carry = up;
          const int j = 64 * n1 + 2 * lane;
          const float2 wv = s_win[n1 * 32 + lane];  // zero beyond L
          float2 d = f2add(v[n1], make_float2(-mu, -mu));
          const float dp = pr - mu;
          if (j >= L) d.x = 0.f;
          if (j + 1 >= L) d.y = 0.f;
          if (p.raw_energy) e = fmaf(d.x, d.x, fmaf(d.y, d.y, e));
          const float2 y = __fmul2_rn(__ffma2_rn(make_float2(dp, d.x), make_float2(-p.preemph, -p.preemph), d), wv);
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
      // ---- the 256-point complex FFT, 8 x 8 x 4, and the split
      f5w_stage1(lane, v, s_tw1, X);
      __syncwarp();
      f5w_stage2_load(lane, X, v);
      __syncwarp();
      f5w_stage2_store(lane, v, tw2, X);
      __syncwarp();
      f5w_stage3(lane, reinterpret_cast<const float4 *>(X), tk, P + f * F5W_PBINS, p.use_mag != 0);
      __syncwarp();
    }

    // ---- epilogue: the warp's (up to) SLOTS frames
    const int nvalid = (int)max((int64_t)0, min((int64_t)SLOTS, T - t0));
    const int nrows = (int)max((int64_t)0, min((int64_t)SLOTS, rows_here - t0));
    float *out = b.out + row0 * p.F;
    if (p.feature == B200FEAT_SPECTROGRAM || p.feature == B200FEAT_LOG_SPECTROGRAM) {
      for (int f = 0; f < nrows; ++f) {
        float *o = out + (int64_t)f * p.F;
        if (f >= nvalid) { for (int k = lane; k < p.F; k += 32) o[k] = post_affine(p, k, b.pad_value); continue; }
        for (int k = lane; k < p.K; k += 32) {
          float x = P[f * F5W_PBINS + k] * (p.use_mag ? 0.5f : 0.25f);
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
      // the exchange tile is idle during the epilogue: [SLOTS][NQ] partial sums of the mel work items, then the log-mel rows (MFCC)
      const int NQ = ft.mel_rounds * 32;
      float *part = reinterpret_cast<float *>(X);
      float *mlog = part + SLOTS * NQ;
      for (int j = 0; j < ft.mel_rounds; ++j) {  // pass 1: one work item (filter, piece of 12 taps) per lane
        const float4 *pa = reinterpret_cast<const float4 *>(P + s_rstart[j * 32 + lane]);
        const float4 *wa = s_mw4 + (j * (F5W_PIECE / 4)) * 32 + lane;  // [round][trip][lane][4]
        float acc[SLOTS];
#pragma unroll
        for (int f = 0; f < SLOTS; ++f) acc[f] = 0.f;
#pragma unroll
        for (int t = 0; t < F5W_PIECE / 4; ++t) {
          const float4 w4 = wa[t * 32];
#pragma unroll
          for (int f = 0; f < SLOTS; ++f) {
            const float4 p4 = pa[f * (F5W_PBINS / 4) + t];
            acc[f] = fmaf(p4.w, w4.w, fmaf(p4.z, w4.z, fmaf(p4.y, w4.y, fmaf(p4.x, w4.x, acc[f]))));
          }
        }
#pragma unroll
        for (int f = 0; f < SLOTS; ++f) part[f * NQ + j * 32 + lane] = acc[f];
      }
      __syncwarp();
      for (int m = lane; m < p.M; m += 32) {  // pass 2: one filter per lane adds its pieces in item order
        const int2 fd = s_fdesc[m];             // {first item, items}
        const int qn = s_fdesc[p.M + (m >> 5)].x;  // uniform bound: the item count of the widest filter among these 32
        float r[SLOTS];
#pragma unroll
        for (int f = 0; f < SLOTS; ++f) r[f] = 0.f;
        for (int q = 0; q < qn; ++q) {
#pragma unroll
          for (int f = 0; f < SLOTS; ++f) r[f] += q < fd.y ? part[f * NQ + fd.x + q] : 0.f;
        }
#pragma unroll
        for (int f = 0; f < SLOTS; ++f) r[f] = fast_lg2_normal(nanmax(r[f], p.mel_floor)) * lgk;
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
struct Fast512wHost {
  Fast512wTables t;
  size_t smem;
  int shape;
};

// launch shapes {warps per CTA, frames per warp, CTAs per SM}: B200FEAT_FAST512W_SHAPE selects (A/B runs)
struct F5wShape { int warps, slots, minb; };
#define F5W_NUM_SHAPES 4
static const F5wShape kF5wShapes[F5W_NUM_SHAPES] = {{8, 4, 3}, {8, 2, 4}, {8, 4, 2}, {10, 4, 2}};

// the constant tables of the FFT stages (also used by scripts/micro/f2k_host_check.cu)
static inline void f5w_fft_tables(std::vector<float2> &tw1, std::vector<float2> &tw2, std::vector<float2> &tk) {
  tw1.resize(64); tw2.resize(32 * 8); tk.resize(32 * 4);
  for (int k1 = 0; k1 < 8; ++k1)
    for (int n2 = 0; n2 < 8; ++n2) {
      const double a = -2.0 * M_PI * (double)((n2 * k1) % 64) / 64.0;
      tw1[k1 * 8 + n2] = make_float2((float)cos(a), (float)sin(a));
    }
  for (int lane = 0; lane < 32; ++lane) {
    const int k1 = lane >> 2, n3 = lane & 3;
    for (int k2 = 0; k2 < 8; ++k2) {
      const double a = -2.0 * M_PI * (double)((n3 * (k1 + 8 * k2)) % 256) / 256.0;
      tw2[lane * 8 + k2] = make_float2((float)cos(a), (float)sin(a));
    }
    for (int s = 0; s < 4; ++s) {
      const double a = -2.0 * M_PI * (double)f5w_kappa(lane, s) / 512.0;
      tk[lane * 4 + s] = make_float2((float)cos(a), (float)sin(a));
    }
  }
}

static inline bool fast512w_supported(const DevPlan &p) {
  return p.N == 512 && p.packed && p.L > 2 && p.L <= 512 && p.C <= 128 && p.M <= 128;
}

template <int DT, int LCT, int WARPS, int SLOTS, int MINB>
static int f5w_go(bool launch, size_t smem, const DevPlan &p, const Fast512wTables &t, const DevBatch &b, dim3 grid, cudaStream_t stream) {
  auto kern = b200feat_fast512w_kernel<DT, LCT, WARPS, SLOTS, MINB>;
  if (!launch)
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess ? 0 : B200FEAT_ECUDA;
  kern<<<grid, dim3(WARPS * 32), smem, stream>>>(p, t, b);
  return 0;
}

template <int DT, int LCT>
static int f5w_shape(int shape, bool launch, size_t smem, const DevPlan &p, const Fast512wTables &t, const DevBatch &b, dim3 grid,
                     cudaStream_t stream) {
  if (shape == 1) return f5w_go<DT, LCT, 8, 2, 4>(launch, smem, p, t, b, grid, stream);
  if (shape == 2) return f5w_go<DT, LCT, 8, 4, 2>(launch, smem, p, t, b, grid, stream);
  if (shape == 3) return f5w_go<DT, LCT, 10, 4, 2>(launch, smem, p, t, b, grid, stream);
  return f5w_go<DT, LCT, 8, 4, 3>(launch, smem, p, t, b, grid, stream);
}

static int f5w_dispatch(int dt, int L, int shape, bool launch, size_t smem, const DevPlan &p, const Fast512wTables &t,
                        const DevBatch &b, dim3 grid, cudaStream_t stream) {
  if (L == 400) return dt == B200FEAT_I16 ? f5w_shape<B200FEAT_I16, 400>(shape, launch, smem, p, t, b, grid, stream)
                                          : f5w_shape<B200FEAT_F32, 400>(shape, launch, smem, p, t, b, grid, stream);
  return dt == B200FEAT_I16 ? f5w_shape<B200FEAT_I16, 0>(shape, launch, smem, p, t, b, grid, stream)
                            : f5w_shape<B200FEAT_F32, 0>(shape, launch, smem, p, t, b, grid, stream);
}

static inline int fast512w_prepare(DevPlan &p, const std::vector<float> &bank, std::vector<void *> &allocs,
                                   int *frames_per_tile, const std::vector<float> &window, Fast512wHost *out) {
  Fast512wHost hst;
  hst.shape = 0;
  if (const char *e = getenv("B200FEAT_FAST512W_SHAPE")) {
    const int v = atoi(e);
    if (v >= 0 && v < F5W_NUM_SHAPES) hst.shape = v;
  }
  const F5wShape shape = kF5wShapes[hst.shape];
  std::vector<float2> win2(8 * 32), tw1, tw2, tk;
  for (int n1 = 0; n1 < 8; ++n1)
    for (int l = 0; l < 32; ++l) {
      const int j = 64 * n1 + 2 * l;
      win2[n1 * 32 + l] = make_float2(j < p.L ? window[j] : 0.f, j + 1 < p.L ? window[j + 1] : 0.f);
    }
  f5w_fft_tables(tw1, tw2, tk);
  int rc;
  if ((rc = f512_upload(tw2, allocs, &hst.t.tw2))) return rc;
  if ((rc = f512_upload(tk, allocs, &hst.t.tk))) return rc;
  MelItems mr = pack_mel_items_T(bank, p.K, p.M, p.use_mag ? 0.5f : 0.25f, 32, 4, F5W_PIECE, /*uniform=*/true);
  if (mr.max_reach > F5W_PBINS) return B200FEAT_EUNSUPPORTED;
  hst.t.mel_rounds = mr.rounds;
  hst.t.xfloats = std::max(F5W_XBUF * 2, shape.slots * (mr.rounds * 32 + ((p.M + 3) & ~3)));
  {
    std::vector<unsigned char> blob;
    auto append = [&](const void *src, size_t bytes) -> int {
      const size_t off = blob.size();
      blob.resize(off + ((bytes + 15) & ~(size_t)15), 0);
      if (bytes) memcpy(blob.data() + off, src, bytes);
      return (int)off;
    };
    append(win2.data(), win2.size() * sizeof(float2));
    hst.t.off_tw1 = append(tw1.data(), tw1.size() * sizeof(float2));
    hst.t.off_rstart = append(mr.rstart.data(), mr.rstart.size() * sizeof(int));
    std::vector<int> fdesc((size_t)(std::max(p.M, 1) + (p.M + 31) / 32 + 1) * 2, 0);  // M x {first item, items}, then per 32 filters {max items, 0}
    for (int m = 0; m < p.M; ++m) {
      fdesc[2 * m] = mr.qfirst[m]; fdesc[2 * m + 1] = mr.qcount[m];
      int &mx = fdesc[2 * (p.M + m / 32)];
      mx = std::max(mx, mr.qcount[m]);
    }
    hst.t.off_fdesc = append(fdesc.data(), fdesc.size() * sizeof(int));
    hst.t.off_mw = append(mr.wdense.data(), mr.wdense.size() * sizeof(float));
    const unsigned char *d = nullptr;
    if ((rc = f512_upload(blob, allocs, &d))) return rc;
    hst.t.cblob = d;
    hst.t.cblob_bytes = (int)blob.size();
  }
  hst.smem = fast512w_smem_bytes(hst.t, shape.warps, shape.slots);
  if (hst.smem > (size_t)(227 * 1024 / shape.minb) - 1024) return B200FEAT_EUNSUPPORTED;  // keep `minb` CTAs per SM
  DevBatch none{};
  for (int dt = 0; dt < 2; ++dt)
    for (int L : {400, 0})
      if (f5w_dispatch(dt, L, hst.shape, false, hst.smem, p, hst.t, none, dim3(1), nullptr)) return B200FEAT_ECUDA;
  *out = hst;
  *frames_per_tile = shape.warps * shape.slots;
  return 0;
}

static inline int fast512w_launch(const DevPlan &p, const Fast512wHost &hst, const DevBatch &b, int dt, int sm_count, cudaStream_t stream) {
  const F5wShape shape = kF5wShapes[hst.shape];
  int64_t blocks = b.num_tiles;
  const int64_t cap = (int64_t)sm_count * shape.minb;
  if (blocks > cap) blocks = cap;
  f5w_dispatch(dt, p.L == 400 ? 400 : 0, hst.shape, true, hst.smem, p, hst.t, b, dim3((unsigned)blocks), stream);
  return (int)cudaGetLastError();
}
