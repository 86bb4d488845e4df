/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * Plain-C second restatement of the reference hot path (float32 data flow; the DFT itself is
 * evaluated in double so that this oracle does not share rounding behaviour with either pocketfft
 * or the CUDA kernels).  Built by `make -C oracle` into oracle/_build/libfbank_oracle.so and called
 * only from tests/ through ctypes.  Tables (window, dense mel bank, dct, lifter) are inputs: how they
 * are built is pinned separately (tests/test_plan_tables.py).
 *
 * Follows, step by step (paths relative to /root/reference):
 *   frame count / reflection ... lhotse/features/kaldi/layers.py:747-772
 *   DC removal ................. layers.py:155-157
 *   raw log-energy ............. layers.py:159-161, :859-870
 *   pre-emphasis ............... layers.py:164-167
 *   window + zero pad .......... layers.py:170-181
 *   rfft, |X|^2 / |X| .......... layers.py:32-42
 *   spectrogram kinds .......... layers.py:392-402, :461-473
 *   mel + log .................. layers.py:565-578
 *   dct + lifter ............... layers.py:708-724
 * and, for the centre-padded STFT front ends (oracle_extract_center below):
 *   WhisperFbank ............... lhotse/features/whisper_fbank.py:16-84 (torch.stft(center=True), |X|^2, mel, log10,
 *                                clamp to the utterance maximum - 8, (x + 4) / 4, zero row up to compute_num_frames)
 *   LibrosaFbank ............... lhotse/features/librosa_fbank.py:64-135 (librosa.stft(pad_mode="reflect"), |X|, mel,
 *                                log10(max(eps, .)), pad_or_truncate_features)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int32_t feature; /* 0 fbank, 1 mfcc, 2 spectrogram, 3 log-spectrogram */
  int32_t L, S, N, M, C;
  int32_t snip_edges, remove_dc, use_energy, raw_energy, use_mag, use_lifter;
  float preemph, energy_floor, mel_floor, log_spec_eps;
} oracle_plan;

int64_t oracle_num_frames(const oracle_plan *p, int64_t n) {
  if (p->snip_edges) return n < p->L ? 0 : 1 + (n - p->L) / p->S;
  return (n + p->S / 2) / p->S;
}

int32_t oracle_feature_dim(const oracle_plan *p) {
  if (p->feature == 0) return p->M + (p->use_energy ? 1 : 0);
  if (p->feature == 1) return p->C;
  return p->N / 2 + 1;
}

static float log_energy(const oracle_plan *p, float e) {
  float le = logf(e + 1e-15f);
  if (p->energy_floor > 0.0f) {
    const float fl = (float)log((double)p->energy_floor);
    if (fl > le) le = fl;
  }
  return le;
}

/* X[k] = sum_j y[j] exp(-2 pi i j k / N), k = 0..N/2, accumulated in double */
static void rdft(const float *y, int N, double *re, double *im, const double *cs, const double *sn) {
  for (int k = 0; k <= N / 2; ++k) {
    double ar = 0.0, ai = 0.0;
    int idx = 0;
    for (int j = 0; j < N; ++j) {
      ar += (double)y[j] * cs[idx];
      ai -= (double)y[j] * sn[idx];
      idx += k;
      if (idx >= N) idx -= N;
    }
    re[k] = ar;
    im[k] = ai;
  }
}

/* returns 0, or -1 when the cut cannot be framed (too short for one reflection per side) */
int oracle_extract(const oracle_plan *p, const float *x, int64_t n, const float *window, const float *mel_bank /* K x M */,
                   const float *dct /* M x C */, const float *lifter /* C */, float *out /* T x F */) {
  const int L = p->L, S = p->S, N = p->N, K = N / 2 + 1, M = p->M, C = p->C;
  const int64_t T = oracle_num_frames(p, n);
  const int F = oracle_feature_dim(p);
  if (T <= 0) return -1;
  const int64_t left = (L - S) / 2;
  if (!p->snip_edges) {
    const int64_t right = (T - 1) * S + L - n - left;
    if (left > n || right > n) return -1;
  }
  float *f = (float *)calloc((size_t)N, sizeof(float));
  float *spec = (float *)malloc(sizeof(float) * (size_t)K);
  float *mel = (float *)malloc(sizeof(float) * (size_t)(M > 0 ? M : 1));
  double *re = (double *)malloc(sizeof(double) * (size_t)K), *im = (double *)malloc(sizeof(double) * (size_t)K);
  double *cs = (double *)malloc(sizeof(double) * (size_t)N), *sn = (double *)malloc(sizeof(double) * (size_t)N);
  for (int j = 0; j < N; ++j) {
    cs[j] = cos(2.0 * M_PI * (double)j / (double)N);
    sn[j] = sin(2.0 * M_PI * (double)j / (double)N);
  }
  for (int64_t t = 0; t < T; ++t) {
    float *o = out + t * F;
    /* gather with symmetric reflection */
    for (int j = 0; j < L; ++j) {
      int64_t i = t * S + j - (p->snip_edges ? 0 : left);
      if (!p->snip_edges) {
        if (i < 0) i = -i - 1;
        if (i >= n) i = 2 * n - 1 - i;
      }
      f[j] = x[i];
    }
    for (int j = L; j < N; ++j) f[j] = 0.0f;
    if (p->remove_dc) {
      float s = 0.0f;
      for (int j = 0; j < L; ++j) s += f[j];
      const float mu = s / (float)L;
      for (int j = 0; j < L; ++j) f[j] -= mu;
    }
    float le = 0.0f;
    if (p->use_energy && p->raw_energy) {
      float e = 0.0f;
      for (int j = 0; j < L; ++j) e += f[j] * f[j];
      le = log_energy(p, e);
    }
    if (p->preemph != 0.0f) {
      for (int j = L - 1; j > 0; --j) f[j] = f[j] - p->preemph * f[j - 1];
      f[0] = f[0] - p->preemph * f[0];
    }
    for (int j = 0; j < L; ++j) f[j] *= window[j];
    if (p->use_energy && !p->raw_energy) {
      float e = 0.0f;
      for (int j = 0; j < N; ++j) e += f[j] * f[j];
      le = log_energy(p, e);
    }
    rdft(f, N, re, im, cs, sn);
    for (int k = 0; k < K; ++k) {
      const double pw = re[k] * re[k] + im[k] * im[k];
      spec[k] = p->use_mag ? (float)sqrt(pw) : (float)pw;
    }
    if (p->feature == 2) {
      for (int k = 0; k < K; ++k) o[k] = spec[k];
      if (p->use_energy) o[0] = le;
    } else if (p->feature == 3) {
      for (int k = 0; k < K; ++k) o[k] = logf(spec[k] + p->log_spec_eps);
      if (p->use_energy) o[0] = le;
    } else {
      for (int m = 0; m < M; ++m) {
        float acc = 0.0f;
        for (int k = 0; k < K; ++k) {
          const float w = mel_bank[(size_t)k * M + m];
          if (w != 0.0f) acc += spec[k] * w;
        }
        mel[m] = logf(acc > p->mel_floor ? acc : p->mel_floor);
      }
      if (p->feature == 0) {
        const int sh = p->use_energy ? 1 : 0;
        for (int m = 0; m < M; ++m) o[m + sh] = mel[m];
        if (sh) o[0] = le;
      } else {
        for (int c = 0; c < C; ++c) {
          float acc = 0.0f;
          for (int m = 0; m < M; ++m) acc += mel[m] * dct[(size_t)m * C + c];
          if (p->use_lifter) acc *= lifter[c];
          o[c] = acc;
        }
        if (p->use_energy) o[0] = le;
      }
    }
  }
  free(f); free(spec); free(mel); free(re); free(im); free(cs); free(sn);
  return 0;
}

/*
 * Centre-padded log10-mel front ends.  kind 4 = whisper-fbank, kind 5 = librosa-fbank; N = n_fft (frame = N samples under
 * `window`, which already holds a shorter window centred and zero-padded), S = hop, reflect padding of N/2 samples per side
 * WITHOUT repeating the edge sample.  `use_mag`: |X| (librosa) or |X|^2 (whisper).  Output rows = (n + S/2) / S.
 * Returns 0, or -1 when n <= N/2 (reflect padding impossible: torch / numpy raise there).
 */
int oracle_extract_center(int32_t kind, int32_t N, int32_t S, int32_t M, int32_t use_mag, float floor_, const float *x,
                          int64_t n, const float *window /* N */, const float *mel_bank /* K x M */, float *out) {
  const int K = N / 2 + 1;
  if (n <= N / 2) return -1;
  const int64_t rows = (n + S / 2) / S;
  const int64_t Tstft = kind == 4 ? n / S : 1 + n / S; /* whisper drops the stft's last frame (whisper_fbank.py:63) */
  const int64_t Tv = Tstft < rows ? Tstft : rows;
  float *f = (float *)malloc(sizeof(float) * (size_t)N);
  double *re = (double *)malloc(sizeof(double) * (size_t)K), *im = (double *)malloc(sizeof(double) * (size_t)K);
  double *cs = (double *)malloc(sizeof(double) * (size_t)N), *sn = (double *)malloc(sizeof(double) * (size_t)N);
  for (int j = 0; j < N; ++j) {
    cs[j] = cos(2.0 * M_PI * (double)j / (double)N);
    sn[j] = sin(2.0 * M_PI * (double)j / (double)N);
  }
  float vmax = -INFINITY;
  for (int64_t t = 0; t < Tv; ++t) {
    for (int j = 0; j < N; ++j) {
      int64_t i = t * S + j - N / 2;
      if (i < 0) i = -i;
      if (i >= n) i = 2 * n - 2 - i;
      f[j] = x[i] * window[j];
    }
    rdft(f, N, re, im, cs, sn);
    for (int m = 0; m < M; ++m) {
      float acc = 0.0f;
      for (int k = 0; k < K; ++k) {
        const float w = mel_bank[(size_t)k * M + m];
        if (w != 0.0f) {
          const double pw = re[k] * re[k] + im[k] * im[k];
          acc += (use_mag ? (float)sqrt(pw) : (float)pw) * w;
        }
      }
      const float v = log10f(acc > floor_ ? acc : floor_);
      out[t * M + m] = v;
      if (v > vmax) vmax = v;
    }
  }
  if (kind == 4)
    for (int64_t i = 0; i < Tv * M; ++i) {
      const float v = out[i] > vmax - 8.0f ? out[i] : vmax - 8.0f;
      out[i] = (v + 4.0f) / 4.0f;
    }
  for (int64_t i = Tv * M; i < rows * M; ++i) out[i] = kind == 4 ? 0.0f : (float)log(1e-10); /* :73-80 / librosa_fbank.py:51-56 */
  free(f); free(re); free(im); free(cs); free(sn);
  return 0;
}

