"""
TEST INFRASTRUCTURE — NOT PRODUCT CODE.

CPU restatement (numpy) of `LibrosaFbank.extract` -> `logmelfilterbank` (lhotse/features/librosa_fbank.py:64-135, :157-159).
Same rules as the other oracles: only ``tests/``, ``__graft_entry__.smoke()`` and the CPU legs of ``bench.py`` may import it.

PARITY PIN — read this.  The arithmetic of this path lives in a third-party dependency that is NOT in the reference tree:
**librosa** (`librosa.stft`, `librosa.filters.mel`; an optional, unpinned extra of the reference — `setup.py` lists it under
no version — and absent from this image, so the reference class itself cannot be constructed here).  What pins this oracle:
  * the reference's OWN code (`logmelfilterbank`: magnitude, `np.dot` with the mel basis, `np.log10(np.maximum(eps, .))`,
    `pad_or_truncate_features`) is executed for the golden vectors (`tests/golden/make_golden_librosa.py`) on a stand-in
    `librosa` module whose `stft` / `filters.mel` are **transformers.audio_utils.spectrogram / mel_filter_bank** — an
    independent implementation that upstream tests against librosa (tests/refshim.py::install_librosa_standin);
  * this file restates librosa's published algorithm directly (below) and must agree with those vectors
    (`tests/test_librosa.py::test_librosa_oracle_matches_golden`, tolerance 2e-5 in log10 units: the two STFTs round
    differently — librosa multiplies by a float64 window and transforms in float64, transformers works in float64 too but
    stores float32 frames).
So: pinned against the reference's call sites + an independent third-party restatement, NOT against librosa itself.

librosa.stft(y, n_fft, hop_length, win_length, window, center=True, pad_mode="reflect") as restated here:
  window = scipy.signal.get_window(window, win_length, fftbins=True) (float64, periodic), zero-padded to n_fft around its
  centre; y padded by n_fft // 2 on both sides (numpy "reflect": the edge sample is not repeated); frame t = padded
  [t * hop, t * hop + n_fft), t = 0 .. n // hop; X = rfft(window * frame) evaluated in float64 and stored as complex64.
librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax): Slaney scale, unit-area triangles, float32 — `oracle/whisper_oracle.py`.
"""
from __future__ import annotations

import numpy as np

from . import whisper_oracle as W


def slaney_mel_filters(sr: int, n_fft: int, n_mels: int, fmin: float, fmax: float) -> np.ndarray:
    """(n_mels, n_fft // 2 + 1) float32, same construction as whisper_oracle.slaney_mel_filters with explicit corners."""
    bins = np.arange(n_fft // 2 + 1, dtype=np.float64) * (sr / n_fft)
    corners = W._mel_to_hz(np.linspace(W._hz_to_mel(fmin), W._hz_to_mel(fmax), n_mels + 2))
    out = np.zeros((n_mels, bins.size), dtype=np.float64)
    for m in range(n_mels):
        lo, ce, hi = corners[m], corners[m + 1], corners[m + 2]
        out[m] = np.maximum(0.0, np.minimum((bins - lo) / (ce - lo), (hi - bins) / (hi - ce))) * (2.0 / (hi - lo))
    return out.astype(np.float32)


def periodic_window(name: str, length: int) -> np.ndarray:
    a = 2.0 * np.pi * np.arange(length, dtype=np.float64) / length
    return {"hann": 0.5 - 0.5 * np.cos(a), "hamming": 0.54 - 0.46 * np.cos(a),
            "blackman": 0.42 - 0.5 * np.cos(a) + 0.08 * np.cos(2 * a), "boxcar": np.ones(length)}[name]


def num_rows(n: int, hop: int) -> int:
    return (n + hop // 2) // hop  # lhotse/utils.py:410-421 with duration = n / sr, frame_shift = hop / sr


def extract(x, sampling_rate=22050, fft_size=1024, hop_size=256, win_length=None, window="hann", num_mel_bins=80,
            fmin=80, fmax=7600, eps=1e-10, float64: bool = False) -> np.ndarray:
    """(n,) or (1, n) -> (num_rows, num_mel_bins).  `float64=True` keeps every step in double (truth for tolerance gates)."""
    a = np.asarray(x).reshape(-1)
    n = a.shape[0]
    wl = fft_size if win_length is None else win_length
    win = np.zeros(fft_size, dtype=np.float64)
    lo = (fft_size - wl) // 2
    win[lo: lo + wl] = periodic_window(window, wl)
    padded = np.pad(a.astype(np.float64), fft_size // 2, mode="reflect")
    T = 1 + n // hop_size
    idx = np.arange(fft_size)[None, :] + hop_size * np.arange(T)[:, None]
    spec = np.fft.rfft(padded[idx] * win[None, :], axis=1)  # (T, K) float64
    fmin = 0.0 if fmin is None else float(fmin)
    fmax = sampling_rate / 2 if fmax is None else float(fmax)
    basis = slaney_mel_filters(sampling_rate, fft_size, num_mel_bins, fmin, fmax)
    if float64:
        feats = np.log10(np.maximum(eps, np.abs(spec) @ basis.astype(np.float64).T))
    else:
        mag = np.abs(spec.astype(np.complex64))  # librosa stores complex64; np.abs -> float32
        feats = np.log10(np.maximum(eps, np.dot(mag, basis.T)))
    rows = num_rows(n, hop_size)
    assert abs(feats.shape[0] - rows) <= 1  # pad_or_truncate_features, librosa_fbank.py:41-61
    if feats.shape[0] > rows:
        feats = feats[:rows]
    elif feats.shape[0] < rows:
        feats = np.pad(feats, ((0, rows - feats.shape[0]), (0, 0)), constant_values=np.log(1e-10))
    return feats if float64 else feats.astype(np.float32)
