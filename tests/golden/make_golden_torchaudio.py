#!/usr/bin/env python
"""Second oracle: golden vectors from `torchaudio.compliance.kaldi.{fbank,mfcc}` called the way
lhotse's TorchaudioFbank / TorchaudioMfcc call them (lhotse/features/base.py:408-424: snip_edges=False, lengths in ms),
to pin the torchaudio / kaldifeat config family: Kaldi log-energy convention, energy placement, 2*pi/(L-1) blackman.
Writes tests/golden/golden_torchaudio_v1.npz.  Build container only (torchaudio is not a runtime dependency)."""
import json
import os

import numpy as np
import torch
import torchaudio.compliance.kaldi as K

HERE = os.path.dirname(os.path.abspath(__file__))

# TorchaudioFbankConfig / TorchaudioMfccConfig field names (lhotse/features/fbank.py:11-39, mfcc.py:9-39)
FBANK_DEFAULTS = dict(dither=0.0, window_type="povey", frame_length=0.025, frame_shift=0.01, remove_dc_offset=True,
                      round_to_power_of_two=True, energy_floor=1e-10, min_duration=0.0, preemphasis_coefficient=0.97,
                      raw_energy=True, low_freq=20.0, high_freq=-400.0, num_mel_bins=80, use_energy=False,
                      vtln_low=100.0, vtln_high=-500.0, vtln_warp=1.0)
MFCC_DEFAULTS = dict(FBANK_DEFAULTS, num_mel_bins=23, num_ceps=13, cepstral_lifter=22.0)

CASES = [
    ("fbank", {}),
    ("fbank", dict(use_energy=True)),
    ("fbank", dict(use_energy=True, raw_energy=False, energy_floor=0.0)),
    ("fbank", dict(window_type="blackman", num_mel_bins=40)),
    ("fbank", dict(window_type="hamming", preemphasis_coefficient=0.0, remove_dc_offset=False)),
    ("mfcc", {}),
    ("mfcc", dict(use_energy=True)),
    ("mfcc", dict(num_ceps=20, num_mel_bins=40, cepstral_lifter=0.0)),
    # VTLN warping of the filter edges (appended in round 2: the earlier cases keep their random draws)
    ("fbank", dict(vtln_warp=1.1)),
    ("fbank", dict(vtln_warp=0.9, num_mel_bins=40, use_energy=True)),
    ("mfcc", dict(vtln_warp=1.15, vtln_low=200.0, vtln_high=-800.0)),
]


def main():
    torch.set_num_threads(1)
    rs = np.random.RandomState(123)
    out, manifest = {}, []
    for i, (kind, over) in enumerate(CASES):
        cfg = dict(MFCC_DEFAULTS if kind == "mfcc" else FBANK_DEFAULTS, **over)
        x = (0.1 * rs.randn(12000)).astype(np.float32)
        if i % 2 == 1:
            x[:3000] *= 1e-4  # a quiet stretch exercises the energy floor / eps paths
        params = dict(cfg, sample_frequency=16000, snip_edges=False)
        params["frame_shift"] *= 1000.0
        params["frame_length"] *= 1000.0
        fn = K.mfcc if kind == "mfcc" else K.fbank
        y = fn(torch.from_numpy(x).unsqueeze(0), **params).to(torch.float32).numpy()
        out[f"x{i}"], out[f"y{i}"] = x, y
        manifest.append(dict(feature=kind, cfg=cfg, shape=list(y.shape)))
        print(i, kind, over, y.shape)
    out["manifest"] = np.frombuffer(json.dumps(manifest).encode(), dtype=np.uint8)
    path = os.path.join(HERE, "golden_torchaudio_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__":
    main()
