#!/usr/bin/env python
"""Golden vectors for Kaldi's `htk_compat=True` layout (kaldifeat FbankOptions / MfccOptions.htk_compat, lhotse/features/kaldifeat.py:158,
:227): kaldifeat is absent here, so the vectors come from torchaudio's Kaldi-compatible implementation of the same option
(`torchaudio.compliance.kaldi.{fbank,mfcc}(htk_compat=True)`: log-energy / C0 column LAST, C0 * sqrt(2) without use_energy), called with
the kaldifeat family's defaults (snip_edges=False, energy_floor 1e-10).  Writes tests/golden/golden_kaldi_htk_v1.npz.  Build container only."""
import json
import os

import numpy as np
import torch
import torchaudio.compliance.kaldi as K

HERE = os.path.dirname(os.path.abspath(__file__))
BASE = dict(dither=0.0, window_type="povey", frame_length=25.0, frame_shift=10.0, remove_dc_offset=True, round_to_power_of_two=True,
            energy_floor=1e-10, preemphasis_coefficient=0.97, raw_energy=True, low_freq=20.0, high_freq=-400.0, snip_edges=False,
            sample_frequency=16000.0, htk_compat=True)
CASES = [  # (feature, kaldifeat-config overrides {use_energy, num_bins, num_ceps, cepstral_lifter}, torchaudio kwargs)
    ("fbank", dict(use_energy=True, num_bins=80)),
    ("fbank", dict(use_energy=False, num_bins=40)),
    ("mfcc", dict(use_energy=False, num_bins=23, num_ceps=13, cepstral_lifter=22.0)),
    ("mfcc", dict(use_energy=True, num_bins=23, num_ceps=13, cepstral_lifter=22.0)),
    ("mfcc", dict(use_energy=False, num_bins=40, num_ceps=20, cepstral_lifter=0.0)),
]


def main():
    torch.set_num_threads(1)
    rs = np.random.RandomState(321)
    out, manifest = {}, []
    for i, (kind, c) in enumerate(CASES):
        x = (0.1 * rs.randn(12000)).astype(np.float32)
        if i % 2 == 1:
            x[:3000] *= 1e-4
        kw = dict(BASE, use_energy=c["use_energy"], num_mel_bins=c["num_bins"])
        if kind == "mfcc":
            kw.update(num_ceps=c["num_ceps"], cepstral_lifter=c["cepstral_lifter"])
        fn = K.mfcc if kind == "mfcc" else K.fbank
        y = fn(torch.from_numpy(x).unsqueeze(0), **kw).to(torch.float32).numpy()
        out[f"x{i}"], out[f"y{i}"] = x, y
        manifest.append(dict(feature=kind, cfg=c, shape=list(y.shape)))
        print(i, kind, c, y.shape)
    out["manifest"] = np.frombuffer(json.dumps(manifest).encode(), dtype=np.uint8)
    path = os.path.join(HERE, "golden_kaldi_htk_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main()
