#!/usr/bin/env python
"""Golden vectors for the TorchaudioSpectrogram family: `torchaudio.compliance.kaldi.spectrogram` called the way lhotse's
TorchaudioFeatureExtractor.extract calls it (lhotse/features/base.py:408-424: snip_edges=False, lengths in ms) with
TorchaudioSpectrogramConfig's fields (lhotse/features/spectrogram.py:11-31).
Writes tests/golden/golden_torchaudio_spec_v1.npz.  Build container only (torchaudio is not a runtime dependency)."""
import json
import os

import numpy as np
import torch
import torchaudio.compliance.kaldi as K

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULTS = dict(dither=0.0, window_type="povey", frame_length=0.025, frame_shift=0.01, remove_dc_offset=True,
                round_to_power_of_two=True, energy_floor=1e-10, min_duration=0.0, preemphasis_coefficient=0.97, raw_energy=True)
CASES = [
    (16000, {}),
    (16000, dict(raw_energy=False, energy_floor=0.0)),
    (16000, dict(window_type="hamming", preemphasis_coefficient=0.0, remove_dc_offset=False)),
    (16000, dict(round_to_power_of_two=False)),
    (8000, dict(window_type="blackman")),
]


def main():
    torch.set_num_threads(1)
    rs = np.random.RandomState(321)
    out, manifest = {}, []
    for i, (sr, over) in enumerate(CASES):
        cfg = dict(DEFAULTS, **over)
        x = (0.1 * rs.randn(int(0.6 * sr))).astype(np.float32)
        if i % 2 == 1:
            x[: len(x) // 3] = 0.0  # exact silence: every bin sits on the eps32 floor, the energy on its own floor
        params = dict(cfg, sample_frequency=sr, snip_edges=False)
        params["frame_shift"] *= 1000.0
        params["frame_length"] *= 1000.0
        y = K.spectrogram(torch.from_numpy(x).unsqueeze(0), **params).to(torch.float32).numpy()
        out[f"x{i}"], out[f"y{i}"] = x, y
        manifest.append(dict(sampling_rate=sr, cfg=cfg, shape=list(y.shape)))
        print(i, sr, over, y.shape)
    out["manifest"] = np.frombuffer(json.dumps(manifest).encode(), dtype=np.uint8)
    path = os.path.join(HERE, "golden_torchaudio_spec_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__":
    main()
