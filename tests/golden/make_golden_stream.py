#!/usr/bin/env python
"""Generates tests/golden/golden_stream_v1.npz by running the REAL reference's streaming entry points
(`Wav2LogFilterBank/Wav2MFCC/Wav2Spec/Wav2LogSpec.online_inference`, lhotse/features/kaldi/layers.py:199-224, :326-333,
framing :775-857) chunk by chunk on seeded inputs, the way test/features/test_kaldi_layers.py:199-235 does.  Build
container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_stream.py

Each case stores the waveform, the chunk boundaries, the per-call frame counts, the concatenated per-call outputs and
the final remainder."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import refshim  # noqa: E402

refshim.import_reference()
import torch  # noqa: E402
from lhotse.features.kaldi.layers import Wav2LogFilterBank, Wav2LogSpec, Wav2MFCC, Wav2Spec  # noqa: E402

LAYERS = {"fbank": Wav2LogFilterBank, "mfcc": Wav2MFCC, "spectrogram": Wav2Spec, "log-spectrogram": Wav2LogSpec}

CASES = [
    # feature, layer kwargs, number of samples, chunk sizes (cycled)
    ("fbank", dict(num_filters=80), 16000, [1200]),
    ("fbank", dict(num_filters=80), 20000, [400, 3000, 161, 159, 5000]),
    ("fbank", dict(num_filters=40, snip_edges=True, use_energy=True), 12345, [1000, 800, 2500]),
    ("fbank", dict(num_filters=40, sampling_rate=8000), 9000, [800]),
    ("mfcc", dict(), 16000, [1600, 777]),
    ("spectrogram", dict(use_energy=False), 8000, [1234]),
    ("log-spectrogram", dict(use_energy=True), 8000, [999, 2001]),
]


def main():
    out = {}
    meta = []
    for i, (feature, kw, n, chunks) in enumerate(CASES):
        rs = np.random.RandomState(100 + i)
        x = (0.1 * rs.randn(1, n)).astype(np.float32)
        layer = LAYERS[feature](**kw)
        bounds, pos, k = [0], 0, 0
        while pos < n:
            pos = min(n, pos + chunks[k % len(chunks)])
            bounds.append(pos)
            k += 1
        feats, counts, rem = [], [], None
        with torch.no_grad():
            for a, b in zip(bounds[:-1], bounds[1:]):
                y, rem = layer.online_inference(torch.from_numpy(x[:, a:b]), context=rem)
                feats.append(y[0].numpy())
                counts.append(y.shape[1])
        out[f"x{i}"] = x[0]
        out[f"y{i}"] = np.concatenate([f for f in feats if f.shape[0]], axis=0)
        out[f"r{i}"] = rem[0].numpy()
        meta.append(dict(feature=feature, cfg=kw, n=n, bounds=bounds, counts=counts))
        print(i, feature, kw, "frames per call", counts, "remainder", rem.shape[1])
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "golden_stream_v1.npz"), **out)


if __name__ == "__main__":
    main()
