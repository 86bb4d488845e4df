"""One small ragged batch through every kernel and output mode — the workload for `compute-sanitizer`
(`--tool memcheck`, `--tool racecheck`, `--tool synccheck`); prints one line per plan and a checksum."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import lhotse_b200 as lb

rs = np.random.RandomState(0)


def waves(sr, lens_s=(0.35, 1.0, 0.61, 2.03)):
    return [(0.1 * rs.randn(int(sr * s) + i)).astype(np.float32) for i, s in enumerate(lens_s)]


PLANS = [
    ("fast512 fbank", lb.B200Fbank(), 16000),
    ("fast512 mfcc+energy", lb.B200Mfcc(lb.B200MfccConfig(use_energy=True)), 16000),
    ("fast512 log-spectrogram", lb.B200LogSpectrogram(), 16000),
    ("tc512 fbank", lb.B200Fbank(lb.B200FbankConfig(kernel="tc")), 16000),
    ("fast256 fbank40 8k", lb.B200Fbank(lb.B200FbankConfig(sampling_rate=8000, num_filters=40)), 8000),
    ("fast1024 fbank 24k", lb.B200Fbank(lb.B200FbankConfig(sampling_rate=24000)), 24000),
    ("fast400 fbank", lb.B200Fbank(lb.B200FbankConfig(round_to_power_of_two=False)), 16000),
    ("fast400 whisper", lb.B200WhisperFbank(), 16000),
    ("fast1024 librosa", lb.B200LibrosaFbank(), 22050),
    ("generic fbank", lb.B200Fbank(lb.B200FbankConfig(kernel="generic")), 16000),
    ("generic whisper", lb.B200WhisperFbank(lb.B200WhisperFbankConfig(kernel="generic")), 16000),
    ("fast2048 fbank 24k/50ms", lb.B200Fbank(lb.B200FbankConfig(sampling_rate=24000, frame_length=0.05)), 24000),
    ("fast2048 mfcc+energy 44.1k", lb.B200Mfcc(lb.B200MfccConfig(sampling_rate=44100, use_energy=True)), 44100),
    ("fast2048 spectrogram 48k", lb.B200Spectrogram(lb.B200SpectrogramConfig(sampling_rate=48000)), 48000),
    ("fast1024 mfcc 22.05k", lb.B200Mfcc(lb.B200MfccConfig(sampling_rate=22050)), 22050),
    ("generic N=2048", lb.B200Fbank(lb.B200FbankConfig(sampling_rate=24000, frame_length=0.05, kernel="generic")), 24000),
    ("generic N=551 (19x29)", lb.B200Spectrogram(lb.B200SpectrogramConfig(sampling_rate=22050, round_to_power_of_two=False)), 22050),
]
total = 0.0
for name, ext, sr in PLANS:
    xs = waves(sr)
    a = ext.extract_batch(xs, sr)                                   # host route (C-ABI extract_host_at)
    b = ext.extract_batch([torch.from_numpy(x) for x in xs], sr)    # device route
    c, lens = ext.extract_batch_padded([torch.from_numpy(x) for x in xs], sr)   # padded output mode
    i16 = ext.extract_batch([np.clip(x * 32768, -32768, 32767).astype(np.int16) for x in xs], sr)  # int16 staging
    s = float(sum(np.asarray(v, dtype=np.float64).sum() for v in a)) + float(sum(t.double().sum().item() for t in b))
    ok = all(np.isfinite(np.asarray(v)).all() for v in a) and bool(torch.isfinite(c).all()) and all(np.isfinite(v).all() for v in i16)
    total += s
    print(f"{name:28s} kernel={ext.engine.kernel:8s} rows={[v.shape[0] for v in a]} finite={ok}", flush=True)
torch.cuda.synchronize()
print("checksum", total)
