"""The plain-C restatement (oracle/fbank_oracle.c) against the golden vectors of the real reference."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

from helpers import gate, load_golden, oracle_cfg
from lhotse_b200 import build_plan
from lhotse_b200.extractors import (B200FbankConfig, B200LogSpectrogramConfig, B200MfccConfig, B200SpectrogramConfig)
from oracle import kaldi_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_build", "libfbank_oracle.so")
KINDS = {"fbank": (0, B200FbankConfig), "mfcc": (1, B200MfccConfig), "spectrogram": (2, B200SpectrogramConfig),
         "log-spectrogram": (3, B200LogSpectrogramConfig)}


class OraclePlan(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("feature", "L", "S", "N", "M", "C", "snip_edges", "remove_dc", "use_energy",
                                         "raw_energy", "use_mag", "use_lifter")] + \
               [(n, C.c_float) for n in ("preemph", "energy_floor", "mel_floor", "log_spec_eps")]


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-s"], check=True)
    lib = C.CDLL(LIB)
    lib.oracle_num_frames.restype = C.c_int64
    lib.oracle_num_frames.argtypes = [C.POINTER(OraclePlan), C.c_int64]
    lib.oracle_extract.restype = C.c_int
    return lib


GOLD = [g for g in load_golden() if g[1]["n"] <= 32000 and g[1]["cfg"].get("sampling_rate", 16000) <= 22050]


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("i,c,x,y", GOLD, ids=[f"{i}-{c['feature']}-{c['kind']}-{c['n']}" for i, c, _, _ in GOLD])
def test_c_oracle_matches_golden(lib, i, c, x, y):
    kind, ccls = KINDS[c["feature"]]
    plan = build_plan(c["feature"], ccls(**c["cfg"]))
    p = OraclePlan(kind, plan.L, plan.S, plan.N, plan.num_filters, plan.num_ceps, int(plan.snip_edges),
                   int(plan.remove_dc_offset), int(plan.use_energy), int(plan.raw_energy), int(plan.use_fft_mag),
                   int(plan.lifter is not None), plan.preemph_coeff, plan.energy_floor, plan.mel_floor, plan.log_spec_eps)
    T = lib.oracle_num_frames(C.byref(p), len(x))
    assert T == y.shape[0]  # frame counts: bit-exact
    out = np.empty(y.shape, dtype=np.float32)
    rc = lib.oracle_extract(C.byref(p), _ptr(np.ascontiguousarray(x)), C.c_int64(len(x)), _ptr(plan.window),
                            _ptr(plan.mel_bank), _ptr(plan.dct), _ptr(plan.lifter), _ptr(out))
    assert rc == 0
    truth = O.extract(x, oracle_cfg(c["feature"], c["cfg"]), dtype=torch.float64)
    ok, msg = gate(out, y, truth, c["feature"], c["cfg"].get("use_energy", False), c["cfg"].get("use_fft_mag", False))
    assert ok, msg


def test_c_oracle_rejects_short_inputs(lib):
    plan = build_plan("fbank", B200FbankConfig())
    p = OraclePlan(0, plan.L, plan.S, plan.N, 80, 0, 0, 1, 0, 1, 0, 0, 0.97, 1e-10, plan.mel_floor, 1e-15)
    x = np.zeros(100, dtype=np.float32)
    out = np.empty((1, 80), dtype=np.float32)
    assert lib.oracle_extract(C.byref(p), _ptr(x), C.c_int64(100), _ptr(plan.window), _ptr(plan.mel_bank), None, None, _ptr(out)) == -1


def _center_goldens():
    import json

    out = []
    here = os.path.dirname(os.path.abspath(__file__))
    for kind, fname in ((4, "golden_whisper_v1.npz"), (5, "golden_librosa_v1.npz")):
        g = np.load(os.path.join(here, "golden", fname))
        for i, c in enumerate(json.loads(bytes(g["manifest"]).decode())):
            if c["n"] <= 32000 and (kind == 4 or c["cfg"]["fft_size"] <= 1024):  # the O(N^2) DFT keeps these in seconds
                out.append((kind, i, c, g[f"x{i}"], g[f"y{i}"]))
    return out


CENTER = _center_goldens()


@pytest.mark.parametrize("kind,i,c,x,y", CENTER, ids=[f"{'whisper' if k == 4 else 'librosa'}-{i}-{c['n']}" for k, i, c, _, _ in CENTER])
def test_c_oracle_center_front_ends_match_golden(lib, kind, i, c, x, y):
    """Third, plain-C restatement of the Whisper / Librosa front ends (oracle_extract_center) against the vectors produced by
    the reference classes; tables come from the product's plan builder (pinned separately in test_whisper / test_librosa)."""
    from lhotse_b200 import B200LibrosaFbankConfig, B200WhisperFbankConfig

    if kind == 4:
        plan = build_plan("whisper-fbank", B200WhisperFbankConfig(num_filters=c["num_filters"]))
    else:
        plan = build_plan("librosa-fbank", B200LibrosaFbankConfig(**c["cfg"]))
    out = np.empty(y.shape, dtype=np.float32)
    lib.oracle_extract_center.restype = C.c_int
    rc = lib.oracle_extract_center(C.c_int32(kind), C.c_int32(plan.N), C.c_int32(plan.S), C.c_int32(plan.num_filters),
                                   C.c_int32(int(plan.use_fft_mag)), C.c_float(plan.mel_floor), _ptr(np.ascontiguousarray(x)),
                                   C.c_int64(len(x)), _ptr(plan.window), _ptr(plan.mel_bank), _ptr(out))
    assert rc == 0
    if kind == 4:  # the reference's fp32 STFT vs this double-precision DFT: last-bits differences only
        np.testing.assert_allclose(out, y, rtol=0, atol=6e-5)
    else:          # float32 window table / float32 product here, float64 in librosa: same amplitude-floor gate as the GPU test
        from test_librosa import librosa_gate

        ok, msg = librosa_gate(out, y.astype(np.float64), c["cfg"])
        assert ok, msg


def test_c_oracle_center_rejects_short_inputs(lib):
    from lhotse_b200 import B200WhisperFbankConfig

    plan = build_plan("whisper-fbank", B200WhisperFbankConfig())
    out = np.empty((2, 80), dtype=np.float32)
    lib.oracle_extract_center.restype = C.c_int
    assert lib.oracle_extract_center(C.c_int32(4), C.c_int32(400), C.c_int32(160), C.c_int32(80), C.c_int32(0), C.c_float(1e-10),
                                     _ptr(np.zeros(200, dtype=np.float32)), C.c_int64(200), _ptr(plan.window),
                                     _ptr(plan.mel_bank), _ptr(out)) == -1
