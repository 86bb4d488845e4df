"""Host-side behaviour of the extractors (container rules, registry, YAML, pickling) exercised on
CPU with the TEST-ONLY oracle-backed engine injected (tests/helpers.py::OracleEngine)."""
import pickle

import numpy as np
import pytest
import torch

import refshim
from helpers import attach_oracle_engine
from lhotse_b200 import (B200Fbank, B200FbankConfig, B200LogSpectrogram, B200Mfcc, B200MfccConfig, B200Spectrogram,
                         LOG_EPSILON)
from lhotse_b200.base import FeatureExtractor, get_extractor_type
from oracle import kaldi_oracle as O


def fb():
    return attach_oracle_engine(B200Fbank())


def test_extract_container_types():
    f = fb()
    x = (0.1 * np.random.RandomState(0).randn(16000)).astype(np.float32)
    y = f.extract(x, 16000)
    assert isinstance(y, np.ndarray) and y.shape == (100, 80) and y.dtype == np.float32
    y2 = f.extract(x[None, :], 16000)  # (1, n)
    assert np.array_equal(y, y2)
    y3 = f.extract(np.stack([x, -x]), 16000)  # multi-channel: channel 0 only (extractors.py:110)
    assert np.array_equal(y, y3)
    yt = f.extract(torch.from_numpy(x), 16000)
    assert isinstance(yt, torch.Tensor) and yt.shape == (100, 80)
    with pytest.raises(AssertionError):
        f.extract(x, 8000)


def test_extract_batch_rules():
    # mirrors test/features/test_kaldifeat_features.py:33-100 and extractors.py:539-554
    f = fb()
    rs = np.random.RandomState(1)
    a = (0.1 * rs.randn(8000)).astype(np.float32)
    b = (0.1 * rs.randn(8000)).astype(np.float32)
    c = (0.1 * rs.randn(12000)).astype(np.float32)
    assert f.extract_batch(a, 16000).shape == (50, 80)  # single 1-D array -> array
    r = f.extract_batch([a], 16000)  # one-item list -> list
    assert isinstance(r, list) and len(r) == 1 and r[0].shape == (50, 80)
    r = f.extract_batch([a, b], 16000)  # equal lengths -> stacked
    assert isinstance(r, np.ndarray) and r.shape == (2, 50, 80)
    r = f.extract_batch(np.stack([a, b]), 16000)  # 2-D array -> stacked
    assert r.shape == (2, 50, 80)
    r = f.extract_batch([a, c], 16000)  # uneven -> list
    assert isinstance(r, list) and [x.shape for x in r] == [(50, 80), (75, 80)]
    r = f.extract_batch([torch.from_numpy(a), torch.from_numpy(c)], 16000)
    assert isinstance(r, list) and all(isinstance(x, torch.Tensor) for x in r)
    r = f.extract_batch([torch.from_numpy(a), torch.from_numpy(b)], 16000)
    assert isinstance(r, torch.Tensor) and r.shape == (2, 50, 80)
    # every batch item equals extract() of that item (ragged semantics)
    r = f.extract_batch([a, c], 16000)
    assert np.array_equal(r[1], f.extract(c, 16000))


def test_extract_batch_with_lengths():
    f = fb()
    rs = np.random.RandomState(2)
    a = (0.1 * rs.randn(8000)).astype(np.float32)
    c = (0.1 * rs.randn(12000)).astype(np.float32)
    padded = torch.zeros(2, 12000)
    padded[0, :8000] = torch.from_numpy(a)
    padded[1] = torch.from_numpy(c)
    r = f.extract_batch(padded, 16000, lengths=torch.tensor([8000, 12000]))
    assert isinstance(r, list) and [tuple(x.shape) for x in r] == [(50, 80), (75, 80)]
    assert np.array_equal(r[0].numpy(), f.extract(a, 16000))
    with pytest.raises(AssertionError):
        f.extract_batch(padded.numpy(), 16000, lengths=[8000, 12000])


def test_padded_collation_mode():
    f = fb()
    rs = np.random.RandomState(3)
    xs = [torch.from_numpy((0.1 * rs.randn(n)).astype(np.float32)) for n in (4000, 8000, 6000)]
    feats, lens = f.extract_batch_padded(xs, 16000)
    assert feats.shape == (3, 50, 80) and lens.tolist() == [25, 50, 38] and lens.dtype == torch.int64
    assert torch.all(feats[0, 25:] == np.float32(LOG_EPSILON))


def test_registry_yaml_pickle(tmp_path):
    for cls, name in ((B200Fbank, "b200-fbank"), (B200Mfcc, "b200-mfcc"), (B200Spectrogram, "b200-spectrogram"),
                      (B200LogSpectrogram, "b200-log-spectrogram")):
        assert get_extractor_type(name) is cls
    f = B200Fbank(B200FbankConfig(num_filters=40, sampling_rate=8000))
    p = tmp_path / "fbank.yml"
    f.to_yaml(p)
    g = FeatureExtractor.from_yaml(p)
    assert type(g) is B200Fbank and g.config == f.config
    d = f.to_dict()
    assert d["feature_type"] == "b200-fbank" and "num_mel_bins" not in d
    h = pickle.loads(pickle.dumps(attach_oracle_engine(f)))  # ProcessPoolExecutor(spawn) path, set.py:2166
    assert h.config == f.config and h._engine is None
    assert f.frame_shift == 0.01 and f.feature_dim(8000) == 40 and B200Mfcc().feature_dim(16000) == 13
    assert B200Spectrogram().feature_dim(16000) == 257


def test_mix_energy_scale_statics():
    a, b = np.log(np.full((3, 4), 2.0)), np.log(np.full((3, 4), 3.0))
    assert np.allclose(B200Fbank.mix(a, b, 0.5), np.log(3.5))
    assert np.isclose(B200Fbank.compute_energy(a), 24.0)
    assert np.allclose(B200Fbank.scale(a, 2.0), np.log(4.0))
    assert np.allclose(B200Spectrogram.mix(np.ones(3), np.ones(3), 0.5), 1.5)
    with pytest.raises(ValueError):
        B200Mfcc.mix(a, b, 1.0)


def test_short_inputs_raise():
    f = fb()
    with pytest.raises(ValueError):
        f.extract(np.zeros(100, dtype=np.float32), 16000)


@pytest.mark.reference
@pytest.mark.skipif(not refshim.reference_available(), reason="reference tree not present")
def test_dropin_through_reference_callers(tmp_path):
    """CutSet.compute_and_store_features / Cut.compute_features / OnTheFlyFeatures accept the B200
    extractor unchanged (set.py:1981, cut/base.py:335, input_strategies.py:410).  The numeric engine is
    the oracle-backed fake here: this checks plumbing, not kernels."""
    refshim.import_reference()
    import importlib
    import wave

    import lhotse_b200.base as lb_base
    import lhotse_b200.extractors as lb_ex

    importlib.reload(lb_base)  # pick up the real lhotse FeatureExtractor now that it is importable
    importlib.reload(lb_ex)
    assert lb_base.HAVE_LHOTSE
    from lhotse import CutSet, MonoCut, Recording
    from lhotse.audio import AudioSource
    from lhotse.audio.backend import AudioBackend, LibsndfileCompatibleAudioInfo, set_current_audio_backend
    from lhotse.dataset.input_strategies import OnTheFlyFeatures
    from lhotse.features.base import FeatureExtractor as RefFE
    from lhotse.features.io import NumpyFilesWriter

    class WaveBackend(AudioBackend):  # stdlib-wave reader: no decoder library is installed
        def read_audio(self, path_or_fd, offset=0.0, duration=None, force_opus_sampling_rate=None):
            with wave.open(str(path_or_fd)) as w:
                sr = w.getframerate()
                w.setpos(int(round(offset * sr)))
                n = w.getnframes() - w.tell() if duration is None else int(round(duration * sr))
                pcm = np.frombuffer(w.readframes(n), dtype="<i2")
            return (pcm.astype(np.float32) / 32768.0)[None, :], sr

        def is_applicable(self, path_or_fd):
            return True

        def supports_info(self):
            return True

        def info(self, path_or_fd, force_opus_sampling_rate=None, force_read_audio=False):
            with wave.open(str(path_or_fd)) as w:
                return LibsndfileCompatibleAudioInfo(channels=1, frames=w.getnframes(), samplerate=w.getframerate(),
                                                     duration=w.getnframes() / w.getframerate())

    set_current_audio_backend(WaveBackend())
    rs = np.random.RandomState(0)
    cuts = []
    NCUTS = 100
    for i in range(NCUTS):  # BASELINE configs[0]: 100 synthetic 1 s MonoCuts @ 16 kHz through compute_and_store_features
        path = tmp_path / f"c{i}.wav"
        pcm = np.clip(rs.randn(16000) * 0.1 * 32768, -32768, 32767).astype("<i2")
        with wave.open(str(path), "wb") as w:
            w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000); w.writeframes(pcm.tobytes())
        rec = Recording(id=f"r{i}", sources=[AudioSource(type="file", channels=[0], source=str(path))],
                        sampling_rate=16000, num_samples=16000, duration=1.0)
        cuts.append(MonoCut(id=f"c{i}", start=0.0, duration=1.0, channel=0, recording=rec))
    cs = CutSet.from_cuts(cuts)
    ext = attach_oracle_engine(lb_ex.B200Fbank())
    assert isinstance(ext, RefFE)
    out = cs.compute_and_store_features(extractor=ext, storage_path=tmp_path / "feats", num_jobs=1,
                                        storage_type=NumpyFilesWriter)
    for cut in out:
        assert cut.features.type == "b200-fbank" and cut.num_frames == 100 and cut.num_features == 80
        feats = cut.load_features()
        assert feats.shape == (100, 80)
        want = O.extract(cut.load_audio()[0], O.OracleConfig())
        assert np.array_equal(feats, want)
    assert cuts[0].compute_features(ext).shape == (100, 80)
    assert len(list(out)) == NCUTS
    feats, lens = OnTheFlyFeatures(ext)(cs)
    assert feats.shape == (NCUTS, 100, 80) and lens.tolist() == [100] * NCUTS
    # registry round trip through the reference's own from_dict
    again = RefFE.from_dict(ext.to_dict())
    assert type(again).__name__ == "B200Fbank"


def test_online_inference_host_logic():
    """Streaming wrapper (buffer/remainder/frame-count logic of layers.py:775-857) over the fake engine, against the
    reference's own streaming runs (tests/golden/make_golden_stream.py)."""
    from helpers import load_golden_stream
    from lhotse_b200 import B200LogSpectrogramConfig, B200SpectrogramConfig
    types = {"fbank": (B200Fbank, B200FbankConfig), "mfcc": (B200Mfcc, B200MfccConfig),
             "spectrogram": (B200Spectrogram, B200SpectrogramConfig),
             "log-spectrogram": (B200LogSpectrogram, B200LogSpectrogramConfig)}
    for i, m, x, y, r in load_golden_stream():
        cls, ccls = types[m["feature"]]
        ext = attach_oracle_engine(cls(ccls(**m["cfg"])))
        xb = torch.from_numpy(np.stack([x, -x]))  # batch of two: the negated signal has the same power spectrum
        rem, feats, counts = None, [], []
        for a, b in zip(m["bounds"][:-1], m["bounds"][1:]):
            f, rem = ext.online_inference(xb[:, a:b], context=rem)
            assert f.dim() == 3 and f.shape[0] == 2
            feats.append(f)
            counts.append(f.shape[1])
        assert counts == m["counts"]
        assert np.array_equal(rem[0].numpy(), r) and np.array_equal(rem[1].numpy(), -r)
        got = torch.cat(feats, dim=1).numpy()
        np.testing.assert_allclose(got[0], y, rtol=1e-4, atol=1e-3 if m["feature"] in ("mfcc", "spectrogram") else 1e-4)
    # a buffer too short for one frame: nothing emitted, everything carried
    ext = attach_oracle_engine(B200Fbank())
    f, rem = ext.online_inference(torch.zeros(1, 100))
    assert f.shape == (1, 0, 80) and rem.shape == (1, 100 + 100)  # 120-sample reflection is cut to the chunk: 100 + 100
