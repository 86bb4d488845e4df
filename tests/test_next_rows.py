"""§8f 'next' rows, host logic through the REAL lhotse callers (build container only): the fused
OnTheFlyFeatures replacement and the rank-sharded CutSet extraction.  Numeric engine = oracle-backed fake."""
import wave

import numpy as np
import pytest
import torch

import refshim

pytestmark = [pytest.mark.reference, pytest.mark.skipif(not refshim.reference_available(), reason="reference tree not present")]


@pytest.fixture(scope="module")
def env(tmp_path_factory):
    refshim.import_reference()
    import importlib

    import lhotse_b200.base as lb_base
    import lhotse_b200.extractors as lb_ex

    importlib.reload(lb_base)
    importlib.reload(lb_ex)
    import lhotse_b200.storage as lb_st

    importlib.reload(lb_st)  # subclass lhotse's FeaturesWriter / FeaturesReader now that they are importable
    from lhotse import CutSet, MonoCut, Recording, SupervisionSegment
    from lhotse.audio import AudioSource
    from lhotse.audio.backend import AudioBackend, LibsndfileCompatibleAudioInfo, set_current_audio_backend

    class WaveBackend(AudioBackend):
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
    root = tmp_path_factory.mktemp("cuts")
    rs = np.random.RandomState(0)
    cuts = []
    for i, dur in enumerate((1.0, 2.5, 1.7, 3.0, 0.8, 2.0, 1.2)):
        n = int(dur * 16000)
        path = root / f"c{i}.wav"
        with wave.open(str(path), "wb") as w:
            w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000)
            w.writeframes(np.clip(rs.randn(n) * 3000, -32768, 32767).astype("<i2").tobytes())
        rec = Recording(id=f"r{i}", sources=[AudioSource(type="file", channels=[0], source=str(path))],
                        sampling_rate=16000, num_samples=n, duration=n / 16000)
        sup = SupervisionSegment(id=f"s{i}", recording_id=f"r{i}", start=0.1, duration=dur - 0.2, text="x")
        cuts.append(MonoCut(id=f"c{i}", start=0.0, duration=n / 16000, channel=0, recording=rec, supervisions=[sup]))
    return CutSet.from_cuts(cuts), lb_ex, root


def test_fused_on_the_fly_matches_reference_strategy(env):
    from helpers import attach_oracle_engine
    from lhotse.dataset.input_strategies import OnTheFlyFeatures

    from lhotse_b200.input_strategies import FusedOnTheFlyFeatures

    cuts, lb_ex, _ = env
    ext = attach_oracle_engine(lb_ex.B200Fbank())
    ref_feats, ref_lens = OnTheFlyFeatures(ext)(cuts)
    fused = FusedOnTheFlyFeatures(ext, return_audio=True)
    feats, lens, audios, audio_lens = fused(cuts)
    assert torch.equal(lens, ref_lens) and lens.dtype == torch.int64
    assert feats.shape == ref_feats.shape and torch.equal(feats, ref_feats)  # same values, same LOG_EPSILON padding
    assert audios.shape == (7, 48000) and audio_lens.tolist() == [c.num_samples for c in cuts]
    a = fused.supervision_intervals(cuts)
    b = OnTheFlyFeatures(ext).supervision_intervals(cuts)
    assert all(torch.equal(a[k], b[k]) for k in b)
    # and inside the reference's dataset class (speech_recognition.py:94-116)
    from lhotse.dataset import K2SpeechRecognitionDataset

    batch = K2SpeechRecognitionDataset(input_strategy=FusedOnTheFlyFeatures(ext))[cuts]
    assert batch["inputs"].shape[0] == 7 and batch["inputs"].shape[2] == 80


def test_rank_sharded_extraction_roundtrip(env):
    from helpers import attach_oracle_engine
    from lhotse.features.io import NumpyFilesWriter

    from lhotse_b200 import dist as lbd

    cuts, lb_ex, root = env
    ext = attach_oracle_engine(lb_ex.B200Fbank())
    out = root / "sharded"
    for r in range(2):
        mine = lbd.compute_and_store_features_sharded(cuts, ext, out, rank=r, world=2, num_workers=0,
                                                      storage_type=NumpyFilesWriter, batch_duration=4.0)
        assert [c.id for c in mine] == [c.id for c in cuts][r::2]
    allc = lbd.combine_shards(out, world=2)
    assert [c.id for c in allc] == [c.id for c in cuts]
    # the same sharding over the fused store (one archive per rank)
    out2 = root / "sharded_fused"
    for r in range(2):
        lbd.compute_and_store_features_sharded(cuts, ext, out2, rank=r, world=2, num_workers=0, batch_duration=4.0,
                                               fused=True, overwrite=True)
    fused = lbd.combine_shards(out2, world=2)
    assert [c.id for c in fused] == [c.id for c in cuts]
    for a, b in zip(fused, allc):
        assert a.features.storage_type == "b200_archive" and np.array_equal(a.load_features(), b.load_features())
    for c in allc:
        assert c.has_features and c.features.type == "b200-fbank"
        f = c.load_features()
        assert f.shape == (c.num_frames, 80)


def test_archive_backend_and_fused_store(env, tmp_path):
    """§8f-3: the b200_archive backend behind lhotse's reader/writer registries, and compute_and_store_features_fused
    against the reference's own compute_and_store_features_batch on the same cuts."""
    from helpers import attach_oracle_engine
    from lhotse import CutSet
    from lhotse.features.io import NumpyFilesWriter, get_reader, get_writer

    from lhotse_b200.storage import B200ArchiveReader, B200ArchiveWriter, compute_and_store_features_fused

    assert get_writer("b200_archive") is B200ArchiveWriter and get_reader("b200_archive") is B200ArchiveReader
    # backend contract: write / write_batch / partial reads / append
    rs = np.random.RandomState(0)
    a, b, c = (rs.randn(t, 80).astype(np.float32) for t in (100, 37, 250))
    with B200ArchiveWriter(tmp_path / "arch") as w:
        ka = w.write("a", a)
        kb, kc = w.write_batch(["b", "c"], np.concatenate([b, c]), [0, 37, 287])
        assert w.storage_path.endswith(".b200feat")
    r = B200ArchiveReader(tmp_path / "arch")
    assert np.array_equal(r.read(ka), a) and np.array_equal(r.read(kb), b) and np.array_equal(r.read(kc), c)
    assert np.array_equal(r.read(kc, left_offset_frames=10, right_offset_frames=60), c[10:60])
    assert r.read(kc, left_offset_frames=250).shape == (0, 80)
    with B200ArchiveWriter(tmp_path / "arch", mode="a") as w:
        kd = w.write("d", a[:5])
    r2 = B200ArchiveReader(tmp_path / "arch")
    assert np.array_equal(r2.read(kd), a[:5]) and np.array_equal(r2.read(ka), a)  # old keys survive an append
    ta = w.__class__(tmp_path / "arr").store_array("k", b, frame_shift=0.01, temporal_dim=0)  # lhotse's Array manifests
    assert np.array_equal(ta.load(), b)

    cuts, lb_ex, root = env
    ext = attach_oracle_engine(lb_ex.B200Fbank())
    want = cuts.compute_and_store_features_batch(ext, root / "ref_store", num_workers=0, batch_duration=4.0,
                                                 storage_type=NumpyFilesWriter, overwrite=True)
    for route in (True, False):
        got = compute_and_store_features_fused(cuts, ext, root / f"fused_{route}", manifest_path=root / f"fused_{route}.jsonl.gz",
                                               batch_duration=4.0, num_workers=2, overwrite=True, pcm16_fast_path=route)
        got = CutSet.from_cuts(list(got))
        assert [c_.id for c_ in got] == [c_.id for c_ in want]
        for g, w_ in zip(got, want):
            assert g.has_features and g.features.storage_type == "b200_archive" and g.features.type == "b200-fbank"
            assert (g.num_frames, g.num_features, g.features.recording_id) == (w_.num_frames, w_.num_features, w_.features.recording_id)
            assert np.array_equal(g.load_features(), w_.load_features())
            part = g.features.load(start=g.start + 0.2, duration=0.3)
            assert np.array_equal(part, w_.features.load(start=w_.start + 0.2, duration=0.3))
    # no manifest path: everything stays in memory (set.py:2290-2294), eager CutSet back
    mem = compute_and_store_features_fused(cuts, ext, root / "fused_mem", batch_duration=100.0, num_workers=0, overwrite=True)
    assert [c_.id for c_ in mem] == [c_.id for c_ in want] and all(np.array_equal(a_.load_features(), b_.load_features()) for a_, b_ in zip(mem, want))
    # resumable like the reference: a second call with the same manifest does not recompute anything
    calls = []
    orig = ext.extract_batch_packed
    ext.extract_batch_packed = lambda *a_, **k_: calls.append(1) or orig(*a_, **k_)
    again = compute_and_store_features_fused(cuts, ext, root / "fused_False", manifest_path=root / "fused_False.jsonl.gz",
                                             batch_duration=4.0, num_workers=0, pcm16_fast_path=False)
    assert len(list(again)) == len(cuts) and not calls


def test_fused_store_manifest_lines_pipeline_and_failures(env, tmp_path, monkeypatch):
    """The fused store writes its manifest lines itself (one write per batch): every line must be the JSON of the cut's own
    `to_dict()`; the three-stage pipeline and the sequential mode give byte-identical manifests and archives; an exception in
    any stage surfaces in the caller and leaves no thread behind."""
    import gzip
    import json
    import threading

    from helpers import attach_oracle_engine
    from lhotse import Recording
    from lhotse.features.base import Features
    from lhotse.utils import fastcopy

    from lhotse_b200.storage import _ManifestLines, compute_and_store_features_fused

    cuts, lb_ex, root = env
    lines = _ManifestLines()
    cl = list(cuts)
    variants = [cl[0], fastcopy(cl[1], custom={"dataloading_info": {"rank": 0, "world_size": 1, "worker_id": None}}),
                fastcopy(cl[2], custom={"note": "x", "nested": {"a": [1, 2.5, None]}}),
                fastcopy(cl[3], custom={"other": cl[3].recording})]  # a Recording in `custom`: the generic path
    for c in variants:
        fm = Features(start=c.start, duration=c.duration, type="b200-fbank", num_frames=123, num_features=80, frame_shift=0.01,
                      sampling_rate=c.sampling_rate, channels=c.channel, storage_type="b200_archive", storage_path="/a/b.b200feat",
                      storage_key="0,123,80", recording_id=c.recording_id)
        want = fastcopy(c, features=fm).to_dict()
        got = lines.cut_dict(c, fm)
        assert got == want and list(got) == list(want) and json.dumps(got) == json.dumps(want)
        assert list(got["features"]) == list(want["features"])

    ext = attach_oracle_engine(lb_ex.B200Fbank())
    outs = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("B200FEAT_STORE_PIPELINE", mode)
        compute_and_store_features_fused(cuts, ext, root / f"pipe{mode}" / "store", manifest_path=root / f"pipe{mode}.jsonl.gz",
                                         batch_duration=4.0, num_workers=2, overwrite=True)
        with gzip.open(root / f"pipe{mode}.jsonl.gz", "rt") as f:
            outs[mode] = f.read().replace(f"pipe{mode}", "pipeX")
        assert (root / f"pipe{mode}" / "store.b200feat").read_bytes() == (root / "pipe0" / "store.b200feat").read_bytes()
    assert outs["0"] == outs["1"] and outs["0"].count("\n") == len(cl)

    class Boom(RuntimeError):
        pass

    for mode in ("0", "1"):
        monkeypatch.setenv("B200FEAT_STORE_PIPELINE", mode)
        n = [0]
        orig = ext.extract_staged_packed

        def flaky(*a, **k):
            n[0] += 1
            if n[0] == 2:
                raise Boom("second batch")
            return orig(*a, **k)

        ext.extract_staged_packed = flaky
        before = threading.active_count()
        try:
            import pytest as _pt

            with _pt.raises(Boom):
                compute_and_store_features_fused(cuts, ext, root / f"boom{mode}", manifest_path=root / f"boom{mode}.jsonl.gz",
                                                 batch_duration=4.0, num_workers=2, overwrite=True)
        finally:
            ext.extract_staged_packed = orig
        assert threading.active_count() <= before
        # what reached the manifest before the failure is loadable and resumable
        done = compute_and_store_features_fused(cuts, ext, root / f"boom{mode}", manifest_path=root / f"boom{mode}.jsonl.gz",
                                                batch_duration=4.0, num_workers=2)
        assert [c.id for c in done] == [c.id for c in cl]
        assert all(c.load_features().shape == (c.num_frames, 80) for c in done)


def test_fused_store_helper_threads_bind_the_callers_gpu(env, monkeypatch):
    """The reader / writer threads of the pipelined store select the GPU the extractor works on — resolved in the calling thread,
    because the default config says "cuda" without an index and `torch.cuda.set_device` refuses an index-less device."""
    import threading

    from helpers import attach_oracle_engine

    from lhotse_b200.storage import compute_and_store_features_fused

    cuts, lb_ex, root = env
    for device, current, want in (("cuda", 3, 3), ("cuda:5", 3, 5), ("cpu", 3, None)):
        ext = attach_oracle_engine(lb_ex.B200Fbank())
        ext.config.device = device
        calls = []

        class _Cuda:  # what storage.py sees as torch.cuda (the rest of the process keeps the real, GPU-less torch)
            is_available = staticmethod(lambda: True)
            current_device = staticmethod(lambda: current)
            set_device = staticmethod(lambda d: calls.append((threading.current_thread().name, d)))

        class _Torch:
            cuda = _Cuda

            def __getattr__(self, name):
                return getattr(torch, name)

        import lhotse_b200.storage as st

        monkeypatch.setattr(st, "torch", _Torch())
        monkeypatch.setenv("B200FEAT_STORE_PIPELINE", "1")
        out = compute_and_store_features_fused(cuts, ext, root / f"bind_{device.replace(':', '_')}", batch_duration=4.0, num_workers=0,
                                               overwrite=True, pcm16_fast_path=False)
        assert len(list(out)) == len(list(cuts))
        if want is None:
            assert calls == []
        else:
            assert sorted(calls) == [("b200feat-reader", want), ("b200feat-writer", want)]
        monkeypatch.undo()


def test_fused_on_the_fly_mixed_sampling_rates_and_family_adapters(env, tmp_path):
    """`use_batch_extract=False` (reference: sequential `extract` so that sampling rates may differ, input_strategies.py:447-459):
    here one padded extraction per sampling rate, through an extractor that takes the rate per call (torchaudio family)."""
    import importlib

    from helpers import attach_oracle_engine
    from lhotse import CutSet, MonoCut, Recording
    from lhotse.audio import AudioSource
    from lhotse.dataset.input_strategies import OnTheFlyFeatures

    import lhotse_b200.families as fam
    from lhotse_b200.input_strategies import FusedOnTheFlyFeatures

    importlib.reload(fam)
    rs = np.random.RandomState(5)
    cuts = []
    for i, (sr, dur) in enumerate(((16000, 1.0), (8000, 1.5), (16000, 0.7), (8000, 0.5))):
        n = int(sr * dur)
        path = tmp_path / f"m{i}.wav"
        with wave.open(str(path), "wb") as w:
            w.setnchannels(1); w.setsampwidth(2); w.setframerate(sr)
            w.writeframes(np.clip(rs.randn(n) * 3000, -32768, 32767).astype("<i2").tobytes())
        rec = Recording(id=f"m{i}", sources=[AudioSource(type="file", channels=[0], source=str(path))],
                        sampling_rate=sr, num_samples=n, duration=n / sr)
        cuts.append(MonoCut(id=f"m{i}", start=0.0, duration=n / sr, channel=0, recording=rec))
    cs = CutSet.from_cuts(cuts)
    ext = fam.B200TorchaudioFbank()
    for sr in (8000, 16000):
        attach_oracle_engine(ext._inner(sr))
    want, want_lens = OnTheFlyFeatures(ext, use_batch_extract=False)(cs)          # the reference strategy, cut by cut
    fused = FusedOnTheFlyFeatures(ext, None, 0, False)                              # same positional arguments
    got, got_lens = fused(cs)
    assert fused.use_batch_extract is False and torch.equal(got_lens, want_lens) and got_lens.tolist() == [100, 150, 70, 50]
    assert got.shape == want.shape and torch.equal(got, want)
    with pytest.raises(AssertionError):
        FusedOnTheFlyFeatures(ext)(cs)                                              # default: one sampling rate per batch
    same = CutSet.from_cuts([cuts[0], cuts[2]])
    a, la = FusedOnTheFlyFeatures(ext)(same)                                        # family adapter on the single-rate route
    b, lb_ = OnTheFlyFeatures(ext)(same)
    assert torch.equal(la, lb_) and torch.equal(a, b)
