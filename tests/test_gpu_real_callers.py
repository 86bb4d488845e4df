"""The CUDA path through the REAL lhotse callers, on the GPU box (SURVEY.md §8 rows a12-a14, f1, and BASELINE configs
[0], [3], [4] at test size): the reference package is the archive `oracle/_ref/lhotse_ref.zip` there
(`oracle/make_ref.py`), the extractor is a B200 extractor with its real engine, and every value is compared per cut
with the reference's own `Fbank().extract` run on the host (lhotse/features/kaldi/extractors.py:92-115).

Callers driven, unchanged: `CutSet.compute_and_store_features` (cut/set.py:1981), `compute_and_store_features_batch`
(:2197), `Cut.compute_features` (cut/base.py:335), `OnTheFlyFeatures.__call__` (dataset/input_strategies.py:410),
`K2SpeechRecognitionDataset.__getitem__` (dataset/speech_recognition.py:94), `DynamicBucketingSampler`
(dataset/sampling/dynamic_bucketing.py:48), `combine` (manipulation.py:18)."""
import numpy as np
import pytest
import torch

import refshim
from helpers import gate
from oracle import kaldi_oracle as O

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not refshim.reference_available(), reason="no reference tree / oracle/_ref archive")]


@pytest.fixture(scope="module")
def env():
    import lhotse_env

    lb_ex = lhotse_env.setup_lhotse()
    from lhotse.features.kaldi.extractors import Fbank, FbankConfig, Mfcc, MfccConfig

    return lb_ex, Fbank, FbankConfig, Mfcc, MfccConfig


def _mk(cls, cfg_cls, **kw):
    """The extractor under test: real engine on cuda:0.  B200_TEST_FAKE_ENGINE=1 swaps in the oracle-backed fake so that
    the test LOGIC can be dry-run in the build container (such a run proves nothing about the kernels)."""
    import os

    if os.environ.get("B200_TEST_FAKE_ENGINE") == "1":
        from helpers import attach_oracle_engine

        return attach_oracle_engine(cls(cfg_cls(**kw)))
    ext = cls(cfg_cls(device="cuda:0", **kw))
    assert ext.engine.kernel in ("fast", "tc", "generic")
    return ext


def _check(got, audio, ref_ext, feature="fbank", ocfg=None):
    """`got` against the reference extractor's output on the same samples (fp32) and the float64 truth."""
    ref = ref_ext.extract(audio, 16000)
    assert got.shape == ref.shape and got.dtype == np.float32, (got.shape, ref.shape)
    truth = O.extract(audio, ocfg or O.OracleConfig(feature=feature), dtype=torch.float64)
    ok, msg = gate(got, ref, truth, feature)
    assert ok, msg


def test_config0_compute_and_store_features(env, tmp_path):
    """BASELINE configs[0] on the CUDA path: 100 synthetic 1 s MonoCuts through CutSet.compute_and_store_features."""
    import lhotse_env
    from lhotse.features.io import NumpyFilesWriter

    lb_ex, Fbank, FbankConfig, _, _ = env
    cuts = lhotse_env.make_cutset(tmp_path, [1.0] * 100, supervisions=False)
    ext = _mk(lb_ex.B200Fbank, lb_ex.B200FbankConfig, num_mel_bins=80)
    out = cuts.compute_and_store_features(extractor=ext, storage_path=tmp_path / "feats", num_jobs=1, storage_type=NumpyFilesWriter)
    ref_ext = Fbank(FbankConfig(num_mel_bins=80))
    n = 0
    for cut in out:
        assert cut.features.type == "b200-fbank" and (cut.num_frames, cut.num_features) == (100, 80)
        _check(cut.load_features(), cut.load_audio()[0], ref_ext)
        n += 1
    assert n == 100
    if ext.engine.kernel != "oracle":
        assert ext.engine.stats()["kernel_launches"] >= 100
    # Cut.compute_features (cut/base.py:335)
    c0 = next(iter(cuts))
    _check(c0.compute_features(ext), c0.load_audio()[0], ref_ext)


def test_compute_and_store_features_batch_and_mfcc(env, tmp_path):
    import lhotse_env
    from lhotse.features.io import NumpyFilesWriter

    lb_ex, Fbank, FbankConfig, Mfcc, MfccConfig = env
    cuts = lhotse_env.make_cutset(tmp_path, [1.0, 2.5, 1.7, 3.0, 0.8, 2.0, 1.2, 10.0, 4.4], supervisions=False)
    ext = _mk(lb_ex.B200Fbank, lb_ex.B200FbankConfig)
    out = cuts.compute_and_store_features_batch(ext, tmp_path / "fb", num_workers=0, batch_duration=6.0, storage_type=NumpyFilesWriter,
                                                overwrite=True)
    ref_ext = Fbank()
    for cut in out:
        assert cut.features.type == "b200-fbank"
        _check(cut.load_features(), cut.load_audio()[0], ref_ext)
    # BASELINE configs[2] through the same caller: Mfcc(num_ceps=13, num_mel_bins=23)
    mext = _mk(lb_ex.B200Mfcc, lb_ex.B200MfccConfig, num_ceps=13, num_mel_bins=23)
    mout = cuts.compute_and_store_features_batch(mext, tmp_path / "mf", num_workers=0, batch_duration=6.0, storage_type=NumpyFilesWriter,
                                                 overwrite=True)
    mref = Mfcc(MfccConfig(num_ceps=13, num_mel_bins=23))
    for cut in mout:
        got, audio = cut.load_features(), cut.load_audio()[0]
        ref = mref.extract(audio, 16000)
        assert got.shape == ref.shape == (cut.num_frames, 13)
        np.testing.assert_allclose(got, ref, rtol=1e-3, atol=1e-3)  # the parity report holds the achieved figures per case


def test_config3_on_the_fly_in_k2_dataset_with_dynamic_bucketing(env, tmp_path):
    """BASELINE configs[3] at test size: mixed 2-30 s cuts, DynamicBucketingSampler -> K2SpeechRecognitionDataset with
    OnTheFlyFeatures(B200Fbank) and with FusedOnTheFlyFeatures(B200Fbank); per-cut comparison with the reference Fbank."""
    import lhotse_env
    from lhotse.dataset import DynamicBucketingSampler, K2SpeechRecognitionDataset
    from lhotse.dataset.input_strategies import OnTheFlyFeatures

    from lhotse_b200 import LOG_EPSILON
    from lhotse_b200.input_strategies import FusedOnTheFlyFeatures

    lb_ex, Fbank, _, _, _ = env
    rs = np.random.RandomState(0)
    cuts = lhotse_env.make_cutset(tmp_path, np.round(rs.uniform(2.0, 30.0, size=40), 2).tolist(), seed=1)
    by_id = {c.id: c for c in cuts}
    ext = _mk(lb_ex.B200Fbank, lb_ex.B200FbankConfig)
    ref_ext = Fbank()
    seen = {}
    for name, strategy in (("reference-strategy", OnTheFlyFeatures(ext)), ("fused", FusedOnTheFlyFeatures(ext))):
        sampler = DynamicBucketingSampler(cuts, max_duration=200.0, num_buckets=5, shuffle=True, seed=0)
        ds = K2SpeechRecognitionDataset(input_strategy=strategy, return_cuts=True)
        count = 0
        for batch_cuts in sampler:
            batch = ds[batch_cuts]
            x, sup = batch["inputs"], batch["supervisions"]
            assert x.dim() == 3 and x.shape[2] == 80 and x.dtype == torch.float32
            xs = x.cpu().numpy()
            for i, cut in enumerate(sup["cut"]):  # one supervision per cut, spanning it
                row = int(sup["sequence_idx"][i])
                T = (by_id[cut.id].num_samples + 80) // 160
                assert int(sup["start_frame"][i]) == 0 and abs(int(sup["num_frames"][i]) - T) <= 1
                assert np.all(xs[row, T:] == np.float32(LOG_EPSILON))
                if name == "fused" or cut.id not in seen:
                    _check(xs[row, :T], cut.load_audio()[0], ref_ext)
                if name == "reference-strategy":
                    seen[cut.id] = xs[row, :T].copy()
                else:
                    assert np.array_equal(seen[cut.id], xs[row, :T])  # fused collation == the reference's collation, bit for bit
                count += 1
        assert count == 40


def test_config4_rank_sharded_cutset_fused_store(env, tmp_path):
    """BASELINE configs[4] at test size, on one GPU: the rank::world split of a lazy CutSet through
    dist.compute_and_store_features_sharded(fused=True), per-rank archive + manifest, combine_shards."""
    import lhotse_env
    from lhotse import CutSet

    from lhotse_b200 import dist as lbd

    lb_ex, Fbank, _, _, _ = env
    cuts = lhotse_env.make_cutset(tmp_path, [10.0] * 6 + [3.3, 7.7, 12.1, 1.0], supervisions=False, seed=3)
    man = tmp_path / "cuts.jsonl.gz"
    cuts.to_file(man)
    lazy = CutSet.from_jsonl_lazy(man)
    ext = _mk(lb_ex.B200Fbank, lb_ex.B200FbankConfig)
    out = tmp_path / "sharded"
    for r in range(2):
        mine = lbd.compute_and_store_features_sharded(lazy, ext, out, rank=r, world=2, num_workers=0, batch_duration=30.0, fused=True,
                                                      overwrite=True)
        assert [c.id for c in mine] == [c.id for c in cuts][r::2]
    allc = lbd.combine_shards(out, world=2)
    assert [c.id for c in allc] == [c.id for c in cuts]
    ref_ext = Fbank()
    for c in allc:
        assert c.has_features and c.features.storage_type == "b200_archive" and c.features.type == "b200-fbank"
        _check(c.load_features(), c.load_audio()[0], ref_ext)


def test_fused_global_mvn_matches_the_reference_transform(env, tmp_path):
    """§8f-1: GlobalMVN (lhotse/dataset/signal_transforms.py:16-58) fused into the kernels' epilogue: FusedOnTheFlyFeatures(
    global_mvn=...) against the reference module applied to the reference strategy's collated batch (padding included)."""
    import lhotse_env
    from lhotse.dataset.input_strategies import OnTheFlyFeatures
    from lhotse.dataset.signal_transforms import GlobalMVN

    from lhotse_b200.input_strategies import FusedOnTheFlyFeatures

    lb_ex, _, _, _, _ = env
    cuts = lhotse_env.make_cutset(tmp_path, [1.0, 2.5, 1.7, 3.0, 0.8], seed=7)
    ext = _mk(lb_ex.B200Fbank, lb_ex.B200FbankConfig)
    if ext.engine.kernel == "oracle":
        pytest.skip("needs the CUDA engine (output affine lives in the kernels)")
    mvn = GlobalMVN(80)
    rs = np.random.RandomState(1)
    mvn.norm_means.copy_(torch.from_numpy(rs.uniform(-12, -4, 80).astype(np.float32)))
    mvn.norm_stds.copy_(torch.from_numpy(rs.uniform(0.5, 4.0, 80).astype(np.float32)))
    plain, lens = OnTheFlyFeatures(ext)(cuts)
    want = mvn(plain.cpu())
    for kernel in ("fast", "tc", "generic"):
        e2 = lb_ex.B200Fbank(lb_ex.B200FbankConfig(device="cuda:0", kernel=kernel))
        got, glens = FusedOnTheFlyFeatures(e2, global_mvn=mvn)(cuts)
        assert torch.equal(glens, lens) and got.shape == want.shape
        ref = mvn(OnTheFlyFeatures(e2)(cuts)[0].cpu())  # the same kernel without the fused affine, then the reference module
        torch.testing.assert_close(got.cpu(), ref, rtol=1e-5, atol=1e-5)
        assert torch.allclose(got.cpu(), want, rtol=1e-4, atol=2e-3)
    # the plain extractor is untouched by the strategy's private normalising handle
    again, _ = OnTheFlyFeatures(ext)(cuts)
    assert torch.equal(again, plain)
