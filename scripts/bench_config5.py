#!/usr/bin/env python
"""BASELINE configs[4] at CutSet level: a synthetic lazy CutSet of 10 s cuts over PCM16 WAV recordings, sharded rank::world
over the GPUs of one box through `lhotse_b200.dist.compute_and_store_features_sharded(fused=True)` (the reference's
`LazySlicer` job split, cut/set.py:2156-2160), one `b200_archive` + manifest per rank, rank 0 `combine_shards`
(manipulation.py:18) and a spot check against a direct extraction.  Reports whole-job hours of audio per second INCLUDING
file reads, PCM staging, H2D, kernel, D2H, archive append and manifest writing.

    python scripts/bench_config5.py --hours-per-rank 1                       # one GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/bench_config5.py --hours-per-rank 2

`bench.py` calls `run_cutset_job` for its `extra.config4_cutset_store` figure, so the driver's 1/2/4/8-GPU scaling run
records it at every N.  lhotse itself comes from the installed package when there is one, else from the reference archive
`oracle/_ref/lhotse_ref.zip` (oracle/make_ref.py) — the callers, manifests and samplers are the reference's own code."""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time
import wave

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SR = 16000


def _ensure_lhotse():
    try:
        import lhotse  # noqa: F401
    except Exception:
        from oracle import refimport

        refimport.import_reference()
    import importlib

    import lhotse_b200.base as lb_base

    if not lb_base.HAVE_LHOTSE:  # lhotse_b200 was imported before lhotse became importable: re-bind its base classes
        import lhotse_b200.extractors as lb_ex
        import lhotse_b200.storage as lb_st

        importlib.reload(lb_base)
        importlib.reload(lb_ex)
        importlib.reload(lb_st)
    import lhotse_b200.extractors as lb_ex

    return lb_ex


def _write_recording(path, seconds, seed):
    import numpy as np

    rs = np.random.RandomState(seed)
    n = int(seconds * SR)
    with wave.open(str(path), "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(SR)
        step = 60 * SR
        for i in range(0, n, step):  # N(0, 0.1^2) noise, one minute at a time
            m = min(step, n - i)
            w.writeframes(np.clip(rs.randn(m) * 0.1 * 32768, -32768, 32767).astype("<i2").tobytes())
    return n


def run_cutset_job(rank, world, local, seconds_of_audio=3600.0, cut_seconds=10.0, num_workers=8, batch_duration=2000.0, keep=False):
    import numpy as np
    import torch

    from lhotse_b200 import dist as lbd

    lb_ex = _ensure_lhotse()
    from lhotse import CutSet, MonoCut, Recording
    from lhotse.audio import AudioSource

    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else tempfile.gettempdir()
    root = os.path.join(base, f"b200feat_config5_{os.environ.get('MASTER_PORT', '0')}_{os.getppid() if world > 1 else os.getpid()}")
    os.makedirs(root, exist_ok=True)
    nper = int(seconds_of_audio // cut_seconds)
    rec_seconds = nper * cut_seconds
    # every rank writes one recording; the manifest (identical on every rank) lists the cuts of ALL recordings round-robin,
    # so that rank r's shard (cuts r::world) is spread over the files like a real corpus
    _write_recording(os.path.join(root, f"rec{rank}.wav"), rec_seconds, seed=100 + rank)
    lbd.barrier()
    recs = [Recording(id=f"rec{r}", sources=[AudioSource(type="file", channels=[0], source=os.path.join(root, f"rec{r}.wav"))],
                      sampling_rate=SR, num_samples=int(rec_seconds * SR), duration=rec_seconds) for r in range(world)]
    man = os.path.join(root, f"cuts-in-{rank}.jsonl.gz")
    CutSet.from_cuts(MonoCut(id=f"c{i:07d}", start=(i // world) * cut_seconds, duration=cut_seconds, channel=0, recording=recs[(i + i // world) % world])
                     for i in range(nper * world)).to_file(man)
    lazy = CutSet.from_jsonl_lazy(man)
    ext = lb_ex.B200Fbank(lb_ex.B200FbankConfig(device=f"cuda:{local}"))
    ext.engine  # handle + tables before the clock starts
    out = os.path.join(root, "store")
    torch.cuda.synchronize()
    lbd.barrier()
    t0 = time.perf_counter()
    mine = lbd.compute_and_store_features_sharded(lazy, ext, out, rank=rank, world=world, num_workers=num_workers, batch_duration=batch_duration,
                                                  fused=True, overwrite=True)
    n_mine = sum(1 for _ in mine)
    mine_s = time.perf_counter() - t0
    lbd.barrier()
    wall = lbd.all_reduce_stats([time.perf_counter() - t0], "max")[0]
    res = {"unit": "h_audio/s", "n_gpus": min(world, torch.cuda.device_count()), "ranks": world, "cuts": nper * world, "cut_seconds": cut_seconds, "hours_of_audio": nper * world * cut_seconds / 3600.0,
           "wall_s": wall, "value": nper * world * cut_seconds / 3600.0 / wall, "rank0_s": mine_s, "num_workers": num_workers,
           "batch_duration_s": batch_duration, "kernel": ext.engine.kernel, "storage": "b200_archive on " + base,
           "note": "lazy CutSet -> rank::world shard -> compute_and_store_features_sharded(fused=True): PCM16 file reads into a pinned ring, H2D, "
                   "kernel, D2H, archive append, manifest; wall clock, max over ranks"}
    assert n_mine == nper
    if rank == 0:
        allc = lbd.combine_shards(out, world)
        ids = [c.id for c in allc]
        assert ids == [f"c{i:07d}" for i in range(nper * world)], "combine_shards must restore the corpus order"
        probe = allc[len(ids) // 2]
        with wave.open(probe.recording.sources[0].source) as w:
            w.setpos(int(round(probe.start * SR)))
            pcm = np.frombuffer(w.readframes(int(round(probe.duration * SR))), dtype="<i2")
        direct = ext.extract(pcm.astype(np.float32) / 32768.0, SR)
        stored = probe.load_features()
        res["spot_check_max_abs_diff"] = float(np.abs(stored - direct).max())
        assert stored.shape == direct.shape and res["spot_check_max_abs_diff"] == 0.0, "archive round trip must be bit-exact"
    lbd.barrier()
    if not keep and rank == 0:
        shutil.rmtree(root, ignore_errors=True)
    return res


def run_onthefly_job(local, seconds_of_audio=2400.0, max_duration=600.0, num_buckets=10, num_workers=4):
    """BASELINE configs[3] as a throughput figure: cuts of U[2, 30] s (seed 0) over one PCM16 WAV recording on tmpfs,
    `DynamicBucketingSampler` (dataset/sampling/dynamic_bucketing.py:48) -> `K2SpeechRecognitionDataset.__getitem__`
    (dataset/speech_recognition.py:94) with `FusedOnTheFlyFeatures(B200Fbank)`: every batch is read (PCM16 -> pinned ring), sent,
    extracted and collated to a padded (B, T_max, 80) device tensor by one launch.  Wall clock over the whole epoch."""
    import numpy as np
    import torch

    lb_ex = _ensure_lhotse()
    from lhotse import CutSet, MonoCut, Recording, SupervisionSegment
    from lhotse.audio import AudioSource
    from lhotse.dataset import DynamicBucketingSampler, K2SpeechRecognitionDataset

    from lhotse_b200.input_strategies import FusedOnTheFlyFeatures

    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else tempfile.gettempdir()
    root = os.path.join(base, f"b200feat_config4_{os.getpid()}")
    os.makedirs(root, exist_ok=True)
    try:
        wav = os.path.join(root, "rec.wav")
        n = _write_recording(wav, seconds_of_audio, seed=7)
        rec = Recording(id="rec", sources=[AudioSource(type="file", channels=[0], source=wav)], sampling_rate=SR, num_samples=n,
                        duration=n / SR)
        rs = np.random.RandomState(0)
        cuts, t, i = [], 0.0, 0
        while True:
            d = float(np.round(rs.uniform(2.0, 30.0), 2))
            if t + d > seconds_of_audio:
                break
            cuts.append(MonoCut(id=f"u{i:06d}", start=t, duration=d, channel=0, recording=rec,
                                supervisions=[SupervisionSegment(id=f"s{i:06d}", recording_id="rec", start=0.0, duration=d, text="x")]))
            t += d; i += 1
        cs = CutSet.from_cuts(cuts)
        ext = lb_ex.B200Fbank(lb_ex.B200FbankConfig(device=f"cuda:{local}"))
        ext.engine
        strategy = FusedOnTheFlyFeatures(ext, num_workers=num_workers)
        ds = K2SpeechRecognitionDataset(input_strategy=strategy)

        def epoch():
            sampler = DynamicBucketingSampler(cs, max_duration=max_duration, num_buckets=num_buckets, shuffle=True, seed=0)
            nb, frames, padded = 0, 0, 0
            for batch_cuts in sampler:
                batch = ds[batch_cuts]
                x = batch["inputs"]
                nb += 1
                frames += int(batch["supervisions"]["num_frames"].sum())
                padded += x.shape[0] * x.shape[1]
            torch.cuda.synchronize()
            return nb, frames, padded

        epoch()  # warm: file cache, pinned ring growth, handle
        t0 = time.perf_counter()
        nb, frames, padded = epoch()
        wall = time.perf_counter() - t0
        return {"unit": "h_audio/s", "value": t / 3600.0 / wall, "wall_s": wall, "cuts": len(cuts), "hours_of_audio": t / 3600.0, "batches": nb,
                "pad_fraction": 1.0 - frames / max(padded, 1), "route": strategy.last_batch_route, "kernel": ext.engine.kernel, "n_gpus": 1,
                "note": f"DynamicBucketingSampler(max_duration={max_duration:g}, num_buckets={num_buckets}) -> K2SpeechRecognitionDataset(FusedOnTheFlyFeatures), "
                        f"cuts U[2,30] s; features stay on the device; wall clock of one epoch incl. sampler, PCM16 reads, H2D"}
    finally:
        shutil.rmtree(root, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--hours-per-rank", type=float, default=1.0)
    ap.add_argument("--num-workers", type=int, default=8)
    ap.add_argument("--batch-duration", type=float, default=2000.0)
    args = ap.parse_args()
    import torch

    from lhotse_b200 import dist as lbd

    rank, world, local = lbd.init_distributed()
    torch.cuda.set_device(local)
    lbd.bind_host_to_gpu_numa(local)
    res = run_cutset_job(rank, world, local, seconds_of_audio=args.hours_per_rank * 3600.0, num_workers=args.num_workers,
                         batch_duration=args.batch_duration)
    if rank == 0:
        print(json.dumps(res), flush=True)
        if world == 1:
            print(json.dumps(run_onthefly_job(local, num_workers=args.num_workers)), flush=True)
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


if __name__ == "__main__":
    main()
