"""Test-only: the pieces every test that drives the REAL lhotse callers needs — the imported reference (tree in the
build container, `oracle/_ref/lhotse_ref.zip` on the GPU box), a stdlib-`wave` audio backend (no decoder library is
installed) and synthetic PCM16 cut sets on disk."""
import importlib
import wave

import numpy as np

import refshim


def setup_lhotse():
    """Imports the reference, re-binds lhotse_b200's base classes to the real lhotse ones, installs the WAV backend.
    Returns the reloaded `lhotse_b200.extractors` module."""
    refshim.import_reference()
    import lhotse_b200.base as lb_base
    import lhotse_b200.extractors as lb_ex

    if not lb_base.HAVE_LHOTSE:
        importlib.reload(lb_base)
        importlib.reload(lb_ex)
        import lhotse_b200.families as fam
        import lhotse_b200.storage as lb_st

        importlib.reload(lb_st)
        importlib.reload(fam)
    assert lb_base.HAVE_LHOTSE
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
    import lhotse_b200.extractors as lb_ex2

    return lb_ex2


def write_wav(path, pcm, sr=16000):
    with wave.open(str(path), "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(sr)
        w.writeframes(np.asarray(pcm, dtype="<i2").tobytes())


def make_cutset(root, durations, sr=16000, seed=0, amp=0.1, supervisions=True, prefix="c"):
    """One PCM16 WAV + MonoCut per duration (seconds): N(0, amp^2) noise as BASELINE.json's configs describe."""
    from lhotse import CutSet, MonoCut, Recording, SupervisionSegment
    from lhotse.audio import AudioSource

    rs = np.random.RandomState(seed)
    cuts = []
    for i, dur in enumerate(durations):
        n = int(round(dur * sr))
        path = root / f"{prefix}{i}.wav"
        write_wav(path, np.clip(rs.randn(n) * amp * 32768, -32768, 32767).astype("<i2"), sr)
        rec = Recording(id=f"{prefix}r{i}", sources=[AudioSource(type="file", channels=[0], source=str(path))],
                        sampling_rate=sr, num_samples=n, duration=n / sr)
        sups = []
        if supervisions:
            sups = [SupervisionSegment(id=f"{prefix}s{i}", recording_id=f"{prefix}r{i}", start=0.0, duration=n / sr, text="x")]
        cuts.append(MonoCut(id=f"{prefix}{i}", start=0.0, duration=n / sr, channel=0, recording=rec, supervisions=sups))
    return CutSet.from_cuts(cuts)
