"""World-size-2 `gloo` tests (CPU) of the multi-GPU host logic: rank::world sharding of cuts
(the reference's LazySlicer split, set.py:2158-2160), bit-identical table broadcast, stat reductions."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lhotse_b200 import B200FbankConfig, build_plan
from lhotse_b200 import dist as lbd


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    r, w, _ = lbd.init_distributed(backend="gloo")
    assert (r, w) == (rank, world)
    plan = build_plan("fbank", B200FbankConfig())
    good = plan.tables_blob().copy()
    if rank != 0:  # corrupt the non-root tables: the broadcast must restore rank 0's bits
        plan.window[:] = 7.0
        plan.mel_bank[:] = -1.0
    lbd.broadcast_plan_tables(plan)
    ok_tables = np.array_equal(plan.tables_blob(), good)
    cuts = list(range(23))
    mine = lbd.shard_slice(cuts, rank, world)
    assert mine == list(lbd.shard_iter(cuts, rank, world))
    frames = sum(100 + c for c in mine)
    tot = lbd.all_reduce_stats([frames, len(mine)], "sum")
    mx = lbd.all_reduce_stats([float(rank + 1)], "max")
    lbd.barrier()
    q.put((rank, ok_tables, mine, tot, mx))
    dist.destroy_process_group()


def test_world2_sharding_and_broadcast():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    shards = [r[2] for r in res]
    assert all(r[1] for r in res), "tables differ after broadcast"
    assert sorted(shards[0] + shards[1]) == list(range(23)) and not set(shards[0]) & set(shards[1])
    assert lbd.unshard(shards) == list(range(23))
    assert res[0][3] == res[1][3] == [float(sum(100 + c for c in range(23))), 23.0]
    assert res[0][4] == res[1][4] == [2.0]


def test_single_process_noops():
    plan = build_plan("fbank", B200FbankConfig())
    assert lbd.broadcast_plan_tables(plan) is plan
    assert lbd.all_reduce_stats([1.0, 2.0]) == [1.0, 2.0]
    assert lbd.shard_slice(list(range(10)), 1, 4) == [1, 5, 9]


def test_numa_binding_helpers_are_safe_without_gpu():
    """`bind_host_to_gpu_numa` must never raise: no GPU / no sysfs / single node all mean "leave the process alone"."""
    from lhotse_b200 import dist as lbd
    assert lbd._parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert lbd._parse_cpulist("") == []
    before = os.sched_getaffinity(0)
    assert lbd.bind_host_to_gpu_numa(0) is None or isinstance(lbd.bind_host_to_gpu_numa(0), int)
    if not torch.cuda.is_available():
        assert os.sched_getaffinity(0) == before


def _cutset_worker(rank, world, port, q):
    """BASELINE configs[4] at test size under a REAL process group (world size 2, gloo): the rank-sharded fused store of
    scripts/bench_config5.py with the oracle-backed engine.  Catches what single-process tests cannot: lhotse's samplers
    shard by the ambient process group on their own, so a shard that is split a second time loses half its cuts."""
    import sys

    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.dirname(here), here):
        if p not in sys.path:
            sys.path.insert(0, p)
    import refshim  # noqa: F401
    import scripts.bench_config5 as bc
    from helpers import attach_oracle_engine

    torch.cuda.synchronize = lambda *a, **k: None
    lbd.init_distributed(backend="gloo")
    lb_ex = bc._ensure_lhotse()
    orig = lb_ex.B200Fbank

    class Fake(orig):
        def __init__(self, cfg=None):
            super().__init__(cfg)
            attach_oracle_engine(self)

        @property
        def engine(self):
            return self._engine

    Fake.name = orig.name
    lb_ex.B200Fbank = Fake
    res = bc.run_cutset_job(rank, world, rank, seconds_of_audio=40.0, num_workers=0, batch_duration=25.0)
    q.put((rank, res["cuts"], res.get("spot_check_max_abs_diff")))
    dist.destroy_process_group()


@pytest.mark.reference
def test_world2_cutset_level_sharded_store():
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import refshim

    if not refshim.reference_available():
        pytest.skip("reference not present")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_cutset_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [8, 8] and res[0][2] == 0.0  # 4 cuts per rank, corpus order restored, archive bit-exact


def test_more_local_ranks_than_gpus_maps_round_robin_and_picks_gloo(monkeypatch):
    """`dist.init_distributed` for the host-bound store job: with more local ranks than GPUs, local rank r works on GPU r mod #GPUs and
    the job-level collectives use gloo (NCCL refuses two ranks on one device); one rank per GPU keeps NCCL."""
    import torch

    import lhotse_b200.dist as lbd

    picked = []
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "device_count", lambda: 4)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(lbd.dist, "is_initialized", lambda: False)
    monkeypatch.setattr(lbd.dist, "init_process_group", lambda backend, rank, world_size, **kw: picked.append((backend, rank, world_size, sorted(kw))))
    for rank, world, local, lws, want_dev, want_backend in ((5, 8, 5, 8, 1, "gloo"), (3, 4, 3, 4, 3, "nccl"), (9, 16, 1, 8, 1, "gloo"), (0, 1, 0, 1, 0, None)):
        for k, v in (("RANK", rank), ("WORLD_SIZE", world), ("LOCAL_RANK", local), ("LOCAL_WORLD_SIZE", lws)):
            monkeypatch.setenv(k, str(v))
        picked.clear()
        r, w, dev = lbd.init_distributed()
        assert (r, w, dev) == (rank, world, want_dev)
        assert (picked[0][0] if picked else None) == want_backend
