"""N > 1 path on CPU: world_size 2, gloo.  Streams are independent replicas; only counters are reduced."""
import os
import socket

import pytest

torch = pytest.importorskip("torch")


def _worker(rank, world, port, q):
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "lf-vio_b200"))
    os.environ["LFVIO_NO_LIB"] = "1"
    import importlib.util

    spec = importlib.util.spec_from_file_location("lfvio_multi", os.path.join(root, "lf-vio_b200", "lfvio", "multi.py"))
    multi = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(multi)
    import torch.distributed as dist

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    # rank r processed 100 frames in (10 + 5 r) ms; each rank owns a different stream seed
    ms, frames = multi.reduce_counters(10.0 + 5.0 * rank, 100)
    fps = multi.aggregate_fps(10.0 + 5.0 * rank, 100)
    q.put((rank, ms, frames, fps, multi.stream_seed(rank)))
    dist.barrier()
    dist.destroy_process_group()


def test_counters_reduce_over_two_ranks():
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ms, frames, fps, seed in out:
        assert ms == 15.0 and frames == 200          # MAX of the times, SUM of the frames
        assert abs(fps - 200 / 0.015) < 1e-6
        assert seed == rank + 1                       # independent streams (no data-path collective)


def test_single_process_passthrough():
    import importlib.util

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("lfvio_multi", os.path.join(root, "lf-vio_b200", "lfvio", "multi.py"))
    multi = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(multi)
    assert multi.reduce_counters(12.5, 7) == (12.5, 7)
