"""N>1 host path on CPU: world_size-2/3 `gloo` processes, contiguous time shards with a
one-frame halo, integer results gathered on rank 0.  The scorer is the oracle-backed fake
(no GPU here); sharding.py, the detectors and the comm plumbing are the product code.  The
sharded run must equal the golden (serial reference) cut list and integer sums exactly."""

import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, case_name, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pyscenedetect_b200.sharding import TorchComm, detect_sharded, shard_bounds
        from tests.fake_engine import OracleEngine, OracleResults
        from tests.golden_util import case_frames, get_case
        from tests.test_gpu_parity import _build
        case = get_case(case_name)
        frames = case_frames(case)
        n = frames.shape[0]
        b = shard_bounds(n, world)
        local = frames[b[rank]:b[rank + 1]]
        det = _build(case)
        cuts, sums = detect_sharded(local, b[rank], n, det, case["fps"], TorchComm(),
                                    engine_factory=OracleEngine, results_factory=OracleResults,
                                    batch_size=16)
        if rank == 0:
            serial = OracleEngine(frames.shape[2], frames.shape[1], det.required_features(),
                                  edge_kernel_size=det.edge_kernel_size_arg())
            serial.submit(frames)
            q.put((cuts, sums.tobytes() == serial.read_sums().tobytes()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("case_name", ["content_default_nostats", "adaptive_w2", "hist_256",
                                       "cfg1_threshold_360p", "content_edges_k3", "hash_default"])
def test_sharded_equals_serial(case_name, world):
    from tests.golden_util import get_case
    if case_name == "cfg1_threshold_360p" and world == 3:
        pytest.skip("one 360p case per world size is enough")
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = 29500 + (os.getpid() + world * 7 + len(case_name)) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, case_name, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    cuts, sums_equal = q.get()
    assert cuts == get_case(case_name)["cuts"]
    assert sums_equal


def test_shard_bounds():
    from pyscenedetect_b200.sharding import shard_bounds
    assert shard_bounds(10, 3) == [0, 3, 6, 10]
    assert shard_bounds(10000, 8)[-1] == 10000
    assert all(b - a in (1250,) for a, b in zip(shard_bounds(10000, 8)[:-1], shard_bounds(10000, 8)[1:]))
    assert isinstance(np.int64(1), np.integer) and torch.__version__
