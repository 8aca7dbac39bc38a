"""world_size-2 test of the N > 1 path on CPU (gloo): image sharding, per-rank decode through
the CPU checker standing in for the GPU, and the final timing / throughput reduce."""
import os
import socket
import sys

import numpy as np
import pytest

import gst_fixtures as fx


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, q):
    sys.path[:0] = [fx.ROOT, os.path.join(fx.ROOT, "tests")]
    import torch.distributed as dist
    from gst_b200.shard import reduce_job, shard_indices
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    mine = shard_indices(n_items, rank, world)
    files = [np.fromfile(os.path.join(fx.GOLDEN_DIR, n), dtype=np.uint8) for n in ("test1.gst", "synth512_s7.gst")]
    digest = []
    for i in mine:  # each rank decodes only its own images
        out = fx.oracle_decode(files[i % 2], taps=False)["out"]
        digest.append((i, fx.sha(out)))
    ms, (texels, images) = reduce_job(10.0 * (rank + 1), [512 * 512 * len(mine), len(mine)])
    q.put((rank, mine, digest, ms, texels, images))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_indices_partition():
    from gst_b200.shard import shard_indices
    for n, world in ((1024, 8), (600, 8), (7, 2), (3, 4), (0, 2)):
        parts = [shard_indices(n, r, world) for r in range(world)]
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        shard_indices(4, 2, 2)


def test_two_ranks_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port, world, n_items = _free_port(), 2, 5
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    [p.start() for p in procs]
    results = [q.get(timeout=120) for _ in range(world)]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    results.sort()
    seen = sorted(i for _, mine, _, _, _, _ in results for i in mine)
    assert seen == list(range(n_items))                      # every image decoded exactly once
    want = {0: fx.sha(fx.golden_test1()[1])}
    for _, _, digest, ms, texels, images in results:
        assert ms == 20.0                                    # max over ranks
        assert images == n_items and texels == 512 * 512 * n_items  # whole-job sums
        for i, h in digest:
            if i % 2 == 0:
                assert h == want[0]                          # rank-local result is the golden one
