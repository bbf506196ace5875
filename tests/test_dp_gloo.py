"""CPU, world_size 2, gloo: the only multi-GPU plumbing of the path -- the load-time weights broadcast and
the image sharding."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from yolo_quantization_b200 import dp, synth


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 128, 1024, 1000):
        for world in (1, 2, 3, 8):
            spans = [dp.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        dp.shard_range(10, 2, 2)


def _worker(rank, world, port, tmp):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    path = os.path.join(tmp, f"rank{rank}.weights")
    layers = synth.single_conv(16, 3) + [synth.LayerSpec("maxpool", size=2, stride=2)]
    if rank == 0:
        synth.write_weights(path, layers, width=8, height=8)
    n = dp.broadcast_weights_file(path, rank, world, None, dist)
    data = np.fromfile(path, dtype=np.uint8)
    assert data.size == n
    # every replica holds identical bytes; shards are disjoint and cover the batch
    import torch
    digest = torch.tensor([int(data.astype(np.int64).sum()), data.size], dtype=torch.int64)
    gathered = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(gathered, digest)
    assert all(torch.equal(g, gathered[0]) for g in gathered)
    lo, hi = dp.shard_range(1024, rank, world)
    cnt = torch.tensor([hi - lo])
    dist.all_reduce(cnt)
    assert int(cnt) == 1024
    dist.destroy_process_group()


def test_weights_broadcast_world2(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
