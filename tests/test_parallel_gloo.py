"""CPU tier: the N>1 host logic with world_size = 2 over gloo (127.0.0.1): unit sharding, pair co-location in the batch
split, gradient averaging, global BatchNorm sums, ordered gather of volume shards."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from superresolution_aniso_mri_b200 import parallel as P


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = {}
        g = torch.Generator().manual_seed(3)
        # ---- inference: volumes sharded with no collective, gathered back in order (ragged: 5 volumes on 2 ranks)
        vols = torch.rand(5, 3, 4, 4, generator=g)
        local = P.shard_volumes(vols, rank, world)
        res["n_local"] = local.shape[0]
        parts = P.gather_volume_shards(local * 2, 5)
        res["gather_ok"] = bool(torch.equal(torch.cat(parts), vols * 2))
        # ---- training batch: B = 6 triplets, image = [from(6); to(6)]
        B = 6
        batch = {"image": torch.arange(2 * B, dtype=torch.float32).view(2 * B, 1, 1, 1),
                 "slice_between": 100 + torch.arange(B, dtype=torch.float32).view(B, 1, 1, 1),
                 "alpha_from": torch.arange(B, dtype=torch.float32).view(B, 1)}
        lb = P.shard_batch_pairs(batch, rank, world)
        b = lb["slice_between"].shape[0]
        ids = (lb["slice_between"].flatten() - 100).long()
        res["pairs_ok"] = bool(torch.equal(lb["image"][:b].flatten().long(), ids) and
                               torch.equal(lb["image"][b:].flatten().long(), ids + B) and
                               torch.equal(lb["alpha_from"].flatten().long(), ids))
        res["ids"] = ids.tolist()
        # ---- gradient mean + BN sums
        grad = torch.full((7,), float(rank + 1))
        P.average_gradients_(grad)
        res["grad"] = grad.tolist()
        stats = torch.tensor([1.0 + rank, 10.0 * (rank + 1)])
        res["count"] = P.sync_bn_sums_(stats, 50)
        res["stats"] = stats.tolist()
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_host_logic():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [out[r]["n_local"] for r in range(2)] == [3, 2]
    assert all(out[r]["gather_ok"] and out[r]["pairs_ok"] for r in range(2))
    assert out[0]["ids"] + out[1]["ids"] == list(range(6))                 # a partition of the global batch
    assert out[0]["grad"] == out[1]["grad"] == [1.5] * 7
    assert out[0]["stats"] == out[1]["stats"] == [3.0, 30.0] and out[0]["count"] == 100


def test_shard_ranges_cover_everything():
    for n in (0, 1, 7, 64, 513):
        for world in (1, 2, 4, 8):
            spans = [P.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(e - s for s, e in spans) - min(e - s for s, e in spans) <= 1
    with pytest.raises(ValueError):
        P.shard_batch_pairs({"image": torch.zeros(5, 1, 2, 2), "slice_between": torch.zeros(3, 1, 2, 2)}, 0, 2)
