"""N>1 host logic on CPU: world_size-2 gloo process group, window-granular sharding and the
ordered gather of labels (the only communication of the multi-GPU path)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helen_b200.sharding import gather_labels, predict_sharded, shard_bounds


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 8, 1000, 3_000_000):
        for world in (1, 2, 3, 8):
            bounds = [shard_bounds(n, world, r) for r in range(world)]
            assert bounds[0][0] == 0 and bounds[-1][1] == n
            assert all(bounds[i][1] == bounds[i + 1][0] for i in range(world - 1))
            sizes = [e - s for s, e in bounds]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_predict(images):
    # deterministic stand-in for the GPU path: labels are a function of the window contents only
    img = images.to(torch.int64)
    base = (img.sum(dim=2) % 5).to(torch.uint8)
    rle = (img[:, :, 0] % 11).to(torch.uint8)
    return base, rle


def _worker(rank, world, port, n_windows, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    gen = torch.Generator().manual_seed(0)
    images = torch.randint(0, 256, (n_windows, 20, 4), dtype=torch.uint8, generator=gen)
    base, rle = predict_sharded(_fake_predict, images)
    if rank == 0:
        ref_b, ref_r = _fake_predict(images)
        ok = np.array_equal(base, ref_b.numpy()) and np.array_equal(rle, ref_r.numpy())
        with open(out_path, "w") as f:
            f.write("ok" if ok else "mismatch")
    else:
        assert base is None and rle is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_windows", [9, 2])
def test_sharded_predict_world_size_2_gloo(tmp_path, n_windows):
    out = tmp_path / "result.txt"
    mp.spawn(_worker, args=(2, _free_port(), n_windows, str(out)), nprocs=2, join=True)
    assert out.read_text() == "ok"


def test_gather_without_process_group_is_identity():
    b = torch.arange(12, dtype=torch.uint8).reshape(3, 4)
    base, rle = gather_labels(b, b + 1, 3)
    assert np.array_equal(base, b.numpy()) and np.array_equal(rle, (b + 1).numpy())
