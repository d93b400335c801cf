"""The N>1 path on CPU: two gloo ranks shard a batch, run a deterministic per-sample stand-in for
the sampler, and gather -- the result must equal the single-process result."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from eo_diffusion_b200.sharding import gather_samples, sample_sharded, shard_bounds


def test_shard_bounds_cover_everything():
    for n in (0, 1, 5, 64, 257):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _fake_sampler(cond_all):
    # per-sample deterministic "sampling": depends only on that sample's conditioning
    def fn(n, cond, y):
        assert cond.shape[0] == n
        return torch.tanh(cond[:, :3] * 3.0 - cond[:, 3:4]) + (0 if y is None else y.reshape(-1, 1, 1, 1).float())
    return fn


def _worker(rank, world, port, n, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        cond = torch.rand((n, 4, 6, 6), generator=g)
        y = torch.arange(n)
        out = sample_sharded(_fake_sampler(cond), n, cond=cond, y=y)
        want = _fake_sampler(cond)(n, cond, y)
        ok = torch.equal(out, want)
        # even split goes through all_gather_into_tensor, ragged through the padded path
        part = gather_samples(torch.full((3, 2), float(rank)), 6)
        ok = ok and torch.equal(part, torch.tensor([[0., 0.]] * 3 + [[1., 1.]] * 3))
        with open(os.path.join(tmp, f"ok{rank}"), "w") as f:
            f.write("1" if ok else "0")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [6, 7])
def test_two_rank_gloo_matches_single_process(tmp_path, n):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, n, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").read_text() == "1" and (tmp_path / "ok1").read_text() == "1"
