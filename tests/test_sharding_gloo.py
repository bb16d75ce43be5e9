"""world_size-2 gloo test (CPU) of the N>1 host logic: shard ranges, G-independent per-sample
noise, 1/B_global gradient scaling and the final all-gather."""
import importlib
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

S = importlib.import_module("3dpointcloudattack_b200.sharding")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, global_batch, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        s, e = S.shard_range(global_batch, rank, world)
        full = torch.arange(global_batch * 6, dtype=torch.float32).view(global_batch, 3, 2)
        local = S.shard(full, rank, world)
        assert local.shape[0] == e - s
        # a per-sample "loss" with the reference's batch mean, scaled so that gradients match the unsharded run
        x = local.clone().requires_grad_(True)
        (x.pow(2).sum((1, 2)).mean() * S.mean_scale(e - s, global_batch)).backward()
        ref = full.clone().requires_grad_(True)
        ref.pow(2).sum((1, 2)).mean().backward()
        assert torch.allclose(x.grad, ref.grad[s:e])
        noise = S.per_sample_noise((3, 2), s, e - s, 0.1, seed=7)
        all_noise = S.gather_batch(noise, global_batch)
        all_clouds = S.gather_batch(local, global_batch)
        all_losses = S.gather_batch(local.sum((1, 2)), global_batch)
        assert torch.equal(all_clouds, full) and torch.equal(all_losses, full.sum((1, 2)))
        assert torch.equal(all_noise, S.per_sample_noise((3, 2), 0, global_batch, 0.1, seed=7))
        out_q.put((rank, "ok"))
    except Exception as exc:            # surface the failure in the parent
        out_q.put((rank, repr(exc)))
    finally:
        dist.destroy_process_group()


def _run(global_batch):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, global_batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_shard_ranges_cover_the_batch():
    for B in (1, 2, 7, 32, 129):
        for G in (1, 2, 4, 8):
            spans = [S.shard_range(B, r, G) for r in range(G)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(e - s for s, e in spans) - min(e - s for s, e in spans) <= 1


def test_two_ranks_even_batch():
    _run(8)


def test_two_ranks_ragged_batch():
    _run(7)
