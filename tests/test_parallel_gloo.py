"""world_size-2 gloo tests (CPU) of the N>1 host logic: ray sharding and the flat-arena gradient all-reduce."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from contrastive_lift_b200 import parallel as par


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 3), torch.nn.Linear(3, 2))
        par.broadcast_parameters(net)
        x = torch.arange(40, dtype=torch.float32).reshape(8, 5) / 10.0
        xs = par.shard_rays(x)                       # each rank renders its own rays
        y = net[:3](xs)                              # the last layer is unused in this pass -> grad None
        (y.pow(2).sum() / x.shape[0]).backward()     # loss is a mean over ALL rays: local sum / global count
        nbytes = par.allreduce_gradients(net.parameters(), average=False)
        assert nbytes == sum(p.numel() for p in net.parameters()) * 4
        assert net[3].weight.grad is None
        got = [p.grad.clone() for p in net[:3].parameters()]
        # single-process gradient of the concatenated batch
        ref = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 3))
        ref.load_state_dict({k: v for k, v in net.state_dict().items() if not k.startswith("3.")})
        (ref(x).pow(2).sum() / x.shape[0]).backward()
        for g, p in zip(got, ref.parameters()):
            assert torch.allclose(g, p.grad, rtol=1e-5, atol=1e-6)
        # mean semantics (DDP): identical per-rank grads stay unchanged
        for p in net[:3].parameters():
            p.grad.fill_(float(rank + 1))
        par.allreduce_gradients(net.parameters(), average=True)
        assert all(torch.allclose(p.grad, torch.full_like(p.grad, (world + 1) / 2.0)) for p in net[:3].parameters())
        # output maps gather back in ray order
        full = par.gather_rays_output(xs * 2.0, x.shape[0])
        assert torch.equal(full, x * 2.0)
        # ragged shards (ray counts that do not divide by the world size), 1-D maps (depth) and an empty frame
        for n in (7, 1, 0, 5 * world + 3):
            m = torch.arange(n * 3, dtype=torch.float32).reshape(n, 3)
            assert torch.equal(par.gather_rays_output(par.shard_rays(m) + 1.0, n), m + 1.0), n
            assert torch.equal(par.gather_rays_output(par.shard_rays(m[:, 0]), n), m[:, 0]), n
        with pytest.raises(ValueError):
            par.gather_rays_output(x, x.shape[0])          # the whole frame is not this rank's shard
        # row-cyclic shards of a frame (balanced cost) gather back in ray order, ragged row counts included
        for h, w in ((6, 4), (7, 3), (1, 5), (2 * world + 1, 2)):
            frame = torch.arange(h * w * 3, dtype=torch.float32).reshape(h * w, 3)
            mine = par.shard_rows_cyclic(frame, h, w)
            assert mine.shape[0] == len(range(rank, h, world)) * w
            assert torch.equal(par.gather_rows_cyclic(mine * 2.0, h, w), frame * 2.0), (h, w)
            assert torch.equal(par.gather_rows_cyclic(mine[:, 0].contiguous(), h, w), frame[:, 0]), (h, w)
        # a None pattern that differs across ranks (one rank skipped a head) is reported by the next reduce
        arena = par.GradientArena(net.parameters())
        for p in net.parameters():
            p.grad = torch.ones_like(p)
        if rank == 0:
            net[3].bias.grad = None
        arena.reduce()
        with pytest.raises(RuntimeError, match="None-pattern"):
            arena.reduce()
        arena.reduce()                                     # the flags of the call that raised were consumed
        out.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_shard_range_covers_everything_in_order():
    for n in (0, 1, 7, 8, 160000, 640001):
        for world in (1, 2, 3, 8):
            spans = [par.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1


@pytest.mark.timeout(180)
@pytest.mark.parametrize("world", [2, 3])
def test_flat_arena_allreduce_matches_single_process_gradient(world):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert sorted(res) == [(r, "ok") for r in range(world)], res
