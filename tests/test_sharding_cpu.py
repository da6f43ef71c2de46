"""world_size-2 gloo tests (CPU) of the multi-GPU plumbing: clip sharding, token gather, the batch-global dB maximum
and the max-over-ranks timing rule.  The data path itself has no collective (clips are independent)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from audiocaption_b200 import sharding


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 64, 129):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, n_clips, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        wav = torch.randn(n_clips, 100, generator=g)
        lens = torch.arange(n_clips) + 50
        w, l = sharding.shard_clips(wav, lens, rank, world)
        a, b = sharding.shard_range(n_clips, rank, world)
        assert w.shape[0] == b - a and (l == lens[a:b]).all()
        # "decode": token ids that encode the global clip index, so the gather order can be checked
        seq = (torch.arange(a, b)[:, None] * 100 + torch.arange(20)[None, :]).to(torch.int64)
        full = sharding.gather_tokens(seq, n_clips)
        if rank == 0:
            want = torch.arange(n_clips)[:, None] * 100 + torch.arange(20)[None, :]
            assert full.shape == (n_clips, 20) and (full == want).all()
        else:
            assert full is None
        # batch-global dB maximum == maximum over the unsharded batch
        gmax = w.max().reshape(1).clone() if w.numel() else torch.full((1,), float("-inf"))
        sharding.global_db_max(gmax)
        assert gmax.item() == wav.max().item()
        # timing rule: every rank reports the slowest rank's time
        assert sharding.max_over_ranks(1.0 + rank) == float(world)
        q.put((rank, "ok"))
    except Exception as e:   # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_clips", [7, 64])
def test_two_rank_gloo(n_clips):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_clips, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


# ------------------------------------------------------------------ training: flat gradient all-reduce (2 gloo ranks)
def _ddp_worker(rank, world, port, q):
    import os
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from audiocaption_b200.train_step import allreduce_gradients, flatten_trainable
    torch.manual_seed(0)                                   # identical replicas
    model = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.LayerNorm(7), torch.nn.Linear(7, 3))
    model[1].bias.requires_grad_(False)                    # a frozen parameter stays out of the flat buffers
    before = {k: v.detach().clone() for k, v in model.state_dict().items()}
    params, flat_p, flat_g = flatten_trainable(model)
    assert len(params) == 5 and flat_p.numel() % 32 == 0 and flat_p.numel() >= sum(p.numel() for p in params)
    assert all((model.state_dict()[k] == v).all() for k, v in before.items())          # values survive the re-homing
    assert all(p.data_ptr() >= flat_p.data_ptr() and p.grad.data_ptr() >= flat_g.data_ptr() for p in params)
    for i, p in enumerate(params):                         # each rank writes its own gradients THROUGH the views
        p.grad.fill_(float(rank + 1) * (i + 1))
    scale = allreduce_gradients(flat_g)
    ok = abs(scale - 1.0 / world) < 1e-12
    for i, p in enumerate(params):                         # sum over ranks, visible through the views
        ok = ok and bool((p.grad == float((1 + 2) * (i + 1))).all())
    with torch.no_grad():
        flat_p.add_(flat_g, alpha=-scale)                  # "optimizer" on the flat buffer updates the module's parameters
    ok = ok and bool(torch.allclose(model[0].weight, before["0.weight"] - 1.5 * 1))
    ok = ok and bool((model[1].bias == before["1.bias"]).all())
    q.put((rank, ok))
    dist.destroy_process_group()


def test_flat_gradient_allreduce_two_ranks():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 200
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


# ------------------------------------------------------------------ training: two-bucket exchange (decoder bucket overlaps the GRU backward)
def _bucket_worker(rank, world, port, q):
    import os
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from audiocaption_b200.train_step import allreduce_bucket_async, allreduce_gradients, bucket_boundary, flatten_trainable

    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.encoder = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 7))
            self.decoder = torch.nn.Sequential(torch.nn.Linear(7, 3), torch.nn.LayerNorm(3))
    torch.manual_seed(0)
    model = M()
    params, flat_p, flat_g = flatten_trainable(model)
    off = bucket_boundary(model, params)
    # encoder tensors: 35, 7, 49, 7 floats, each rounded up to 32 -> 64 + 32 + 64 + 32
    ok = off == 192 and params[4] is model.decoder[0].weight and flat_g.numel() > off
    flat_g.fill_(float(rank + 1))
    work = allreduce_bucket_async(flat_g[off:])            # TrainStep.step: right after the decoder's backward pass
    scale = allreduce_gradients(flat_g[:off])              # ... and after the GRU's
    work.wait()
    ok = ok and abs(scale - 0.5) < 1e-12 and bool((flat_g == 3.0).all())
    # a model whose flat order interleaves the two halves has no clean boundary: the step falls back to one all-reduce
    shuffled = [params[4], params[0], params[5]]
    ok = ok and bucket_boundary(model, shuffled) is None and allreduce_bucket_async(flat_g[:0]) is None
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_two_bucket_gradient_exchange_two_ranks():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29850 + os.getpid() % 100
    procs = [ctx.Process(target=_bucket_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
