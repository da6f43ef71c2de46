"""Multi-GPU plumbing of the inference path: clips are independent, so every rank decodes its own contiguous
slice of the batch (one process per GPU, `torch.distributed`); there is NO collective on the data path.

Two optional collectives exist around it:
  * `gather_tokens`  -- rank 0 collects the token ids of all ranks (20 int64 per clip), after the timed region;
  * `global_db_max`  -- AmplitudeToDB(top_db=120) clamps against the maximum of the WHOLE batch
                        (reference: captioning/models/hf_wrapper.py:292-293, torchaudio amplitude_to_DB on a 3-D
                        input); when a batch is split across ranks one scalar all-reduce(max) restores exactly the
                        single-GPU semantics.  It only matters for batches whose dynamic range exceeds 120 dB.
Works with the nccl backend on GPUs and the gloo backend on CPU (tests)."""
import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int):
    """[start, stop) of the contiguous slice owned by `rank` (sizes differ by at most one)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_clips(wav: torch.Tensor, wav_len, rank: int, world: int):
    """This rank's clips of a [B, N] batch and their lengths."""
    a, b = shard_range(wav.shape[0], rank, world)
    lens = torch.as_tensor(wav_len)
    return wav[a:b], lens[a:b]


def global_db_max(gmax: torch.Tensor, group=None) -> torch.Tensor:
    """In-place all-reduce(max) of the per-rank dB maximum (a 1-element tensor)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(gmax, op=dist.ReduceOp.MAX, group=group)
    return gmax


def gather_tokens(seq_local: torch.Tensor, n_total: int, group=None, dst: int = 0):
    """Collect the [n_local, max_len] int64 token ids of every rank on `dst` in clip order; other ranks get None."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return seq_local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    max_len = seq_local.shape[1]
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    pad = max(b - a for a, b in sizes)
    buf = torch.zeros(pad, max_len, dtype=seq_local.dtype, device=seq_local.device)
    buf[: seq_local.shape[0]] = seq_local
    out = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, out, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([o[: b - a] for o, (a, b) in zip(out, sizes)], dim=0)


def max_over_ranks(ms: float, device=None, group=None) -> float:
    """The slowest rank's time (multi-GPU numbers are the max over ranks, never wall clock)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return ms
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
