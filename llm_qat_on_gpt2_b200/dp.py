"""Batch-sharded data parallelism for the switchable-precision path (one process per GPU).

The reference has no distributed code; BASELINE.json's north_star asks for
batch sharding over the GPUs of a node with an all-reduce of only the trainable gradients.  Every
op on the path is independent per token except

  (1) calibration min/max (a reduction over tokens): `sync_calibration_stats` MIN/MAX all-reduces
      the collected statistics, so that N ranks calibrating on N shards end up with exactly the
      parameters one process would compute on the whole batch (min/max are order independent);
  (2) parameter gradients (a sum over tokens): `allreduce_gradients` packs the gradients of the
      parameters that actually received one (active-precision LoRA A/B, active LayerNorm pair,
      anything else with requires_grad) into one flat fp32 bucket, all-reduces it once over
      NCCL (NVLink / NVSwitch) and averages, matching the mean-reduced losses of the reference.

There is no data-path collective in the forward or backward of a layer.  Works with any
torch.distributed backend (nccl on GPUs; the CPU tests use gloo).
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def _world(group=None) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def allreduce_gradients(params: Iterable[torch.nn.Parameter], group=None, average: bool = True,
                        async_op: bool = False):
    """Sum (or average) the .grad of every parameter that has one, in a single flat bucket.

    All ranks must hold gradients for the same set of parameters (they do: every rank runs the
    same precision).  Returns the number of elements reduced (or, with async_op, a callable that
    waits for the collective and scatters the result back)."""
    ps: List[torch.nn.Parameter] = [p for p in params if p.requires_grad and p.grad is not None]
    if not ps:
        return 0
    world = _world(group)
    flat = torch.cat([p.grad.reshape(-1).float() for p in ps])
    work = None
    if world > 1:
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)

    def finish():
        if work is not None and async_op:
            work.wait()
        if average and world > 1:
            flat.div_(world)
        off = 0
        for p in ps:
            n = p.grad.numel()
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
            off += n
        return flat.numel()

    return finish if async_op else finish()


def trainable_parameters(model: torch.nn.Module):
    return [p for p in model.parameters() if p.requires_grad]


def sync_calibration_stats(quantizers, group=None) -> None:
    """MIN / MAX all-reduce of the statistics collected so far by `quantizers` (objects with
    temp_min / temp_max / _stat_state), batched into two collectives plus one for the flags."""
    qs = [q for q in quantizers if q.temp_min is not None]
    if _world(group) == 1:
        return
    # every rank must bring the same quantisers (same count, same total width): a rank that collected nothing for
    # one of them would otherwise hang or silently mis-align the bucket
    mine = torch.tensor([len(qs), sum(q.temp_min.numel() for q in qs)], dtype=torch.int64,
                        device=qs[0].temp_min.device if qs else _any_device(quantizers))
    lo, hi = mine.clone(), mine.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    if not torch.equal(lo, hi):
        raise RuntimeError(f"sync_calibration_stats: ranks disagree on the collected quantisers "
                           f"(this rank {mine.tolist()}, min {lo.tolist()}, max {hi.tolist()})")
    if not qs:
        return
    mins = torch.cat([q.temp_min.reshape(-1) for q in qs])
    maxs = torch.cat([q.temp_max.reshape(-1) for q in qs])
    dist.all_reduce(mins, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(maxs, op=dist.ReduceOp.MAX, group=group)
    flags = [q._stat_state for q in qs if getattr(q, "_stat_state", None) is not None]
    if flags:
        fl = torch.cat(flags)
        dist.all_reduce(fl, op=dist.ReduceOp.MAX, group=group)
    off = 0
    foff = 0
    for q in qs:
        n = q.temp_min.numel()
        q.temp_min.copy_(mins[off:off + n].view_as(q.temp_min))
        q.temp_max.copy_(maxs[off:off + n].view_as(q.temp_max))
        off += n
        if getattr(q, "_stat_state", None) is not None:
            q._stat_state.copy_(fl[foff:foff + 1])
            foff += 1


def _any_device(quantizers):
    for q in quantizers:
        st = getattr(q, "_stat_state", None)
        if st is not None:
            return st.device
        buf = getattr(q, "scale", None)
        if buf is not None:
            return buf.device
    return torch.device("cpu")


def install_calibration_sync(model: torch.nn.Module, group=None) -> int:
    """Make every ACTIVATION quantiser of `model` (is_input=True: the only statistics that depend on the batch
    shard) all-reduce its statistics inside its own finish_calibration, so the reference's unmodified
    CalibrationManager becomes DP-correct.  Weight and LoRA quantisers see replicated parameters and need no
    exchange -- leaving them unhooked also keeps them eligible for the one-launch `calibrate_many` /
    `training.LoRARefresher` path.  One pair of small collectives per quantiser; `finish_calibration_many` is the
    batched form."""
    n = 0
    for m in model.modules():
        if m.__class__.__name__ == "LearnableFakeQuantize" and getattr(m, "is_input", False):
            m.stats_sync_hook = lambda q, tmin, tmax, _g=group: sync_calibration_stats([q], _g)
            n += 1
    return n


def finish_calibration_many(quantizers, group=None, debug: bool = False) -> None:
    """finish_calibration() for a list of quantisers with ONE statistics exchange and ONE host read
    of the log-mode `any(|x| > eps)` flags, instead of one of each per quantiser."""
    qs = list(quantizers)
    sync_calibration_stats(qs, group)
    live = [q for q in qs if q.num_batches_collected > 0 and q.temp_min is not None
            and q.quantizer_type == 'log' and q._stat_state is not None]
    if live:
        flags = torch.cat([q._stat_state for q in live]).tolist()      # one device->host copy
        for q, f in zip(live, flags):
            q._stat_flag_host = int(f)
    for q in qs:
        hook, q.stats_sync_hook = q.stats_sync_hook, None
        try:
            q.finish_calibration(debug=debug)
        finally:
            q.stats_sync_hook = hook
            q._stat_flag_host = None


def shard_batch(t: torch.Tensor, rank: Optional[int] = None, world: Optional[int] = None) -> torch.Tensor:
    """This rank's contiguous slice of the batch dimension (dim 0)."""
    if world is None:
        world = _world()
    if rank is None:
        rank = dist.get_rank() if world > 1 else 0
    b = t.shape[0]
    if b % world:
        raise ValueError(f"global batch {b} is not divisible by world size {world}")
    per = b // world
    return t[rank * per:(rank + 1) * per]
