"""World-size-2 CPU (gloo) tests of the data-parallel host logic in llm_qat_on_gpt2_b200/dp.py:
gradient bucket all-reduce over only the gradients that exist, MIN/MAX exchange of calibration
statistics, batch sharding.  No GPU, no kernels: the quantiser objects are stand-ins that carry
the attributes dp.py touches."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


class _FakeQuantizer:
    def __init__(self, tmin, tmax, flag):
        self.temp_min, self.temp_max = tmin, tmax
        self._stat_state = torch.tensor([flag], dtype=torch.int32)
        self.num_batches_collected = 1
        self.quantizer_type = "log"
        self.stats_sync_hook = None
        self._stat_flag_host = None
        self.finished_with = None

    def finish_calibration(self, debug=False):
        self.finished_with = (self.temp_min.clone(), self.temp_max.clone(), self._stat_flag_host)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from llm_qat_on_gpt2_b200 import dp
    try:
        # --- gradients: only params with a grad take part; result is the mean over ranks
        torch.manual_seed(0)
        lin = torch.nn.Linear(4, 3)
        frozen = torch.nn.Parameter(torch.ones(5), requires_grad=False)
        unused = torch.nn.Parameter(torch.ones(2))                      # requires_grad but never gets a grad
        lin.weight.grad = torch.full((3, 4), float(rank + 1))
        lin.bias.grad = torch.arange(3, dtype=torch.float32) * (rank + 1)
        n = dp.allreduce_gradients([lin.weight, frozen, unused, lin.bias])
        assert n == 15
        assert torch.allclose(lin.weight.grad, torch.full((3, 4), 1.5))
        assert torch.allclose(lin.bias.grad, torch.arange(3, dtype=torch.float32) * 1.5)
        assert unused.grad is None
        fin = dp.allreduce_gradients([lin.weight], average=False, async_op=True)
        assert fin() == 12 and torch.allclose(lin.weight.grad, torch.full((3, 4), 3.0))
        # --- calibration statistics: elementwise MIN / MAX, flags OR-ed
        a = _FakeQuantizer(torch.tensor([[1.0, -2.0 - rank]]), torch.tensor([[3.0 + rank, 0.5]]), flag=rank)
        b = _FakeQuantizer(torch.tensor([float(rank)]), torch.tensor([10.0 * rank]), flag=0)
        dp.finish_calibration_many([a, b])
        tmin, tmax, fl = a.finished_with
        assert torch.equal(tmin, torch.tensor([[1.0, -3.0]])) and torch.equal(tmax, torch.tensor([[4.0, 0.5]])) and fl == 1
        tmin, tmax, fl = b.finished_with
        assert torch.equal(tmin, torch.tensor([0.0])) and torch.equal(tmax, torch.tensor([10.0])) and fl == 0
        assert a._stat_flag_host is None                                 # reset after use
        # --- batch sharding
        x = torch.arange(8).view(4, 2)
        assert torch.equal(dp.shard_batch(x), x[rank * 2:(rank + 1) * 2])
        with pytest.raises(ValueError):
            dp.shard_batch(torch.zeros(3, 2))
        q.put((rank, "ok"))
    except Exception as e:                                              # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_dp_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}: {msg}"


def test_single_process_is_a_noop():
    from llm_qat_on_gpt2_b200 import dp
    p = torch.nn.Parameter(torch.ones(3)); p.grad = torch.full((3,), 2.0)
    assert dp.allreduce_gradients([p]) == 3 and torch.equal(p.grad, torch.full((3,), 2.0))
    q = _FakeQuantizer(torch.tensor([1.0]), torch.tensor([2.0]), 1)
    dp.sync_calibration_stats([q])
    assert torch.equal(q.temp_min, torch.tensor([1.0]))
    assert torch.equal(dp.shard_batch(torch.arange(4)), torch.arange(4))
