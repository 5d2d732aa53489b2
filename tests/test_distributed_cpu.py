"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: window sharding, rank-0 parameter
broadcast after lazy creation, and the single flattened gradient all-reduce of the training step."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world_size, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    from temporal_latticenet_b200 import parallel
    try:
        # sharding: a partition of the windows, round robin
        mine = parallel.shard_windows(11)
        gathered = [None] * world_size
        dist.all_gather_object(gathered, mine)
        assert sorted(sum(gathered, [])) == list(range(11))
        assert mine == list(range(rank, 11, world_size))
        # broadcast after "lazy creation" with different seeds per rank
        torch.manual_seed(100 + rank)
        m = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.GroupNorm(1, 7), torch.nn.Linear(7, 3))
        parallel.broadcast_parameters(m)
        flat = torch.cat([p.detach().reshape(-1) for p in m.parameters()])
        ref = [torch.zeros_like(flat) for _ in range(world_size)]
        dist.all_gather(ref, flat)
        assert all(torch.equal(ref[0], r) for r in ref)
        # one flattened all-reduce == per-tensor mean of the ranks' gradients
        torch.manual_seed(7 + rank)
        x = torch.randn(4, 5)
        m(x).square().sum().backward()
        local = [p.grad.clone() for p in m.parameters()]
        n = parallel.FlatGradAllReduce(m.parameters())()
        assert n == sum(p.numel() for p in m.parameters())
        for p, g in zip(m.parameters(), local):
            both = [torch.zeros_like(g) for _ in range(world_size)]
            dist.all_gather(both, g)
            assert torch.allclose(p.grad, sum(both) / world_size, atol=1e-6)
        # the two-bucket form: gradients are views of one flat buffer, the `early` parameters are exchanged from their
        # post-accumulate hooks DURING backward, the rest afterwards; same result
        torch.manual_seed(50)
        m2 = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.GroupNorm(1, 7), torch.nn.Linear(7, 3))
        ar = parallel.FlatGradAllReduce(m2.parameters(), early=list(m2[2].parameters()))
        for step in range(2):      # second step: the views are reused, zero_grad(set_to_none=False) keeps them
            for q in m2.parameters():
                if q.grad is not None:
                    q.grad.zero_()
            torch.manual_seed(11 + rank + 10 * step)
            x2 = torch.randn(4, 5)
            m2r = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.GroupNorm(1, 7), torch.nn.Linear(7, 3))
            m2r.load_state_dict(m2.state_dict())
            m2r(x2).square().sum().backward()
            ar.prepare()
            m2(x2).square().sum().backward()
            assert ar._early_left == 0 and ar.n_early == sum(q.numel() for q in m2[2].parameters())
            ar()
            flat_ptr = ar._flat.data_ptr()
            assert all(flat_ptr <= q.grad.data_ptr() < flat_ptr + 4 * ar.n for q in m2.parameters())
            for q, qr in zip(m2.parameters(), m2r.parameters()):
                both = [torch.zeros_like(qr.grad) for _ in range(world_size)]
                dist.all_gather(both, qr.grad)
                assert torch.allclose(q.grad, sum(both) / world_size, atol=1e-6)
        # identical AdamW steps on every rank keep the replicas in lock-step
        opt = torch.optim.AdamW(m.parameters(), lr=1e-3, weight_decay=1e-3, amsgrad=True)
        opt.step()
        flat = torch.cat([p.detach().reshape(-1) for p in m.parameters()])
        dist.all_gather(ref, flat)
        assert all(torch.equal(ref[0], r) for r in ref)
        out.put((rank, "ok"))
    except Exception as e:  # surfaced by the parent
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding_broadcast_and_flat_allreduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_single_process_defaults():
    from temporal_latticenet_b200 import parallel
    assert parallel.world() == (0, 1)
    assert parallel.shard_windows(5) == [0, 1, 2, 3, 4]
    assert parallel.shard_windows(5, 1, 2) == [1, 3]
    m = torch.nn.Linear(3, 2)
    m(torch.ones(1, 3)).sum().backward()
    g = m.weight.grad.clone()
    assert parallel.FlatGradAllReduce(m.parameters())() == 8
    assert torch.equal(m.weight.grad, g)
