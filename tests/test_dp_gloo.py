"""world_size-2 gloo tests (CPU) of the data-parallel host logic: bucketed gradient all-reduce over the
contiguous gradient arena, rank-shared randomness (teacher-forcing coin, past/gap/future split)."""
import os
import random
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from inpaintnet_b200.arena import ParamArena
        from inpaintnet_b200.trainer import allreduce_grads, LatentRNNTrainer
        from inpaintnet_b200.data import SyntheticFolkDataset
        torch.manual_seed(0)
        random.seed(0)
        m = torch.nn.Sequential(torch.nn.Linear(37, 19), torch.nn.Linear(19, 5))
        for p in m[1].parameters():
            p.requires_grad = False          # frozen tail: must stay out of the reduced range
        a = ParamArena(m)
        a.zero_grad()
        a.grad[: a.n_trainable] = torch.arange(a.n_trainable, dtype=torch.float32) * (rank + 1)
        a.grad[a.n_trainable:] = 7.0 * (rank + 1)
        allreduce_grads(a, bucket_bytes=256)  # many small buckets
        exp = torch.arange(a.n_trainable, dtype=torch.float32) * sum(r + 1 for r in range(world))
        ok = torch.equal(a.grad[: a.n_trainable], exp) and bool((a.grad[a.n_trainable:] == 7.0 * (rank + 1)).all())
        ok = ok and m[0].weight.grad.data_ptr() == a.grad.data_ptr()
        # overlapped exchange: a parameter group announced early + the rest at finish(), small buckets, two rounds
        from inpaintnet_b200.trainer import GradExchange
        m2 = torch.nn.ModuleDict(dict(enc=torch.nn.Linear(11, 13), dec=torch.nn.Linear(13, 7), frozen=torch.nn.Linear(7, 3)))
        for p in m2["frozen"].parameters():
            p.requires_grad = False
        a2 = ParamArena(m2)
        ex = GradExchange(bucket_bytes=64)
        for rnd in range(2):
            a2.zero_grad()
            a2.grad[: a2.n_trainable] = torch.arange(a2.n_trainable, dtype=torch.float32) * (rank + 1 + rnd)
            ex.ready(a2, "frozen.")                      # no trainable parameter: ignored
            ex.ready(a2, "dec.")
            ex.ready(a2, "dec.")                         # announced twice: reduced once
            early = ex.n_early
            ex.finish(a2)
            exp2 = torch.arange(a2.n_trainable, dtype=torch.float32) * sum(r + 1 + rnd for r in range(world))
            ok = ok and torch.equal(a2.grad[: a2.n_trainable], exp2) and early > 0 and not ex.done and not ex.works
        lo, hi = a2.trainable_range("dec.")
        ok = ok and (lo, hi) == (a2.offset["dec.weight"], a2.n_trainable) and a2.trainable_range(("enc.weight", "dec.bias")) is None
        # rank-shared seeds -> identical TF coins and splits on every rank
        coins = [random.random() < 0.5 for _ in range(8)]
        ds = SyntheticFolkDataset(num_notes=20)

        class _T(LatentRNNTrainer):
            def __init__(self):   # split logic only
                self.dataset = ds
                self.min_num_measures_target, self.max_num_measure_target = 2, 6
                self.measure_seq_len = 24

        import inpaintnet_b200.trainer as T
        T.to_cuda_variable_long = lambda t: t.long()
        splits = []
        for _ in range(5):
            p_, f_, t_ = _T().split_score_stochastic(torch.zeros(2, 1, 384, dtype=torch.int32))
            splits.append((p_.shape[1], t_.shape[1], f_.shape[1]))
        gathered = [None] * world
        dist.all_gather_object(gathered, (coins, splits))
        ok = ok and all(g == gathered[0] for g in gathered)
        ok = ok and all(2 <= s[1] <= 6 and s[0] >= 1 and s[2] >= 2 and sum(s) == 16 for s in splits)
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_bucketed_allreduce_and_shared_randomness_world2():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r for r, _ in res) == [0, 1]
    assert all(ok for _, ok in res), res
