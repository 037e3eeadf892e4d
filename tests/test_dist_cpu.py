"""CPU: the N>1 host logic (shard ranges, final all-gather) under gloo with world_size 2."""
import os
import socket

import torch
import torch.multiprocessing as mp

from cobel_rl_b200 import dist as cdist


def test_shard_ranges_partition():
    for n in (1, 7, 4096, 1000003):
        for w in (1, 2, 3, 8):
            r = [cdist.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_total, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    cdist.init_from_env('gloo')
    lo, hi = cdist.shard_range(n_total, rank, world)
    ids = torch.arange(lo, hi, dtype=torch.int64)
    res = {'trial_steps': (ids.reshape(-1, 1) * 10 + torch.arange(3)).to(torch.int32),
           'trial_reward': ids.reshape(-1, 1).to(torch.float64) * 0.5 + torch.zeros(1, 3, dtype=torch.float64),
           'n_steps': ids * 2, 'n_replay': ids * 64}
    full = cdist.gather_results(res, n_total)
    ok = (torch.equal(full['n_steps'], torch.arange(n_total) * 2)
          and torch.equal(full['trial_steps'][:, 1], (torch.arange(n_total) * 10 + 1).to(torch.int32))
          and full['trial_reward'].shape == (n_total, 3))
    tmax = cdist.max_over_ranks(1.0 + rank, torch.device('cpu'))
    tsum = cdist.sum_over_ranks(hi - lo, torch.device('cpu'))
    cdist.barrier()
    q.put((rank, bool(ok), tmax, tsum))
    torch.distributed.destroy_process_group()


def test_gather_world2_gloo():
    _gather_world2(11)      # odd: ranks hold 6 and 5 agents (padded all-gather)


def test_gather_world2_gloo_equal_shards():
    _gather_world2(12)      # equal shards: one all_gather_into_tensor of the packed per-agent records


def _gather_world2(n_total):
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert out == [(0, True, 2.0, float(n_total)), (1, True, 2.0, float(n_total))]
