"""Multi-GPU plumbing: agents are independent, so they shard across ranks as contiguous
ranges with NO per-step communication; the only collective is one final all-gather of
per-agent statistics.  One process per GPU, ``torch.distributed``
(NCCL on GPUs, gloo in the CPU tests).  Stream ids are global agent ids, so results do not
depend on the number of ranks.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise the default process group from torchrun's environment (no-op for 1 process).
    Returns (rank, world_size, local_rank)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        kw = {}
        if backend == 'nccl':
            torch.cuda.set_device(local)
            kw['device_id'] = torch.device('cuda', local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def shard_range(n_total, rank, world):
    """Contiguous agent range ``[lo, hi)`` of ``rank``: sizes differ by at most one."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_agents(t, n_total=None):
    """All-gather a per-agent tensor ``[n_local, ...]`` into ``[n_total, ...]`` in global agent
    order (ranks may hold ranges that differ by one agent)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return t
    world = dist.get_world_size()
    if n_total is None:
        cnt = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
        dist.all_reduce(cnt)
        n_total = int(cnt.item())
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    pad = max(hi - lo for lo, hi in sizes)
    buf = t
    if t.shape[0] < pad:
        buf = torch.cat([t, t.new_zeros((pad - t.shape[0],) + tuple(t.shape[1:]))], dim=0)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf.contiguous())
    return torch.cat([o[:hi - lo] for o, (lo, hi) in zip(out, sizes)], dim=0)


def gather_results(res, n_total=None, keys=('trial_steps', 'trial_reward', 'n_steps', 'n_replay')):
    """The final collective: all-gather the per-agent statistics of a RunResult -- ONE collective for all keys (the
    fields of an agent are packed into one byte record, ``all_gather_into_tensor`` when the shards are equal)."""
    keys = [k for k in keys if k in res]
    if not dist.is_initialized() or dist.get_world_size() == 1 or not keys:
        return {k: res[k] for k in keys}
    n_local = res[keys[0]].shape[0]
    parts = [res[k].contiguous().reshape(n_local, -1).view(torch.uint8) for k in keys]
    widths = [p.shape[1] for p in parts]
    rec = torch.cat(parts, dim=1) if len(parts) > 1 else parts[0]
    world = dist.get_world_size()
    if n_total is not None and n_total == n_local * world:
        out = torch.empty((n_total, rec.shape[1]), dtype=torch.uint8, device=rec.device)
        dist.all_gather_into_tensor(out, rec.contiguous())
    else:
        out = gather_agents(rec, n_total)
    got, off = {}, 0
    for k, w in zip(keys, widths):
        t = res[k]
        got[k] = out[:, off:off + w].contiguous().view(t.dtype).reshape((out.shape[0],) + tuple(t.shape[1:]))
        off += w
    return got


def max_over_ranks(x, device):
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x, device):
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
