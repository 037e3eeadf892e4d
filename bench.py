#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native tabular simulator.

Metric (BASELINE.json): agent-steps/sec including replay.  Workload at N GPUs
(weak scaling): BASELINE.json configs[1] per GPU -- 4096 independent Dyna-Q
agents on the 5x5 open gridworld, the reference demo's setting
(demo/gridworld/demo_dyna_q.py:36-56: 500 trials x <=50 steps, replay batch 32,
epsilon 0.1, lr 0.99, gamma 0.99).  One "step" of this benchmark = one complete
``DynaQ.train(env, 500, 50, 32)`` of all agents from freshly initialised tables.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload dynaq|...]

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

SEED = 0x5EED
# algorithmic bytes per unit, SURVEY.md section 8d / DESIGN.md
DYNAQ_BYTES_PER_STEP = 2066

WORKLOADS = {
    'dynaq': dict(desc='C2: 4096 Dyna-Q agents/GPU, 5x5 open field, 500 trials x <=50 steps, batch 32',
                  agents_per_gpu=4096, trials=500, steps=50, batch=32, S=25, A=4),
}


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs, burst copy)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(smax) if smax else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


# --------------------------------------------------------------------------- #
# CPU baseline / reference arm: the reference's algorithm (oracle port) on host cores
# --------------------------------------------------------------------------- #

def _cpu_worker(args):
    agent_ids, trials, steps, batch = args
    os.environ['OPENBLAS_NUM_THREADS'] = '1'
    from oracle import tabular as tb
    from oracle.philox import LazyStream
    from oracle.cases import world_args
    from cobel_rl_b200.misc.gridworld_tools import make_gridworld
    h, w, kw = world_args('open5')
    W = tb.compile_gridworld(make_gridworld(h, w, **kw))
    total = 0
    for g in agent_ids:
        rng = tb.Draws(LazyStream(SEED, g), 1)
        st = tb.dynaq_init(W['S'], W['A'])
        rec = tb.dynaq_train(W, st, rng, trials, steps, batch)
        total += len(rec.s)
    return total


def cpu_dynaq(wl, agents_per_core=2, pool=None, cores=None):
    """Oracle port of the reference loop, one process per host core, each running a slice of
    agents of the SAME workload (same world, hyper-parameters and per-agent streams)."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    ids = [list(range(c * agents_per_core, (c + 1) * agents_per_core)) for c in range(cores)]
    own = pool is None
    if own:
        pool = mp.get_context('spawn').Pool(cores)
        pool.map(_cpu_worker, [([0], 2, 5, 4)] * cores)       # import / warm the workers
    t0 = time.perf_counter()
    steps = sum(pool.map(_cpu_worker, [(i, wl['trials'], wl['steps'], wl['batch']) for i in ids]))
    dt = time.perf_counter() - t0
    if own:
        pool.close()
    return {'value': steps / dt, 'unit': 'agent-steps/s', 'cores': cores, 'kind': 'port',
            'sample': '%d agents (%d per core) of the same workload, full %d trials x <=%d steps, batch %d; '
                      '%d agent-steps in %.1f s; oracle/tabular.py (NumPy restatement validated bit-exact '
                      'against the reference)' % (cores * agents_per_core, agents_per_core, wl['trials'],
                                                   wl['steps'], wl['batch'], steps, dt),
            'seconds': dt, 'agent_steps': steps}


def run_reference(args, wl):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    pool = mp.get_context('spawn').Pool(cores)
    pool.map(_cpu_worker, [([0], 2, 5, 4)] * cores)
    for _ in range(args.warmup):
        cpu_dynaq(wl, 1, pool, cores)
    t_tot, s_tot, last = 0.0, 0, None
    for _ in range(args.steps):
        last = cpu_dynaq(wl, 1, pool, cores)
        t_tot += last['seconds']; s_tot += last['agent_steps']
    pool.close()
    val = s_tot / t_tot
    cb = dict(last); cb['value'] = val
    cb.pop('seconds'); cb.pop('agent_steps')
    print(json.dumps({
        'impl': 'reference', 'metric': 'agent-steps/sec incl. replay', 'value': val, 'unit': 'agent-steps/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * t_tot / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': wl['desc'], 'note': 'each step = one agent per host core run through the full workload'},
        'cpu_baseline': cb,
        'e2e': {'value': val, 'unit': 'agent-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }))


# --------------------------------------------------------------------------- #
# our arm
# --------------------------------------------------------------------------- #

def run_ours(args, wl):
    import torch
    import cobel_rl_b200 as cb
    from cobel_rl_b200 import _lib, dist as cdist
    from cobel_rl_b200.agent import DynaQ
    from cobel_rl_b200.interface import Gridworld
    from cobel_rl_b200.memory import DynaQMemory
    from cobel_rl_b200.misc.gridworld_tools import make_open_field
    from cobel_rl_b200.policy import EpsilonGreedy

    rank, world, local = cdist.init_from_env()
    assert world == args.gpus, 'launch with torchrun --nproc-per-node %d for --gpus %d' % (args.gpus, args.gpus)
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    L = _lib.lib()
    n_local = wl['agents_per_gpu']
    n_total = n_local * world
    lo, hi = cdist.shard_range(n_total, rank, world)
    S, A, trials, steps, batch = wl['S'], wl['A'], wl['trials'], wl['steps'], wl['batch']

    stream = cb.BatchStream(n_local, seed=SEED, device=dev, agent_id_base=lo)
    env = Gridworld(make_open_field(5, 5, 0, 1), rng=stream)
    mem = DynaQMemory(S, A, 0.9, rng=stream)
    agent = DynaQ(env.observation_space, env.action_space, EpsilonGreedy(0.1, rng=stream), None, 0.99, 0.99, mem)

    # pristine tables: device copies for the kernel-only loop, pinned host copies for the e2e loop
    init = {'Q': agent._Q.clone(), 'Mr': mem._rewards.clone(), 'Ms': mem._states.clone(), 'Mt': mem._terminals.clone()}
    host_in = {k: v.cpu().pin_memory() for k, v in init.items()}
    live = {'Q': agent._Q, 'Mr': mem._rewards, 'Ms': mem._states, 'Mt': mem._terminals}
    host_out = {'Q': torch.empty_like(host_in['Q']).pin_memory(),
                'trial_steps': torch.empty((n_local, trials), dtype=torch.int32).pin_memory(),
                'trial_reward': torch.empty((n_local, trials), dtype=torch.float64).pin_memory()}
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def reset_device():
        for k in live:
            live[k].copy_(init[k])
        stream.draw_count.fill_(1)          # draw 0 was consumed by the environment constructor

    def step_device():
        reset_device()
        return agent.train(env, trials, steps, batch)

    def step_e2e():
        for k in live:
            live[k].copy_(host_in[k], non_blocking=True)
        stream.draw_count.fill_(1)
        res = agent.train(env, trials, steps, batch)
        host_out['Q'].copy_(agent._Q, non_blocking=True)
        host_out['trial_steps'].copy_(res['trial_steps'], non_blocking=True)
        host_out['trial_reward'].copy_(res['trial_reward'], non_blocking=True)
        return res

    def timed(fn, k):
        """K steps; device time per step from CUDA events on the launching (current) stream,
        L2 flushed between steps outside the event pairs; wall clock kept as a cross-check."""
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
        res = None
        cdist.barrier(); torch.cuda.synchronize(dev)
        w0 = time.perf_counter()
        for s, e in evs:
            flush.fill_(1)
            s.record()
            res = fn()
            e.record()
        torch.cuda.synchronize(dev); cdist.barrier()
        wall = time.perf_counter() - w0
        ms = [s.elapsed_time(e) for s, e in evs]
        return ms, wall, res

    for _ in range(max(args.warmup, 3)):
        step_device()
    torch.cuda.synchronize(dev)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = L.cobel_launch_count()
    ms, wall, res = timed(step_device, args.steps)
    launches = L.cobel_launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    steps_local = float(res['n_steps'].sum().item())           # agent-steps of ONE bench step on this rank
    replays_local = float(res['n_replay'].sum().item())

    for _ in range(2):
        step_e2e()
    torch.cuda.synchronize(dev)
    ms_e2e, wall_e2e, res_e = timed(step_e2e, args.steps)
    assert float(res_e['n_steps'].sum().item()) == steps_local

    # the only collective of the path: final all-gather of per-agent statistics
    torch.cuda.synchronize(dev)
    g0 = time.perf_counter()
    gathered = cdist.gather_results(res, n_total)
    torch.cuda.synchronize(dev)
    gather_ms = 1e3 * (time.perf_counter() - g0)
    assert gathered['n_steps'].shape[0] == n_total

    t_step = cdist.max_over_ranks(sum(ms) / len(ms), dev)              # ms, max over ranks
    t_e2e = cdist.max_over_ranks(sum(ms_e2e) / len(ms_e2e), dev)
    steps_total = cdist.sum_over_ranks(steps_local, dev)
    replays_total = cdist.sum_over_ranks(replays_local, dev)
    wall_max = cdist.max_over_ranks(wall, dev)

    if rank != 0:
        return
    value = steps_total / (t_step * 1e-3)
    peak, peak_src = peaks()
    # roofline of the dominant (only) kernel, dynaq_smem_kernel: one launch per bench step per GPU
    achieved = DYNAQ_BYTES_PER_STEP * steps_local / (ms and (sum(ms) / len(ms)) * 1e-3) / 1e9
    h2d = sum(v.numel() * v.element_size() for v in host_in.values())
    d2h = sum(v.numel() * v.element_size() for v in host_out.values())
    out = {
        'metric': 'agent-steps/sec incl. replay', 'value': value, 'unit': 'agent-steps/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': t_step, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': wl['desc'], 'agents_total': n_total, 'agent_steps_per_bench_step': steps_total,
                   'replay_updates_per_bench_step': replays_total,
                   'l2': 'flushed between steps with a 256 MiB write (outside the event pairs); inputs are 7 MB',
                   'timing': 'CUDA events on the launching stream per step, mean of K, max over ranks',
                   'wall_s_incl_flush': wall_max, 'final_all_gather_ms': gather_ms, 'seed': SEED},
        'roofline': {'bound': 'hbm', 'kernel': 'dynaq_smem_kernel<4>', 'achieved': achieved, 'peak': peak,
                     'unit': 'GB/s', 'frac': achieved / peak, 'traffic': None,
                     'peak_source': peak_src, 'algorithmic_bytes_per_agent_step': DYNAQ_BYTES_PER_STEP,
                     'note': 'tables live in shared memory for the whole launch: real DRAM traffic is the one-off '
                             'stage-in/out; the kernel is bound by the serial fp64 TD-update chain (see DESIGN.md)'},
        'e2e': {'value': steps_total / (t_e2e * 1e-3), 'unit': 'agent-steps/s', 'h2d_bytes_per_step': h2d,
                'd2h_bytes_per_step': d2h, 'ms_per_step': t_e2e},
        'gpu_launches': int(launches), 'clocks': clocks,
    }
    if world == 1 and not args.no_cpu:
        cbase = cpu_dynaq(wl, 2)
        cbase.pop('seconds'); cbase.pop('agent_steps')
        out['cpu_baseline'] = cbase
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='dynaq', choices=sorted(WORKLOADS))
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == 'reference':
        run_reference(args, wl)
    else:
        run_ours(args, wl)
    try:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == '__main__':
    main()
