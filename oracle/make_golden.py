"""Generate tests/golden/*.npz from the UNMODIFIED reference (build container only).

Test infrastructure (see oracle/__init__.py).  Usage:  python -m oracle.make_golden
Each fixture holds the reference's trajectory (states, actions, rewards), replay
indices, per-trial stats, draw count and final fp64 tables for one case of
oracle/cases.py, produced by driving /root/reference under the Philox stream
(seed, agent) through oracle/ref_runs.py.  Inputs that are expensive to recompute
and numerically library-dependent (the DR similarity matrix D) are stored too.
"""
import os
import sys

import numpy as np

from . import cases, ref_loader, ref_runs
from .philox import LazyStream

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def build_world(cobel, name):
    h, w, kw = cases.world_args(name)
    kw = dict(kw)
    slip = kw.pop('slippery', None)
    world = cobel.misc.gridworld_tools.make_gridworld(h, w, **kw)
    return cases.make_slippery(world, slip) if slip else world


def main(only=None):
    cobel = ref_loader.load()
    os.makedirs(OUT, exist_ok=True)
    for name, (kind, wname, agent, args) in cases.CASES.items():
        if only and name not in only:
            continue
        u = LazyStream(cases.SEED, agent)
        a = dict(args)
        trials, steps = a.pop('trials'), a.pop('steps')
        extra = {}
        if a.pop('valid_mask', False):       # action mask: moves that change the state
            from .tabular import valid_move_mask
            succ = np.argmax(build_world(cobel, wname)['sas'], axis=2)
            a['action_mask'] = valid_move_mask(succ)
        if kind == 'dynaq':
            out = ref_runs.run_dynaq(build_world(cobel, wname), u, trials, steps, a.pop('batch'),
                                     policy_test=('eps', 0.0), test_trials=5, **a)
        elif kind == 'q_grid':
            out = ref_runs.run_q_gridworld(build_world(cobel, wname), u, trials, steps, a.pop('batch'), **a)
        elif kind == 'q_topo':
            fn, fargs = wname
            nodes, starts = getattr(cobel.misc.topology_tools, fn)(*fargs)
            out = ref_runs.run_q_topology(nodes, starts, u, trials, steps, a.pop('batch'), **a)
        elif kind == 'sr':
            out = ref_runs.run_sr(build_world(cobel, wname), u, trials, steps, **a)
        elif kind == 'sfma':
            world = build_world(cobel, wname)
            D = cobel.memory.utils.metrics.DR(world['width'], world['height'], world['sas'], 0.9,
                                              world['invalid_transitions']).D
            extra['D'] = D
            out = ref_runs.run_sfma(world, D, u, trials, steps, a.pop('batch'), **a)
        elif kind == 'pma':
            out = ref_runs.run_pma(build_world(cobel, wname), u, trials, steps, a.pop('batch'), **a)
        else:
            raise ValueError(kind)
        out.update(extra)
        np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
        print('%-32s steps=%5d replays=%6d draws=%7d' % (name, len(out['states']), len(out['replay']), out['draws']))


if __name__ == '__main__':
    main(set(sys.argv[1:]) or None)
