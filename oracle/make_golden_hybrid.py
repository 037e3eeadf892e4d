"""Generate tests/golden/hybrid_*.npz from the UNMODIFIED reference's DynaDQN / DynaDSR (build container only).

Usage:  python -m oracle.make_golden_hybrid
Each fixture drives agent/dyna_q.py:333-1150 with the reference's TorchNetwork (CPU, fp64, Adam) under the Philox
stream of one agent id and stores the trajectory, the replayed experiences, the draw count and the final Q-value
predictions for all states.  The initial weights are ``oracle.hybrid_models.seeded(out, seed)``.
"""
import os

import numpy as np

from . import cases, ref_loader
from .hybrid_models import seeded
from .philox import LazyStream
from .ref_runs import _Capture, _gridworld, _policy
from .stream_rng import StreamRNG

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')

# name -> (kind, agent id, kwargs)
HYBRID_CASES = {
    'hybrid_dqn_soft': ('dqn', 3, dict(trials=4, steps=12, batch=8, target_update=0.01, ddqn=False, mask=False)),
    'hybrid_dqn_ddqn_hard': ('dqn', 5, dict(trials=4, steps=12, batch=8, target_update=3, ddqn=True, mask=True)),
    'hybrid_dsr_sr': ('dsr', 7, dict(trials=3, steps=10, batch=8, target_update=0.01, use_DR=False,
                                     use_follow_up_state=False, ignore_terminality=True)),
    'hybrid_dsr_dr_hard': ('dsr', 9, dict(trials=3, steps=10, batch=8, target_update=2, use_DR=True,
                                          use_follow_up_state=True, ignore_terminality=False)),
}
WORLD = 'open5'
MODEL_SEED = 1234


def run(name):
    import torch
    torch.set_num_threads(1)
    cobel = ref_loader.load()
    from cobel.network import TorchNetwork
    kind, agent_id, kw = HYBRID_CASES[name]
    h, w, wkw = cases.world_args(WORLD)
    world = cobel.misc.gridworld_tools.make_gridworld(h, w, **wkw)
    rng = StreamRNG(LazyStream(cases.SEED, agent_id))
    env = _gridworld(cobel, world, rng)
    cap = _Capture()
    mem = cobel.memory.DynaQMemory(world['states'], 4, 0.9, rng=rng)
    orig = mem.retrieve_batch

    def retrieve_batch(n=32):
        b = orig(n)
        cap.replay.extend(int(e['state']) * 4 + int(e['action']) for e in b)
        cap.replay_len.append(len(b))
        return b
    mem.retrieve_batch = retrieve_batch
    pol, pol_test = _policy(cobel, ('eps', 0.1), rng), _policy(cobel, ('eps', 0.0), rng)
    if kind == 'dqn':
        agent = cobel.agent.DynaDQN(env.observation_space, env.action_space, pol, TorchNetwork(seeded(4, MODEL_SEED + agent_id)),
                                    policy_test=pol_test, gamma=0.9, memory=mem, custom_callbacks=cap.callbacks())
        agent.DDQN = kw['ddqn']
        agent.mask_actions = kw['mask']
        if kw['mask']:
            from .tabular import valid_move_mask
            agent.action_mask = valid_move_mask(np.argmax(world['sas'], axis=2)).astype(bool)
    else:
        agent = cobel.agent.DynaDSR(env.observation_space, env.action_space, pol,
                                    TorchNetwork(seeded(25, MODEL_SEED + agent_id)), TorchNetwork(seeded(1, MODEL_SEED + 100 + agent_id)),
                                    policy_test=pol_test, gamma=0.9, memory=mem, custom_callbacks=cap.callbacks())
        agent.use_DR, agent.use_follow_up_state = kw['use_DR'], kw['use_follow_up_state']
        agent.ignore_terminality = kw['ignore_terminality']
    agent.target_update = kw['target_update']
    agent.train(env, kw['trials'], kw['steps'], kw['batch'])
    q_train = np.array(agent.predict_on_batch(np.arange(world['states'])))
    draws_train = rng.k
    n_train = len(cap.s)
    agent.test(env, 2, kw['steps'])
    out = cap.arrays()
    out.update(q_train=q_train, draws_train=draws_train, n_train_steps=n_train, draws=rng.k,
               Mr=mem.rewards, Ms=mem.states, Mt=mem.terminals)
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    for name in HYBRID_CASES:
        out = run(name)
        np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
        print('%-24s steps=%4d replays=%5d draws=%6d |Q|max=%.4f' % (name, len(out['states']), len(out['replay']), out['draws'],
                                                                  np.abs(out['q_train']).max()))


if __name__ == '__main__':
    main()
