"""Shared definitions of the small parity cases (worlds, hyper-parameters, seeds).

Test infrastructure (see oracle/__init__.py).  Used by make_golden.py (reference
side, build container) and by the tests (oracle / CUDA side), so both always run
the same configuration.  Worlds are described by builder arguments only, which
keeps the fixtures small.
"""
import numpy as np

SEED = 0x5EED

WALLS_5x5 = [(1, 2), (2, 1), (6, 7), (7, 6), (11, 12), (12, 11), (13, 8), (8, 13)]
# asymmetric wall set of config C3
WALLS_10x10 = [(4, 5), (5, 4), (14, 15), (15, 14), (24, 25), (25, 24), (34, 35), (35, 34),
               (62, 72), (72, 62), (63, 73), (73, 63)]


def world_args(name):
    """(height, width, kwargs) of make_gridworld for a named world."""
    if name == 'open5':      # demo/gridworld/demo_dyna_q.py:41
        return 5, 5, dict(terminals=[0], rewards=np.array([[0, 1]]), goals=[0])
    if name == 'walls5':     # demo/gridworld/demo_pma.py:35-60 shape
        return 5, 5, dict(terminals=[4], rewards=np.array([[4, 10]]), goals=[4], starting_states=[12],
                          invalid_transitions=WALLS_5x5)
    if name == 'open13':
        return 13, 13, dict(terminals=[0], rewards=np.array([[0, 1]]), goals=[0])
    if name == 'walls10':    # config C3
        return 10, 10, dict(terminals=[9], rewards=np.array([[9, 10]]), goals=[9], starting_states=[57],
                            invalid_transitions=WALLS_10x10)
    if name == 'track10x2':  # config C4's linear track as a gridworld
        return 2, 10, dict(terminals=[9, 19], rewards=np.array([[9, 1], [19, 1]]), goals=[9, 19],
                           starting_states=[0, 10])
    if name == 'slip5':      # non-deterministic transitions (interface/gridworld.py:118-123)
        return 5, 5, dict(terminals=[0], rewards=np.array([[0, 1]]), goals=[0], slippery=0.25)
    if name == 'open8':
        return 8, 8, dict(terminals=[0], rewards=np.array([[0, 1]]), goals=[0])
    raise KeyError(name)


def make_slippery(world, slip=0.2):
    """Turn a deterministic WorldDict into a non-deterministic one (in place): with probability
    ``slip`` the move of a uniformly random action is executed instead of the chosen one."""
    sas = np.asarray(world['sas'])
    world['sas'] = (1.0 - slip) * sas + slip * np.mean(sas, axis=1, keepdims=True)
    world['deterministic'] = False
    return world


# name -> (kind, world, agent index in the stream, run arguments)
CASES = {
    # Dyna-Q (demo/gridworld/demo_dyna_q.py:36-56; unit_tests/test_dyna_q.py:13-38)
    'dynaq_open5_eps':      ('dynaq', 'open5', 0, dict(trials=40, steps=50, batch=32, policy=('eps', 0.1))),
    'dynaq_open5_softmax':  ('dynaq', 'open5', 1, dict(trials=15, steps=50, batch=32, policy=('softmax', 2.0))),
    'dynaq_open5_xeps':     ('dynaq', 'open5', 2, dict(trials=25, steps=50, batch=8, policy=('xeps', 0.2))),
    'dynaq_walls5_mask':    ('dynaq', 'walls5', 3, dict(trials=25, steps=30, batch=32, policy=('eps', 0.1),
                                                        mask_actions=True, valid_mask=True, lr=0.9, gamma=0.95,
                                                        mem_lr=0.5)),
    'dynaq_open5_episodic': ('dynaq', 'open5', 4, dict(trials=25, steps=20, batch=16, policy=('eps', 0.3),
                                                       episodic_replay=True)),
    'dynaq_open5_noreplay': ('dynaq', 'open5', 5, dict(trials=25, steps=20, batch=32, policy=('eps', 0.1),
                                                       no_replay=True)),
    'dynaq_slip5':          ('dynaq', 'slip5', 23, dict(trials=25, steps=40, batch=32, policy=('eps', 0.1))),
    # QAgent (unit_tests/test_q.py:46-77; demo/topology/demo.py:40-103)
    'q_open5':              ('q_grid', 'open5', 6, dict(trials=25, steps=50, batch=32, policy=('eps', 0.1))),
    'q_track':              ('q_topo', ('linear_track', (10, 2, 1.0, 20.0, 'right')), 7,
                             dict(trials=25, steps=50, batch=8, policy=('eps', 0.1))),
    'q_tmaze_nobatch':      ('q_topo', ('t_maze', (6, 3, 2)), 8, dict(trials=25, steps=50, batch=0, policy=('eps', 0.1))),
    'q_grid5_softmax':      ('q_topo', ('grid', (5,)), 9, dict(trials=20, steps=40, batch=16, policy=('softmax', 3.0))),
    # SR (demo/gridworld/demo_sr.py:36-56; unit_tests/test_sr.py:13-37)
    'sr_open5':             ('sr', 'open5', 10, dict(trials=30, steps=50, policy=('eps', 0.1))),
    'sr_open13':            ('sr', 'open13', 11, dict(trials=10, steps=120, policy=('eps', 0.1))),
    'sr_walls5_mask':       ('sr', 'walls5', 12, dict(trials=20, steps=40, policy=('eps', 0.2), mask_actions=True,
                                                      valid_mask=True, lr=0.3, gamma=0.9)),
    'sr_slip5':             ('sr', 'slip5', 24, dict(trials=20, steps=40, policy=('eps', 0.1))),
    # SFMA (demo/gridworld/demo_sfma.py:36-80; unit_tests/test_sfma.py:20-97)
    'sfma_walls5_default':  ('sfma', 'walls5', 13, dict(trials=25, steps=50, batch=32, mode='default', mask_actions=True,
                                                        valid_mask=True)),
    'sfma_walls5_reverse':  ('sfma', 'walls5', 14, dict(trials=25, steps=50, batch=32, mode='reverse', mask_actions=True,
                                                        valid_mask=True)),
    'sfma_walls5_timeout':  ('sfma', 'walls5', 15, dict(trials=20, steps=6, batch=16, mode='default', mask_actions=False)),
    'sfma_track_reverse':   ('sfma', 'track10x2', 16, dict(trials=20, steps=40, batch=32, mode='reverse', mask_actions=True)),
    'sfma_open8_forward':   ('sfma', 'open8', 17, dict(trials=12, steps=80, batch=32, mode='forward', mask_actions=False)),
    'sfma_walls5_sweeping_recency': ('sfma', 'walls5', 18, dict(trials=20, steps=50, batch=32, mode='sweeping',
                                                               mask_actions=True, recency=True)),
    'sfma_walls5_random':   ('sfma', 'walls5', 27, dict(trials=15, steps=30, batch=32, mode='default', mask_actions=True,
                                                       valid_mask=True, random_replay=True)),
    'sfma_slip5':           ('sfma', 'slip5', 25, dict(trials=15, steps=40, batch=16, mode='default', mask_actions=False)),
    # agent.dynamic (agent/sfma.py:311-318): 'reverse' / 'default' chosen per trial from the accumulated TD error
    'sfma_walls5_dynamic':  ('sfma', 'walls5', 31, dict(trials=25, steps=50, batch=32, mode='default', mask_actions=True,
                                                       valid_mask=True, dynamic=True)),
    # PMA (demo/gridworld/demo_pma.py:29-72; unit_tests/test_pma.py:15-78)
    'pma_walls5':           ('pma', 'walls5', 19, dict(trials=6, steps=50, batch=16, gamma_q=0.99, mask_actions=True,
                                                       valid_mask=True)),
    'pma_walls5_timeout':   ('pma', 'walls5', 20, dict(trials=6, steps=6, batch=16, gamma_q=0.99, mask_actions=True)),
    'pma_walls5_prefill':   ('pma', 'walls5', 21, dict(trials=6, steps=30, batch=16, gamma_q=0.99, mask_actions=True,
                                                       prefill=True)),
    'pma_slip5':            ('pma', 'slip5', 26, dict(trials=5, steps=30, batch=16, gamma_q=0.99, mask_actions=False)),
    'pma_walls10':          ('pma', 'walls10', 22, dict(trials=3, steps=100, batch=32, gamma_q=0.99, mask_actions=True,
                                                        valid_mask=True)),
}
