"""Gridworld builders (reference: misc/gridworld_tools.py:10-234).

Produces the same ``WorldDict`` (keys, dtypes and state numbering
``state = row*width + column``; actions 0 left, 1 up, 2 right, 3 down) as the
reference, built with vectorised index arithmetic instead of the reference's
S x 4 Python loop.  The successor table ``succ[S,4]`` is always included;
the dense one-hot ``sas[S,4,S]`` (needed by ``PMAMemory`` / the DR and SR
metrics) can be skipped with ``dense_sas=False`` for very large worlds
(100 x 100 would need 3.2 GB).
"""
import numpy as np


def make_gridworld(height, width, terminals=None, rewards=None, goals=None, starting_states=None,
                   invalid_states=None, invalid_transitions=None, wind=None, deterministic=True,
                   dense_sas=True):
    S = height * width
    world = {'height': height, 'width': width, 'states': S, 'goals': [] if goals is None else goals}
    world['terminals'] = np.zeros(S).astype(int)
    if terminals is not None:
        world['terminals'][terminals] = 1
    world['rewards'] = np.zeros(S).astype(float)
    if rewards is not None:
        rewards = np.asarray(rewards)
        world['rewards'][rewards[:, 0].astype(int)] = rewards[:, 1]
    # starting states default to all non-terminal states (gridworld_tools.py:76-82)
    if starting_states is not None and len(starting_states) > 0:
        world['starting_states'] = np.array(starting_states)
    else:
        term = set([] if terminals is None else terminals)
        world['starting_states'] = np.array(list(set(range(S)) - term))
    world['wind'] = np.zeros((S, 2)).astype(int)
    if wind is not None:
        wind = np.asarray(wind)
        world['wind'][wind[:, 0].astype(int)] = wind[:, 1:].astype(int)
    world['invalid_states'] = [] if invalid_states is None else invalid_states
    world['invalid_transitions'] = [] if invalid_transitions is None else invalid_transitions
    # coordinates: x = column, y counted from the bottom row (gridworld_tools.py:98-102)
    idx = np.arange(S)
    row, col = idx // width, idx % width
    world['coordinates'] = np.stack([col, height - 1 - row], axis=1).astype(float)
    # successor of every (state, action): move, apply wind, clip, then undo forbidden moves
    dh = np.array([0, -1, 0, 1])
    dw = np.array([-1, 0, 1, 0])
    h = np.clip(row[:, None] + dh[None, :], 0, height - 1) + world['wind'][:, 0:1]
    w = np.clip(col[:, None] + dw[None, :], 0, width - 1) + world['wind'][:, 1:2]
    succ = np.clip(h, 0, height - 1) * width + np.clip(w, 0, width - 1)
    blocked = np.isin(succ, np.asarray(world['invalid_states'], dtype=int))
    if len(world['invalid_transitions']) > 0:
        pairs = np.asarray(world['invalid_transitions'], dtype=np.int64).reshape(-1, 2)
        code = pairs[:, 0] * S + pairs[:, 1]
        blocked |= np.isin(idx[:, None] * S + succ, code)
    succ = np.where(blocked, idx[:, None], succ).astype(np.int32)
    world['succ'] = succ
    if dense_sas:
        sas = np.zeros((S, 4, S))
        sas[idx[:, None], np.arange(4)[None, :], succ] = 1
        world['sas'] = sas
    else:
        world['sas'] = None
    world['deterministic'] = deterministic
    return world


def make_open_field(height, width, goal_state=0, reward=1, dense_sas=True):
    """Open field with one terminal goal state (gridworld_tools.py:139-167)."""
    return make_gridworld(height, width, terminals=[goal_state], rewards=np.array([[goal_state, reward]]),
                          goals=[goal_state], dense_sas=dense_sas)


def make_empty_field(height, width):
    """Open field without goal (gridworld_tools.py:170-186)."""
    return make_gridworld(height, width)


def make_windy_gridworld(height, width, columns, goal_state=0, reward=1, direction='up'):
    """Column-wise wind applied to the row coordinate (gridworld_tools.py:189-234)."""
    sign = {'up': 1, 'down': -1}[direction]
    idx = np.arange(height * width)
    wind = np.stack([idx, np.asarray(columns)[idx % width] * sign, np.zeros_like(idx)], axis=1)
    return make_gridworld(height, width, terminals=[goal_state], rewards=np.array([[goal_state, reward]]),
                          goals=[goal_state], wind=wind)


# ---------------------------------------------------------------------------
# Maze templates (reference: misc/gridworld_tools.py:237-507).  A maze is a set of corridor cells on
# the full rectangular grid; the walls are the transitions, in both directions, between a corridor
# cell and an adjacent cell outside the corridor (the outside cells stay valid but unreachable
# states, as in the reference).  The templates produce the same WorldDict as the reference's; the
# order of `invalid_transitions` (a set semantically: make_gridworld only tests membership) is
# row-major by corridor cell here.
# ---------------------------------------------------------------------------
def _corridor_walls(height, width, corridor):
    """Both directions of every transition between a corridor cell and a 4-neighbour outside it."""
    inside = np.zeros((height, width), dtype=bool)
    for r, c in corridor:
        inside[r, c] = True
    walls = []
    for r, c in sorted(set(corridor)):
        for dr, dc in ((0, -1), (-1, 0), (0, 1), (1, 0)):
            rr, cc = r + dr, c + dc
            if 0 <= rr < height and 0 <= cc < width and not inside[rr, cc]:
                a, b = r * width + c, rr * width + cc
                walls += [(a, b), (b, a)]
    return walls


def _maze(height, width, corridor, goal_state, start_state, reward):
    return make_gridworld(height, width, [goal_state], np.array([[goal_state, reward]]), [goal_state], [start_state],
                          invalid_transitions=_corridor_walls(height, width, corridor))


def make_t_maze(stem_length, arm_length, goal_arm='right', reward=1):
    """T-maze (misc/gridworld_tools.py:237-298): the arms are the top row, the stem hangs from its middle."""
    assert stem_length > 0 and arm_length > 0, 'Stem and arm length must be greater than zero!'
    height, width = stem_length + 1, arm_length * 2 + 1
    corridor = [(0, c) for c in range(width)] + [(r, arm_length) for r in range(1, height)]
    goal = 0 if goal_arm == 'left' else width - 1
    return _maze(height, width, corridor, goal, height * width - arm_length - 1, reward)


def make_double_t_maze(stem_length, arm_length, goal_arm='right-right', reward=1):
    """Double T-maze (misc/gridworld_tools.py:301-432): a stem, a cross bar, and a T on either end of the bar."""
    assert stem_length > 0 and arm_length > 0, 'Stem and arm length must be greater than zero!'
    height, width = stem_length * 2 + 2, arm_length * 4 + 3
    mid, left, right, bar = 2 * arm_length + 1, arm_length, width - 1 - arm_length, stem_length + 1
    corridor = [(r, mid) for r in range(bar + 1, height)]                     # lower stem
    corridor += [(bar, c) for c in range(left, right + 1)]                    # cross bar
    corridor += [(r, c) for r in range(1, bar) for c in (left, right)]        # the two upper stems
    corridor += [(0, c) for c in range(width) if c != mid]                    # the four arms (top row without its middle)
    goal = {'left-left': 0, 'left-right': arm_length * 2, 'right-left': arm_length * 2 + 2,
            'right-right': arm_length * 4 + 2}.get(goal_arm, 0)
    return _maze(height, width, corridor, goal, height * width - arm_length * 2 - 2, reward)


def make_two_sided_t_maze(stem_length, arm_length, goal_arm='right-right', reward=1):
    """Two-sided T-maze (misc/gridworld_tools.py:435-506): a horizontal stem with vertical arms at both ends."""
    assert stem_length > 0 and arm_length > 0, 'Stem and arm length must be greater than zero!'
    height, width = arm_length * 2 + 1, stem_length + 2
    corridor = [(arm_length, c) for c in range(width)] + [(r, c) for r in range(height) for c in (0, width - 1)]
    goal = {'right-left': width - 1, 'left-left': width * (height - 1), 'right-right': width * height - 1}.get(goal_arm, 0)
    return _maze(height, width, corridor, goal, arm_length * width + int(stem_length / 2), reward)


def _ring(height, width):
    return [(r, c) for r in range(height) for c in range(width) if r in (0, height - 1) or c in (0, width - 1)]


def make_8_maze(center_height, lap_width, goal_location='right', reward=1):
    """Figure-8 maze (misc/gridworld_tools.py:669-735): the outer ring of the grid plus its centre column."""
    assert center_height > 0 and lap_width > 0, 'Center height and lap width must be greater than zero!'
    height, width = center_height + 2, lap_width * 2 + 3
    corridor = _ring(height, width) + [(r, lap_width + 1) for r in range(height)]
    goal = (height // 2) * width + (width - 1 if goal_location == 'right' else 0)
    return _maze(height, width, corridor, goal, lap_width + 2, reward)


def make_two_choice_t_maze(center_height, lap_width, arm_length, chirality='right', goal_location='right', reward=1):
    """Two-choice T-maze (misc/gridworld_tools.py:509-666): an outer ring whose centre column only reaches down to the
    arm row of an inner T; the T's stem rises from the bottom of the ring on the `chirality` side of the centre column
    and one of its arms ends in that column.  (Like the reference, the reward at the goal is 1 whatever `reward` says.)"""
    assert arm_length > 0, '!'
    assert center_height > 2, '!'
    assert lap_width >= arm_length * 2 + 1, '!'
    assert chirality in ['left', 'right'], 'Invalid chirality!'
    height, width = center_height + 2, lap_width * 2 + 3
    centre, arm_row = lap_width + 1, (center_height - 1) // 2 + 1
    if chirality == 'left':
        first, last, stem = centre - 2 * arm_length, centre, centre - arm_length
    else:
        first, last, stem = centre, centre + 2 * arm_length, centre + arm_length
    corridor = _ring(height, width) + [(r, centre) for r in range(1, arm_row + 1)]
    corridor += [(arm_row, c) for c in range(first, last + 1)] + [(r, stem) for r in range(arm_row + 1, height - 1)]
    goal = (height // 2) * width + (width - 1 if goal_location == 'right' else 0)
    return _maze(height, width, corridor, goal, width * (height - 1) + lap_width + arm_length, 1.0)


def make_detour_maze(width_small, height_small, width_large, height_large, reward=1):
    """Detour maze (misc/gridworld_tools.py:738-895): a straight stem from the start (bottom) to the goal (top),
    a large rectangular loop to its right and a small one to its left, which meet on one crossing row."""
    assert width_small > 0 and height_small > 0, 'Width and height of the small side piece must be greater than zero!'
    assert width_large > width_small and height_large > height_small, \
        'Width and height of the large side piece must be greater than those of the small side piece!'
    width, height = width_small + width_large + 3, height_small + height_large + 5
    stem, cross_row = width_small + 1, height_large + 2

    def rectangle(r0, r1, c0, c1):
        return [(r, c) for r in range(r0, r1 + 1) for c in range(c0, c1 + 1) if r in (r0, r1) or c in (c0, c1)]
    corridor = [(r, stem) for r in range(height)]
    corridor += rectangle(1, cross_row, stem, width - 1) + rectangle(cross_row, height - 2, 0, stem)
    return _maze(height, width, corridor, stem, stem + width * (height - 1), reward)


def make_cross_maze(arm_length, arm_width, goal_arm='top', reward=1.0):
    """Cross maze (misc/gridworld_tools.py:898-974): two bands of width `arm_width` crossing in the middle; the agent
    starts anywhere on the central square, the goal is the far end of one arm.  (Like the reference, the reward at
    the goal cells is 1 whatever `reward` says.)"""
    assert arm_length > 0 and arm_width > 0
    assert goal_arm in ('left', 'top', 'right', 'bottom'), 'Invalid goal arm!'
    size = arm_length * 2 + arm_width
    band = range(arm_length, arm_length + arm_width)
    corridor = [(r, c) for r in range(size) for c in range(size) if r in band or c in band]
    starting = [r * size + c for r in band for c in band]
    ends = {'top': [(0, c) for c in band], 'bottom': [(size - 1, c) for c in band],
            'left': [(r, 0) for r in band], 'right': [(r, size - 1) for r in band]}[goal_arm]
    terminals = [np.int64(r * size + c) for r, c in ends]
    rewards = np.ones((arm_width, 2))
    rewards[:, 0] = terminals
    return make_gridworld(size, size, terminals, rewards, terminals, starting_states=starting,
                          invalid_transitions=_corridor_walls(size, size, corridor))


def load_world(file_name):
    """A WorldDict pickled by the reference's gridworld editor (misc/gridworld_gui.py:203-239 saves / loads
    ``self.world`` with pickle) or by ``save_world``; the successor table is added if the file has none."""
    import pickle
    with open(file_name, 'rb') as f:
        world = pickle.load(f)
    if 'succ' not in world and world.get('sas') is not None:
        world['succ'] = np.argmax(np.asarray(world['sas']), axis=2).astype(np.int32)    # interface/gridworld.py:116-117
    return world


def save_world(world, file_name):
    """The counterpart of ``load_world``, in the editor's format (a pickled WorldDict)."""
    import pickle
    with open(file_name, 'wb') as f:
        pickle.dump(world, f)
