"""CPU (build container only: needs /root/reference): the product's world / graph builders produce
exactly the reference's WorldDicts and node dictionaries (keys, dtypes, values, neighbour order)."""
import numpy as np
import pytest

from cobel_rl_b200.misc import gridworld_tools as mg, topology_tools as mt
from cobel_rl_b200.memory.utils import metrics as mm

WALLS = [(1, 2), (2, 1), (6, 7), (7, 6), (11, 12), (12, 11), (13, 8), (8, 13)]


def same_world(a, b):
    for k in b:
        x, y = a[k], b[k]
        if isinstance(y, np.ndarray):
            assert x.dtype == y.dtype and np.array_equal(x, y), k
        else:
            assert x == y, (k, x, y)


def test_gridworld_builders(reference):
    from cobel.misc import gridworld_tools as rg
    kw = dict(terminals=[4], rewards=np.array([[4, 10]]), goals=[4], starting_states=[12], invalid_transitions=WALLS)
    same_world(mg.make_gridworld(5, 5, **kw), rg.make_gridworld(5, 5, **kw))
    same_world(mg.make_open_field(7, 4, 3, 2.5), rg.make_open_field(7, 4, 3, 2.5))
    same_world(mg.make_empty_field(3, 6), rg.make_empty_field(3, 6))
    cols = np.array([0, 0, 0, 1, 1, 1, 2, 2, 1, 0])
    same_world(mg.make_windy_gridworld(7, 10, cols, 37, 1., 'up'), rg.make_windy_gridworld(7, 10, cols, 37, 1., 'up'))
    same_world(mg.make_windy_gridworld(7, 10, cols, 37, 1., 'down'), rg.make_windy_gridworld(7, 10, cols, 37, 1., 'down'))
    same_world(mg.make_gridworld(6, 6, invalid_states=[7, 8, 20]), rg.make_gridworld(6, 6, invalid_states=[7, 8, 20]))
    big = mg.make_open_field(100, 100, 0, 1, dense_sas=False)
    assert big['sas'] is None and big['succ'].shape == (10000, 4)


def same_maze(a, b):
    """Like same_world, with `invalid_transitions` compared as the set it is (make_gridworld only tests membership)."""
    for k in b:
        x, y = a[k], b[k]
        if k == 'invalid_transitions':
            as_set = lambda w: {(int(s), int(t)) for s, t in w}
            assert as_set(x) == as_set(y) and len(x) == len(y)
        elif isinstance(y, np.ndarray):
            assert x.dtype == y.dtype and np.array_equal(x, y), k
        else:
            assert x == y, (k, x, y)


def test_maze_templates(reference):
    from cobel.misc import gridworld_tools as rg
    for stem in (1, 2, 3, 4):
        for arm in (1, 2, 3):
            for goal in ('left', 'right'):
                same_maze(mg.make_t_maze(stem, arm, goal, 2.5), rg.make_t_maze(stem, arm, goal, 2.5))
            for goal in ('left-left', 'left-right', 'right-left', 'right-right'):
                same_maze(mg.make_double_t_maze(stem, arm, goal), rg.make_double_t_maze(stem, arm, goal))
                same_maze(mg.make_two_sided_t_maze(stem, arm, goal), rg.make_two_sided_t_maze(stem, arm, goal))
    for ch in (1, 2, 3, 4, 5):
        for lw in (1, 2, 3):
            for goal in ('left', 'right'):
                same_maze(mg.make_8_maze(ch, lw, goal, 3.0), rg.make_8_maze(ch, lw, goal, 3.0))
    for ch in (3, 4, 5, 6):
        for lw, arm in ((3, 1), (4, 1), (5, 2), (7, 2), (7, 3)):
            for chirality in ('left', 'right'):
                for goal in ('left', 'right'):
                    same_maze(mg.make_two_choice_t_maze(ch, lw, arm, chirality, goal, 2.0),
                              rg.make_two_choice_t_maze(ch, lw, arm, chirality, goal, 2.0))
    for ws, hs, dw, dh in ((1, 1, 1, 1), (2, 2, 2, 1), (1, 3, 4, 2), (3, 2, 1, 3)):
        same_maze(mg.make_detour_maze(ws, hs, ws + dw, hs + dh, 2.0), rg.make_detour_maze(ws, hs, ws + dw, hs + dh, 2.0))
    for arm in (1, 2, 3, 4):
        for aw in (1, 2, 3):
            for goal in ('left', 'top', 'right', 'bottom'):
                same_maze(mg.make_cross_maze(arm, aw, goal, 2.0), rg.make_cross_maze(arm, aw, goal, 2.0))


def test_maze_templates_known_answers():
    """Runs everywhere (no reference needed): the T-maze of the reference's demo, drawn from the successor table."""
    w = mg.make_t_maze(3, 2)
    assert (w['height'], w['width']) == (4, 5) and list(w['starting_states']) == [17] and w['goals'] == [4]
    assert w['rewards'][4] == 1 and w['terminals'][4] == 1 and len(w['invalid_transitions']) == 20
    succ = w['succ']
    # stem: only up / down moves; arms: left / right along the top row, down only at the junction
    assert list(succ[17]) == [17, 12, 17, 17] and list(succ[12]) == [12, 7, 12, 17] and list(succ[7]) == [7, 2, 7, 12]
    assert list(succ[2]) == [1, 2, 3, 7] and list(succ[0]) == [0, 0, 1, 0] and list(succ[4]) == [3, 4, 4, 4]
    d = mg.make_double_t_maze(2, 1)
    assert (d['height'], d['width']) == (6, 7) and list(d['starting_states']) == [38] and d['goals'] == [6]
    assert len(d['invalid_transitions']) == 54
    t = mg.make_two_sided_t_maze(2, 2)
    assert (t['height'], t['width']) == (5, 4) and list(t['starting_states']) == [9] and t['goals'] == [19]
    assert len(t['invalid_transitions']) == 24
    e = mg.make_8_maze(3, 2)
    assert (e['height'], e['width']) == (5, 7) and list(e['starting_states']) == [4] and e['goals'] == [20]
    assert len(e['invalid_transitions']) == 40
    c = mg.make_two_choice_t_maze(3, 3, 1)
    assert (c['height'], c['width']) == (5, 9) and list(c['starting_states']) == [40] and c['goals'] == [26]
    assert len(c['invalid_transitions']) == 56
    d = mg.make_detour_maze(2, 2, 4, 3)
    assert (d['height'], d['width']) == (10, 9) and list(d['starting_states']) == [84] and d['goals'] == [3]
    assert len(d['invalid_transitions']) == 98
    x = mg.make_cross_maze(2, 2, 'left')
    assert x['height'] == x['width'] == 6 and list(x['starting_states']) == [14, 15, 20, 21] and list(x['goals']) == [12, 18]
    assert len(x['invalid_transitions']) == 32 and x['rewards'][12] == 1 and x['terminals'][18] == 1


def test_topology_builders(reference):
    from cobel.misc import topology_tools as rt
    for args in [(10, 2, 1.0, 20., 'right'), (5, 1, 0.5, 1., 'left'), (4, 3)]:
        assert mt.linear_track(*args) == rt.linear_track(*args), args
    for args in [(5,), ((3, 4),), (4, (0., 2.), 3., '5'), ((2, 5), ((0., 1.), (1., 3.)))]:
        assert mt.grid(*args) == rt.grid(*args), args
    for args in [(4, 3, 1), (6, 3, 2), (2, 2, 3, 0.5, 2., 'left')]:
        assert mt.t_maze(*args) == rt.t_maze(*args), args
    for args in [(10,), (6,), (5, (0., 2.), 3., '7'), (2,), (9, (1.0, 4.0))]:
        assert mt.hexagonal(*args) == rt.hexagonal(*args), args
    for args in [(3, 1), (2, 2, 0.5, 30.0), (4, 3, 2.0, 90.0), (1, 1)]:
        assert mt.cross(*args) == rt.cross(*args), args


def test_metrics(reference):
    from cobel.memory.utils import metrics as rm
    from cobel.misc import gridworld_tools as rg
    world = rg.make_gridworld(5, 5, terminals=[4], invalid_transitions=WALLS)
    np.testing.assert_array_equal(mm.DR(5, 5, world['sas'], 0.9, WALLS).D, rm.DR(5, 5, world['sas'], 0.9, WALLS).D)
    np.testing.assert_array_equal(mm.DR(5, 5, world['sas'], 0.9, []).D, rm.DR(5, 5, world['sas'], 0.9, []).D)
    np.testing.assert_array_equal(mm.SR(world['sas'], 0.8).D, rm.SR(world['sas'], 0.8).D)
    np.testing.assert_allclose(mm.Euclidean(4, 3).D, rm.Euclidean(4, 3).D, rtol=1e-15)


def test_pickled_world_round_trip(reference, tmp_path):
    """The gridworld editor's file format (misc/gridworld_gui.py:203-239: a pickled WorldDict)."""
    import pickle
    from cobel.misc import gridworld_tools as rg
    from cobel_rl_b200.interface.gridworld import successor_table
    ref_world = rg.make_t_maze(3, 2)                 # as the reference's editor would have saved it (no 'succ' key)
    path = tmp_path / 'maze.pkl'
    with open(path, 'wb') as f:
        pickle.dump(ref_world, f)
    world = mg.load_world(path)
    same_world(world, ref_world)
    assert np.array_equal(world['succ'], mg.make_t_maze(3, 2)['succ']) and np.array_equal(successor_table(world), world['succ'])
    mg.save_world(world, tmp_path / 'again.pkl')
    again = mg.load_world(tmp_path / 'again.pkl')
    same_world(again, world)


def test_remove_obstructed_neighbors_known_answers():
    """misc/topology_tools.py:472-505 (the reference needs shapely, which this image does not have: the predicate
    is checked against hand-derived answers on a 5 x 5 grid graph)."""
    from cobel_rl_b200.misc import topology_tools as tt
    nodes, _ = tt.grid(5, (0.0, 4.0))
    wall = [[1.4, -0.5], [1.6, -0.5], [1.6, 2.5], [1.4, 2.5]]        # between the columns x = 1 and x = 2, rows y <= 2
    upd = tt.remove_obstructed_neighbors(nodes, [wall])
    cut = {(n, i) for n in nodes for i in range(4) if nodes[n]['neighbors'][i] != upd[n]['neighbors'][i]}
    by_pos = {nodes[n]['pose'][:2]: n for n in nodes}
    want = set()
    for y in (0.0, 1.0, 2.0):
        want |= {(by_pos[(1.0, y)], 2), (by_pos[(2.0, y)], 0)}          # right of x = 1, left of x = 2
    assert cut == want
    assert all(upd[n]['neighbors'][i] == n for n, i in cut)             # obstructed edges point back to the node
    assert nodes[by_pos[(1.0, 0.0)]]['neighbors'][2] == by_pos[(2.0, 0.0)]      # the input graph is not modified
    # the edge (1,3)-(2,3) passes the wall's top edge at distance 0.5: cut from buffer 0.5 on, not below
    e = (by_pos[(1.0, 3.0)], 2)
    assert tt.remove_obstructed_neighbors(nodes, [wall], 0.49)[e[0]]['neighbors'][2] == by_pos[(2.0, 3.0)]
    assert tt.remove_obstructed_neighbors(nodes, [wall], 0.5)[e[0]]['neighbors'][2] == e[0]
    # an edge that ends inside an obstacle, a shapely-like polygon object with a hole, and the assertion
    class Ring:
        def __init__(self, c): self.coords = c
    class Poly:
        exterior = Ring([(2.5, 2.5), (4.5, 2.5), (4.5, 4.5), (2.5, 4.5), (2.5, 2.5)])
        interiors = [Ring([(3.6, 3.6), (4.4, 3.6), (4.4, 4.4), (3.6, 4.4), (3.6, 3.6)])]
    upd = tt.remove_obstructed_neighbors(nodes, [Poly()])
    assert upd[by_pos[(2.0, 3.0)]]['neighbors'][2] == by_pos[(2.0, 3.0)]        # (2,3) -> (3,3) enters the polygon
    assert upd[by_pos[(3.0, 3.0)]]['neighbors'][1] == by_pos[(3.0, 3.0)]        # inside -> inside
    assert upd[by_pos[(0.0, 0.0)]]['neighbors'] == nodes[by_pos[(0.0, 0.0)]]['neighbors']
    with pytest.raises(AssertionError):
        tt.remove_obstructed_neighbors(nodes, [wall], -1.0)
