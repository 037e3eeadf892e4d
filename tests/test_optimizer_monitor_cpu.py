"""CPU: host logic of the batched grid-search back-end and the monitors (no GPU needed)."""
import numpy as np
import pytest
import torch

from cobel_rl_b200.monitor import EscapeLatencyMonitor, RewardMonitor
from cobel_rl_b200.optimizer import GridSearchOptimizer

PARAMS = {'x_1': [0, 1, 2, 3, 4], 'x_2': [0., 0.1, 0.2, 0.3, 0.4], 'x_3': np.array([0.9, 0.5, 0.6, 0.7, 0.8])}


@pytest.mark.parametrize('order', ['nested', 'systematic'])
def test_parameter_combinations_match_reference(order, reference, tmp_path):
    from cobel.optimizer import GridSearchOptimizer as Ref
    ours = GridSearchOptimizer(str(tmp_path) + '/', PARAMS, 2, order, rng=np.random.default_rng(4))
    ref = Ref(str(tmp_path) + '/', PARAMS, 2, order, rng=np.random.default_rng(4))
    assert list(ours.parameter_combinations.items()) == list(ref.parameter_combinations.items())


def test_shuffled_order_is_a_permutation(tmp_path):
    # (the reference's 'shuffled' order raises AttributeError: it uses self.rng before assigning it,
    #  optimizer/grid_search.py:103 vs 107/161)
    nested = GridSearchOptimizer(str(tmp_path) + '/', PARAMS, 1, 'nested')
    shuffled = GridSearchOptimizer(str(tmp_path) + '/', PARAMS, 1, 'shuffled', rng=np.random.default_rng(4))
    assert set(shuffled.parameter_combinations) == set(nested.parameter_combinations)
    assert list(shuffled.parameter_combinations) != list(nested.parameter_combinations)


def test_batched_fit_equals_per_run_fit(tmp_path):
    """The doc example of the reference (grid_search.py:66-83): same fit dict as the sequential loop."""
    calls = []

    def sim_batch(task, params):
        calls.append(len(params['_run']))
        return params['x_1'] + params['x_2'] ** 2 + params['x_3'] ** 3 + 0.0 * params['_run']

    def loss(data_sim, data_exp):
        return sum((np.mean(data_sim[t]) - data_exp[t]) ** 2 for t in data_sim) / len(data_sim)

    data = {'task_1': 2 + 0.1 ** 2 + 0.7 ** 3}
    opt = GridSearchOptimizer(str(tmp_path) + '/', PARAMS, nb_runs=3)
    fit = opt.fit(sim_batch, {'task_1': {}}, data, loss, store_simulation_data=True)
    assert calls == [125 * 3]                               # one batched call: all combinations x runs
    assert len(fit) == 125 and min(fit, key=fit.get) == (2, 0.1, 0.7) and fit[(2, 0.1, 0.7)] < 1e-30
    for combo, f in fit.items():
        expect = (combo[0] + combo[1] ** 2 + combo[2] ** 3 - data['task_1']) ** 2
        assert f == pytest.approx(expect, rel=1e-12, abs=1e-30)
    # resume: a second optimizer on the same directory finds fit.pkl and runs nothing
    calls.clear()
    opt2 = GridSearchOptimizer(str(tmp_path) + '/', PARAMS, nb_runs=3)
    assert opt2.fit(sim_batch, {'task_1': {}}, data, loss) == fit and calls == []
    assert opt2.recompute_fit(data, loss, overwrite=True) == fit
    # chunked calls
    opt3 = GridSearchOptimizer(str(tmp_path) + '/', PARAMS, nb_runs=3, max_agents=60)
    fit3 = opt3.fit(sim_batch, {'task_1': {}}, data, loss, overwrite=True)
    assert fit3 == fit and max(calls) <= 60 and sum(calls) == 375


def test_monitors_from_callbacks_and_results():
    el, rw = EscapeLatencyMonitor(4, 50, n_agents=3), RewardMonitor(4, n_agents=3)
    for t in range(4):
        logs = {'trial': t, 'steps': torch.tensor([t, 2 * t, 49]), 'trial_reward': torch.tensor([0.0, 1.0, t / 2])}
        el.update(logs); rw.update(logs)
    assert el.get_trace().tolist() == [[0, 1, 2, 3], [0, 2, 4, 6], [49] * 4]
    assert rw.get_trace()[2].tolist() == [0.0, 0.5, 1.0, 1.5]
    res = {'trial_steps': torch.arange(12).reshape(3, 4).int(), 'trial_reward': torch.ones(3, 4, dtype=torch.float64)}
    assert EscapeLatencyMonitor(4, 50, n_agents=3).from_result(res).get_trace()[1].tolist() == [4, 5, 6, 7]
    single = EscapeLatencyMonitor(2, 50)
    single.update({'trial': 1, 'steps': 7})
    assert single.get_trace().tolist() == [0, 7]


def test_response_and_q_monitors_match_reference(reference):
    """monitor/behavior.py:212-301, 388-464: same traces as the reference's monitors fed the same logs."""
    import importlib
    import numpy as np
    from cobel_rl_b200.monitor import QMonitor, ResponseMonitor
    rb = importlib.import_module('cobel.monitor.behavior')
    rewards = [0.0, 1.0, 0.0, 2.5, 0.0, 1.0]
    ref, mine = rb.ResponseMonitor(6), ResponseMonitor(6)
    for t, r in enumerate(rewards):
        logs = {'trial': t, 'trial_reward': r}
        if t == 4:
            logs['response'] = 3
        ref.update(dict(logs)); mine.update(dict(logs))
    assert np.array_equal(ref.get_trace(), mine.get_trace()) and np.array_equal(ref.CRC, mine.get_cumulative())

    class _Agent:
        def __init__(self):
            self.k = 0

        def predict_on_batch(self, batch):
            self.k += 1
            return np.arange(len(batch) * 4, dtype=float).reshape(len(batch), 4) * self.k
    a1, a2 = _Agent(), _Agent()
    ref, mine = rb.QMonitor(3, [0, 2, 5]), QMonitor(3, [0, 2, 5])
    for t in range(3):
        ref.update({'trial': t, 'agent': a1}); mine.update({'trial': t, 'agent': a2})
    assert all(np.array_equal(x, y) for x, y in zip(ref.get_trace(), mine.get_trace())) and len(mine.get_trace()) == 3
    # batched response monitor: one row per agent
    m = ResponseMonitor(3, n_agents=2)
    for t, r in enumerate(([0.0, 1.0], [2.0, 0.0], [0.0, 0.0])):
        m.update({'trial': t, 'trial_reward': np.array(r)})
    assert np.array_equal(m.get_trace(), [[0, 1, 0], [1, 0, 0]]) and np.array_equal(m.get_cumulative(), [[0, 1, 1], [1, 1, 1]])


def test_occupancy_map_and_match(reference):
    """analysis/behavior_spatial.py:9-111 against the reference over random inputs, plus known answers."""
    import importlib
    import numpy as np
    import torch
    from cobel_rl_b200.analysis import get_occupancy_map, match
    ra = importlib.import_module('cobel.analysis.behavior_spatial')
    rs = np.random.default_rng(0)
    for k in range(120):
        w, h = rs.uniform(1, 6, 2)
        b = rs.uniform(0.2, min(w, h))
        if k % 4 == 0:
            b, w, h = [0.5, 1.0, 0.25][k % 3], float(int(w) + 1), float(int(h) + 1)
        margins = ['expand', 'include', 'ignore'][k % 3]
        trajs = [np.round(rs.uniform(0, max(w, h) + 0.5, (int(rs.integers(1, 30)), 2)), 1 if k % 2 else 6) for _ in range(3)]
        want = ra.get_occupancy_map(trajs, w, h, b, margins)
        got = get_occupancy_map(trajs, w, h, b, margins).numpy()
        assert got.shape == want.shape and np.array_equal(got, want), (k, w, h, b, margins)
    for k in range(60):
        seq, tpl = rs.integers(0, 4, int(rs.integers(1, 20))), rs.integers(0, 4, int(rs.integers(1, 8)))
        assert np.array_equal(ra.match(seq, tpl), match(seq, tpl))
    # a NaN-padded [N, T, 2] tensor gives the same map as the list of its valid rows
    pos = torch.tensor([[[0.5, 0.5], [1.5, 0.5], [float('nan')] * 2], [[0.5, 1.5], [0.5, 1.5], [1.5, 1.5]]], dtype=torch.float64)
    occ = get_occupancy_map(pos, 2.0, 2.0, 1.0)
    assert occ.tolist() == [[1.0, 2.0], [1.0, 1.0]]


def test_match_known_answers():
    import numpy as np
    from cobel_rl_b200.analysis import match
    assert match(np.array([1, 2, 3, 1, 2]), np.array([1, 2])).tolist() == [2, 0, 0, 2, 0]
    assert match(np.array([5]), np.array([5, 5, 5])).tolist() == [1]
