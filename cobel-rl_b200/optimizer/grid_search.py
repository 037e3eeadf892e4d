"""Grid-search optimizer with a batched back-end (reference: optimizer/grid_search.py:16-312).

The reference runs ``simulation(task, params)`` once per parameter combination and run,
optionally through a process pool (grid_search.py:239-252) -- the only place where it runs
"N independent agents".  Here all combinations x ``nb_runs`` of a task are mapped onto the
agent axis of ONE batched simulation:

    simulation(task, params) -> results

where ``params`` maps every parameter name to an array with one entry per agent
(``params['_combination']`` / ``params['_run']`` give the combination and run index of each
agent) and ``results`` is indexable along that agent axis.  A typical simulation builds a
``BatchStream(len(params['_run']))``, passes the parameter arrays as per-agent
hyper-parameters and returns e.g. ``res.trial_steps``.

Everything else follows the reference: the same constructor, parameter-combination orders,
``fit.pkl`` / ``sim_<values>.pkl`` resume files and the same ``loss(simulation_data, data)``
contract (``simulation_data[task]`` is the list of the ``nb_runs`` results of a combination).
"""
import copy
import pickle
from itertools import product
from os import listdir
from os.path import isfile, join

import numpy as np


class GridSearchOptimizer:
    def __init__(self, file_path, parameters, nb_runs=1, order='nested', rng=None, max_agents=None):
        self.parameters = copy.deepcopy(parameters)
        for name in self.parameters:                      # grid_search.py:99-102
            if type(self.parameters[name]) is np.ndarray:
                self.parameters[name] = np.sort(self.parameters[name])
        self.rng = np.random.default_rng() if rng is None else rng
        self.prepare_parameter_combinations(self.parameters, order)
        self.file_path = file_path
        self.nb_runs = nb_runs
        self.max_agents = max_agents                      # upper bound on agents per batched call (None: all)
        self.present_files = [f for f in listdir(self.file_path) if isfile(join(self.file_path, f))]

    def prepare_parameter_combinations(self, parameters, order='nested'):
        """Ordered dict of combinations (grid_search.py:113-171): 'nested' = Cartesian product in the
        given value order, 'shuffled' = the same after shuffling each value list, 'systematic' =
        coarse-to-fine sub-grids (stride halving per level), duplicates dropped."""
        assert order in ['nested', 'shuffled', 'systematic'], 'Invalid order!'
        names = list(parameters.keys())
        values = [list(parameters[n]) for n in names]
        grids = []
        if order == 'systematic':
            strides = []
            for v in values:
                levels = int(np.ceil(np.sqrt(len(v))))
                st = [max(int(len(v) / (2 ** (lvl + 1))), 1) for lvl in range(levels)]
                if 1 not in st:
                    st.append(1)
                strides.append(st)
            depth = max(len(st) for st in strides)
            strides = [st + [1] * (depth - len(st)) for st in strides]
            for lvl in range(depth):
                grids.append(product(*[[v[k * st[lvl]] for k in range(len(v) // st[lvl])]
                                       for v, st in zip(values, strides)]))
        else:
            if order == 'shuffled':
                for n in names:
                    self.rng.shuffle(parameters[n])
                values = [list(parameters[n]) for n in names]
            grids.append(product(*values))
        self.parameter_combinations = {}
        for grid in grids:
            for combo in grid:
                if tuple(combo) not in self.parameter_combinations:
                    self.parameter_combinations[tuple(combo)] = dict(zip(names, combo))

    def _file_name(self, combo):
        return 'sim' + ('_%s' * len(combo)) % combo + '.pkl'

    def fit(self, simulation, tasks, data, loss, overwrite=False, store_simulation_data=False, pool=None):
        """grid_search.py:173-262 with all pending combinations x runs of a task in one batched call."""
        assert tasks.keys() == data.keys(), 'Task mismatch!'
        assert pool is None, 'the batched back-end replaces the process pool'
        fit = {}
        if 'fit.pkl' in self.present_files:
            fit = pickle.load(open(self.file_path + 'fit.pkl', 'rb'))
        pending, loaded = [], {}
        for combo in self.parameter_combinations:
            if combo in fit and not overwrite:
                continue
            if self._file_name(combo) in self.present_files and not overwrite:
                loaded[combo] = pickle.load(open(self.file_path + self._file_name(combo), 'rb'))
            else:
                pending.append(combo)
        sim_data = {combo: {} for combo in pending}
        names = list(self.parameters.keys())
        per_call = len(pending) if not self.max_agents else max(1, self.max_agents // self.nb_runs)
        for c0 in range(0, len(pending), max(per_call, 1)):
            chunk = pending[c0:c0 + per_call]
            comb_idx = np.repeat(np.arange(len(chunk)), self.nb_runs)
            params = {n: np.array([self.parameter_combinations[chunk[c]][n] for c in comb_idx]) for n in names}
            params['_combination'] = comb_idx
            params['_run'] = np.tile(np.arange(self.nb_runs), len(chunk))
            for task in tasks:
                results = simulation(tasks[task], params)
                for c, combo in enumerate(chunk):
                    sim_data[combo][task] = [results[c * self.nb_runs + r] for r in range(self.nb_runs)]
        for combo in self.parameter_combinations:          # losses in combination order, like the reference
            if combo in sim_data or combo in loaded:
                sd = sim_data.get(combo, loaded.get(combo))
                if combo in sim_data and store_simulation_data:
                    pickle.dump(sd, open(self.file_path + self._file_name(combo), 'wb'))
                fit[combo] = loss(sd, data)
                pickle.dump(fit, open(self.file_path + 'fit.pkl', 'wb'))
        return fit

    def recompute_fit(self, data, loss, overwrite=False):
        """grid_search.py:264-312: recompute the losses from stored simulation data."""
        fit = {}
        if 'fit.pkl' in self.present_files and not overwrite:
            fit = pickle.load(open(self.file_path + 'fit.pkl', 'rb'))
        present = [f for f in listdir(self.file_path) if isfile(join(self.file_path, f))]
        for combo in self.parameter_combinations:
            if self._file_name(combo) in present:
                sd = pickle.load(open(self.file_path + self._file_name(combo), 'rb'))
                fit[combo] = loss(sd, data)
        pickle.dump(fit, open(self.file_path + 'fit.pkl', 'wb'))
        return fit
