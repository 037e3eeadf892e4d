"""Evolutionary-algorithm optimizer with a batched back-end (reference: optimizer/evolution.py:17-196).

The reference's ``EAOptimizer`` evaluates every individual of a generation by ``population_size`` calls of
``simulation(task, individual)`` per task, optionally through a process pool (evolution.py:151-175), keeps the
best individual and mutates it into the next generation (evolution.py:176-197).  Here a generation is ONE batched
simulation per task: all individuals x ``population_size`` repetitions are mapped onto the agent axis,

    simulation(task, params) -> results

with ``params[name]`` an array holding one value per agent (``params['_individual']`` / ``params['_run']`` give the
individual and repetition of each agent) and ``results`` indexable along that axis -- the convention of
``GridSearchOptimizer`` in this package.  Constructor arguments, the ``parameters`` schema (``param_type``,
``init_range``, ``param_range``, ``mutator = (callable, kwargs)``), the draw order from ``rng``, the ``run_<r>.pkl``
resume files (``file_path + 'run_%d.pkl'``, a list of ``(best individual, fitness)`` per generation) and the
returned ``Fit`` follow the reference.

Two facts about the reference's implementation (both pinned by tests/test_optimizer_monitor_cpu.py against the
unmodified class):

* its constructor cannot run: the default ``mutator`` is written as a set literal containing a dict
  (evolution.py:73-75), which raises ``TypeError: unhashable type`` while the argument is evaluated, whether or not
  the caller supplied a mutator.  Here the default is the tuple the code that consumes it expects
  (evolution.py:183-186): ``(rng.normal, {'scale': 0.1})``, i.e. ``value' = rng.normal(value, 0.1)``.
* in ``fit`` the statement that records an individual's loss sits outside the loop over the individuals
  (evolution.py:175) and the selection / mutation block outside the loop over the generations (evolution.py:176-197):
  the first population is evaluated ``generations`` times, the one recorded loss is the LAST individual's,
  ``argmin`` is 0, so a run returns its first random individual with the last one's loss.
  ``bookkeeping='reference'`` reproduces exactly that (same files, same return value as the reference's ``fit`` method run on an object whose
  constructor was bypassed); the default ``bookkeeping='intended'`` records one loss per individual and selects the
  arg-min.  Likewise the reference looks for a resume file under its full path in a list of bare file names
  (evolution.py:135), so it never resumes; both modes here resume from ``run_<r>.pkl`` unless ``overwrite``.
"""
import copy
import pickle
from os import listdir
from os.path import isfile, join

import numpy as np


class EAOptimizer:
    def __init__(self, file_path, parameters, nb_runs=1, population_size=1, rng=None, bookkeeping='intended'):
        assert bookkeeping in ('intended', 'reference')
        self.rng = np.random.default_rng() if rng is None else rng
        self.parameters = copy.deepcopy(parameters)
        for name in self.parameters:                       # evolution.py:68-75
            self.parameters[name].setdefault('init_range', {'low': -1, 'high': 1})
            self.parameters[name].setdefault('param_range', {'a_min': None, 'a_max': None})
            self.parameters[name].setdefault('mutator', (self.rng.normal, {'scale': 0.1}))
        self.file_path = file_path
        self.nb_runs = nb_runs
        self.population_size = population_size
        self.bookkeeping = bookkeeping
        self.present_files = [f for f in listdir(self.file_path) if isfile(join(self.file_path, f))]

    def _evaluate(self, simulation, tasks, population):
        """simulation_data[i][task] = list of the ``population_size`` results of individual i: one batched call
        per task for the whole generation."""
        names = list(self.parameters.keys())
        reps = self.population_size
        ind = np.repeat(np.arange(len(population)), reps)
        params = {n: np.array([population[i][n] for i in ind]) for n in names}
        params['_individual'] = ind
        params['_run'] = np.tile(np.arange(reps), len(population))
        data = [dict() for _ in population]
        for t, task in tasks.items():
            results = simulation(task, params)
            for i in range(len(population)):
                data[i][t] = [results[i * reps + r] for r in range(reps)]
        return data

    def _select_and_mutate(self, population, fit, history, file_name, individuals):
        """evolution.py:176-197: record the best individual, then the next generation = it and individuals - 1
        mutations of it, clipped to ``param_range``."""
        best = population[int(np.argmin(fit))]
        history.append((best, np.amin(fit)))
        pickle.dump(history, open(file_name, 'wb'))
        population = [best]
        for _ in range(individuals - 1):
            population.append({p: self.parameters[p]['mutator'][0](best[p], **self.parameters[p]['mutator'][1])
                               for p in self.parameters})
            for p in self.parameters:
                population[-1][p] = np.clip(self.parameters[p]['param_type'](population[-1][p]),
                                            **self.parameters[p]['param_range'])
        return population

    def fit(self, simulation, tasks, data, loss, overwrite=False, generations=100, individuals=10, pool=None):
        """evolution.py:88-207 with every generation as one batched simulation per task."""
        assert tasks.keys() == data.keys(), 'Task mismatch!'
        assert pool is None, 'the batched back-end replaces the process pool'
        best_fits = []
        for run in range(self.nb_runs):
            file_name = self.file_path + 'run_%d.pkl' % run
            if 'run_%d.pkl' % run in self.present_files and not overwrite:
                best_fits.append(pickle.load(open(file_name, 'rb')))
                continue
            best_fits.append([])
            # first generation: one uniform draw per individual and parameter, in that order (evolution.py:140-148)
            population = [{p: self.parameters[p]['param_type'](self.rng.uniform(**self.parameters[p]['init_range']))
                           for p in self.parameters} for _ in range(individuals)]
            if self.bookkeeping == 'reference':
                # evolution.py:149-197 as indented there: the generation loop only re-evaluates the first population,
                # selection and mutation follow it once, and the one recorded loss is the last individual's
                for _ in range(generations):
                    sim = self._evaluate(simulation, tasks, population)
                fit = [loss(sim[-1], data)]
                self._select_and_mutate(population, fit, best_fits[-1], file_name, individuals)
                continue
            for _ in range(generations):
                sim = self._evaluate(simulation, tasks, population)
                fit = [loss(sd, data) for sd in sim]
                population = self._select_and_mutate(population, fit, best_fits[-1], file_name, individuals)
        final_fit = {}
        for f in best_fits:                                 # evolution.py:199-201
            final_fit[tuple(list(f[-1][0].values()))] = f[-1][1]
        return final_fit
