"""NumPy / pure-Python restatement of CoBeL-RL's tabular closed loop (the oracle).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  One agent at a time, one
step at a time, in the reference's own operation order, so that under an
identical uniform stream it reproduces the reference bit-for-bit (integer
trajectories, replay indices, draw counts and fp64 tables).  Pinned against the
reference itself by tests/test_oracle_vs_reference.py and against the golden
vectors under tests/golden/.

Every function cites the reference file:line (relative to /root/reference/src/cobel)
it restates.  Conventions shared with the CUDA path:

* world tables: ``succ[S,A]`` int, ``reward[S]`` f64, ``terminal[S]`` 0/1,
  ``starts[K]`` int  (interface/gridworld.py:115-145 compiled to tables)
* the experience field called ``terminal`` by the reference holds
  ``1 - end_trial``; it is called ``nt`` ("non-terminal") here
* flat experience index: Dyna-Q ``i = s*A + a`` (C order, memory/dyna_q.py:137-142);
  PMA / SFMA ``i = a*S + s`` (F order, memory/pma.py:206, memory/sfma.py:206-215)
* ``logs['steps']`` is the index of the last step of a trial (agent/dyna_q.py:208)
"""
import numpy as np

# --------------------------------------------------------------------------- #
# random stream
# --------------------------------------------------------------------------- #


class Draws:
    """Program-order reader of one agent's uniform stream (Appendix A.2)."""

    def __init__(self, u, k=0):
        self.u, self.k = u, k

    def next(self):
        v = float(self.u[self.k])
        self.k += 1
        return v


def draw_integer(u, n):
    """``Generator.integers(n)`` / ``choice(a)`` from one uniform: min(floor(u*n), n-1)."""
    return min(int(np.floor(u * n)), n - 1)


def draw_categorical(p, u):
    """``Generator.choice(len(p), p=p)`` from one uniform.

    NumPy: ``cdf = cumsum(p); cdf /= cdf[-1]; searchsorted(cdf, u, side='right')``
    (sequential cumsum).  Used at policy/greedy.py:58, policy/softmax.py:58,
    memory/pma.py:252-254, memory/sfma.py:272,326.
    """
    cdf = np.cumsum(np.asarray(p, dtype=np.float64))
    cdf /= cdf[-1]
    return int(cdf.searchsorted(u, side='right'))


# --------------------------------------------------------------------------- #
# world compilers (host side, run once)
# --------------------------------------------------------------------------- #


def compile_gridworld(world):
    """Dense ``sas`` -> successor table; interface/gridworld.py:115-129.

    Deterministic worlds only: ``s' = argmax(sas[s, a, :])`` (first maximum).
    Reward and terminal flag are looked up at the arrival state (125-126).
    """
    sas = np.asarray(world['sas'])
    succ = np.argmax(sas, axis=2).astype(np.int32)
    return {
        'sas': None if world['deterministic'] else sas,
        'S': int(world['states']), 'A': 4, 'succ': succ,
        'reward': np.asarray(world['rewards'], dtype=np.float64).copy(),
        'terminal': np.asarray(world['terminals']).astype(np.uint8),
        'starts': np.asarray(world['starting_states']).astype(np.int32),
    }


def compile_topology(nodes, starting_nodes=None):
    """Node dictionary -> tables; interface/topology.py:86-112,149-172.

    Node ids are mapped to their insertion index.  ``starting_nodes=None`` means
    all non-terminal nodes (topology.py:100-107).  The action count is taken from
    the first node (all templates give every node the same number of neighbours;
    the reference takes it from the randomly chosen initial node, 110-112).
    """
    ids = list(nodes.keys())
    index = {n: i for i, n in enumerate(ids)}
    A = len(nodes[ids[0]]['neighbors'])
    succ = np.array([[index[m] for m in nodes[n]['neighbors']] for n in ids], dtype=np.int32)
    if starting_nodes is None:
        starting_nodes = [n for n in ids if not nodes[n]['terminal']]
    return {
        'S': len(ids), 'A': A, 'succ': succ,
        'reward': np.array([float(nodes[n]['reward']) for n in ids]),
        'terminal': np.array([1 if nodes[n]['terminal'] else 0 for n in ids], dtype=np.uint8),
        'starts': np.array([index[n] for n in starting_nodes], dtype=np.int32),
        'ids': ids,
        'pose': np.array([nodes[n]['pose'] for n in ids], dtype=np.float64),
    }


def valid_move_mask(succ):
    """Action mask used by the parity cases: an action is valid iff it changes the state."""
    succ = np.asarray(succ)
    return succ != np.arange(succ.shape[0]).reshape(-1, 1)


def env_reset(W, rng):
    """interface/gridworld.py:131-145, interface/topology.py:159-172: one draw."""
    return int(W['starts'][draw_integer(rng.next(), len(W['starts']))])


def env_step(W, s, a, rng=None):
    """interface/gridworld.py:115-129 / interface/topology.py:149-157.  Non-deterministic worlds
    (gridworld.py:118-123) draw the next state with one uniform: ``choice(arange(S), p=sas[s,a])``."""
    if W.get('sas') is not None:
        s2 = draw_categorical(W['sas'][s, a], rng.next())
    else:
        s2 = int(W['succ'][s, a])
    return s2, float(W['reward'][s2]), int(W['terminal'][s2])


# --------------------------------------------------------------------------- #
# policies
# --------------------------------------------------------------------------- #


def action_probs(policy, v, mask=None):
    """Action probabilities of a policy spec ``(kind, parameter)``.

    kind 'eps'     EpsilonGreedy.get_action_probs          policy/greedy.py:60-88
    kind 'xeps'    ExclusiveEpsilonGreedy.get_action_probs policy/greedy.py:117-147
    kind 'softmax' Softmax.get_action_probs                policy/softmax.py:60-88
    ``v`` is a 1-D float array, ``mask`` a bool array or None.
    """
    kind, par = policy
    v = np.asarray(v, dtype=np.float64)
    A = v.shape[0]
    valid = np.arange(A) if mask is None else np.arange(A)[np.asarray(mask, dtype=bool)]
    assert len(valid) > 0, 'The action mask masks all actions!'
    vals = v[valid]
    p = np.zeros(A)
    if kind == 'eps':
        ties = vals == np.amax(vals)
        p[valid] = par / len(valid)
        p[valid] += (1.0 - par) * ties / np.sum(ties)
    elif kind == 'xeps':
        ties = vals == np.amax(vals)
        p[valid] = (1.0 - par) * ties / np.sum(ties)
        p[valid] += par * (ties == False) / max(len(valid) - np.sum(ties), 1)  # noqa: E712
    elif kind == 'softmax':
        vals = vals - np.amax(vals)
        p[valid] = np.exp(vals * par)
        p[valid] /= np.sum(p[valid])
    else:
        raise ValueError(kind)
    return p


def select_action(policy, v, mask, rng):
    """Policy.select_action: probabilities then one categorical draw (greedy.py:40-58)."""
    return draw_categorical(action_probs(policy, v, mask), rng.next())


# --------------------------------------------------------------------------- #
# trajectory recorder
# --------------------------------------------------------------------------- #


class Record:
    """Collects what the CUDA path also emits: steps, per-trial stats, replay indices."""

    def __init__(self):
        self.s, self.a, self.s2, self.r = [], [], [], []
        self.trial_steps, self.trial_reward = [], []
        self.replay, self.replay_len = [], []

    def step(self, s, a, s2, r):
        self.s.append(s); self.a.append(a); self.s2.append(s2); self.r.append(r)

    def arrays(self):
        return {
            'states': np.array(self.s, dtype=np.int32), 'actions': np.array(self.a, dtype=np.int32),
            'next_states': np.array(self.s2, dtype=np.int32), 'rewards': np.array(self.r, dtype=np.float64),
            'trial_steps': np.array(self.trial_steps, dtype=np.int32),
            'trial_reward': np.array(self.trial_reward, dtype=np.float64),
            'replay': np.array(self.replay, dtype=np.int32),
            'replay_len': np.array(self.replay_len, dtype=np.int32),
        }


def _td_update(Q, s, a, r, s2, nt, lr, gamma, row_mask=None):
    """One-step TD update shared by DynaQ.update_q (agent/dyna_q.py:275-301),
    QAgent.update_q (agent/q.py:297-322) and SFMA.update_q (agent/sfma.py:423-458).

    ``td = r; td += gamma * nt * max(Q[s2]); td -= Q[s,a]; Q[s,a] += lr * td``
    evaluated left to right; ``row_mask`` restricts the max (SFMA only, 440-449).
    """
    row = Q[s2] if row_mask is None else Q[s2][row_mask]
    td = r
    td += gamma * nt * np.amax(row)
    td -= Q[s, a]
    Q[s, a] += lr * td
    return td


# --------------------------------------------------------------------------- #
# Dyna-Q  (agent/dyna_q.py:140-330, memory/dyna_q.py:62-157)
# --------------------------------------------------------------------------- #


def dynaq_init(S, A):
    """agent/dyna_q.py:128-138 and memory/dyna_q.py:73-75 (states init to self-loops)."""
    return {
        'Q': np.zeros((S, A)), 'Mr': np.zeros((S, A)),
        'Ms': np.tile(np.arange(S).reshape(S, 1), A).astype(np.int32),
        'Mt': np.zeros((S, A), dtype=np.int32),
        'action_mask': np.ones((S, A), dtype=bool),
    }


def dynaq_train(W, st, rng, trials, steps, batch, *, policy=('eps', 0.1), lr=0.99, gamma=0.99,
                mem_lr=0.9, mask_actions=False, no_replay=False, episodic_replay=False, rec=None):
    """DynaQ.train, agent/dyna_q.py:140-215 (loop order: select, env step, M.store,
    update_q, replay; the replay also runs after the terminal step)."""
    S, A = W['S'], W['A']
    Q, Mr, Ms, Mt = st['Q'], st['Mr'], st['Ms'], st['Mt']
    rec = rec if rec is not None else Record()

    def replay():
        # memory/dyna_q.py:137-157: B draws, C-order unravel, then sequential updates
        idx = [draw_integer(rng.next(), S * A) for _ in range(batch)]
        for i in idx:
            rs, ra = divmod(i, A)
            _td_update(Q, rs, ra, Mr[rs, ra], int(Ms[rs, ra]), int(Mt[rs, ra]), lr, gamma)
        rec.replay.extend(idx); rec.replay_len.append(len(idx))

    for _ in range(trials):
        s = env_reset(W, rng)
        treward, step = 0.0, 0
        for step in range(steps):
            a = select_action(policy, Q[s], st['action_mask'][s] if mask_actions else None, rng)
            s2, r, end = env_step(W, s, a, rng)
            nt = 1 - end
            # memory/dyna_q.py:92-96 (store before update)
            Mr[s, a] += mem_lr * (r - Mr[s, a]); Ms[s, a] = s2; Mt[s, a] = nt
            _td_update(Q, s, a, r, s2, nt, lr, gamma)
            rec.step(s, a, s2, r)
            s = s2
            if not no_replay and not episodic_replay:
                replay()
            treward += r
            if end:
                break
        rec.trial_steps.append(step); rec.trial_reward.append(treward)
        if not no_replay and episodic_replay:
            replay()
    return rec


def tabular_test(W, Q, rng, trials, steps, *, policy=('eps', 0.0), action_mask=None, rec=None):
    """DynaQ.test / PMA.test / SFMA.test / QAgent.test: act, never learn
    (agent/dyna_q.py:217-273, agent/pma.py:260-317, agent/sfma.py:330-390, agent/q.py:230-295)."""
    rec = rec if rec is not None else Record()
    for _ in range(trials):
        s = env_reset(W, rng)
        treward, step = 0.0, 0
        for step in range(steps):
            a = select_action(policy, Q[s], None if action_mask is None else action_mask[s], rng)
            s2, r, end = env_step(W, s, a, rng)
            rec.step(s, a, s2, r)
            s = s2
            treward += r
            if end:
                break
        rec.trial_steps.append(step); rec.trial_reward.append(treward)
    return rec


# --------------------------------------------------------------------------- #
# QAgent  (agent/q.py:160-354) on a table-compiled environment
# --------------------------------------------------------------------------- #


def q_init(S, A):
    """agent/q.py:142-143: Q rows are created lazily as zeros (equivalent to a zero
    table indexed by observation id); memory is an append-only experience log."""
    return {'Q': np.zeros((S, A)), 'log': []}


def q_train(W, st, rng, trials, steps, batch, *, policy=('eps', 0.1), lr=0.9, gamma=0.8, rec=None):
    """QAgent.train, agent/q.py:160-228: append experience, online update, then a
    replay of ``batch`` uniform draws over the whole log (344-354; the log already
    contains the step just taken).  ``batch == 0`` draws nothing (demo/topology/demo.py:76)."""
    Q, log = st['Q'], st['log']
    rec = rec if rec is not None else Record()
    for _ in range(trials):
        s = env_reset(W, rng)
        treward, step = 0.0, 0
        for step in range(steps):
            a = select_action(policy, Q[s], None, rng)
            s2, r, end = env_step(W, s, a, rng)
            nt = 1 - end
            log.append((s, a, r, s2, nt))
            _td_update(Q, s, a, r, s2, nt, lr, gamma)
            rec.step(s, a, s2, r)
            s = s2
            idx = [draw_integer(rng.next(), len(log)) for _ in range(batch)]
            for i in idx:
                es, ea, er, es2, ent = log[i]
                _td_update(Q, es, ea, er, es2, ent, lr, gamma)
            rec.replay.extend(idx); rec.replay_len.append(len(idx))
            treward += r
            if end:
                break
        rec.trial_steps.append(step); rec.trial_reward.append(treward)
    return rec


# --------------------------------------------------------------------------- #
# SR agent  (agent/sr.py:142-308)
# --------------------------------------------------------------------------- #


def sr_init(S, A):
    """agent/sr.py:130-136: SR = I, one-hot transition model = self-loops, rewards = 0.
    The dense (S,A,S) one-hot model is held as the successor index ``model[S,A]``."""
    return {'SR': np.eye(S), 'model': np.tile(np.arange(S).reshape(S, 1), A).astype(np.int32),
            'rew': np.zeros(S), 'action_mask': np.ones((S, A), dtype=bool)}


def sr_retrieve_q(st, s):
    """SR.retrieve_q, agent/sr.py:288-308.  The reference forms
    ``values = np.sum(SR * rewards, axis=1)`` (pairwise sum per row) and reads it at
    the modelled successor of each action (mean over exactly one element); only
    those A rows are evaluated here."""
    A = st['model'].shape[1]
    return np.array([np.sum(st['SR'][int(st['model'][s, a])] * st['rew']) for a in range(A)])


def sr_update(st, s, a, r, s2, nt, lr, gamma):
    """SR.update, agent/sr.py:255-286."""
    SRm, rew = st['SR'], st['rew']
    S = SRm.shape[0]
    td_reward = r - rew[s2]
    rew[s2] += td_reward * lr
    st['model'][s, a] = s2
    td = np.zeros(S); td[s] = 1.0
    if nt > 0:
        td += gamma * np.copy(SRm[s2])
    else:
        e = np.zeros(S); e[s2] = 1.0
        td += gamma * e
    td -= np.copy(SRm[s])
    SRm[s] += lr * td


def sr_train(W, st, rng, trials, steps, *, policy=('eps', 0.1), lr=0.1, gamma=0.99,
             mask_actions=False, learn=True, rec=None):
    """SR.train / SR.test, agent/sr.py:142-253."""
    rec = rec if rec is not None else Record()
    for _ in range(trials):
        s = env_reset(W, rng)
        treward, step = 0.0, 0
        for step in range(steps):
            mask = st['action_mask'][s] if mask_actions else None
            a = select_action(policy, sr_retrieve_q(st, s), mask, rng)
            s2, r, end = env_step(W, s, a, rng)
            if learn:
                sr_update(st, s, a, r, s2, 1 - end, lr, gamma)
            rec.step(s, a, s2, r)
            s = s2
            treward += r
            if end:
                break
        rec.trial_steps.append(step); rec.trial_reward.append(treward)
    return rec


# --------------------------------------------------------------------------- #
# SFMA  (agent/sfma.py:233-458, memory/sfma.py:195-373)
# --------------------------------------------------------------------------- #

SFMA_MODES = ('default', 'forward', 'reverse', 'blend_forward', 'blend_reverse', 'interpolate', 'sweeping')


def sfma_init(S, A):
    """agent/sfma.py:214-231, memory/sfma.py:162-172."""
    return {
        'Q': np.zeros((S, A)), 'Mr': np.zeros((S, A)),
        'Ms': np.tile(np.arange(S).reshape(S, 1), A).astype(np.int32),
        'Mt': np.zeros((S, A), dtype=np.int32),
        'C': np.zeros(S * A), 'T': np.zeros(S * A), 'I': np.zeros(S),
        'action_mask': np.ones((S, A), dtype=bool), 'td_acc': 0.0,
    }


def sfma_memory_replay(st, D, rng, length, current_state, *, mode='default', beta=20.0,
                       decay_inhibition=0.9, threshold=1e-6, recency=False, deterministic=False,
                       blend=0.1, interp=(0.5, 0.5), i_step=1.0, c_normalize=False, d_normalize=False,
                       r_normalize=True):
    """SFMAMemory.replay, memory/sfma.py:238-347.  Returns the flat indices
    ``a*S + s`` of the reactivated experiences."""
    S, A = st['Q'].shape
    C, I, Ms = st['C'], st['I'], st['Ms']
    action = draw_integer(rng.next(), A)                       # 264
    if current_state is None:                                   # 267-274
        Pc = np.clip(C, a_min=0, a_max=None)
        P = Pc / np.sum(Pc)
        e = draw_categorical(P, rng.next())
        current_state = e % S
        action = int(e / S)
    next_state = int(Ms[current_state, action])
    I *= 0                                                      # 277
    out = []
    statesF = Ms.flatten(order='F')
    for _ in range(length):
        Cc = np.copy(C)
        if c_normalize:                                         # 283-284
            Cc /= np.amax(Cc)
        Dv = np.tile(D[current_state], A)
        if d_normalize:                                         # 287-288
            Dv /= np.amax(Dv)
        if mode == 'forward':
            Dv = np.tile(D[next_state], A)
        elif mode == 'reverse':
            Dv = Dv[statesF]
        elif mode == 'blend_forward':
            Dv += blend * np.tile(D[next_state], A)
        elif mode == 'blend_reverse':
            Dv += blend * Dv[statesF]
        elif mode == 'interpolate':
            Dv = interp[0] * np.tile(D[next_state], A) + interp[1] * Dv[statesF]
        elif mode == 'sweeping':
            Dv = np.tile(D[next_state], A)[statesF]
        R = Cc * Dv * (1 - np.tile(I, A))                       # 311
        if recency:
            R *= st['T']
        R[R < threshold] = 0.0
        if np.sum(R) == 0.0:                                    # 316
            break
        if r_normalize:                                         # 319-320
            R /= np.amax(R)
        e = int(np.argmax(R))
        if not deterministic:
            ex = np.exp(R * beta) + (-1)                        # softmax(R, -1, beta), 349-373
            if np.sum(ex) == 0:
                ex.fill(1)
            else:
                ex /= np.sum(ex)
            probs = ex / np.sum(ex)
            e = draw_categorical(probs, rng.next())
        action = int(e / S)
        current_state = e - action * S
        next_state = int(Ms[current_state, action])
        I *= decay_inhibition
        I[current_state] = min(float(I[current_state] + i_step), 1.0)
        out.append(action * S + current_state)
    return out


def sfma_train(W, st, D, rng, trials, steps, batch, *, policy=('eps', 0.1), lr=0.99, gamma=0.99,
               mem_lr=0.9, mask_actions=False, mode='default', decay_strength=1.0,
               decay_recency=0.9, no_replay=False, nb_replays=1, start_replay=False,
               random_replay=False, dynamic=False, reward_mod_local=False, reward_mod=False, state_mod=False,
               reward_modulation=1.0, replay_kwargs=None, rec=None):
    """SFMA.train, agent/sfma.py:233-328 (store before update_q; replay at trial end
    from the terminal state, or from a strength-sampled experience when the trial
    timed out; ``M.T`` zeroed after every trial, 324).  ``random_replay`` restates ``agent.random``;
    ``dynamic`` restates the per-trial choice between 'reverse' and 'default' from the accumulated
    |TD error| (agent/sfma.py:311-318; chosen modes are returned in ``st['modes']``)."""
    S, A = W['S'], W['A']
    Q, Mr, Ms, Mt, C, T = st['Q'], st['Mr'], st['Ms'], st['Mt'], st['C'], st['T']
    rec = rec if rec is not None else Record()
    kw = dict(replay_kwargs or {})
    st.setdefault('modes', [])

    def replay(state, apply=True):
        # agent/sfma.py:392-421: the batch is sampled first, then applied in order.  The replay at
        # trial start calls the memory only (agent/sfma.py:272-275): a trace is generated and the
        # inhibition changes, but Q is not updated.
        if random_replay and apply:
            # agent/sfma.py:408-414 + SFMAMemory.retrieve_random_batch (memory/sfma.py:375-416): `batch`
            # draws in ONE Generator.choice call over the (masked) experiences, F-order unravel
            mask = st['action_mask'].flatten(order='F') if mask_actions else np.ones(S * A)
            probs = np.ones(S * A) * mask.astype(int)
            probs /= np.sum(probs)
            cdf = np.cumsum(probs)
            cdf /= cdf[-1]
            idx = [int(cdf.searchsorted(rng.next(), side='right')) for _ in range(batch)]
        else:
            idx = sfma_memory_replay(st, D, rng, batch, state, mode=mode, **kw)
        for i in (idx if apply else []):
            ea, es = divmod(i, S)
            es2 = int(Ms[es, ea])
            td = _td_update(Q, es, ea, Mr[es, ea], es2, int(Mt[es, ea]), lr, gamma,
                            st['action_mask'][es2] if mask_actions else None)
            st['td_acc'] += np.abs(td)
        rec.replay.extend(idx); rec.replay_len.append(len(idx))

    for _ in range(trials):
        last = None
        s = env_reset(W, rng)
        if start_replay:
            replay(s, apply=False)
        treward, step = 0.0, 0
        for step in range(steps):
            a = select_action(policy, Q[s], st['action_mask'][s] if mask_actions else None, rng)
            s2, r, end = env_step(W, s, a, rng)
            nt = 1 - end
            # SFMAMemory.store, memory/sfma.py:206-215
            Mr[s, a] += mem_lr * (r - Mr[s, a]); Ms[s, a] = s2; Mt[s, a] = nt
            C *= decay_strength; C[S * a + s] += 1.0
            T *= decay_recency; T[S * a + s] = 1.0
            if reward_mod_local:                                # memory/sfma.py:216-220
                C[S * a + s] += r * reward_modulation
            if reward_mod:                                      # 221-224
                C += r * np.tile(D[s], A) * reward_modulation
            if state_mod:                                       # 235-236
                C[[s + S * b for b in range(A)]] += 1.0
            td = _td_update(Q, s, a, r, s2, nt, lr, gamma,
                            st['action_mask'][s2] if mask_actions else None)
            st['td_acc'] += np.abs(td)
            rec.step(s, a, s2, r)
            s = s2
            treward += r
            if end:
                last = s2
                break
        rec.trial_steps.append(step); rec.trial_reward.append(treward)
        if not no_replay:
            if dynamic:
                # agent/sfma.py:311-318: p('reverse') is a logistic function of the TD error accumulated since
                # the last choice; one Generator.choice draw; the accumulator restarts
                p_mode = 1 / (1 + np.exp(-(st['td_acc'] * 5 - 2)))
                pick = draw_categorical(np.array([p_mode, 1 - p_mode]), rng.next())
                mode = ['reverse', 'default'][pick]
                st['modes'].append(pick)
                st['td_acc'] = 0.0
            for _r in range(nb_replays):
                replay(last)
            T.fill(0)
    return rec


# --------------------------------------------------------------------------- #
# PMA  (agent/pma.py:167-353, memory/pma.py:104-496)
# --------------------------------------------------------------------------- #


def pma_init(W_sas_T0, S, A, gamma_sr=0.9):
    """agent/pma.py:160-165, memory/pma.py:134-146.  ``W_sas_T0`` is
    ``np.sum(sas, axis=1) / A`` (139).  Note ``states`` initialises to 0 (136), and
    ``update_mask`` is computed once from that (144-146)."""
    T = np.array(W_sas_T0, dtype=np.float64)
    st = {
        'Q': np.zeros((S, A)), 'Mr': np.zeros((S, A)),
        'Ms': np.zeros((S, A), dtype=np.int32), 'Mt': np.zeros((S, A), dtype=np.int32),
        'T': T, 'SR': np.linalg.inv(np.eye(S) - gamma_sr * T),
        'action_mask': np.ones((S, A), dtype=bool),
    }
    st['update_mask'] = pma_compute_update_mask(st)
    return st


def pma_compute_update_mask(st):
    """PMAMemory.compute_update_mask, memory/pma.py:417-421."""
    S, A = st['Q'].shape
    return st['Ms'].flatten(order='F') != np.tile(np.arange(S), A)


def t0_from_succ(succ):
    """``np.sum(sas, axis=1) / A`` for a deterministic one-hot ``sas`` (memory/pma.py:139):
    the per-row sum over 4 one-hot vectors is exact in any order."""
    S, A = succ.shape
    T = np.zeros((S, S))
    for s in range(S):
        for a in range(A):
            T[s, succ[s, a]] += 1.0
    return T / A


def _probs_rows(policy, Qrows, mask_rows):
    """Row-wise ``get_action_probs`` (memory/pma.py:440-449).  The epsilon-greedy kinds are
    evaluated for all rows at once with the same element-wise IEEE operations as the per-row
    call (``eps/n + ((1-eps)*tie)/k``), which keeps the result bit-identical and the oracle usable
    as a CPU baseline; other kinds fall back to the per-row restatement."""
    kind, par = policy
    if kind not in ('eps', 'xeps'):
        return np.array([action_probs(policy, q, None if mask_rows is None else mask_rows[j])
                         for j, q in enumerate(Qrows)])
    Qrows = np.asarray(Qrows, dtype=np.float64)
    valid = np.ones(Qrows.shape, dtype=bool) if mask_rows is None else np.asarray(mask_rows, dtype=bool)
    vmax = np.max(np.where(valid, Qrows, -np.inf), axis=1, keepdims=True)
    ties = (Qrows == vmax) & valid
    nv = valid.sum(axis=1, keepdims=True)
    k = ties.sum(axis=1, keepdims=True)
    if kind == 'eps':
        p = par / nv + ((1.0 - par) * ties) / k
    else:
        p = ((1.0 - par) * ties) / k + (par * (~ties & valid)) / np.maximum(nv - k, 1)
    return np.where(valid, p, 0.0)


def pma_gain_batch(st, Qx, mask, policy, lr_q, gamma_q, min_gain):
    """PMAMemory.compute_gain_batch, memory/pma.py:333-386 (+ action_probs_batch 423-450)."""
    S, A = Qx.shape
    updates = np.tile(Qx, (A, 1))
    next_states = st['Ms'].flatten(order='F')
    targets = Qx[next_states]
    target_mask = np.zeros(updates.shape)
    for a in range(A):
        target_mask[S * a:S * (a + 1), a] = 1.0
    q_new = np.copy(updates)
    q_new += (lr_q * target_mask * (
        np.tile(st['Mr'], (A, 1))
        + gamma_q * np.amax(targets, axis=1).reshape(-1, 1)
        * st['Mt'].flatten(order='F').reshape(-1, 1)
        - q_new))
    m = np.tile(mask, (A, 1)) if mask is not None else None
    p_old = _probs_rows(policy, updates, m)
    p_old = p_old / np.sum(p_old, axis=1).reshape(-1, 1)
    p_new = _probs_rows(policy, q_new, m)
    p_new = p_new / np.sum(p_new, axis=1).reshape(-1, 1)
    gain = np.sum(p_new * q_new, axis=1) - np.sum(p_old * q_new, axis=1)
    return np.clip(gain, a_min=min_gain, a_max=None)


def pma_gain_sequence(st, Qx, mask, policy, seq, lr_q, gamma_sr, gamma_q, min_gain, original=True):
    """PMAMemory.compute_gain for one n-step candidate, memory/pma.py:269-331.
    Rewards are discounted with the *memory* gamma (310), the bootstrap with
    ``gamma_q`` (315); probabilities are not re-normalised here."""
    S, A = Qx.shape
    n = len(seq)
    last = seq[-1]
    ls, la = last % S, last // S
    fv = np.amax(Qx[int(st['Ms'][ls, la])]) * st['Mt'][ls, la]
    gain = 0.0
    for j, i in enumerate(seq):
        s, a = i % S, i // S
        m = mask[s] if mask is not None else None
        p_before = action_probs(policy, Qx[s], m)
        r = 0.0
        f = 0
        for f in range(n - j):
            k = seq[j + f]
            r += st['Mr'][k % S, k // S] * (gamma_sr ** f)
        q_target = np.copy(Qx[s])
        q_target[a] = r + fv * (gamma_q ** (f + 1))
        q_new = Qx[s] + lr_q * (q_target - Qx[s])
        p_after = action_probs(policy, q_new, m)
        step_gain = np.sum(q_new * p_after) - np.sum(q_new * p_before)
        if original:
            step_gain = max(step_gain, min_gain)
        gain += step_gain
    return max(gain, min_gain)


def pma_need(st, current_state):
    """PMAMemory.compute_need, memory/pma.py:388-411."""
    A = st['Q'].shape[1]
    if current_state is None:
        from scipy import linalg
        eig, vec = linalg.eig(st['T'], left=True, right=False)
        best = np.argmin(np.abs(eig - 1))
        return np.tile(np.abs(vec[:, best].T), A)
    return np.tile(st['SR'][current_state], A)


def pma_update_q_sequence(st, Qx, seq, lr_q, gamma_q):
    """PMAMemory.update_q, memory/pma.py:452-496 (n-step; aborts at an intermediate
    terminal / never-experienced transition)."""
    S, A = Qx.shape
    n = len(seq)
    last = seq[-1]
    ls, la = last % S, last // S
    fv = np.amax(Qx[int(st['Ms'][ls, la])]) * st['Mt'][ls, la]
    for j, i in enumerate(seq):
        s, a = i % S, i // S
        r = 0.0
        ok = 1
        f = 0
        for f in range(n - j):
            k = seq[j + f]
            if st['Mt'][k % S, k // S] == 0 and j != n - 1:
                ok = 0
                break
            r += st['Mr'][k % S, k // S] * (gamma_q ** f)
        if ok == 0:
            break
        td = r + fv * (gamma_q ** (f + 1))
        td -= Qx[s][a]
        Qx[s][a] += lr_q * td
    return Qx


def pma_memory_replay(st, Q, mask, length, current_state, rng, *, policy=('eps', 0.1), lr_q=0.9,
                      gamma_sr=0.9, gamma_q=0.9, min_gain=1e-6, original=True, allow_loops=False,
                      equal_need=False, equal_gain=False, ignore_barriers=True, cert=None):
    """PMAMemory.replay, memory/pma.py:168-267.  Returns (performed flat indices, new Q).
    ``cert`` (optional list) receives the relative gap between the two largest
    distinct utilities of every selection."""
    S, A = Q.shape
    Qx = np.copy(Q)
    performed = []
    last_seq = 0
    for it in range(length):
        ext, cand = -1, None
        if len(performed) > 0:
            lp = performed[-1]
            ext = int(st['Ms'][lp % S, lp // S])
            loop = any(ext == (p % S) for p in performed[last_seq:])
            if not loop or allow_loops:
                ea = select_action(policy, Qx[ext], mask[ext] if mask is not None else None, rng)
                ext += int(ea) * S
                cand = performed[last_seq:] + [ext]
        gain = pma_gain_batch(st, Qx, mask, policy, lr_q, gamma_q, min_gain)
        if ext != -1:
            seq = cand if cand is not None else [ext]
            gain[ext] = pma_gain_sequence(st, Qx, mask, policy, seq, lr_q, gamma_sr, gamma_q,
                                          min_gain, original)
        if equal_gain:
            gain.fill(1)
        need = pma_need(st, current_state)
        if equal_need:
            need.fill(1)
        utility = gain * need
        if ignore_barriers:
            utility *= st['update_mask']
        umaxv = np.amax(utility)
        ties = utility == umaxv
        if cert is not None:
            rest = utility[~ties]
            cert.append(float((umaxv - np.amax(rest)) / abs(umaxv)) if rest.size and umaxv != 0 else np.inf)
        umax = draw_categorical(ties / np.sum(ties), rng.next())
        chosen = cand if (cand is not None and umax == ext) else [umax]
        Qx = pma_update_q_sequence(st, Qx, chosen, lr_q, gamma_q)
        performed.append(umax)
        if ext != umax:
            last_seq = it
    return performed, Qx


def pma_store(st, s, a, r, s2, nt, mem_lr, lr_T=0.9):
    """PMAMemory.store, memory/pma.py:148-166."""
    S = st['T'].shape[0]
    st['Mr'][s][a] += mem_lr * (r - st['Mr'][s][a])
    st['Ms'][s][a] = s2
    st['Mt'][s][a] = nt
    st['T'][s] += lr_T * ((np.arange(S) == s2) - st['T'][s])


def pma_train(W, st, rng, trials, steps, batch, *, policy=('eps', 0.1), mem_policy=('eps', 0.1),
              lr=0.9, gamma=0.99, mem_lr=0.9, lr_q=0.9, gamma_sr=0.9, gamma_q=0.9,
              mask_actions=False, no_replay=False, replay_kwargs=None, rec=None, cert=None):
    """PMA.train, agent/pma.py:167-258: replay at trial start (need = SR[start]),
    online one-step update *then* store, ``update_sr`` and replay at trial end
    (need from the terminal state, or the stationary distribution on time-out)."""
    S, A = W['S'], W['A']
    rec = rec if rec is not None else Record()
    kw = dict(policy=mem_policy, lr_q=lr_q, gamma_sr=gamma_sr, gamma_q=gamma_q, cert=cert)
    kw.update(replay_kwargs or {})

    def replay(cur):
        perf, Qn = pma_memory_replay(st, st['Q'], st['action_mask'] if mask_actions else None,
                                     batch, cur, rng, **kw)
        st['Q'] = Qn
        rec.replay.extend(perf); rec.replay_len.append(len(perf))

    for _ in range(trials):
        last = None
        s = env_reset(W, rng)
        if not no_replay:
            replay(s)
        treward, step = 0, 0
        for step in range(steps):
            Q = st['Q']
            a = select_action(policy, Q[s], st['action_mask'][s] if mask_actions else None, rng)
            s2, r, end = env_step(W, s, a, rng)
            nt = 1 - end
            # PMA.update_q([experience]), agent/pma.py:319-353 with a one-element list
            fv = np.amax(Q[s2]) * nt
            rr = 0.0
            rr += r * (gamma ** 0)
            td = rr + fv * (gamma ** 1)
            td -= Q[s][a]
            Q[s][a] += lr * td
            pma_store(st, s, a, r, s2, nt, mem_lr)
            rec.step(s, a, s2, r)
            s = s2
            treward += r
            if end:
                last = s2
                break
        rec.trial_steps.append(step); rec.trial_reward.append(treward)
        if not no_replay:
            st['SR'] = np.linalg.inv(np.eye(S) - gamma_sr * st['T'])     # update_sr, 413-415
            replay(last)
    return rec
