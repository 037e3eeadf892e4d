// thread_agent.cuh -- building blocks of the ONE-THREAD-PER-AGENT kernels: the stand-alone methods of the class
// API (ops.cu: env.step, policy.select_action, M.store, agent.update_q, ... for callers that drive the loop
// themselves) and the online-step phase of SFMA (sfma.cu), where a step touches a handful of table entries and
// 32 agents per warp beat one warp per agent.  Everything is thread-local and follows the reference's operation
// order exactly (same helpers as the warp kernels: common.cuh).
#pragma once
#include "common.cuh"

// t[idx] by a select chain: a dynamically indexed register array would live in local memory
template <int A>
COBEL_DEV double pick_reg(const double (&t)[A], int idx) {
  double v = t[0];
#pragma unroll
  for (int a = 1; a < A; ++a) v = idx == a ? t[a] : v;
  return v;
}

// Policy.get_action_probs for one row: policy/greedy.py:60-88 (EpsilonGreedy), 117-147
// (ExclusiveEpsilonGreedy), policy/softmax.py:60-88 (Softmax).  mask = valid-action bits.
template <int A>
COBEL_DEV void action_probs_thread(const double (&v)[A], uint32_t mask, int kind, double par, double (&p)[A]) {
  double m = -__longlong_as_double(0x7FF0000000000000ll);
#pragma unroll
  for (int a = 0; a < A; ++a) m = (mask >> a & 1u) ? xmax(m, v[a]) : m;
  const int nv = __popc(mask);
  if (kind == COBEL_POLICY_SOFTMAX) {
    double sum = 0.0;
#pragma unroll
    for (int a = 0; a < A; ++a) {
      p[a] = 0.0;
      if (mask >> a & 1u) { p[a] = exp(xmul(xsub(v[a], m), par)); sum = xadd(sum, p[a]); }
    }
    if (A == 8 && nv == 8) sum = np_sum<A>(p);             // np.sum over exactly 8 values is a tree, not a loop
#pragma unroll
    for (int a = 0; a < A; ++a)
      if (mask >> a & 1u) p[a] = xdiv(p[a], sum);
    return;
  }
  uint32_t ties = 0;
#pragma unroll
  for (int a = 0; a < A; ++a) ties |= (v[a] == m ? 1u : 0u) << a;
  ties &= mask;
  const int k = __popc(ties);
  const double tie = xdiv(xsub(1.0, par), (double)k);
  double top, low;
  if (kind == COBEL_POLICY_EPS_GREEDY) {
    const double base = xdiv(par, (double)nv);
    top = xadd(base, tie); low = xadd(base, 0.0);
  } else {
    const int d = nv - k > 1 ? nv - k : 1;
    top = xadd(tie, 0.0); low = xadd(0.0, xdiv(par, (double)d));
  }
#pragma unroll
  for (int a = 0; a < A; ++a) p[a] = (mask >> a & 1u) ? ((ties >> a & 1u) ? top : low) : 0.0;
}

// Generator.choice(A, p=p) from one uniform: searchsorted(cumsum(p) / cumsum(p)[-1], u, 'right')  (policy/greedy.py:58)
template <int A>
COBEL_DEV int draw_categorical_thread(const double (&p)[A], double u) {
  double cdf[A];
  double c = p[0];
  cdf[0] = c;
#pragma unroll
  for (int a = 1; a < A; ++a) { c = xadd(c, p[a]); cdf[a] = c; }
  int n = 0;
#pragma unroll
  for (int a = 0; a < A - 1; ++a) n += (c != 1.0 ? xdiv(cdf[a], c) : cdf[a]) <= u ? 1 : 0;
  return n;
}

template <int A>
COBEL_DEV int select_action_thread(const double (&v)[A], uint32_t mask, int kind, double par, double u) {
  double p[A];
  action_probs_thread<A>(v, mask, kind, par, p);
  return draw_categorical_thread<A>(p, u);
}

// row of A doubles (A even and 16-byte aligned rows -> 128-bit loads)
template <int A>
COBEL_DEV void load_row_t(const double* r, double (&v)[A]) {
  if constexpr (A % 2 == 0) {
#pragma unroll
    for (int x = 0; x < A; x += 2) {
      const double2 t = *reinterpret_cast<const double2*>(r + x);
      v[x] = t.x; v[x + 1] = t.y;
    }
  } else {
#pragma unroll
    for (int x = 0; x < A; ++x) v[x] = r[x];
  }
}

// valid-action bits of state s from an action mask [S, A] (NULL = every action valid)
template <int A>
COBEL_DEV uint32_t mask_bits(const uint8_t* amask, int s) {
  uint32_t mb = (1u << A) - 1u;
  if (amask) {
    mb = 0;
#pragma unroll
    for (int a = 0; a < A; ++a) mb |= (amask[s * A + a] ? 1u : 0u) << a;
  }
  return mb;
}

// Next state of a non-deterministic gridworld: Generator.choice(arange(S), p=sas[s,a,:])
// (interface/gridworld.py:118-123) = first index whose sequential cumsum / total exceeds u.
COBEL_DEV int stochastic_successor_t(const CobelWorld& w, int sa, double u) {
  const int lo = __ldg(w.tp_off + sa), hi = __ldg(w.tp_off + sa + 1);
  double tot = 0.0;
  for (int j = lo; j < hi; ++j) tot = xadd(tot, __ldg(w.tp_prob + j));
  double c = 0.0;
  for (int j = lo; j < hi; ++j) {
    c = xadd(c, __ldg(w.tp_prob + j));
    if (xdiv(c, tot) > u) return __ldg(w.tp_next + j);
  }
  return __ldg(w.tp_next + hi - 1);
}

// One-step TD update shared by DynaQ / QAgent / SFMA.update_q (agent/dyna_q.py:275-301, agent/q.py:297-322,
// agent/sfma.py:423-458): td = r + (gamma * terminal) * max_{valid} Q[s'] - Q[s,a]; Q[s,a] += lr * td.
// `mb` = bits of the actions that enter the max (SFMA honours the action mask of s').  Returns td.
template <int A>
COBEL_DEV double td_update_thread(double* Q, int s, int a, double r, int s2, int nt, double lr, double gamma, uint32_t mb,
                                  bool apply = true) {
  double row[A];
  load_row_t<A>(Q + (size_t)s2 * A, row);
  double mx = -__longlong_as_double(0x7FF0000000000000ll);
#pragma unroll
  for (int x = 0; x < A; ++x) mx = (mb >> x & 1u) ? xmax(mx, row[x]) : mx;
  const double q = Q[(size_t)s * A + a];
  double td = xadd(r, xmul(nt ? gamma : 0.0, mx));
  td = xsub(td, q);
  if (apply) Q[(size_t)s * A + a] = xadd(q, xmul(lr, td));
  return td;
}
