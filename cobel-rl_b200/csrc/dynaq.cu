// dynaq.cu -- K1: the whole DynaQ.train()/test() loop of N independent agents in one launch.
//
// Reference: agent/dyna_q.py:140-330 (trial/step loop, update_q, replay),
// memory/dyna_q.py:62-157 (store, retrieve_batch), policy/greedy.py, policy/softmax.py,
// interface/gridworld.py:92-145.  Semantics: SURVEY.md Appendix A.4.
//
// Mapping: ONE THREAD PER AGENT.  The 33 TD updates of a step (1 online + 32
// replayed) form a strictly sequential fp64 dependency chain per agent, so the
// only parallelism is across agents; a thread per agent keeps 32 agents per warp
// busy instead of one.  Each warp stages the tables of its 32 agents in shared
// memory in a lane-interleaved layout (see SmemTables): whatever state each of
// the 32 lanes looks up, lane l always hits "its" banks, so the per-lane random
// row accesses of 32 different agents are conflict-free (a row of A=4 doubles is
// two LDS.128).  HBM is touched once per launch (stage in / stage out, coalesced).
#include "common.cuh"

namespace {

constexpr int kWarp = 32;

// ---- table accessors -------------------------------------------------------
// Shared-memory resident block of 32 agents.  Q rows are split into 16-byte
// vectors laid out [state][vector][lane] (A even) so that a warp-wide LDS.128
// of "my agent's row of state s_lane" is conflict-free whatever s_lane is; Mr is
// [state*A+action][lane] doubles, and the memory's (next state, non-terminal
// flag) pair is packed into one u16 per (s,a), [state*A+action][lane].
template <int A>
struct SmemTables {
  static constexpr int V = (A % 2 == 0) ? 2 : 1;   // doubles per vector
  static constexpr int NV = A / V;                 // vectors per row
  double* q;
  double* mr;
  uint16_t* mx;
  int lane;
  COBEL_DEV int qidx(int s, int a) const { return ((s * NV + a / V) * kWarp + lane) * V + a % V; }
  COBEL_DEV void load_qrow(int s, double (&v)[A]) const {
    if constexpr (V == 2) {
#pragma unroll
      for (int c = 0; c < NV; ++c) {
        const double2 t = *reinterpret_cast<const double2*>(q + ((s * NV + c) * kWarp + lane) * 2);
        v[2 * c] = t.x; v[2 * c + 1] = t.y;
      }
    } else {
#pragma unroll
      for (int a = 0; a < A; ++a) v[a] = q[(s * A + a) * kWarp + lane];
    }
  }
  COBEL_DEV double q_get(int s, int a) const { return q[qidx(s, a)]; }
  COBEL_DEV void q_set(int s, int a, double x) const { q[qidx(s, a)] = x; }
  COBEL_DEV double mr_get(int s, int a) const { return mr[(s * A + a) * kWarp + lane]; }
  COBEL_DEV void mr_set(int s, int a, double x) const { mr[(s * A + a) * kWarp + lane] = x; }
  COBEL_DEV void mem_get(int s, int a, int& s2, int& nt) const {
    const uint16_t v = mx[(s * A + a) * kWarp + lane];
    s2 = v & 0x7FFF; nt = v >> 15;
  }
  COBEL_DEV void mem_set(int s, int a, int s2, int nt) const {
    mx[(s * A + a) * kWarp + lane] = (uint16_t)(s2 | (nt << 15));
  }
  // index of element (s,a) of local agent `al` in the staged Q / flat layouts
  COBEL_DEV static int stage_q(int s, int a, int al) { return ((s * NV + a / V) * kWarp + al) * V + a % V; }
  COBEL_DEV static int stage_flat(int s, int a, int al) { return (s * A + a) * kWarp + al; }
};

// Tables left in global memory (state spaces too large for the staged layout).
template <int A>
struct GmemTables {
  double* q;      // this agent's [S,A]
  double* mr;
  int32_t* ms;
  int32_t* mt;
  COBEL_DEV void load_qrow(int s, double (&v)[A]) const {
#pragma unroll
    for (int a = 0; a < A; ++a) v[a] = q[s * A + a];
  }
  COBEL_DEV double q_get(int s, int a) const { return q[s * A + a]; }
  COBEL_DEV void q_set(int s, int a, double x) const { q[s * A + a] = x; }
  COBEL_DEV double mr_get(int s, int a) const { return mr[s * A + a]; }
  COBEL_DEV void mr_set(int s, int a, double x) const { mr[s * A + a] = x; }
  COBEL_DEV void mem_get(int s, int a, int& s2, int& nt) const { s2 = ms[s * A + a]; nt = mt[s * A + a]; }
  COBEL_DEV void mem_set(int s, int a, int s2, int nt) const { ms[s * A + a] = s2; mt[s * A + a] = nt; }
};

// One-step TD update, agent/dyna_q.py:275-301:
//   td = r; td += gamma * nt * max(Q[s2]); td -= Q[s,a]; Q[s,a] += lr * td
template <int A, class Tab>
COBEL_DEV void td_update(const Tab& t, int s, int a, double r, int s2, int nt, double lr, double gamma) {
  double row[A];
  t.load_qrow(s2, row);
  const double q = t.q_get(s, a);
  const double g = nt ? gamma : 0.0;                  // gamma * nt, nt in {0,1}
  double td = xadd(r, xmul(g, row_max<A>(row)));
  td = xsub(td, q);
  t.q_set(s, a, xadd(q, xmul(lr, td)));
}

struct WorldView {
  const int32_t* succ; const double* reward; const uint8_t* terminal; const int32_t* starts;
  int S, K;
};

template <int A, class Tab>
COBEL_DEV void run_agent(const CobelDynaQParams& p, const WorldView& w, const Tab& t, int64_t n) {
  Rng rng; rng.init(p.stream, n);
  const double lr = p.lr[n], gamma = p.gamma[n], mlr = p.mem_lr[n];
  const double par = p.policy.param[n];
  const int kind = p.policy.kind;
  const int S = w.S, B = p.batch;
  const uint8_t* amask = p.action_mask ? p.action_mask + n * p.mask_agent_stride : nullptr;
  const bool learn = p.learn != 0;
  const bool step_replay = learn && !p.no_replay && !p.episodic_replay && B > 0;
  const bool trial_replay = learn && !p.no_replay && p.episodic_replay && B > 0;
  int64_t nsteps = 0, nrep = 0, ncalls = 0;
  int flags = 0;
  const CobelTrace& tr = p.trace;

  auto replay = [&]() {
    // memory/dyna_q.py:137-157 then agent/dyna_q.py:329-330: B uniform draws over
    // S*A (C order), applied strictly in order.  Experiences are fetched 8 at a
    // time (the memory does not change during a replay), the Q chain is serial.
    for (int b0 = 0; b0 < B; b0 += 8) {
      int rs[8], ra[8], rs2[8], rnt[8];
      double rr[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (b0 + j < B) {
          const int i = draw_integer(rng.next(), S * A);
          rs[j] = i / A; ra[j] = i - rs[j] * A;
          rr[j] = t.mr_get(rs[j], ra[j]);
          t.mem_get(rs[j], ra[j], rs2[j], rnt[j]);
          if (tr.replay_idx) {
            if (nrep + j < tr.replay_cap) tr.replay_idx[n * tr.replay_cap + nrep + j] = i;
            else flags |= COBEL_FLAG_TRACE_OVERFLOW;
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (b0 + j < B) td_update<A>(t, rs[j], ra[j], rr[j], rs2[j], rnt[j], lr, gamma);
      nrep += (B - b0 < 8 ? B - b0 : 8);
    }
    if (tr.replay_len) {
      if (ncalls < tr.replay_calls_cap) tr.replay_len[n * tr.replay_calls_cap + ncalls] = B;
      else flags |= COBEL_FLAG_TRACE_OVERFLOW;
    }
    ++ncalls;
  };

  for (int trial = 0; trial < p.trials; ++trial) {
    int s = w.starts[draw_integer(rng.next(), w.K)];       // interface/gridworld.py:142
    double treward = 0.0;
    int step = 0;
    for (; step < p.steps; ++step) {
      double row[A];
      t.load_qrow(s, row);
      uint32_t mask = (1u << A) - 1u;
      if (amask) {
        mask = 0;
#pragma unroll
        for (int a = 0; a < A; ++a) mask |= (amask[s * A + a] ? 1u : 0u) << a;
      }
      const int a = select_action<A>(row, mask, kind, par, rng.next());
      const int s2 = w.succ[s * A + a];
      const double r = w.reward[s2];
      const int end = w.terminal[s2];
      const int nt = 1 - end;
      if (tr.step_sa) {
        if (nsteps < tr.step_cap) tr.step_sa[n * tr.step_cap + nsteps] = s * A + a;
        else flags |= COBEL_FLAG_TRACE_OVERFLOW;
      }
      ++nsteps;
      if (learn) {
        // memory/dyna_q.py:92-96 -- store first, then the online update
        const double m0 = t.mr_get(s, a);
        t.mr_set(s, a, xadd(m0, xmul(mlr, xsub(r, m0))));
        t.mem_set(s, a, s2, nt);
        td_update<A>(t, s, a, r, s2, nt, lr, gamma);
      }
      s = s2;
      if (step_replay) replay();
      treward = xadd(treward, r);
      if (end) break;
    }
    if (step == p.steps) step = p.steps - 1;               // logs['steps'] = last loop index
    tr.trial_steps[n * p.trials + trial] = step;
    tr.trial_reward[n * p.trials + trial] = treward;
    if (trial_replay) replay();
  }
  p.stream.draw_count[n] = (int64_t)rng.k;
  tr.n_steps[n] += nsteps;
  tr.n_replay[n] += nrep;
  if (tr.flags && flags) tr.flags[n] |= flags;
}

// ---- staged (shared-memory) kernel: one warp per CTA, 32 agents ------------
template <int A>
__global__ void __launch_bounds__(kWarp) dynaq_smem_kernel(const __grid_constant__ CobelDynaQParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int S = p.world.n_states, K = p.world.n_starts, SA = S * A;
  const int lane = threadIdx.x;
  const int64_t base = (int64_t)blockIdx.x * kWarp;
  const int64_t n = base + lane;
  const int nloc = (int)((p.n_agents - base) < kWarp ? (p.n_agents - base) : kWarp);

  double* q_s = reinterpret_cast<double*>(smem);
  double* mr_s = q_s + (size_t)SA * kWarp;
  double* rew_s = mr_s + (size_t)SA * kWarp;
  int32_t* succ_s = reinterpret_cast<int32_t*>(rew_s + S);
  int32_t* starts_s = succ_s + SA;
  uint16_t* mx_s = reinterpret_cast<uint16_t*>(starts_s + K);
  uint8_t* term_s = reinterpret_cast<uint8_t*>(mx_s + (size_t)SA * kWarp);

  // stage in: coalesced reads of [agent][s][a], scattered into [s][lane][a]
  for (int e = lane; e < SA * kWarp; e += kWarp) {
    const int al = e / SA, idx = e - al * SA;
    const int s = idx / A, a = idx - s * A;
    const int dq = SmemTables<A>::stage_q(s, a, al), df = SmemTables<A>::stage_flat(s, a, al);
    if (al < nloc) {
      const size_t g = (size_t)(base + al) * SA + idx;
      q_s[dq] = p.Q[g];
      mr_s[df] = p.Mr[g];
      mx_s[df] = (uint16_t)(p.Ms[g] | ((p.Mt[g] ? 1 : 0) << 15));
    } else {
      q_s[dq] = 0.0; mr_s[df] = 0.0; mx_s[df] = 0;
    }
  }
  for (int e = lane; e < SA; e += kWarp) succ_s[e] = p.world.succ[e];
  for (int e = lane; e < S; e += kWarp) { rew_s[e] = p.world.reward[e]; term_s[e] = p.world.terminal[e]; }
  for (int e = lane; e < K; e += kWarp) starts_s[e] = p.world.starts[e];
  __syncwarp();

  if (n < p.n_agents) {
    SmemTables<A> t{q_s, mr_s, mx_s, lane};
    WorldView w{succ_s, rew_s, term_s, starts_s, S, K};
    run_agent<A>(p, w, t, n);
  }
  __syncwarp();

  if (p.learn) {
    for (int e = lane; e < SA * kWarp; e += kWarp) {
      const int al = e / SA, idx = e - al * SA;
      if (al >= nloc) break;
      const int s = idx / A, a = idx - s * A;
      const int dq = SmemTables<A>::stage_q(s, a, al), df = SmemTables<A>::stage_flat(s, a, al);
      const size_t g = (size_t)(base + al) * SA + idx;
      p.Q[g] = q_s[dq];
      p.Mr[g] = mr_s[df];
      p.Ms[g] = mx_s[df] & 0x7FFF;
      p.Mt[g] = mx_s[df] >> 15;
    }
  }
}

// ---- fallback: tables stay in global memory (S*A too large to stage) --------
template <int A>
__global__ void __launch_bounds__(128) dynaq_gmem_kernel(const __grid_constant__ CobelDynaQParams p) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= p.n_agents) return;
  const size_t SA = (size_t)p.world.n_states * A;
  GmemTables<A> t{p.Q + n * SA, p.Mr + n * SA, p.Ms + n * SA, p.Mt + n * SA};
  WorldView w{p.world.succ, p.world.reward, p.world.terminal, p.world.starts, p.world.n_states, p.world.n_starts};
  run_agent<A>(p, w, t, n);
}

size_t smem_bytes(int S, int A, int K) {
  const size_t SA = (size_t)S * A;
  size_t b = 2 * SA * kWarp * sizeof(double);      // Q, Mr
  b += (size_t)S * sizeof(double);                 // reward
  b += SA * sizeof(int32_t) + (size_t)K * sizeof(int32_t);
  b += SA * kWarp * sizeof(uint16_t);              // packed memory
  b += (size_t)S;                                  // terminal
  return (b + 15) & ~(size_t)15;
}

template <int A>
int launch(const CobelDynaQParams& p, cudaStream_t st) {
  const int S = p.world.n_states, K = p.world.n_starts;
  const size_t sm = smem_bytes(S, A, K);
  if (sm <= 227 * 1024 && S <= 0x7FFF) {
    COBEL_CUDA_OK(cudaFuncSetAttribute(dynaq_smem_kernel<A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    const unsigned grid = (unsigned)((p.n_agents + kWarp - 1) / kWarp);
    dynaq_smem_kernel<A><<<grid, kWarp, sm, st>>>(p);
  } else {
    const unsigned grid = (unsigned)((p.n_agents + 127) / 128);
    dynaq_gmem_kernel<A><<<grid, 128, 0, st>>>(p);
  }
  cobel_count_launch();
  COBEL_CUDA_OK(cudaGetLastError());
  return COBEL_OK;
}

}  // namespace

int cobel_validate_common(int64_t n_agents, const CobelWorld& w, const CobelStream& s, const CobelPolicy& pol,
                          const CobelTrace& tr, int trials, int steps) {
  COBEL_REQUIRE(n_agents > 0, COBEL_EINVAL, "n_agents must be positive");
  COBEL_REQUIRE(w.n_states > 0 && w.n_actions > 0 && w.n_starts > 0, COBEL_EINVAL, "empty world");
  COBEL_REQUIRE(w.succ && w.reward && w.terminal && w.starts, COBEL_EINVAL, "world tables missing");
  COBEL_REQUIRE(s.draw_count, COBEL_EINVAL, "stream.draw_count missing");
  COBEL_REQUIRE(pol.param && pol.kind >= 0 && pol.kind <= 2, COBEL_EINVAL, "bad policy");
  COBEL_REQUIRE(trials >= 0 && steps > 0, COBEL_EINVAL, "trials must be >= 0 and steps > 0");
  COBEL_REQUIRE(tr.trial_steps && tr.trial_reward && tr.n_steps && tr.n_replay, COBEL_EINVAL,
                "trace.trial_steps/trial_reward/n_steps/n_replay are mandatory");
  return COBEL_OK;
}

extern "C" int cobel_dynaq_run(const CobelDynaQParams* pp, void* stream) {
  COBEL_REQUIRE(pp != nullptr, COBEL_EINVAL, "null params");
  const CobelDynaQParams& p = *pp;
  int rc = cobel_validate_common(p.n_agents, p.world, p.stream, p.policy, p.trace, p.trials, p.steps);
  if (rc) return rc;
  COBEL_REQUIRE(p.Q && p.Mr && p.Ms && p.Mt && p.lr && p.gamma && p.mem_lr, COBEL_EINVAL, "agent tables missing");
  COBEL_REQUIRE(p.batch >= 0, COBEL_EINVAL, "batch must be >= 0");
  if (p.trials == 0) return COBEL_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (p.world.n_actions) {
    case 2: return launch<2>(p, st);
    case 3: return launch<3>(p, st);
    case 4: return launch<4>(p, st);
    case 6: return launch<6>(p, st);
    case 8: return launch<8>(p, st);
    default:
      cobel_set_error("unsupported number of actions %d (built for 2,3,4,6,8)", p.world.n_actions);
      return COBEL_EUNSUPPORTED;
  }
}
