// util.cu -- error plumbing, launch counter and the stream-contract probe.
#include "common.cuh"
#include <atomic>
#include <cstdarg>

namespace {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

__global__ void draw_uniforms_kernel(uint64_t seed, int64_t agent_id_base, int64_t n_agents,
                                     int64_t first, int64_t n_draws, double* out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_agents * n_draws) return;
  const int64_t a = i / n_draws, k = i - a * n_draws;
  CobelStream s{seed, agent_id_base, nullptr, nullptr, 0};
  Rng rng;
  rng.agent = (uint64_t)(s.agent_id_base + a);
  rng.key0 = (uint32_t)seed; rng.key1 = (uint32_t)(seed >> 32);
  rng.user = nullptr; rng.user_len = 0; rng.have = false; rng.k = 0;
  out[i] = rng.at((uint64_t)(first + k));
}
__global__ void stream_next_kernel(CobelStream s, int64_t n_agents, int64_t n_draws, double* out) {
  const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_agents) return;
  Rng rng; rng.init(s, a);
  for (int64_t k = 0; k < n_draws; ++k) out[a * n_draws + k] = rng.next();
  s.draw_count[a] = (int64_t)rng.k;
}
}  // namespace

void cobel_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}

void cobel_count_launch(int n) { g_launches += n; }

extern "C" int cobel_abi_version(void) { return COBEL_ABI_VERSION; }

extern "C" void cobel_last_error(char* buf, size_t len) {
  if (!buf || !len) return;
  strncpy(buf, g_err, len - 1);
  buf[len - 1] = 0;
}

extern "C" int64_t cobel_launch_count(void) { return g_launches.load(); }

extern "C" int cobel_draw_uniforms(uint64_t seed, int64_t agent_id_base, int64_t n_agents, int64_t first,
                                   int64_t n_draws, double* out, void* stream) {
  COBEL_REQUIRE(n_agents > 0 && n_draws > 0 && out, COBEL_EINVAL, "bad arguments to cobel_draw_uniforms");
  const int64_t total = n_agents * n_draws;
  draw_uniforms_kernel<<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      seed, agent_id_base, n_agents, first, n_draws, out);
  cobel_count_launch();
  COBEL_CUDA_OK(cudaGetLastError());
  return COBEL_OK;
}

extern "C" int cobel_stream_next(const CobelStream* s, int64_t n_agents, int64_t n_draws, double* out, void* stream) {
  COBEL_REQUIRE(s && s->draw_count && n_agents > 0 && n_draws > 0 && out, COBEL_EINVAL,
                "bad arguments to cobel_stream_next");
  stream_next_kernel<<<(unsigned)((n_agents + 127) / 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      *s, n_agents, n_draws, out);
  cobel_count_launch();
  COBEL_CUDA_OK(cudaGetLastError());
  return COBEL_OK;
}

extern "C" size_t cobel_sizeof(const char* name) {
  if (!name) return 0;
#define COBEL_SZ(T) if (!strcmp(name, #T)) return sizeof(T);
  COBEL_SZ(CobelWorld) COBEL_SZ(CobelStream) COBEL_SZ(CobelPolicy) COBEL_SZ(CobelTrace)
  COBEL_SZ(CobelDynaQParams) COBEL_SZ(CobelQParams) COBEL_SZ(CobelSRParams) COBEL_SZ(CobelSRCompactParams) COBEL_SZ(CobelSFMAParams) COBEL_SZ(CobelPMAParams) COBEL_SZ(CobelExperiences)
#undef COBEL_SZ
  return 0;
}
