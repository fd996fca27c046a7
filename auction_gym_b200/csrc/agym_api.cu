// C ABI of libagym (include/agym.h): handle lifetime, configuration, argument checking, dispatch.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "agym_common.cuh"

namespace agym {

static std::string g_create_error;

int set_error(agym_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg; else g_create_error = msg;
  return code;
}

int check_cuda(agym_handle* h, cudaError_t e, const char* what) {
  if (e == cudaSuccess) return AGYM_OK;
  return set_error(h, AGYM_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

SimParams make_params(const agym_handle* h) {
  SimParams p{};
  const agym_shape& s = h->shape;
  p.R = s.R; p.A = s.A; p.I = s.I; p.D = s.D; p.Do = s.Do; p.K = h->K; p.P = s.P;
  p.mechanism = s.mechanism;
  p.max_slots = s.max_slots > 1 ? s.max_slots : 1;
  p.run_offset = s.run_offset;
  p.embedding_var = s.embedding_var;
  p.n_items = h->d_n_items; p.alloc_kind = h->d_alloc_kind; p.bidder_kind = h->d_bidder_kind;
  p.E64 = h->d_E64; p.V64 = h->d_V64; p.E32 = h->d_E32; p.V32 = h->d_V32;
  p.m = h->m; p.sigma = h->sigma;
  p.pk = h->pk_valid ? h->d_pk : nullptr; p.cat8 = h->d_cat8;
  p.bidder_d = h->bidder_d; p.bidder_w = h->bidder_w;
  p.acc = h->acc; p.revenue = h->revenue;
  p.fit_ctx = h->fit_ctx; p.fit_meta = h->fit_meta; p.Tcap = h->Tcap;
  p.bid_rows = h->bid_rows; p.bid_meta = h->bid_meta; p.bid_Tcap = h->bid_Tcap;
  p.terms = h->terms; p.log_base = h->log_base;
  p.round0 = h->rounds_in_iter;
  p.run0 = 0; p.n_runs = s.R;
  return p;
}

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

static int ready_for_rounds(agym_handle* h, const char* who) {
  if (!h->agents_set || !h->catalog_set) return set_error(h, AGYM_ERR_STATE, std::string(who) + ": agym_set_agents / agym_set_catalog not called");
  if (!h->acc || !h->revenue) return set_error(h, AGYM_ERR_STATE, std::string(who) + ": metrics not bound (agym_bind_metrics)");
  if (h->any_learnt && (!h->m || !h->sigma)) return set_error(h, AGYM_ERR_STATE, std::string(who) + ": learnt allocators need agym_bind_allocator_state");
  if (h->any_shaded && !h->bidder_d) return set_error(h, AGYM_ERR_STATE, std::string(who) + ": shaded bidders need agym_bind_bidder_state");
  return AGYM_OK;
}

}  // namespace agym

using namespace agym;

extern "C" {

int agym_abi_version(void) { return AGYM_ABI_VERSION; }

const char* agym_last_error(const agym_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int agym_create(const agym_shape* shape, int device, agym_handle** out) {
  if (!shape || !out) return set_error(nullptr, AGYM_ERR_INVALID, "agym_create: null argument");
  const agym_shape& s = *shape;
  if (s.R < 1 || s.A < 1 || s.I < 1 || s.D < 1 || s.Do < 0 || s.Do > s.D || s.P < 1 || s.P > s.A)
    return set_error(nullptr, AGYM_ERR_INVALID, "agym_create: need R,A,I,D >= 1, 0 <= Do <= D, 1 <= P <= A");
  if (s.A > 4096 || s.I > 4096) return set_error(nullptr, AGYM_ERR_INVALID, "agym_create: A and I are limited to 4096 (fit_meta packing)");
  if (s.mechanism != AGYM_SECOND_PRICE && s.mechanism != AGYM_FIRST_PRICE) return set_error(nullptr, AGYM_ERR_INVALID, "agym_create: unknown mechanism");
  if (s.precision != AGYM_FP32 && s.precision != AGYM_FP64) return set_error(nullptr, AGYM_ERR_INVALID, "agym_create: unknown precision");
  if (s.max_slots < 0 || s.max_slots > kMaxP) return set_error(nullptr, AGYM_ERR_INVALID, "agym_create: max_slots must be in [0, 32]");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return set_error(nullptr, AGYM_ERR_CUDA, std::string("agym_create: no CUDA device (") + cudaGetErrorString(e) + "); this engine has no CPU fallback");
  if (device < 0 || device >= ndev) return set_error(nullptr, AGYM_ERR_INVALID, "agym_create: bad device index");
  DeviceGuard g(device);
  agym_handle* h = new agym_handle();
  h->shape = s;
  h->device = device;
  h->K = s.Do + 1;
  cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device);
  const size_t nE = (size_t)s.A * s.I * (s.D + 1), nV = (size_t)s.A * s.I;
  bool ok = cudaMalloc(&h->d_n_items, s.A * sizeof(int)) == cudaSuccess && cudaMalloc(&h->d_alloc_kind, s.A * sizeof(int)) == cudaSuccess &&
            cudaMalloc(&h->d_bidder_kind, s.A * sizeof(int)) == cudaSuccess && cudaMalloc(&h->d_bidder_fit, s.A * sizeof(int)) == cudaSuccess && cudaMalloc(&h->d_E64, nE * sizeof(double)) == cudaSuccess &&
            cudaMalloc(&h->d_V64, nV * sizeof(double)) == cudaSuccess && cudaMalloc(&h->d_E32, nE * sizeof(float)) == cudaSuccess &&
            cudaMalloc(&h->d_V32, nV * sizeof(float)) == cudaSuccess &&
            (s.D != 5 || cudaMalloc(&h->d_cat8, (size_t)s.A * tiles_of(s.I) * kCatTile) == cudaSuccess) &&
            cudaMalloc(&h->d_adam_sz0, kAdamTable * sizeof(double)) == cudaSuccess &&
            cudaMalloc(&h->d_adam_bc2s, kAdamTable * sizeof(float)) == cudaSuccess &&
            cudaMalloc(&h->d_adam_ep, kAdamTable * sizeof(float2)) == cudaSuccess &&
            cudaMalloc(&h->d_adam_bc1, kAdamTable2 * sizeof(double)) == cudaSuccess &&
            cudaMalloc(&h->d_adam_bc2s2, kAdamTable2 * sizeof(float)) == cudaSuccess;
  if (ok) {
    // torch.optim.Adam's per-step scalars, computed the way torch does (Python floats: beta ** step)
    std::vector<double> sz0(kAdamTable);
    std::vector<float> bc2s(kAdamTable);
    for (int e = 0; e < kAdamTable; ++e) {
      const double t = double(e + 1);
      sz0[e] = 2e-3 / (1.0 - std::pow(0.9, t));
      bc2s[e] = float(std::sqrt(1.0 - std::pow(0.999, t)));
    }
    std::vector<float2> ep(kAdamTable);
    for (int e = 0; e < kAdamTable; ++e) ep[e] = make_float2(float(sz0[e]), bc2s[e]);
    std::vector<double> bc1(kAdamTable2);
    std::vector<float> bc2s2(kAdamTable2);
    for (int e = 0; e < kAdamTable2; ++e) {
      const double t = double(e + 1);
      bc1[e] = 1.0 - std::pow(0.9, t);
      bc2s2[e] = float(std::sqrt(1.0 - std::pow(0.999, t)));
    }
    ok = cudaMemcpy(h->d_adam_bc1, bc1.data(), kAdamTable2 * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(h->d_adam_bc2s2, bc2s2.data(), kAdamTable2 * sizeof(float), cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(h->d_adam_sz0, sz0.data(), kAdamTable * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(h->d_adam_bc2s, bc2s.data(), kAdamTable * sizeof(float), cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(h->d_adam_ep, ep.data(), kAdamTable * sizeof(float2), cudaMemcpyHostToDevice) == cudaSuccess;
  }
  if (!ok) {
    set_error(nullptr, AGYM_ERR_CUDA, std::string("agym_create: cudaMalloc failed: ") + cudaGetErrorString(cudaGetLastError()));
    agym_destroy(h);
    return AGYM_ERR_CUDA;
  }
  *out = h;
  return AGYM_OK;
}

int agym_destroy(agym_handle* h) {
  if (!h) return AGYM_OK;
  DeviceGuard g(h->device);
  cudaFree(h->d_n_items); cudaFree(h->d_alloc_kind); cudaFree(h->d_bidder_kind); cudaFree(h->d_bidder_fit);
  cudaFree(h->d_E64); cudaFree(h->d_V64); cudaFree(h->d_E32); cudaFree(h->d_V32); cudaFree(h->d_cat8); cudaFree(h->d_pk);
  cudaFree(h->d_memory); cudaFree(h->d_mem_off);
  destroy_comm(h);
  cudaFree(h->d_fit_epochs); cudaFree(h->est_scratch);
  if (h->aux_stream) cudaStreamDestroy(h->aux_stream);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  cudaFree(h->d_adam_sz0); cudaFree(h->d_adam_bc2s); cudaFree(h->d_adam_ep); cudaFree(h->k4_scratch); cudaFree(h->d_adam_bc1); cudaFree(h->d_adam_bc2s2);
  delete h;
  return AGYM_OK;
}

int agym_set_agents(agym_handle* h, const int32_t* n_items, const int32_t* alloc_kind, const int32_t* bidder_kind) {
  if (!h || !n_items || !alloc_kind || !bidder_kind) return set_error(h, AGYM_ERR_INVALID, "agym_set_agents: null argument");
  DeviceGuard g(h->device);
  const int A = h->shape.A;
  h->any_learnt = h->any_shaded = h->any_search = h->any_unbuilt_fit = false;
  h->max_items = 0;
  for (int a = 0; a < A; ++a) {
    if (n_items[a] < 1 || n_items[a] > h->shape.I) return set_error(h, AGYM_ERR_INVALID, "agym_set_agents: n_items out of [1, I]");
    if (alloc_kind[a] < AGYM_ALLOC_ORACLE || alloc_kind[a] > AGYM_ALLOC_MAP) return set_error(h, AGYM_ERR_INVALID, "agym_set_agents: unknown allocator kind");
    if (bidder_kind[a] < AGYM_BID_TRUTHFUL || bidder_kind[a] > AGYM_BID_POLICY) return set_error(h, AGYM_ERR_INVALID, "agym_set_agents: unknown bidder kind");
    h->any_learnt |= alloc_kind[a] != AGYM_ALLOC_ORACLE;
    h->any_shaded |= bidder_kind[a] != AGYM_BID_TRUTHFUL;
    h->any_search |= bidder_kind[a] == AGYM_BID_SEARCH;
    h->any_unbuilt_fit |= bidder_kind[a] == AGYM_BID_BANDIT || bidder_kind[a] == AGYM_BID_POLICY;
    if (n_items[a] > h->max_items) h->max_items = n_items[a];
  }
  h->alloc_kind_host.assign(alloc_kind, alloc_kind + A);
  cudaError_t e = cudaMemcpy(h->d_n_items, n_items, A * sizeof(int), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(h->d_alloc_kind, alloc_kind, A * sizeof(int), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(h->d_bidder_kind, bidder_kind, A * sizeof(int), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return check_cuda(h, e, "agym_set_agents");
  h->agents_set = true;
  std::vector<int32_t> fits(A);
  for (int a = 0; a < A; ++a)
    fits[a] = bidder_kind[a] == AGYM_BID_SEARCH ? AGYM_BFIT_VL_SEARCH : bidder_kind[a] == AGYM_BID_POLICY ? AGYM_BFIT_VL_POLICY : AGYM_BFIT_NONE;
  return agym_set_bidder_fits(h, fits.data());
}

int agym_set_bidder_fits(agym_handle* h, const int32_t* fit_kind) {
  if (!h || !fit_kind) return set_error(h, AGYM_ERR_INVALID, "agym_set_bidder_fits: null argument");
  if (!h->agents_set) return set_error(h, AGYM_ERR_STATE, "agym_set_bidder_fits: call agym_set_agents first");
  DeviceGuard g(h->device);
  const int A = h->shape.A;
  std::vector<int32_t> kinds(A);
  cudaError_t e = cudaMemcpy(kinds.data(), h->d_bidder_kind, A * sizeof(int), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return check_cuda(h, e, "agym_set_bidder_fits");
  h->any_winrate_fit = h->any_policy_fit = h->any_unassigned_bandit = h->any_empirical_fit = false;
  for (int a = 0; a < A; ++a) {
    const int f = fit_kind[a], k = kinds[a];
    if (f < AGYM_BFIT_NONE || f > AGYM_BFIT_EMPIRICAL) return set_error(h, AGYM_ERR_INVALID, "agym_set_bidder_fits: unknown fit kind");
    const bool ok = (f == AGYM_BFIT_NONE) || (f == AGYM_BFIT_VL_SEARCH && k == AGYM_BID_SEARCH) || (f == AGYM_BFIT_VL_POLICY && k == AGYM_BID_POLICY) ||
                    (f >= AGYM_BFIT_PL_REINFORCE && f <= AGYM_BFIT_DR && k == AGYM_BID_BANDIT) ||
                    (f == AGYM_BFIT_EMPIRICAL && k == AGYM_BID_GAUSS_CLIP);
    if (!ok) return set_error(h, AGYM_ERR_INVALID, "agym_set_bidder_fits: fit kind does not match the agent's bid kind");
    h->any_winrate_fit |= f == AGYM_BFIT_VL_SEARCH || f == AGYM_BFIT_VL_POLICY || f == AGYM_BFIT_DR;
    h->any_policy_fit |= f >= AGYM_BFIT_VL_POLICY && f <= AGYM_BFIT_DR;
    h->any_empirical_fit |= f == AGYM_BFIT_EMPIRICAL;
    h->any_unassigned_bandit |= (k == AGYM_BID_BANDIT || k == AGYM_BID_POLICY || k == AGYM_BID_SEARCH) && f == AGYM_BFIT_NONE;
  }
  e = cudaMemcpy(h->d_bidder_fit, fit_kind, A * sizeof(int), cudaMemcpyHostToDevice);
  return check_cuda(h, e, "agym_set_bidder_fits");
}

int agym_set_catalog(agym_handle* h, const double* E, const double* V) {
  if (!h || !E || !V) return set_error(h, AGYM_ERR_INVALID, "agym_set_catalog: null argument");
  DeviceGuard g(h->device);
  const agym_shape& s = h->shape;
  const size_t nE = (size_t)s.A * s.I * (s.D + 1), nV = (size_t)s.A * s.I;
  std::vector<float> e32(nE), v32(nV);
  for (size_t i = 0; i < nE; ++i) e32[i] = float(E[i]);
  for (size_t i = 0; i < nV; ++i) v32[i] = float(V[i]);
  cudaError_t e = cudaMemcpy(h->d_E64, E, nE * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(h->d_V64, V, nV * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(h->d_E32, e32.data(), nE * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(h->d_V32, v32.data(), nV * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess && h->d_cat8) {  // D == 5: tiles of 8 items, [8 x {e0 e1 e2 e3}] [8 x {e4 e5 V 0}] (SimParams::cat8)
    const int NT = tiles_of(s.I);
    std::vector<float> c8((size_t)s.A * NT * (kCatTile / 4), 0.f);
    for (int a = 0; a < s.A; ++a)
      for (int i = 0; i < s.I; ++i) {
        const float* r = &e32[((size_t)a * s.I + i) * 6];
        float* t = &c8[((size_t)a * NT + i / kTile) * (kCatTile / 4)];
        float* f0 = t + (i % kTile) * 4;
        float* f1 = t + 32 + (i % kTile) * 4;
        f0[0] = r[0]; f0[1] = r[1]; f0[2] = r[2]; f0[3] = r[3];
        f1[0] = r[4]; f1[1] = r[5]; f1[2] = v32[(size_t)a * s.I + i]; f1[3] = 0.f;
      }
    e = cudaMemcpy(h->d_cat8, c8.data(), c8.size() * sizeof(float), cudaMemcpyHostToDevice);
  }
  if (e != cudaSuccess) return check_cuda(h, e, "agym_set_catalog");
  h->catalog_set = true;
  return AGYM_OK;
}

int agym_bind_allocator_state(agym_handle* h, float* m, float* q, float* m_prev, float* sigma) {
  if (!h || !m || !q || !m_prev || !sigma) return set_error(h, AGYM_ERR_INVALID, "agym_bind_allocator_state: null argument");
  h->m = m; h->q = q; h->m_prev = m_prev; h->sigma = sigma;
  h->pk_valid = false;
  if (h->shape.Do == 4 && h->shape.D == 5 && !h->d_pk) {  // standard shape: packed {m, 1/q} for the production round loop
    DeviceGuard g(h->device);
    const size_t bytes = (size_t)h->shape.R * h->shape.A * tiles_of(h->shape.I) * kPkTile;
    if (cudaMalloc(&h->d_pk, bytes) != cudaSuccess) {
      h->d_pk = nullptr;
      return set_error(h, AGYM_ERR_CUDA, std::string("agym_bind_allocator_state: cudaMalloc of the packed state failed: ") + cudaGetErrorString(cudaGetLastError()));
    }
  }
  return AGYM_OK;
}

int agym_refresh_sigma(agym_handle* h, void* stream) {
  if (!h) return AGYM_ERR_INVALID;
  DeviceGuard g(h->device);
  return launch_refresh_sigma(h, (cudaStream_t)stream);
}

int agym_bind_bidder_state(agym_handle* h, double* bidder_d, float* bidder_w) {
  if (!h || !bidder_d || !bidder_w) return set_error(h, AGYM_ERR_INVALID, "agym_bind_bidder_state: null argument");
  h->bidder_d = bidder_d; h->bidder_w = bidder_w;
  return AGYM_OK;
}

int agym_bind_metrics(agym_handle* h, double* acc, double* revenue) {
  if (!h || !acc || !revenue) return set_error(h, AGYM_ERR_INVALID, "agym_bind_metrics: null argument");
  h->acc = acc; h->revenue = revenue;
  return AGYM_OK;
}

int agym_bind_fit_log(agym_handle* h, float* fit_ctx, uint32_t* fit_meta, int64_t Tcap) {
  if (!h) return AGYM_ERR_INVALID;
  if ((fit_ctx == nullptr) != (fit_meta == nullptr) || Tcap < 0) return set_error(h, AGYM_ERR_INVALID, "agym_bind_fit_log: bad arguments");
  if (h->log_base > 0 && (!fit_ctx || Tcap <= h->log_base + h->rounds_in_iter))
    return set_error(h, AGYM_ERR_INVALID, "agym_bind_fit_log: with log retention the log must hold the retained rows and the rounds recorded so far");
  h->fit_ctx = fit_ctx; h->fit_meta = fit_meta; h->Tcap = fit_ctx ? Tcap : 0;
  return AGYM_OK;
}

int agym_bind_bid_log(agym_handle* h, float* bid_rows, uint32_t* bid_meta, int64_t Tcap) {
  if (!h) return AGYM_ERR_INVALID;
  if ((bid_rows == nullptr) != (bid_meta == nullptr) || Tcap < 0) return set_error(h, AGYM_ERR_INVALID, "agym_bind_bid_log: bad arguments");
  if (h->log_base > 0 && (!bid_rows || Tcap <= h->log_base + h->rounds_in_iter))
    return set_error(h, AGYM_ERR_INVALID, "agym_bind_bid_log: with log retention the log must hold the retained rows and the rounds recorded so far");
  h->bid_rows = bid_rows; h->bid_meta = bid_meta; h->bid_Tcap = bid_rows ? Tcap : 0;
  return AGYM_OK;
}

size_t agym_bidder_workspace_bytes(const agym_handle* h, int64_t Tcap) { return h ? bidder_workspace_bytes(h, Tcap) : 0; }

int agym_bind_bidder_workspace(agym_handle* h, void* ws, size_t bytes) {
  if (!h) return AGYM_ERR_INVALID;
  h->bws = ws; h->bws_bytes = ws ? bytes : 0;
  return AGYM_OK;
}

int agym_update_bidders(agym_handle* h, uint64_t seed, int32_t iter, int32_t max_epochs, float* fit_info, void* stream) {
  if (!h) return AGYM_ERR_INVALID;
  if (h->any_unassigned_bandit)
    return set_error(h, AGYM_ERR_STATE, "agym_update_bidders: an agent bids with a learnt model but has no fit kind (agym_set_bidder_fits)");
  if (!h->any_winrate_fit && !h->any_policy_fit && !h->any_empirical_fit) return AGYM_OK;
  if (!h->bidder_d || !h->bidder_w) return set_error(h, AGYM_ERR_STATE, "agym_update_bidders: bidder state not bound");
  if (!h->bid_rows) return set_error(h, AGYM_ERR_STATE, "agym_update_bidders: bid log not bound (agym_bind_bid_log)");
  if (max_epochs < 0) return set_error(h, AGYM_ERR_INVALID, "agym_update_bidders: max_epochs < 0");
  DeviceGuard g(h->device);
  return launch_update_bidders(h, seed, iter, max_epochs, fit_info, (cudaStream_t)stream);
}

size_t agym_workspace_bytes(const agym_handle* h, int64_t Tcap) { return h ? fit_workspace_bytes(h, Tcap) : 0; }

int agym_bind_workspace(agym_handle* h, void* ws, size_t bytes) {
  if (!h) return AGYM_ERR_INVALID;
  h->ws = ws; h->ws_bytes = ws ? bytes : 0;
  return AGYM_OK;
}

int64_t agym_rounds_in_iteration(const agym_handle* h) { return h ? h->rounds_in_iter : -1; }

int agym_set_rounds_in_iteration(agym_handle* h, int64_t n) {
  if (!h) return AGYM_ERR_INVALID;
  if (n < 0 || (h->fit_ctx && h->log_base + n > h->Tcap) || (h->bid_rows && h->log_base + n > h->bid_Tcap)) return set_error(h, AGYM_ERR_INVALID, "agym_set_rounds_in_iteration: out of range");
  h->rounds_in_iter = n;
  return AGYM_OK;
}

int agym_simulate_rounds(agym_handle* h, uint64_t seed, int32_t iter, int64_t T, const agym_round_log* log, void* stream) {
  if (!h) return AGYM_ERR_INVALID;
  if (T < 0) return set_error(h, AGYM_ERR_INVALID, "agym_simulate_rounds: T < 0");
  int rc = ready_for_rounds(h, "agym_simulate_rounds");
  if (rc) return rc;
  const int64_t S = h->shape.max_slots > 1 ? h->shape.max_slots : 1;  // winner-log rows per round
  if (S > 1 && h->log_base > 0) return set_error(h, AGYM_ERR_UNSUPPORTED, "several slots per round together with log retention (Agent memory) is not built");
  if (h->any_learnt && h->fit_ctx && h->log_base + (h->rounds_in_iter + T) * S > h->Tcap)
    return set_error(h, AGYM_ERR_INVALID, "agym_simulate_rounds: fit log capacity exceeded (call agym_clear_iteration or bind a larger log)");
  if (h->bid_rows && h->log_base + h->rounds_in_iter + T > h->bid_Tcap)
    return set_error(h, AGYM_ERR_INVALID, "agym_simulate_rounds: bid log capacity exceeded");
  if (T == 0) return AGYM_OK;
  DeviceGuard g(h->device);
  SimParams p = make_params(h);
  p.T = T; p.seed = seed; p.iter = iter;
  rc = launch_simulate(h, p, nullptr, log, (cudaStream_t)stream);
  if (rc == AGYM_OK) h->rounds_in_iter += T;
  return rc;
}

int agym_replay_rounds(agym_handle* h, int32_t run0, int32_t n_runs, int64_t T, const agym_replay_inputs* in,
                       const agym_round_log* log, void* stream) {
  if (!h || !in) return set_error(h, AGYM_ERR_INVALID, "agym_replay_rounds: null argument");
  if (run0 < 0 || n_runs < 1 || run0 + n_runs > h->shape.R || T < 0) return set_error(h, AGYM_ERR_INVALID, "agym_replay_rounds: bad run range or T");
  if (!in->ctx || !in->parts || !in->u) return set_error(h, AGYM_ERR_INVALID, "agym_replay_rounds: ctx, parts and u are required");
  int rc = ready_for_rounds(h, "agym_replay_rounds");
  if (rc) return rc;
  if (h->any_shaded && !in->gamma_z && !(h->any_search && in->grid_u && in->grid_n > 0))
    return set_error(h, AGYM_ERR_INVALID, "agym_replay_rounds: shaded bidders need gamma_z (and grid_u once a win-rate model is fitted)");
  const int64_t S = h->shape.max_slots > 1 ? h->shape.max_slots : 1;  // winner-log rows per round
  if (S > 1 && h->log_base > 0) return set_error(h, AGYM_ERR_UNSUPPORTED, "several slots per round together with log retention (Agent memory) is not built");
  if (h->fit_ctx && h->log_base + (h->rounds_in_iter + T) * S > h->Tcap) return set_error(h, AGYM_ERR_INVALID, "agym_replay_rounds: fit log capacity exceeded");
  if (h->bid_rows && h->log_base + h->rounds_in_iter + T > h->bid_Tcap) return set_error(h, AGYM_ERR_INVALID, "agym_replay_rounds: bid log capacity exceeded");
  // (the 128-point search grid grid_u is only read once some win-rate model is initialised)
  if (T == 0) return AGYM_OK;
  DeviceGuard g(h->device);
  SimParams p = make_params(h);
  p.T = T; p.run0 = run0; p.n_runs = n_runs;
  rc = launch_simulate(h, p, in, log, (cudaStream_t)stream);
  if (rc == AGYM_OK && run0 + n_runs == h->shape.R) h->rounds_in_iter += T;  // advance once the last run range was replayed
  return rc;
}

int agym_clear_iteration(agym_handle* h, void* stream) {
  if (!h) return AGYM_ERR_INVALID;
  if (!h->acc || !h->revenue) return set_error(h, AGYM_ERR_STATE, "agym_clear_iteration: metrics not bound");
  DeviceGuard g(h->device);
  const agym_shape& s = h->shape;
  cudaError_t e = cudaMemsetAsync(h->acc, 0, (size_t)s.R * s.A * kNumMetrics * sizeof(double), (cudaStream_t)stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(h->revenue, 0, (size_t)s.R * sizeof(double), (cudaStream_t)stream);
  h->rounds_in_iter = 0;
  if (e == cudaSuccess && h->log_base > 0) return clear_retained_rows(h, (cudaStream_t)stream);
  return check_cuda(h, e, "agym_clear_iteration");
}

int agym_set_log_retention(agym_handle* h, const int32_t* memory, double* terms) {
  if (!h) return AGYM_ERR_INVALID;
  DeviceGuard g(h->device);
  const int A = h->shape.A;
  int64_t total = 0;
  std::vector<int32_t> off(A, 0);
  if (memory)
    for (int a = 0; a < A; ++a) {
      if (memory[a] < 0) return set_error(h, AGYM_ERR_INVALID, "agym_set_log_retention: memory < 0");
      off[a] = int32_t(total);
      total += memory[a];
    }
  if (h->rounds_in_iter != 0 && total != h->log_base)  // the rows already recorded would shift
    return set_error(h, AGYM_ERR_STATE, "agym_set_log_retention: sum(memory) can only change between iterations (rounds are recorded)");
  if (total == 0) {  // retention off
    h->log_base = 0; h->terms = nullptr;
    return AGYM_OK;
  }
  if (!terms) return set_error(h, AGYM_ERR_INVALID, "agym_set_log_retention: terms buffer required");
  if (!h->bid_rows || h->bid_Tcap <= total) return set_error(h, AGYM_ERR_STATE, "agym_set_log_retention: bid log not bound or not larger than sum(memory)");
  if (h->any_learnt && (!h->fit_ctx || h->Tcap <= total)) return set_error(h, AGYM_ERR_STATE, "agym_set_log_retention: winner log not bound or not larger than sum(memory)");
  if (total > 0x7fffffffLL / h->shape.P) return set_error(h, AGYM_ERR_INVALID, "agym_set_log_retention: sum(memory) too large");
  if (!h->d_memory) {
    if (cudaMalloc(&h->d_memory, A * sizeof(int)) != cudaSuccess || cudaMalloc(&h->d_mem_off, A * sizeof(int)) != cudaSuccess)
      return check_cuda(h, cudaGetLastError(), "agym_set_log_retention");
  }
  cudaError_t e = cudaMemcpy(h->d_memory, memory, A * sizeof(int), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(h->d_mem_off, off.data(), A * sizeof(int), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return check_cuda(h, e, "agym_set_log_retention");
  h->log_base = total;
  h->terms = terms;
  return AGYM_OK;
}

int64_t agym_retained_capacity(const agym_handle* h) { return h ? h->log_base : -1; }

int agym_retain_logs(agym_handle* h, void* stream) {
  if (!h) return AGYM_ERR_INVALID;
  if (h->log_base <= 0) return set_error(h, AGYM_ERR_STATE, "agym_retain_logs: log retention not configured (agym_set_log_retention)");
  if (!h->acc) return set_error(h, AGYM_ERR_STATE, "agym_retain_logs: metrics not bound");
  DeviceGuard g(h->device);
  const int rc = launch_retain_logs(h, (cudaStream_t)stream);
  if (rc == AGYM_OK) h->rounds_in_iter = 0;
  return rc;
}

uint64_t agym_launch_count(const agym_handle* h) { return h ? h->launches : 0; }

int agym_set_option(agym_handle* h, const char* name, double value) {
  if (!h || !name) return AGYM_ERR_INVALID;
  static const char* const known[] = {"fit_dense", "fit_nt", "fit_ncap", "fit_warp", "fit_heavy", "sim_g", "sim_cat_smem", "bidfit_wide"};
  for (const char* k : known)
    if (std::string(k) == name) { h->options[name] = value; return AGYM_OK; }
  return set_error(h, AGYM_ERR_INVALID, std::string("agym_set_option: unknown option '") + name + "'");
}

int agym_update_allocators(agym_handle* h, int32_t fit_mode, int32_t max_epochs, float* fit_info, void* stream) {
  if (!h) return AGYM_ERR_INVALID;
  if (fit_mode != AGYM_FIT_ADAM_REF && fit_mode != AGYM_FIT_ADAM_FAST && fit_mode != AGYM_FIT_NEWTON) return set_error(h, AGYM_ERR_INVALID, "agym_update_allocators: unknown fit mode");
  if (!h->any_learnt) return AGYM_OK;
  if (!h->m || !h->q || !h->m_prev || !h->sigma) return set_error(h, AGYM_ERR_STATE, "agym_update_allocators: allocator state not bound");
  if (!h->fit_ctx) return set_error(h, AGYM_ERR_STATE, "agym_update_allocators: fit log not bound");
  DeviceGuard g(h->device);
  const int rc = launch_update_allocators(h, fit_mode, max_epochs, fit_info, (cudaStream_t)stream);
  return rc ? rc : launch_pack_state(h, (cudaStream_t)stream);  // the round loop's packed copy of {m, 1/q} follows every update
}

static int staged_params(agym_handle* h, uint64_t seed, int32_t iter, int64_t T, SimParams* p, const char* who) {
  if (!h->agents_set || !h->catalog_set) return set_error(h, AGYM_ERR_STATE, std::string(who) + ": configuration incomplete");
  if (T < 1) return set_error(h, AGYM_ERR_INVALID, std::string(who) + ": T < 1");
  if (h->shape.A > 256 || h->shape.I > 256) return set_error(h, AGYM_ERR_UNSUPPORTED, std::string(who) + ": staged kernels use uint8 ids (A, I <= 256)");
  *p = make_params(h);
  p->T = T; p->seed = seed; p->iter = iter; p->round0 = 0;
  return AGYM_OK;
}

int agym_estimate_ctr(agym_handle* h, int32_t run, int32_t agent, const double* context, int32_t sample, const float* eps, uint64_t seed,
                      int32_t iter, int64_t query, double* out, void* stream) {
  if (!h || !context || !out) return set_error(h, AGYM_ERR_INVALID, "agym_estimate_ctr: null argument");
  const agym_shape& sh = h->shape;
  if (run < 0 || run >= sh.R || agent < 0 || agent >= sh.A) return set_error(h, AGYM_ERR_INVALID, "agym_estimate_ctr: run / agent out of range");
  if (!h->agents_set || !h->catalog_set) return set_error(h, AGYM_ERR_STATE, "agym_estimate_ctr: configuration incomplete");
  if (h->any_learnt && (!h->m || !h->sigma)) return set_error(h, AGYM_ERR_STATE, "agym_estimate_ctr: allocator state not bound");
  DeviceGuard g(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t nctx = size_t(sh.D + 1), neps = size_t(sh.I) * h->K, nout = size_t(sh.I);
  const size_t bytes = (nctx + nout) * sizeof(double) + neps * sizeof(float);
  if (!h->est_scratch) {
    cudaError_t e = cudaMalloc(&h->est_scratch, bytes);
    if (e != cudaSuccess) return check_cuda(h, e, "agym_estimate_ctr: scratch");
  }
  double* d_ctx = static_cast<double*>(h->est_scratch);
  double* d_out = d_ctx + nctx;
  float* d_eps = reinterpret_cast<float*>(d_out + nout);
  std::vector<double> cpad(nctx, 0.0);
  const bool oracle = h->alloc_kind_host[agent] == AGYM_ALLOC_ORACLE;
  for (size_t i = 0; i < (oracle ? nctx : size_t(h->K)); ++i) cpad[i] = context[i];  // Oracle agents see the true context, learnt ones the observed one (Auction.py:46-49)
  cudaError_t e = cudaMemcpyAsync(d_ctx, cpad.data(), nctx * sizeof(double), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && eps) e = cudaMemcpyAsync(d_eps, eps, neps * sizeof(float), cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return check_cuda(h, e, "agym_estimate_ctr: upload");
  SimParams p = make_params(h);
  p.seed = seed; p.iter = iter; p.round0 = query;
  int rc = launch_estimate(h, p, run, agent, d_ctx, sample, eps ? d_eps : nullptr, d_out, st);
  if (rc) return rc;
  e = cudaMemcpyAsync(out, d_out, nout * sizeof(double), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // the reference's estimate_CTR returns a host array
  return check_cuda(h, e, "agym_estimate_ctr: download");
}

int agym_k1_contexts(agym_handle* h, uint64_t seed, int32_t iter, int64_t T, float* ctx, uint8_t* parts, void* stream) {
  if (!h || !ctx || !parts) return set_error(h, AGYM_ERR_INVALID, "agym_k1_contexts: null argument");
  SimParams p;
  int rc = staged_params(h, seed, iter, T, &p, "agym_k1_contexts");
  if (rc) return rc;
  DeviceGuard g(h->device);
  return launch_k1(h, p, ctx, parts, (cudaStream_t)stream);
}

int agym_k2_allocate(agym_handle* h, uint64_t seed, int32_t iter, int64_t T, const float* ctx, const uint8_t* parts,
                     uint8_t* item, float* est, float* true_ctr, float* best_ev, float* value, void* stream) {
  if (!h || !ctx || !parts || !item || !est || !true_ctr || !best_ev || !value) return set_error(h, AGYM_ERR_INVALID, "agym_k2_allocate: null argument");
  SimParams p;
  int rc = staged_params(h, seed, iter, T, &p, "agym_k2_allocate");
  if (rc) return rc;
  if (h->any_learnt && (!h->m || !h->sigma)) return set_error(h, AGYM_ERR_STATE, "agym_k2_allocate: allocator state not bound");
  DeviceGuard g(h->device);
  return launch_k2(h, p, ctx, parts, item, est, true_ctr, best_ev, value, (cudaStream_t)stream);
}

int agym_k3_bids(agym_handle* h, uint64_t seed, int32_t iter, int64_t T, const uint8_t* parts, const float* est,
                 const float* value, float* bid, float* gamma, float* propensity, void* stream) {
  if (!h || !parts || !est || !value || !bid) return set_error(h, AGYM_ERR_INVALID, "agym_k3_bids: null argument");
  SimParams p;
  int rc = staged_params(h, seed, iter, T, &p, "agym_k3_bids");
  if (rc) return rc;
  if (h->any_shaded && !h->bidder_d) return set_error(h, AGYM_ERR_STATE, "agym_k3_bids: bidder state not bound");
  DeviceGuard g(h->device);
  return launch_k3(h, p, parts, est, value, bid, gamma, propensity, (cudaStream_t)stream);
}

int agym_k4_resolve(agym_handle* h, uint64_t seed, int32_t iter, int64_t T, const float* bid, const float* true_ctr,
                    const float* value, const uint8_t* parts, uint8_t* winner, float* price, float* second,
                    uint8_t* outcome, int32_t accumulate, void* stream) {
  if (!h || !bid || !true_ctr || !value || !parts || !winner || !price || !second || !outcome)
    return set_error(h, AGYM_ERR_INVALID, "agym_k4_resolve: null argument");
  SimParams p;
  int rc = staged_params(h, seed, iter, T, &p, "agym_k4_resolve");
  if (rc) return rc;
  if (accumulate && (!h->acc || !h->revenue)) return set_error(h, AGYM_ERR_STATE, "agym_k4_resolve: metrics not bound");
  DeviceGuard g(h->device);
  return launch_k4(h, p, bid, true_ctr, value, parts, winner, price, second, outcome, accumulate, (cudaStream_t)stream);
}

}  // extern "C"
