// Shared device/host definitions for libagym (sm_100a).  Internal -- the ABI is include/agym.h.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "agym.h"

namespace agym {

constexpr int kAdamTable = 16384;  // BidderAllocation.py:38  epochs = 8192 * 2
constexpr int kAdamTable2 = 32768;  // Bidder.py:240  epochs = 8192 * 4 (win-rate fits)
constexpr uint32_t kBidValid = 1u << 31, kBidWon = 1u << 30, kBidClick = 1u << 29;
constexpr int kMaxP = 32;        // participants per round handled by one lane group (P <= group width)
constexpr int kNumMetrics = AGYM_NUM_METRICS;

// fit_meta packing (agym.h: agym_bind_fit_log)
constexpr uint32_t kMetaValid = 1u << 31;
constexpr uint32_t kMetaClick = 1u << 30;
__host__ __device__ inline uint32_t pack_meta(int agent, int item, bool click) {
  return kMetaValid | (click ? kMetaClick : 0u) | (uint32_t(agent) << 12) | uint32_t(item);
}
__host__ __device__ inline int meta_agent(uint32_t m) { return int((m >> 12) & 0xFFFu); }
__host__ __device__ inline int meta_item(uint32_t m) { return int(m & 0xFFFu); }

// Everything a round-loop kernel needs, passed by value.
constexpr int kTile = 8, kPkTile = 320, kCatTile = 256;
__host__ __device__ inline int tiles_of(int I) { return (I + kTile - 1) / kTile; }

struct SimParams {
  int R, A, I, D, Do, K, P, mechanism;
  int max_slots;     // >= 1; rows of the winner log per round
  int run_offset;
  double embedding_var;
  // static per-agent configuration [A]
  const int* n_items;
  const int* alloc_kind;
  const int* bidder_kind;
  // catalog, both precisions: E [A][I][D+1], V [A][I]
  const double* E64;
  const double* V64;
  const float* E32;
  const float* V32;
  // learnt allocator state [R][A][I][K]
  const float* m;
  const float* sigma;
  // production-mode copies for the standard shape (D = 5, Do = 4), null when not applicable / not current.  Items go in tiles of
  // kTile = 8, field-major inside a tile, so that the 8 lanes of a group read one field of items i .. i+7 as 128 contiguous
  // bytes (16 sectors per warp request instead of the 32 of an array of records):
  //   pk   per (run, agent), tiles_of(I) tiles of kPkTile = 320 B: [8 x float4 {m0 m1 m2 m3}] [8 x float4 {v0 v1 v2 v3}]
  //        [8 x float2 {m4 v4}], v = 1 / q                                            (written by pack_state_kernel)
  //   cat8 per agent, tiles of kCatTile = 256 B: [8 x float4 {e0 e1 e2 e3}] [8 x float4 {e4 e5 V 0}]   (agym_set_catalog)
  const unsigned char* pk;
  const unsigned char* cat8;
  // bidder state
  const double* bidder_d;  // [R][A][AGYM_BIDDER_D]
  const float* bidder_w;   // [R][A][AGYM_BIDDER_W]
  // accumulators
  double* acc;      // [R][A][kNumMetrics]
  double* revenue;  // [R]
  // winner records for the allocator fit
  float* fit_ctx;      // [R][Tcap][Do]
  uint32_t* fit_meta;  // [R][Tcap]
  long long Tcap;
  // per-(round, slot) bid records for the bidder fits (optional)
  float* bid_rows;     // [R][bid_Tcap][P][AGYM_BID_ROW]
  uint32_t* bid_meta;  // [R][bid_Tcap][P]
  long long bid_Tcap;
  double* terms;       // [R][bid_Tcap][P][AGYM_TERM_ROW] metric summands per record (log retention only)
  long long log_base;  // rows reserved at the head of both logs for retained records
  long long round0;  // rounds already simulated in this iteration (append offset and RNG counter base)
  long long T;       // rounds in this launch
  int run0, n_runs;  // runs covered by this launch
  int chunk;         // rounds per CTA
  // production noise
  uint64_t seed;
  int iter;
};

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11), counter-based: key = (seed, run), counter = (round, iter,
// purpose|slot, index).  Written out here; no curand dependency.
// ------------------------------------------------------------------------------------------------
enum Purpose : uint32_t { kPurposeCtx = 0, kPurposePart = 1, kPurposeClick = 2, kPurposeTS = 3, kPurposeGamma = 4, kPurposeGrid = 5, kPurposeSlots = 6 };

struct PhiloxKey {
  uint32_t k0, k1;
};

__host__ __device__ inline PhiloxKey make_key(uint64_t seed, uint32_t global_run) {
  PhiloxKey k;
  k.k0 = uint32_t(seed) ^ (global_run * 0x9E3779B9u);
  k.k1 = uint32_t(seed >> 32) ^ (global_run + 0x7F4A7C15u);
  return k;
}

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, PhiloxKey key) {
  uint32_t k0 = key.k0, k1 = key.k1;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// u32 -> (0, 1]
__device__ __forceinline__ float u32_to_unit_open0(uint32_t x) { return (float(x >> 8) + 1.0f) * (1.0f / 16777216.0f); }
// u32 -> [0, 1)
__device__ __forceinline__ float u32_to_unit(uint32_t x) { return float(x >> 8) * (1.0f / 16777216.0f); }
__device__ __forceinline__ double u32x2_to_unit_d(uint32_t hi, uint32_t lo) {
  return double((uint64_t(hi) << 21) ^ uint64_t(lo >> 11)) * (1.0 / 9007199254740992.0);
}

// Box-Muller: two u32 -> two standard normals (float).
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
  const float u1 = u32_to_unit_open0(a);
  const float r = sqrtf(-2.0f * __logf(u1));
  float s, c;
  __sincosf(6.283185307179586f * u32_to_unit(b), &s, &c);
  return make_float2(r * c, r * s);
}

__device__ __forceinline__ float4 philox_normal4(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, PhiloxKey key) {
  const uint4 w = philox4x32_10(c0, c1, c2, c3, key);
  const float2 a = box_muller(w.x, w.y), b = box_muller(w.z, w.w);
  return make_float4(a.x, a.y, b.x, b.y);
}

// ------------------------------------------------------------------------------------------------
// precision policies
// ------------------------------------------------------------------------------------------------
template <typename Real>
struct Arith;

template <>
struct Arith<float> {
  static constexpr bool kExact = false;
  __device__ static __forceinline__ float sigmoid(float z) { return __fdividef(1.0f, 1.0f + __expf(-z)); }
  // learnt CTR estimate: float either way
  __device__ static __forceinline__ float sigmoid32(float z) { return __fdividef(1.0f, 1.0f + __expf(-z)); }
  __device__ static __forceinline__ float ts_weight(float m, float eps, float s) { return fmaf(eps, s, m); }
  __device__ static __forceinline__ float mac(float w, float x, float acc) { return fmaf(w, x, acc); }
  __device__ static __forceinline__ float neg_inf() { return -INFINITY; }
};

template <>
struct Arith<double> {
  static constexpr bool kExact = true;
  __device__ static __forceinline__ double sigmoid(double z) { return 1.0 / (1.0 + exp(-z)); }
  __device__ static __forceinline__ float sigmoid32(float z) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-z))); }
  // Models.py:31 -- separately rounded multiply and add, as torch's CPU kernels do
  __device__ static __forceinline__ float ts_weight(float m, float eps, float s) { return __fadd_rn(m, __fmul_rn(eps, s)); }
  __device__ static __forceinline__ float mac(float w, float x, float acc) { return __fadd_rn(acc, __fmul_rn(w, x)); }
  __device__ static __forceinline__ double neg_inf() { return -INFINITY; }
};

template <typename Real>
struct Catalog;
template <>
struct Catalog<float> {
  __device__ static __forceinline__ const float* E(const SimParams& p) { return p.E32; }
  __device__ static __forceinline__ const float* V(const SimParams& p) { return p.V32; }
};
template <>
struct Catalog<double> {
  __device__ static __forceinline__ const double* E(const SimParams& p) { return p.E64; }
  __device__ static __forceinline__ const double* V(const SimParams& p) { return p.V64; }
};

// shuffles for both precisions inside a lane group of width G
template <int G>
__device__ __forceinline__ float shfl_xor(float v, int off) { return __shfl_xor_sync(0xffffffffu, v, off, G); }
template <int G>
__device__ __forceinline__ double shfl_xor(double v, int off) { return __shfl_xor_sync(0xffffffffu, v, off, G); }
template <int G>
__device__ __forceinline__ int shfl_xor(int v, int off) { return __shfl_xor_sync(0xffffffffu, v, off, G); }
template <int G, typename T>
__device__ __forceinline__ T shfl_idx(T v, int src) { return __shfl_sync(0xffffffffu, v, src, G); }

}  // namespace agym

// ------------------------------------------------------------------------------------------------
// host-side handle
// ------------------------------------------------------------------------------------------------
struct agym_handle {
  agym_shape shape;
  int device;
  int K;
  // owned device memory (configuration + catalog, tiny)
  int* d_n_items = nullptr;
  int* d_alloc_kind = nullptr;
  int* d_bidder_kind = nullptr;
  int* d_bidder_fit = nullptr;
  int bidder_fit_host[4096];
  std::vector<int32_t> alloc_kind_host;
  bool any_winrate_fit = false, any_policy_fit = false, any_unassigned_bandit = false, any_empirical_fit = false;
  double* d_E64 = nullptr;
  double* d_V64 = nullptr;
  float* d_E32 = nullptr;
  float* d_V32 = nullptr;
  unsigned char* d_cat8 = nullptr;  // tiled float catalog (D == 5), see SimParams
  unsigned char* d_pk = nullptr;    // tiled {m, 1/q} (Do == 4), current while pk_valid
  bool pk_valid = false;
  double* d_adam_bc1 = nullptr;  // [kAdamTable2] 1 - 0.9^t
  float* d_adam_bc2s2 = nullptr; // [kAdamTable2] sqrt(1 - 0.999^t)
  double* d_adam_sz0 = nullptr;  // [kAdamTable] 2e-3 / (1 - 0.9^t), t = epoch + 1 (torch Adam step size at lr 2e-3)
  float* d_adam_bc2s = nullptr;  // [kAdamTable] sqrt(1 - 0.999^t)
  float2* d_adam_ep = nullptr;   // [kAdamTable] {float(d_adam_sz0), d_adam_bc2s}: one 64-bit load per epoch
  void* est_scratch = nullptr;   // device staging of agym_estimate_ctr: context, eps, output (lazy)
  float* k4_scratch = nullptr;   // [R][chunks][A * 5 + 1] float partial sums of the staged resolution kernel (lazy)
  size_t k4_scratch_bytes = 0;
  bool agents_set = false, catalog_set = false;
  bool any_learnt = false, any_shaded = false;
  int max_items = 0;
  // borrowed
  float *m = nullptr, *q = nullptr, *m_prev = nullptr, *sigma = nullptr;
  double* bidder_d = nullptr;
  float* bidder_w = nullptr;
  double *acc = nullptr, *revenue = nullptr;
  float* fit_ctx = nullptr;
  uint32_t* fit_meta = nullptr;
  int64_t Tcap = 0;
  void* ws = nullptr;
  size_t ws_bytes = 0;
  float* bid_rows = nullptr;
  uint32_t* bid_meta = nullptr;
  int64_t bid_Tcap = 0;
  void* bws = nullptr;
  size_t bws_bytes = 0;
  bool any_search = false, any_unbuilt_fit = false;
  int64_t rounds_in_iter = 0;
  // log retention (Agent.memory)
  int* d_memory = nullptr;   // [A]
  int* d_mem_off = nullptr;  // [A] first retained row of each agent
  double* terms = nullptr;   // borrowed
  int64_t log_base = 0;      // sum(memory)
  int num_sms = 148;
  unsigned long long launches = 0;  // kernels of this library launched through this handle (agym_launch_count)
  // second stream for kernels that run beside each other inside one call (fork / join with events; created on first use)
  void* nccl_comm = nullptr;     // ncclComm_t of agym_comm_init (agym_nccl.cu), one per handle
  int nccl_rank = 0, nccl_world = 1;
  int* d_fit_epochs = nullptr;   // [R*A] epochs of each allocator fit in the previous update (launch-order hint, owned)
  size_t fit_epochs_len = 0;
  cudaStream_t aux_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // kernel-selection / tuning overrides set through agym_set_option (tests and experiments; never the environment)
  std::map<std::string, double> options;
  bool has_option(const char* k) const { return options.count(k) != 0; }
  double option(const char* k, double dflt) const { auto it = options.find(k); return it == options.end() ? dflt : it->second; }
  std::string err;
};

namespace agym {
int set_error(agym_handle* h, int code, const std::string& msg);
void destroy_comm(agym_handle* h);
int check_cuda(agym_handle* h, cudaError_t e, const char* what);
SimParams make_params(const agym_handle* h);

// kernels' host launchers (one per translation unit)
int launch_simulate(agym_handle* h, const SimParams& p, const agym_replay_inputs* in, const agym_round_log* log, cudaStream_t s);
// Lane-group width of the round loop (one group per opportunity), the same for the fused and the staged kernels: the Thompson
// noise of an item is addressed through its lane's position.  Option "sim_g" overrides the default (tests run every width).
int sim_group_width(const agym_handle* h, int P, int DMAX);
int launch_refresh_sigma(agym_handle* h, cudaStream_t s);
int launch_pack_state(agym_handle* h, cudaStream_t s);  // rebuilds SimParams::pk from m and q (no-op for other shapes)
int launch_retain_logs(agym_handle* h, cudaStream_t s);
int clear_retained_rows(agym_handle* h, cudaStream_t s);
int launch_update_allocators(agym_handle* h, int fit_mode, int max_epochs, float* fit_info, cudaStream_t s);
size_t fit_workspace_bytes(const agym_handle* h, int64_t Tcap);
int launch_update_bidders(agym_handle* h, uint64_t seed, int iter, int max_epochs, float* fit_info, cudaStream_t s);
size_t bidder_workspace_bytes(const agym_handle* h, int64_t Tcap);
int launch_k1(agym_handle* h, const SimParams& p, float* ctx, uint8_t* parts, cudaStream_t s);
int launch_k2(agym_handle* h, const SimParams& p, const float* ctx, const uint8_t* parts, uint8_t* item, float* est,
              float* true_ctr, float* best_ev, float* value, cudaStream_t s);
int launch_k3(agym_handle* h, const SimParams& p, const uint8_t* parts, const float* est, const float* value, float* bid,
              float* gamma, float* propensity, cudaStream_t s);
int launch_estimate(agym_handle* h, const SimParams& p, int run, int a, const double* ctx, int sample, const float* eps, double* out,
                    cudaStream_t s);
int launch_k4(agym_handle* h, const SimParams& p, const float* bid, const float* true_ctr, const float* value,
              const uint8_t* parts, uint8_t* winner, float* price, float* second, uint8_t* outcome, int accumulate,
              cudaStream_t s);
}  // namespace agym
