// K7 (first part): ValueLearningBidder.update -- the win-rate model P(win | estimated CTR, value, gamma)
// (reference src/Bidder.py:210-260, src/Models.py:51-62): logistic regression with 3 weights + bias on the agent's logged
// rows plus the augmentation "with gamma = 0 you would have lost" (Bidder.py:223-236), BCELoss(mean),
// Adam(lr 3e-3, weight_decay 1e-6, amsgrad=True), ReduceLROnPlateau(patience 100, factor 0.1, min_lr 1e-7), at most
// 32 768 epochs, early stop after 512 epochs without a 1e-6 improvement.  An agent that won nothing this iteration gets
// `initialised = 0` instead (Bidder.py:213-216).
//
//   bidrows_bucket_kernel  one CTA per run: stable counting sort of the (round, slot) bid records by agent
//   winrate_fit_kernel     one CTA per (run, agent): rows staged in shared memory as float4 {est, value, gamma, won};
//                          every epoch is one row-parallel pass (both the logged and the augmented row of a record are
//                          evaluated together), a 5-value block reduction (loss + 4 gradient components), and the
//                          4-parameter Adam / scheduler / stop state machine replicated in every thread.
#include <math_constants.h>

#include "agym_common.cuh"

namespace agym {

struct BidFitParams {
  int R, A, P;
  long long Tcap, Tn;
  const int* bidder_kind;
  const float* rows;        // [R][Tcap][P][AGYM_BID_ROW]
  const uint32_t* meta;     // [R][Tcap][P]
  uint32_t* srt_idx;        // [R][Tcap*P] record indices grouped by agent (stable)
  int* aoff;                // [R][A+1]
  float4* spill;            // [R][Tcap*P] rows that overflow shared memory
  double* bidder_d;         // [R][A][AGYM_BIDDER_D]
  float* bidder_w;          // [R][A][AGYM_BIDDER_W]
  float* info;              // [R][A][4] or null
  const double* bc1;        // [kAdamTable2]
  const float* bc2s;        // [kAdamTable2]
  int max_epochs, ncap;
};

__device__ __forceinline__ bool learns_winrate(int kind) { return kind == AGYM_BID_SEARCH; }

__global__ void __launch_bounds__(256) bidrows_bucket_kernel(const BidFitParams p) {
  extern __shared__ int sm_i[];
  int* hist = sm_i;              // [A+1]
  int* cursor = sm_i + p.A + 1;  // [A]
  const int run = blockIdx.x;
  const long long NR = p.Tn * p.P;
  const uint32_t* __restrict__ meta = p.meta + (size_t)run * p.Tcap * p.P;
  for (int a = threadIdx.x; a <= p.A; a += blockDim.x) hist[a] = 0;
  __syncthreads();
  for (long long j = threadIdx.x; j < NR; j += blockDim.x) {
    const uint32_t mt = meta[j];
    if ((mt & kBidValid) && learns_winrate(p.bidder_kind[mt & 0xFFFu])) atomicAdd(&hist[mt & 0xFFFu], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int sum = 0;
    for (int a = 0; a < p.A; ++a) { const int c = hist[a]; hist[a] = sum; cursor[a] = sum; sum += c; }
    hist[p.A] = sum;
  }
  __syncthreads();
  for (int a = threadIdx.x; a <= p.A; a += blockDim.x) p.aoff[(size_t)run * (p.A + 1) + a] = hist[a];
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    uint32_t* __restrict__ out = p.srt_idx + (size_t)run * p.Tcap * p.P;
    for (long long base = 0; base < NR; base += 32) {
      const long long j = base + lane;
      int keyv = -1;
      if (j < NR) {
        const uint32_t mt = meta[j];
        if ((mt & kBidValid) && learns_winrate(p.bidder_kind[mt & 0xFFFu])) keyv = int(mt & 0xFFFu);
      }
      const unsigned peers = __match_any_sync(0xffffffffu, keyv);
      const int rank = __popc(peers & ((1u << lane) - 1u));
      if (keyv >= 0) out[cursor[keyv] + rank] = uint32_t(j);
      __syncwarp();
      if (keyv >= 0 && rank == 0) cursor[keyv] += __popc(peers);
      __syncwarp();
    }
  }
}

__device__ __forceinline__ float clamp_log(float x) { return fmaxf(logf(x), -100.f); }  // BCELoss clamps its logs at -100

__global__ void __launch_bounds__(256) winrate_fit_kernel(const BidFitParams p) {
  constexpr int NT = 256, NW = NT / 32;
  extern __shared__ __align__(16) float4 srow[];  // [ncap] {est, value, gamma, won}
  __shared__ float red[2][NW][5];
  __shared__ int wins_s;
  const int run = blockIdx.x / p.A, a = blockIdx.x % p.A;
  if (!learns_winrate(p.bidder_kind[a])) return;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int* __restrict__ aoff = p.aoff + (size_t)run * (p.A + 1);
  const int row0 = aoff[a], n = aoff[a + 1] - row0;
  double* __restrict__ bd = p.bidder_d + ((size_t)run * p.A + a) * AGYM_BIDDER_D;
  float* __restrict__ bw = p.bidder_w + ((size_t)run * p.A + a) * AGYM_BIDDER_W;
  float* info = p.info ? p.info + ((size_t)run * p.A + a) * 4 : nullptr;
  // ---- stage the agent's rows ----
  if (tid == 0) wins_s = 0;
  __syncthreads();
  const uint32_t* __restrict__ idx = p.srt_idx + (size_t)run * p.Tcap * p.P + row0;
  const uint32_t* __restrict__ meta = p.meta + (size_t)run * p.Tcap * p.P;
  const float* __restrict__ rows = p.rows + (size_t)run * p.Tcap * p.P * AGYM_BID_ROW;
  float4* __restrict__ spill = p.spill + (size_t)run * p.Tcap * p.P + row0;
  int my_wins = 0;
  for (int j = tid; j < n; j += NT) {
    const uint32_t r = idx[j];
    const float* __restrict__ src = rows + (size_t)r * AGYM_BID_ROW;
    const bool won = (meta[r] & kBidWon) != 0;
    my_wins += won;
    const float4 v = make_float4(src[0], src[1], src[2], won ? 1.f : 0.f);
    if (j < p.ncap) srow[j] = v; else spill[j] = v;
  }
  if (my_wins) atomicAdd(&wins_s, my_wins);
  __syncthreads();
  if (wins_s == 0) {  // Bidder.py:213-216 -- lost every auction: fall back to the un-shaded Gaussian logging policy
    if (tid == 0) {
      bd[2] = 0.0;
      if (info) { info[0] = -1.f; info[1] = 0.f; info[2] = CUDART_NAN_F; info[3] = float(n); }
    }
    return;
  }
  float w[4], ea[4] = {0, 0, 0, 0}, es[4] = {0, 0, 0, 0}, mx[4] = {0, 0, 0, 0};
#pragma unroll
  for (int k = 0; k < 4; ++k) w[k] = bw[k];
  const float invN = 1.0f / float(2 * n);
  double lr = 3e-3, best_sched = INFINITY, best_loss = INFINITY;
  int bad = 0, best_epoch = -1, stop_epoch = -1, epochs_run = 0;
  float last_loss = 0.f;
  for (int epoch = 0; epoch < p.max_epochs; ++epoch) {
    float part[5] = {0, 0, 0, 0, 0};  // loss, dL/dw0, dL/dw1, dL/dw2, dL/db  (sums; divided by N below)
    for (int j = tid; j < n; j += NT) {
      const float4 r = j < p.ncap ? srow[j] : spill[j];
      const float base = fmaf(r.y, w[1], fmaf(r.x, w[0], w[3]));
      const float p1 = __fdiv_rn(1.0f, 1.0f + expf(-fmaf(r.z, w[2], base)));  // the logged row
      const float p0 = __fdiv_rn(1.0f, 1.0f + expf(-base));                    // its gamma = 0 copy, labelled lost
      part[0] -= (r.w > 0.5f ? clamp_log(p1) : clamp_log(1.0f - p1)) + clamp_log(1.0f - p0);
      const float g1 = p1 - r.w, gs = g1 + p0;
      part[1] = fmaf(gs, r.x, part[1]);
      part[2] = fmaf(gs, r.y, part[2]);
      part[3] = fmaf(g1, r.z, part[3]);
      part[4] += gs;
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) part[k] += __shfl_xor_sync(0xffffffffu, part[k], off);
    }
    float(*r2)[5] = red[epoch & 1];
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < 5; ++k) r2[wid][k] = part[k];
    }
    __syncthreads();
    float tot[5] = {0, 0, 0, 0, 0};
#pragma unroll
    for (int v = 0; v < NW; ++v) {
#pragma unroll
      for (int k = 0; k < 5; ++k) tot[k] += r2[v][k];
    }
    const float loss = tot[0] * invN;
    // Adam with L2 weight decay and amsgrad (torch/optim/adam.py, single-tensor path)
    const float alpha = float(-(lr / p.bc1[epoch]));
    const float bc2s = p.bc2s[epoch];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float g = fmaf(1e-6f, w[k], tot[k + 1] * invN);
      ea[k] = fmaf(g - ea[k], 0.1f, ea[k]);
      es[k] = fmaf(0.001f * g, g, es[k] * 0.999f);
      mx[k] = fmaxf(mx[k], es[k]);
      w[k] += __fdiv_rn(alpha * ea[k], __fdiv_rn(__fsqrt_rn(mx[k]), bc2s) + 1e-8f);
    }
    epochs_run = epoch + 1;
    last_loss = loss;
    const double cur = double(loss);
    // ReduceLROnPlateau('min', patience=100, factor=0.1, min_lr=1e-7)  (Bidder.py:243)
    if (cur < best_sched * (1.0 - 1e-4)) { best_sched = cur; bad = 0; } else { ++bad; }
    if (bad > 100) {
      const double new_lr = fmax(lr * 0.1, 1e-7);
      if (lr - new_lr > 1e-8) lr = new_lr;
      bad = 0;
    }
    // Bidder.py:255-260
    if (best_loss - cur > 1e-6) { best_epoch = epoch; best_loss = cur; }
    else if (epoch - best_epoch > 512) { stop_epoch = epoch; break; }
  }
  if (tid == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) bw[k] = w[k];
    bd[2] = 1.0;  // model_initialised = True (Bidder.py:325)
    if (info) { info[0] = float(stop_epoch); info[1] = float(epochs_run); info[2] = last_loss; info[3] = float(n); }
  }
}

size_t bidder_workspace_bytes(const agym_handle* h, int64_t Tcap) {
  const agym_shape& s = h->shape;
  const size_t NR = (size_t)s.R * Tcap * s.P;
  return NR * sizeof(uint32_t) + (size_t)s.R * (s.A + 1) * sizeof(int) + 256 + NR * sizeof(float4) + 256;
}

int launch_update_bidders(agym_handle* h, int max_epochs, float* fit_info, cudaStream_t s) {
  const agym_shape& sh = h->shape;
  const int64_t Tn = h->rounds_in_iter;
  if (Tn <= 0) return AGYM_OK;
  if (h->bws == nullptr || h->bws_bytes < bidder_workspace_bytes(h, h->bid_Tcap))
    return set_error(h, AGYM_ERR_STATE, "agym_update_bidders: workspace not bound or too small (agym_bidder_workspace_bytes)");
  BidFitParams bp{};
  bp.R = sh.R; bp.A = sh.A; bp.P = sh.P;
  bp.Tcap = h->bid_Tcap; bp.Tn = Tn;
  bp.bidder_kind = h->d_bidder_kind;
  bp.rows = h->bid_rows; bp.meta = h->bid_meta;
  const size_t NR = (size_t)sh.R * h->bid_Tcap * sh.P;
  unsigned char* w = static_cast<unsigned char*>(h->bws);
  w = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(w) + 255) & ~uintptr_t(255));
  bp.srt_idx = reinterpret_cast<uint32_t*>(w); w += NR * sizeof(uint32_t);
  bp.aoff = reinterpret_cast<int*>(w); w += (size_t)sh.R * (sh.A + 1) * sizeof(int);
  w = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(w) + 255) & ~uintptr_t(255));
  bp.spill = reinterpret_cast<float4*>(w);
  bp.bidder_d = h->bidder_d; bp.bidder_w = h->bidder_w;
  bp.info = fit_info;
  bp.bc1 = h->d_adam_bc1; bp.bc2s = h->d_adam_bc2s2;
  bp.max_epochs = max_epochs > 0 ? max_epochs : kAdamTable2;  // Bidder.py:240  epochs = 8192 * 4
  bidrows_bucket_kernel<<<sh.R, 256, (2 * sh.A + 1) * sizeof(int), s>>>(bp);
  int rc = check_cuda(h, cudaGetLastError(), "bidrows_bucket_kernel");
  if (rc) return rc;
  long long ncap = 2 * (Tn * sh.P / sh.A) + 64;  // expected rows per agent x 2
  if (ncap > Tn * sh.P) ncap = Tn * sh.P;
  if (ncap > 12000) ncap = 12000;  // 192 KB of float4
  bp.ncap = int(ncap);
  const size_t smem = (size_t)bp.ncap * sizeof(float4);
  cudaError_t e = cudaFuncSetAttribute(winrate_fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  if (e != cudaSuccess) return check_cuda(h, e, "winrate_fit_kernel attribute");
  winrate_fit_kernel<<<unsigned(sh.R) * unsigned(sh.A), 256, smem, s>>>(bp);
  return check_cuda(h, cudaGetLastError(), "winrate_fit_kernel");
}

}  // namespace agym
