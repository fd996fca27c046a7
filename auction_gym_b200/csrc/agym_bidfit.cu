// K7: the bidder fits of Agent.update -> bidder.update (reference src/Bidder.py:210-325, 369-431, 477-615 and
// src/Models.py:51-218), one CTA per (run, agent), all epochs on-chip.
//
//   ValueLearningBidder('search')   win-rate fit                                              Bidder.py:210-260
//   ValueLearningBidder('policy')   win-rate fit, then the policy maximises W * (V - gamma V)  Bidder.py:278-316
//   PolicyLearningBidder(loss)      [initialise_policy], then REINFORCE / off-policy / TRPO / PPO  Bidder.py:369-431
//   DoublyRobustBidder              win-rate fit, u_hat = W (V - P), [initialise_policy], DR loss   Bidder.py:477-615
//
// Kernels (launched in this order by launch_update_bidders; agents that do not need a stage exit at once):
//   bidrows_bucket_kernel  CTA per run: stable counting sort of the (round, slot) bid records by agent
//   gather_rows_kernel     CTA per (run, agent): compact rows {est, value, gamma, propensity | utility, won, u_hat, -}
//   winrate_fit_kernel     logistic regression (3 weights + bias) with the gamma = 0 augmentation; BCELoss(mean);
//                          Adam(lr 3e-3, wd 1e-6, amsgrad) + ReduceLROnPlateau + "no 1e-6 improvement" stop
//   policy_fit_kernel      the 12-parameter Gaussian policy (2 -> 2 -> {mu, sigma}, softplus) with hand-written
//                          back-propagation; stage IMITATE = initialise_policy (Models.py:110-133), stage MAIN = the
//                          bidder's loss.  Every epoch: one row-parallel pass, a 13-value block reduction (loss + 12
//                          gradient components), Adam(amsgrad, weight decay) + scheduler + stop rule replicated per thread.
// The stochastic objectives (Doubly Robust, 'policy') draw their rsample noise from Philox keyed by
// (seed, run, iteration, epoch, row), so a fit is reproducible and independent of the launch geometry.
#include <math_constants.h>

#include <cstdlib>

#include "agym_common.cuh"

namespace agym {

constexpr int kRowF = 8;  // floats per compact row
struct BidFitParams {
  int R, A, P;
  long long Tcap, Tn;
  const int* fit_kind;      // [A] agym_bidder_fit
  const float* rows;        // [R][Tcap][P][AGYM_BID_ROW]
  const uint32_t* meta;     // [R][Tcap][P]
  uint32_t* srt_idx;        // [R][Tcap*P] record indices grouped by agent (stable)
  int* aoff;                // [R][A+1]
  int* wins;                // [R][A] won rows; -1 = skip the remaining stages (Bidder.py:213-216 fallback)
  float* crow;              // [R][Tcap*P][kRowF] compact rows, grouped by agent
  double* bidder_d;         // [R][A][AGYM_BIDDER_D]
  float* bidder_w;          // [R][A][AGYM_BIDDER_W]
  float* info;              // [R][A][4] or null (last stage that ran for the agent)
  const double* bc1;        // [kAdamTable2]  1 - 0.9^t
  const float* bc2s;        // [kAdamTable2]  sqrt(1 - 0.999^t)
  int max_epochs, ncap, stage;
  int run_offset;
  uint64_t seed;
  int iter;
};

__device__ __forceinline__ bool fits_winrate(int k) { return k == AGYM_BFIT_VL_SEARCH || k == AGYM_BFIT_VL_POLICY || k == AGYM_BFIT_DR; }
__device__ __forceinline__ bool fits_policy(int k) { return k >= AGYM_BFIT_VL_POLICY && k <= AGYM_BFIT_DR; }
__device__ __forceinline__ bool needs_imitation(int k) { return k >= AGYM_BFIT_PL_REINFORCE && k <= AGYM_BFIT_DR; }

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
    if ((mt & kBidValid) && p.fit_kind[mt & 0xFFFu] != AGYM_BFIT_NONE) atomicAdd(&hist[mt & 0xFFFu], 1);
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
        if ((mt & kBidValid) && p.fit_kind[mt & 0xFFFu] != AGYM_BFIT_NONE) keyv = int(mt & 0xFFFu);
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

// Agent.update's gathering of the logged arrays (Agent.py:81-90) + the utilities of Bidder.py:219-220,371-372,479-480
__global__ void __launch_bounds__(256) gather_rows_kernel(const BidFitParams p) {
  __shared__ int wins_s;
  const int run = blockIdx.x / p.A, a = blockIdx.x % p.A;
  const int kind = p.fit_kind[a];
  if (kind == AGYM_BFIT_NONE) return;
  const int* __restrict__ aoff = p.aoff + (size_t)run * (p.A + 1);
  const int row0 = aoff[a], n = aoff[a + 1] - row0;
  if (threadIdx.x == 0) wins_s = 0;
  __syncthreads();
  const uint32_t* __restrict__ idx = p.srt_idx + (size_t)run * p.Tcap * p.P + row0;
  const uint32_t* __restrict__ meta = p.meta + (size_t)run * p.Tcap * p.P;
  const float* __restrict__ rows = p.rows + (size_t)run * p.Tcap * p.P * AGYM_BID_ROW;
  float4* __restrict__ out = reinterpret_cast<float4*>(p.crow + ((size_t)run * p.Tcap * p.P + row0) * kRowF);
  int my = 0;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const uint32_t r = idx[j], mt = meta[r];
    const float* __restrict__ s = rows + (size_t)r * AGYM_BID_ROW;
    const bool won = (mt & kBidWon) != 0;
    my += won;
    const float util = won ? ((mt & kBidClick) ? s[1] : 0.0f) - s[4] : 0.0f;  // value * outcome - price on won rows
    out[2 * j] = make_float4(s[0], s[1], s[2], fmaxf(s[3], 1e-15f));            // propensities clipped at 1e-15 (Bidder.py:385,571)
    out[2 * j + 1] = make_float4(util, won ? 1.f : 0.f, 0.f, 0.f);
  }
  if (my) atomicAdd(&wins_s, my);
  __syncthreads();
  if (threadIdx.x == 0) {
    int w = wins_s;
    if (w == 0 && (kind == AGYM_BFIT_VL_SEARCH || kind == AGYM_BFIT_VL_POLICY)) {
      // Bidder.py:213-216 -- lost every auction: fall back to the Gaussian logging policy, fit nothing
      p.bidder_d[((size_t)run * p.A + a) * AGYM_BIDDER_D + 2] = 0.0;
      w = -1;
      if (p.info) { float* f = p.info + ((size_t)run * p.A + a) * 12; f[0] = -1.f; f[1] = 0.f; f[2] = CUDART_NAN_F; f[3] = float(n); }
    }
    p.wins[(size_t)run * p.A + a] = w;
  }
}

__device__ __forceinline__ float clamp_log(float x) { return fmaxf(logf(x), -100.f); }  // BCELoss clamps its logs at -100
__device__ __forceinline__ float sigmoidf_rn(float z) { return __fdiv_rn(1.0f, 1.0f + expf(-z)); }
__device__ __forceinline__ float softplusf(float x) { return x > 20.f ? x : log1pf(expf(x)); }   // torch.nn.Softplus(beta=1, threshold=20)
__device__ __forceinline__ float dsoftplusf(float x) { return x > 20.f ? 1.f : sigmoidf_rn(x); }

template <int NV, int NT>
__device__ __forceinline__ void block_sum(float (&v)[NV], float* red /*[2][NT/32][NV]*/, int epoch, int tid) {
#pragma unroll
  for (int k = 0; k < NV; ++k) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], off);
  }
  float* r = red + (size_t)(epoch & 1) * (NT / 32) * NV;
  if ((tid & 31) == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) r[(tid >> 5) * NV + k] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) t += r[w * NV + k];
    v[k] = t;
  }
}

// Adam(weight_decay, amsgrad=True), ReduceLROnPlateau('min') and the "no 1e-6 improvement for `stop_after` epochs" rule
// of the reference's training loops, replicated in every thread (all inputs are block-uniform).
template <int NPAR>
struct Trainer {
  float ea[NPAR], es[NPAR], mx[NPAR];
  double lr, wd, best_sched = INFINITY, best_loss = INFINITY, factor, min_lr, threshold;
  int bad = 0, best_epoch = -1, patience, stop_after;
  bool use_sched;
  __device__ Trainer(double lr_, double wd_, bool sched, int patience_, double factor_, double min_lr_, double threshold_, int stop_after_)
      : lr(lr_), wd(wd_), factor(factor_), min_lr(min_lr_), threshold(threshold_), patience(patience_), stop_after(stop_after_), use_sched(sched) {
#pragma unroll
    for (int k = 0; k < NPAR; ++k) ea[k] = es[k] = mx[k] = 0.f;
  }
  __device__ __forceinline__ void adam(float (&w)[NPAR], const float* grad, double bc1, float bc2s) {
    const float alpha = float(-(lr / bc1)), wdf = float(wd);
#pragma unroll
    for (int k = 0; k < NPAR; ++k) {
      const float g = fmaf(wdf, w[k], grad[k]);
      ea[k] = fmaf(g - ea[k], 0.1f, ea[k]);
      es[k] = fmaf(0.001f * g, g, es[k] * 0.999f);
      mx[k] = fmaxf(mx[k], es[k]);
      w[k] += __fdiv_rn(alpha * ea[k], __fdiv_rn(__fsqrt_rn(mx[k]), bc2s) + 1e-8f);
    }
  }
  // returns true when training stops at this epoch
  __device__ __forceinline__ bool after_epoch(double cur, int epoch) {
    if (use_sched) {
      if (cur < best_sched * (1.0 - threshold)) { best_sched = cur; bad = 0; } else { ++bad; }
      if (bad > patience) {
        const double new_lr = fmax(lr * factor, min_lr);
        if (lr - new_lr > 1e-8) lr = new_lr;
        bad = 0;
      }
    }
    if (best_loss - cur > 1e-6) { best_epoch = epoch; best_loss = cur; return false; }
    return epoch - best_epoch > stop_after;
  }
};

// ------------------------------------------------------------------------------------------------
// win-rate model (Models.py:51-62)
// ------------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(NT) winrate_fit_kernel(const BidFitParams p) {
  extern __shared__ __align__(16) float4 srow[];  // [ncap] {est, value, gamma, won}
  __shared__ float red[2 * (NT / 32) * 5];
  const int run = blockIdx.x / p.A, a = blockIdx.x % p.A;
  const int kind = p.fit_kind[a];
  if (!fits_winrate(kind)) return;
  if (p.wins[(size_t)run * p.A + a] < 0) return;
  const int tid = threadIdx.x;
  const int* __restrict__ aoff = p.aoff + (size_t)run * (p.A + 1);
  const int row0 = aoff[a], n = aoff[a + 1] - row0;
  if (n == 0) return;
  double* __restrict__ bd = p.bidder_d + ((size_t)run * p.A + a) * AGYM_BIDDER_D;
  float* __restrict__ bw = p.bidder_w + ((size_t)run * p.A + a) * AGYM_BIDDER_W;
  float* info = p.info ? p.info + ((size_t)run * p.A + a) * 12 : nullptr;  // stage slot 0
  float4* __restrict__ crow = reinterpret_cast<float4*>(p.crow + ((size_t)run * p.Tcap * p.P + row0) * kRowF);
  for (int j = tid; j < n && j < p.ncap; j += NT) {
    const float4 x = crow[2 * j], y = crow[2 * j + 1];
    srow[j] = make_float4(x.x, x.y, x.z, y.y);
  }
  __syncthreads();
  float w[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) w[k] = bw[k];
  const bool dr = kind == AGYM_BFIT_DR;
  // Bidder.py:242-243 (ValueLearning) / :520-521 (DoublyRobust)
  Trainer<4> tr(3e-3, 1e-6, true, dr ? 256 : 100, dr ? 0.2 : 0.1, 1e-7, 1e-4, dr ? 1024 : 512);
  const float invN = 1.0f / float(2 * n);
  int stop_epoch = -1, epochs_run = 0;
  float last_loss = 0.f;
  for (int epoch = 0; epoch < p.max_epochs; ++epoch) {
    float part[5] = {0, 0, 0, 0, 0};  // loss, dL/dw0, dL/dw1, dL/dw2, dL/db  (sums; divided by N below)
    for (int j = tid; j < n; j += NT) {
      float4 r;
      if (j < p.ncap) r = srow[j];
      else { const float4 x = crow[2 * j], y = crow[2 * j + 1]; r = make_float4(x.x, x.y, x.z, y.y); }
      const float base = fmaf(r.y, w[1], fmaf(r.x, w[0], w[3]));
      const float p1 = sigmoidf_rn(fmaf(r.z, w[2], base));  // the logged row
      const float p0 = sigmoidf_rn(base);                    // its gamma = 0 copy, labelled lost (Bidder.py:227-236)
      part[0] -= (r.w > 0.5f ? clamp_log(p1) : clamp_log(1.0f - p1)) + clamp_log(1.0f - p0);
      const float g1 = p1 - r.w, gs = g1 + p0;
      part[1] = fmaf(gs, r.x, part[1]);
      part[2] = fmaf(gs, r.y, part[2]);
      part[3] = fmaf(g1, r.z, part[3]);
      part[4] += gs;
    }
    block_sum<5, NT>(part, red, epoch, tid);
    const float loss = part[0] * invN;
    float grad[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) grad[k] = part[k + 1] * invN;
    tr.adam(w, grad, p.bc1[epoch], p.bc2s[epoch]);
    epochs_run = epoch + 1;
    last_loss = loss;
    if (tr.after_epoch(double(loss), epoch)) { stop_epoch = epoch; break; }
  }
  __syncthreads();
  if (dr) {  // Bidder.py:545-551  u_hat = W (V - P) on the logged gammas
    for (int j = tid; j < n; j += NT) {
      const float4 x = crow[2 * j];
      float4 y = crow[2 * j + 1];
      const float W = sigmoidf_rn(fmaf(x.z, w[2], fmaf(x.y, w[1], fmaf(x.x, w[0], w[3]))));
      const float V = x.x * x.y;
      y.z = W * (V - V * x.z);
      crow[2 * j + 1] = y;
    }
  }
  if (tid == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) bw[k] = w[k];
    if (kind == AGYM_BFIT_VL_SEARCH) bd[2] = 1.0;  // model_initialised = True (Bidder.py:325)
    if (info) { info[0] = float(stop_epoch); info[1] = float(epochs_run); info[2] = last_loss; info[3] = float(n); }
  }
}

// ------------------------------------------------------------------------------------------------
// Gaussian bid-shading policy (Models.py:65-218)
// ------------------------------------------------------------------------------------------------
struct PolicyFwd {
  float h0, h1, s0, s1, a_mu, a_sg, mu, sg_raw, sigma;
};

__device__ __forceinline__ PolicyFwd policy_forward(const float (&th)[12], float x0, float x1) {
  PolicyFwd f;
  f.h0 = fmaf(x1, th[1], fmaf(x0, th[0], th[4]));
  f.h1 = fmaf(x1, th[3], fmaf(x0, th[2], th[5]));
  f.s0 = softplusf(f.h0);
  f.s1 = softplusf(f.h1);
  f.a_mu = fmaf(f.s1, th[7], fmaf(f.s0, th[6], th[8]));
  f.a_sg = fmaf(f.s1, th[10], fmaf(f.s0, th[9], th[11]));
  f.mu = softplusf(f.a_mu);
  f.sg_raw = softplusf(f.a_sg);
  f.sigma = f.sg_raw + 1e-2f;  // min_sigma (Models.py:80,104)
  return f;
}

__device__ __forceinline__ void policy_backward(const float (&th)[12], const PolicyFwd& f, float x0, float x1, float d_mu, float d_sg,
                                                float (&acc)[13]) {
  const float d_amu = d_mu * dsoftplusf(f.a_mu), d_asg = d_sg * dsoftplusf(f.a_sg);
  const float d_h0 = fmaf(d_asg, th[9], d_amu * th[6]) * dsoftplusf(f.h0);
  const float d_h1 = fmaf(d_asg, th[10], d_amu * th[7]) * dsoftplusf(f.h1);
  acc[1] = fmaf(d_h0, x0, acc[1]);
  acc[2] = fmaf(d_h0, x1, acc[2]);
  acc[3] = fmaf(d_h1, x0, acc[3]);
  acc[4] = fmaf(d_h1, x1, acc[4]);
  acc[5] += d_h0;
  acc[6] += d_h1;
  acc[7] = fmaf(d_amu, f.s0, acc[7]);
  acc[8] = fmaf(d_amu, f.s1, acc[8]);
  acc[9] += d_amu;
  acc[10] = fmaf(d_asg, f.s0, acc[10]);
  acc[11] = fmaf(d_asg, f.s1, acc[11]);
  acc[12] += d_asg;
}

enum { kStageImitate = 0, kStageMain = 1 };

template <int NT>
__global__ void __launch_bounds__(NT) policy_fit_kernel(const BidFitParams p) {
  __shared__ float red[2 * (NT / 32) * 13];
  const int run = blockIdx.x / p.A, a = blockIdx.x % p.A;
  const int kind = p.fit_kind[a];
  if (!fits_policy(kind)) return;
  if (p.wins[(size_t)run * p.A + a] < 0) return;
  double* __restrict__ bd = p.bidder_d + ((size_t)run * p.A + a) * AGYM_BIDDER_D;
  const bool imitate = p.stage == kStageImitate;
  if (imitate && (!needs_imitation(kind) || bd[2] != 0.0)) return;  // Bidder.py:381-382,567-568: first update only
  const int tid = threadIdx.x;
  const int* __restrict__ aoff = p.aoff + (size_t)run * (p.A + 1);
  const int row0 = aoff[a], n = aoff[a + 1] - row0;
  if (n == 0) return;
  float* __restrict__ bw = p.bidder_w + ((size_t)run * p.A + a) * AGYM_BIDDER_W;
  float* info = p.info ? p.info + ((size_t)run * p.A + a) * 12 + (p.stage == kStageImitate ? 4 : 8) : nullptr;
  const float4* __restrict__ crow = reinterpret_cast<const float4*>(p.crow + ((size_t)run * p.Tcap * p.P + row0) * kRowF);
  float th[12], ww[4];
#pragma unroll
  for (int k = 0; k < 12; ++k) th[k] = bw[4 + k];
#pragma unroll
  for (int k = 0; k < 4; ++k) ww[k] = bw[k];
  // hyper-parameters: Models.py:113-115 | Bidder.py:283-286 | :389-392 | :575-578
  double lr = 2e-3, wd = 1e-4, factor = 0.2, min_lr = 1e-8, threshold = 1e-4;
  int stop_after = 512, cap = kAdamTable;
  bool sched = true;
  if (imitate) { lr = 1e-3; sched = false; }
  else if (kind == AGYM_BFIT_VL_POLICY) { wd = 1e-6; factor = 0.1; min_lr = 1e-7; stop_after = 256; }
  else if (kind == AGYM_BFIT_DR) { lr = 7e-3; threshold = 5e-3; cap = kAdamTable2; }
  Trainer<12> tr(lr, wd, sched, 100, factor, min_lr, threshold, stop_after);
  const int max_epochs = p.max_epochs > 0 && p.max_epochs < cap ? p.max_epochs : cap;
  const float invn = 1.0f / float(n);
  const PhiloxKey key = make_key(p.seed, uint32_t(p.run_offset + run));
  const float lo = 1.0f / 50.0f, hi = 50.0f;  // importance_weight_clipping_eps = 50 (Bidder.py:398,584)
  int stop_epoch = -1, epochs_run = 0;
  float last_loss = 0.f;
  bool nan_seen = false;
  for (int epoch = 0; epoch < max_epochs; ++epoch) {
    float acc[13];
#pragma unroll
    for (int k = 0; k < 13; ++k) acc[k] = 0.f;
    for (int j = tid; j < n; j += NT) {
      const float4 x = crow[2 * j], y = crow[2 * j + 1];  // {est, value, gamma, prop}, {utility, won, u_hat, -}
      const PolicyFwd f = policy_forward(th, x.x, x.y);
      float d_mu = 0.f, d_sg = 0.f;
      if (imitate) {  // Models.py:122-124: MSE(mu, gamma) + MSE(sigma without min_sigma, 0.05)
        const float e_mu = f.mu - x.z, e_sg = f.sg_raw - 0.05f;
        acc[0] = fmaf(e_mu, e_mu, fmaf(e_sg, e_sg, acc[0]));
        d_mu = 2.0f * e_mu;
        d_sg = 2.0f * e_sg;
      } else {
        float t = 0.f, dt_dmu = 0.f, dt_dsg = 0.f, iw = 0.f;
        if (kind != AGYM_BFIT_VL_POLICY) {  // normal_pdf (Models.py:157-165), clipped at 1e-30
          const float dm = f.mu - x.z, inv_s = __fdiv_rn(1.0f, f.sigma), z = dm * inv_s;
          const float t_raw = expf(-0.5f * z * z) * inv_s * 0.3989422804014327f;
          if (t_raw > 1e-30f) {
            t = t_raw;
            dt_dmu = -t * dm * inv_s * inv_s;
            dt_dsg = t * (dm * dm * inv_s * inv_s * inv_s - inv_s);
          } else {
            t = 1e-30f;
          }
          iw = __fdiv_rn(t, x.w);
        }
        const float u = y.x;
        if (kind == AGYM_BFIT_PL_REINFORCE) {            // Models.py:173-174
          acc[0] -= t * u;
          d_mu = -u * dt_dmu; d_sg = -u * dt_dsg;
        } else if (kind == AGYM_BFIT_PL_OFFPOLICY) {     // Models.py:176-178
          acc[0] -= iw * u;
          const float c = -__fdiv_rn(u, x.w);
          d_mu = c * dt_dmu; d_sg = c * dt_dsg;
        } else if (kind == AGYM_BFIT_PL_TRPO) {          // Models.py:180-187, KL_weight 5e-2
          const float dm = f.mu - x.z, s2 = f.sigma * f.sigma;
          acc[0] += 0.05f * (__fdiv_rn(s2 + dm * dm, 2.0f * s2) - 0.5f) - iw * u;
          const float c = -__fdiv_rn(u, x.w);
          d_mu = fmaf(c, dt_dmu, 0.05f * __fdiv_rn(dm, s2));
          d_sg = fmaf(c, dt_dsg, -0.05f * __fdiv_rn(dm * dm, s2 * f.sigma));
        } else if (kind == AGYM_BFIT_PL_PPO) {           // Models.py:189-196
          const float cl = fminf(fmaxf(iw, lo), hi);
          acc[0] -= fminf(iw * u, cl * u);
          const bool pass = (iw >= lo && iw <= hi) || (iw > hi && u < 0.f) || (iw < lo && u > 0.f);
          const float c = pass ? -__fdiv_rn(u, x.w) : 0.f;
          d_mu = c * dt_dmu; d_sg = c * dt_dsg;
        } else {                                         // Doubly Robust (Models.py:198-218) / 'policy' (Bidder.py:292-302)
          const float eps = philox_normal4(uint32_t(j), uint32_t(epoch), (6u << 16) | uint32_t(a), uint32_t(p.iter), key).x;
          const float raw = fmaf(f.sigma, eps, f.mu);
          const float gs = fminf(fmaxf(raw, 0.f), 1.f);
          const float W = sigmoidf_rn(fmaf(gs, ww[2], fmaf(x.y, ww[1], fmaf(x.x, ww[0], ww[3]))));
          const float V = x.x * x.y;
          acc[0] -= W * (V - V * gs);
          if (raw > 0.f && raw < 1.f) {
            const float dd = -V * (W * (1.0f - W) * ww[2] * (1.0f - gs) - W);
            d_mu = dd; d_sg = dd * eps;
          }
          if (kind == AGYM_BFIT_DR) {
            const float du = u - y.z, cl = fminf(fmaxf(iw, lo), hi);
            acc[0] -= du * cl;
            if (iw >= lo && iw <= hi) {
              const float c = -__fdiv_rn(du, x.w);
              d_mu = fmaf(c, dt_dmu, d_mu); d_sg = fmaf(c, dt_dsg, d_sg);
            }
          }
        }
      }
      policy_backward(th, f, x.x, x.y, d_mu, d_sg, acc);
    }
    block_sum<13, NT>(acc, red, epoch, tid);
    const float loss = acc[0] * invn;
    float grad[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) grad[k] = acc[k + 1] * invn;
    tr.adam(th, grad, p.bc1[epoch], p.bc2s[epoch]);
    epochs_run = epoch + 1;
    last_loss = loss;
    nan_seen |= !(loss == loss);
    if (tr.after_epoch(double(loss), epoch)) { stop_epoch = epoch; break; }
  }
  if (tid == 0) {
#pragma unroll
    for (int k = 0; k < 12; ++k) bw[4 + k] = th[k];
    if (!imitate) bd[2] = 1.0;  // model_initialised = True (Bidder.py:325,430,614)
    if (info) { info[0] = float(stop_epoch); info[1] = float(epochs_run); info[2] = nan_seen ? CUDART_NAN_F : last_loss; info[3] = float(n); }
  }
}


// ------------------------------------------------------------------------------------------------
// EmpiricalShadedBidder.update (Bidder.py:60-125): bucketise the logged gammas (bucket width 0.005 between the smallest
// and the largest gamma), estimate mean utility and its standard error per bucket, move prev_gamma to the centre of the
// bucket with the best lower confidence bound (mean - 1.96 stderr; the highest bucket among ties), clipped to [0, 1].
// ------------------------------------------------------------------------------------------------
constexpr int kEmpMaxBuckets = 2048;
__global__ void __launch_bounds__(256) empirical_update_kernel(const BidFitParams p) {
  constexpr int NT = 256;
  __shared__ float cnt[kEmpMaxBuckets], sum[kEmpMaxBuckets], dev2[kEmpMaxBuckets];
  __shared__ float redmin[NT / 32], redmax[NT / 32];
  const int run = blockIdx.x / p.A, a = blockIdx.x % p.A;
  if (p.fit_kind[a] != AGYM_BFIT_EMPIRICAL) return;
  const int tid = threadIdx.x;
  const int* __restrict__ aoff = p.aoff + (size_t)run * (p.A + 1);
  const int row0 = aoff[a], n = aoff[a + 1] - row0;
  float* info = p.info ? p.info + ((size_t)run * p.A + a) * 12 : nullptr;
  if (n == 0) return;
  const float4* __restrict__ crow = reinterpret_cast<const float4*>(p.crow + ((size_t)run * p.Tcap * p.P + row0) * kRowF);
  float lo = INFINITY, hi = -INFINITY;
  for (int j = tid; j < n; j += NT) { const float g = crow[2 * j].z; lo = fminf(lo, g); hi = fmaxf(hi, g); }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) { lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, off)); hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, off)); }
  if ((tid & 31) == 0) { redmin[tid >> 5] = lo; redmax[tid >> 5] = hi; }
  __syncthreads();
  for (int w = 0; w < NT / 32; ++w) { lo = fminf(lo, redmin[w]); hi = fmaxf(hi, redmax[w]); }
  const double dlo = double(lo), dhi = double(hi);
  // num_buckets = int((max - min) // grid_delta) + 1  (Bidder.py:82).  Python's // is the floor of the EXACT quotient
  // (1.0 // 0.005 == 199 although 1.0 / 0.005 rounds to 200.0), so the rounded quotient is corrected with an fma remainder.
  double qd = floor((dhi - dlo) / 0.005);
  const double rem = fma(-qd, 0.005, dhi - dlo);
  if (rem < 0.0) qd -= 1.0; else if (rem >= 0.005) qd += 1.0;
  int nb = int(qd) + 1;
  if (nb > kEmpMaxBuckets) nb = kEmpMaxBuckets;
  const int nbk = nb - 1;                                  // np.linspace(min, max, nb) has nb - 1 intervals
  const double step = nbk > 0 ? (dhi - dlo) / double(nbk) : 0.0;
  for (int b = tid; b < nbk; b += NT) { cnt[b] = 0.f; sum[b] = 0.f; dev2[b] = 0.f; }
  __syncthreads();
  auto bucket_of = [&](float g) -> int {                   // bucket_lo <= gamma < bucket_hi  (Bidder.py:92)
    if (nbk <= 0) return -1;
    int b = int((double(g) - dlo) / step);
    if (b >= nbk) b = nbk - 1;
    while (b > 0 && double(g) < dlo + step * b) --b;
    while (b < nbk - 1 && double(g) >= dlo + step * (b + 1)) ++b;
    const double e_lo = dlo + step * b, e_hi = (b + 1 == nbk) ? dhi : dlo + step * (b + 1);
    return (double(g) >= e_lo && double(g) < e_hi) ? b : -1;
  };
  for (int j = tid; j < n; j += NT) {
    const int b = bucket_of(crow[2 * j].z);
    if (b >= 0) { atomicAdd(&cnt[b], 1.0f); atomicAdd(&sum[b], crow[2 * j + 1].x); }
  }
  __syncthreads();
  for (int j = tid; j < n; j += NT) {  // second pass: squared deviations from the bucket mean (np.std, population)
    const int b = bucket_of(crow[2 * j].z);
    if (b >= 0) { const float d = crow[2 * j + 1].x - sum[b] / cnt[b]; atomicAdd(&dev2[b], d * d); }
  }
  __syncthreads();
  if (tid == 0) {
    int best = -1;
    float best_lb = -INFINITY;
    for (int b = 0; b < nbk; ++b) {
      if (cnt[b] > 1.5f) {  // num_samples > 1  (Bidder.py:95)
        const float mean = sum[b] / cnt[b], se = sqrtf(dev2[b] / cnt[b]) / sqrtf(cnt[b]);
        const float lb = mean - 1.96f * se;
        if (lb >= best_lb) { best_lb = lb; best = b; }      // reversed nanargmax: the highest bucket among ties (Bidder.py:119)
      }
    }

    if (best >= 0) {
      const double b_lo = dlo + step * best, b_hi = (best + 1 == nbk) ? dhi : dlo + step * (best + 1);
      double g = (b_hi - b_lo) / 2.0 + b_lo;                // Bidder.py:90
      g = g < 0.0 ? 0.0 : (g > 1.0 ? 1.0 : g);
      p.bidder_d[((size_t)run * p.A + a) * AGYM_BIDDER_D + 0] = g;  // self.prev_gamma = best_gamma
    }
    if (info) { info[0] = float(best); info[1] = float(nbk); info[2] = best_lb; info[3] = float(n); }
  }
}

// ------------------------------------------------------------------------------------------------
size_t bidder_workspace_bytes(const agym_handle* h, int64_t Tcap) {
  const agym_shape& s = h->shape;
  const size_t NR = (size_t)s.R * Tcap * s.P;
  return NR * sizeof(uint32_t) + (size_t)s.R * (s.A + 1) * sizeof(int) + (size_t)s.R * s.A * sizeof(int) + NR * kRowF * sizeof(float) + 1024;
}

int launch_update_bidders(agym_handle* h, uint64_t seed, int iter, int max_epochs, float* fit_info, cudaStream_t s) {
  const agym_shape& sh = h->shape;
  if (h->rounds_in_iter <= 0 && h->log_base <= 0) return AGYM_OK;
  const int64_t Tn = h->log_base + h->rounds_in_iter;  // retained rows (if any) come first
  if (h->bws == nullptr || h->bws_bytes < bidder_workspace_bytes(h, h->bid_Tcap))
    return set_error(h, AGYM_ERR_STATE, "agym_update_bidders: workspace not bound or too small (agym_bidder_workspace_bytes)");
  BidFitParams bp{};
  bp.R = sh.R; bp.A = sh.A; bp.P = sh.P;
  bp.Tcap = h->bid_Tcap; bp.Tn = Tn;
  bp.fit_kind = h->d_bidder_fit;
  bp.rows = h->bid_rows; bp.meta = h->bid_meta;
  const size_t NR = (size_t)sh.R * h->bid_Tcap * sh.P;
  auto align = [](unsigned char* w) { return reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(w) + 255) & ~uintptr_t(255)); };
  unsigned char* w = align(static_cast<unsigned char*>(h->bws));
  bp.srt_idx = reinterpret_cast<uint32_t*>(w); w = align(w + NR * sizeof(uint32_t));
  bp.aoff = reinterpret_cast<int*>(w); w = align(w + (size_t)sh.R * (sh.A + 1) * sizeof(int));
  bp.wins = reinterpret_cast<int*>(w); w = align(w + (size_t)sh.R * sh.A * sizeof(int));
  bp.crow = reinterpret_cast<float*>(w);
  bp.bidder_d = h->bidder_d; bp.bidder_w = h->bidder_w;
  bp.info = fit_info;
  bp.bc1 = h->d_adam_bc1; bp.bc2s = h->d_adam_bc2s2;
  bp.max_epochs = max_epochs;
  bp.run_offset = sh.run_offset; bp.seed = seed; bp.iter = iter;
  const unsigned grid = unsigned(sh.R) * unsigned(sh.A);
  bidrows_bucket_kernel<<<sh.R, 256, (2 * sh.A + 1) * sizeof(int), s>>>(bp);
  gather_rows_kernel<<<grid, 256, 0, s>>>(bp);
  h->launches += 2;
  int rc = check_cuda(h, cudaGetLastError(), "bidder fit prologue");
  if (rc) return rc;
  // few fits (the shipped configs have 6 - 18): every CTA has an SM to itself and an epoch is a latency chain over the
  // fit's rows, so wider CTAs shorten it; with many fits 256 threads keep more of them resident
  bool wide = grid <= unsigned(h->num_sms);
  if (h->has_option("bidfit_wide")) wide = h->option("bidfit_wide", 0) != 0;
  if (h->any_winrate_fit) {
    long long ncap = 2 * (Tn * sh.P / sh.A) + 64;  // expected rows per agent x 2
    if (ncap > Tn * sh.P) ncap = Tn * sh.P;
    if (ncap > 12000) ncap = 12000;  // 192 KB of float4
    BidFitParams wp = bp;
    wp.ncap = int(ncap);
    wp.max_epochs = max_epochs > 0 && max_epochs < kAdamTable2 ? max_epochs : kAdamTable2;  // Bidder.py:240  epochs = 8192 * 4
    const size_t smem = (size_t)wp.ncap * sizeof(float4);
    cudaError_t e;
    if (wide) {
      e = cudaFuncSetAttribute(winrate_fit_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
      if (e == cudaSuccess) winrate_fit_kernel<512><<<grid, 512, smem, s>>>(wp);
    } else {
      e = cudaFuncSetAttribute(winrate_fit_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
      if (e == cudaSuccess) winrate_fit_kernel<256><<<grid, 256, smem, s>>>(wp);
    }
    if (e != cudaSuccess) return check_cuda(h, e, "winrate_fit_kernel attribute");
    h->launches += 1;
    if ((rc = check_cuda(h, cudaGetLastError(), "winrate_fit_kernel"))) return rc;
  }
  if (h->any_policy_fit) {
    BidFitParams pp = bp;
    for (int stage = kStageImitate; stage <= kStageMain; ++stage) {
      pp.stage = stage;
      if (wide) policy_fit_kernel<512><<<grid, 512, 0, s>>>(pp);
      else policy_fit_kernel<256><<<grid, 256, 0, s>>>(pp);
    }
    h->launches += 2;
    if ((rc = check_cuda(h, cudaGetLastError(), "policy_fit_kernel"))) return rc;
  }
  if (h->any_empirical_fit) {
    empirical_update_kernel<<<grid, 256, 0, s>>>(bp);
    h->launches += 1;
    if ((rc = check_cuda(h, cudaGetLastError(), "empirical_update_kernel"))) return rc;
  }
  return AGYM_OK;
}

}  // namespace agym
