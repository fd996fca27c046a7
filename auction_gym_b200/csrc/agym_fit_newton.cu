// K6, OPT-IN fit mode AGYM_FIT_NEWTON -- a different algorithm from the reference's, never the default and never
// part of a parity or headline number.
//
// The reference (src/BidderAllocation.py:29-65) minimises  BCE_sum(sigmoid(x . m[item]), y) + 0.5 * sum q (m_prev - m)^2
// (src/Models.py:39-41) with up to 16 384 full-batch Adam epochs and stops where its scheduler says, short of the
// optimum -- and for an item whose rows are all clicks or all non-clicks that optimum does not exist: the reference leaves
// the intercept column out of the prior (q[:, :-1]), so only its early stop keeps the intercept finite.
// This mode therefore solves the regularised logistic regression the reference's source cites (Chapelle & Li 2011,
// algorithm 3): the same likelihood with the Gaussian prior N(m_prev, 1/q) on ALL columns, the intercept included (its q is
// the one the reference's Laplace update already maintains, Models.py:43-45).  The objective is separable over items (a
// row only touches the weights of its own item) and strongly convex, so each item gets a damped Newton iteration on its
// 5 x 5 system  (X^T W X + diag q) d = g  with step halving whenever the objective does not decrease, until the Newton
// decrement g . d drops below 1e-9: a handful of passes over an item's rows instead of thousands of epochs over all of them.
// The Laplace update (including its literal exp(1 - z)) and update_prior (Models.py:47-48) are the reference's.
//
// One warp per (run, agent); the agent's rows are scattered (stable) into the workspace grouped by item, lanes stride
// the rows of one item at a time, sums cross lanes in FP64 by butterfly (bit-identical in every lane), every lane
// solves the same 5 x 5 Cholesky system.
#include "agym_fit.cuh"

namespace agym {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kNewtonDefaultPasses = 50;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(kFull, v, off);
  return v;
}

// Solves (H + jitter) d = g for the symmetric positive definite K x K matrix H (upper triangle, row-major packed).
template <int K>
__device__ __forceinline__ void chol_solve(const double (&Hp)[K * (K + 1) / 2], const double (&g)[K], double (&d)[K]) {
  double L[K][K];
  int o = 0;
#pragma unroll
  for (int r = 0; r < K; ++r)
#pragma unroll
    for (int c = r; c < K; ++c) L[c][r] = Hp[o++];  // lower triangle
#pragma unroll
  for (int j = 0; j < K; ++j) {
    double s = L[j][j] * (1.0 + 1e-12) + 1e-12;
#pragma unroll
    for (int k = 0; k < j; ++k) s -= L[j][k] * L[j][k];
    s = s > 1e-300 ? s : 1e-300;
    const double inv = rsqrt(s);
    L[j][j] = inv;  // reciprocal of the pivot
#pragma unroll
    for (int i = j + 1; i < K; ++i) {
      double t = L[i][j];
#pragma unroll
      for (int k = 0; k < j; ++k) t -= L[i][k] * L[j][k];
      L[i][j] = t * inv;
    }
  }
  double yv[K];
#pragma unroll
  for (int i = 0; i < K; ++i) {
    double t = g[i];
#pragma unroll
    for (int k = 0; k < i; ++k) t -= L[i][k] * yv[k];
    yv[i] = t * L[i][i];
  }
#pragma unroll
  for (int i = K - 1; i >= 0; --i) {
    double t = yv[i];
#pragma unroll
    for (int k = i + 1; k < K; ++k) t -= L[k][i] * d[k];
    d[i] = t * L[i][i];
  }
}

template <int K>
__global__ void __launch_bounds__(128) fit_newton_kernel(const FitParams p, const int max_passes) {
  extern __shared__ int sm_int[];
  constexpr int Do = K - 1, NH = K * (K + 1) / 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int fit = blockIdx.x * 4 + warp;
  if (fit >= p.R * p.A) return;  // no CTA-wide barrier below
  const int I = p.I;
  int* seg = sm_int + warp * (2 * I + 2);  // [I + 1] first sorted row of every item
  int* cur = seg + I + 1;                  // [I] scatter cursors
  const int run = fit / p.A, a = fit % p.A;
  const int* __restrict__ aoff = p.aoff + (size_t)run * (p.A + 1);
  const int row0 = aoff[a], n = aoff[a + 1] - row0;
  float* info = p.fit_info ? p.fit_info + (size_t)fit * 4 : nullptr;
  if (p.alloc_kind[a] == AGYM_ALLOC_ORACLE) return;
  if (n < 2) {  // BidderAllocation.py:33
    if (lane == 0 && info) { info[0] = -1.f; info[1] = 0.f; info[2] = CUDART_NAN_F; info[3] = float(n); }
    return;
  }
  const int nI = p.n_items[a];
  const size_t soff = ((size_t)run * p.A + a) * I * K;
  const uint32_t* __restrict__ idx = p.srt_idx + (size_t)run * p.Tcap + row0;
  const uint32_t* __restrict__ meta = p.fit_meta + (size_t)run * p.Tcap;
  float* xs = p.srt_x + ((size_t)run * p.Tcap + row0) * K;  // [n][K]: observed context, click

  // ---- rows grouped by item (stable counting sort) ----
  for (int i = lane; i <= I; i += 32) seg[i] = 0;
  __syncwarp();
  for (int j = lane; j < n; j += 32) atomicAdd(&seg[meta_item(meta[idx[j]]) + 1], 1);
  __syncwarp();
  if (lane == 0) {
    int s = 0;
    for (int i = 0; i < I; ++i) { s += seg[i + 1]; seg[i + 1] = s; }
  }
  __syncwarp();
  for (int i = lane; i < I; i += 32) cur[i] = seg[i];
  __syncwarp();
  for (int base = 0; base < n; base += 32) {
    const int j = base + lane;
    int item = -1;
    uint32_t t = 0, mt = 0;
    if (j < n) { t = idx[j]; mt = meta[t]; item = meta_item(mt); }
    const unsigned peers = __match_any_sync(kFull, item);
    const int rank = __popc(peers & ((1u << lane) - 1u));
    if (item >= 0) {
      float* d = xs + (size_t)(cur[item] + rank) * K;
      const float* __restrict__ c = p.fit_ctx + ((size_t)run * p.Tcap + t) * Do;
#pragma unroll
      for (int k = 0; k < Do; ++k) d[k] = c[k];
      d[Do] = (mt & kMetaClick) ? 1.f : 0.f;
    }
    __syncwarp();
    if (item >= 0 && rank == 0) cur[item] += __popc(peers);
    __syncwarp();
  }

  double loss_total = 0.0;
  int passes_max = 0, passes_total = 0;
  for (int i = 0; i < nI; ++i) {
    const int r0 = seg[i], r1 = seg[i + 1];
    const size_t io = soff + (size_t)i * K;
    if (r1 == r0) {  // no rows: m and q untouched, update_prior still copies m
      if (lane < K) p.m_prev[io + lane] = p.m[io + lane];
      continue;
    }
    float macc[K], mtry[K], m0[K], qk[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      macc[k] = mtry[k] = p.m[io + k];
      m0[k] = p.m_prev[io + k];
      qk[k] = p.q[io + k];  // prior on every column, the intercept included (see the header)
    }
    double loss_acc = INFINITY, delta[K], alpha = 1.0;
#pragma unroll
    for (int k = 0; k < K; ++k) delta[k] = 0.0;
    int passes = 0;
    while (true) {
      float g[K], H[NH], ls = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) g[k] = 0.f;
#pragma unroll
      for (int k = 0; k < NH; ++k) H[k] = 0.f;
      for (int r = r0 + lane; r < r1; r += 32) {
        const float* xr = xs + (size_t)r * K;
        float x[K];
#pragma unroll
        for (int k = 0; k < Do; ++k) x[k] = xr[k];
        x[Do] = 1.f;
        const bool y = xr[Do] > 0.5f;
        float z = mtry[Do];
#pragma unroll
        for (int k = 0; k < Do; ++k) z = fmaf(mtry[k], x[k], z);
        // e = exp(-|z|): no cancellation in 1 - P, P (1 - P) or the log-likelihood however large |z| is
        const float e = expf(-fabsf(z)), inv = __fdiv_rn(1.0f, 1.0f + e), ei = e * inv;
        const bool pos = z >= 0.f;
        const float p1 = pos ? inv : ei, p0 = pos ? ei : inv;  // P(click), 1 - P(click)
        const float w = ei * inv;
        const float gz = y ? -p0 : p1;
        ls += log1pf(e) + (y == pos ? 0.f : fabsf(z));  // -log of the row's likelihood
        int o = 0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          g[k] = fmaf(gz, x[k], g[k]);
          const float wx = w * x[k];
#pragma unroll
          for (int l = k; l < K; ++l) { H[o] = fmaf(wx, x[l], H[o]); ++o; }
        }
      }
      ++passes;
      double gd[K], Hd[NH];
      double lsd = warp_sum(double(ls));
#pragma unroll
      for (int k = 0; k < K; ++k) gd[k] = warp_sum(double(g[k]));
#pragma unroll
      for (int k = 0; k < NH; ++k) Hd[k] = warp_sum(double(H[k]));
      {
        int o = 0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const double dk = double(mtry[k]) - double(m0[k]), qd = double(qk[k]);
          lsd += 0.5 * qd * dk * dk;
          gd[k] += qd * dk;
          Hd[o] += qd;
          o += K - k;
        }
      }
      // accept unless the objective went up (beyond what float32 row arithmetic can resolve) or is not a number
      if (!(lsd <= loss_acc + 1e-6 * (1.0 + fabs(loss_acc)))) {
        alpha *= 0.5;
        if (alpha < 1.0 / 1024 || passes >= max_passes) break;
#pragma unroll
        for (int k = 0; k < K; ++k) mtry[k] = float(double(macc[k]) - alpha * delta[k]);
        continue;
      }
#pragma unroll
      for (int k = 0; k < K; ++k) macc[k] = mtry[k];
      loss_acc = lsd;
      chol_solve<K>(Hd, gd, delta);
      double dec = 0.0;
#pragma unroll
      for (int k = 0; k < K; ++k) dec += gd[k] * delta[k];  // Newton decrement squared
      if (dec < 1e-9 || passes >= max_passes) break;
      alpha = 1.0;
#pragma unroll
      for (int k = 0; k < K; ++k) mtry[k] = float(double(macc[k]) - delta[k]);
    }
    // ---- Laplace approximation at the fitted m: q += sum P (1 - P) x^2 with P = 1 / (1 + exp(1 - z)) (Models.py:43-45) ----
    float qi[K];
#pragma unroll
    for (int k = 0; k < K; ++k) qi[k] = 0.f;
    for (int r = r0 + lane; r < r1; r += 32) {
      const float* xr = xs + (size_t)r * K;
      float x[K];
#pragma unroll
      for (int k = 0; k < Do; ++k) x[k] = xr[k];
      x[Do] = 1.f;
      float z = macc[Do];
#pragma unroll
      for (int k = 0; k < Do; ++k) z = fmaf(macc[k], x[k], z);
      const float t = fminf(1.0f + expf(1.0f - z), 1e38f);
      const float P = __fdiv_rn(1.0f, t);
      const float w = P * (1.0f - P);
#pragma unroll
      for (int k = 0; k < K; ++k) qi[k] = fmaf(w * x[k], x[k], qi[k]);
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const float inc = float(warp_sum(double(qi[k])));
      if (lane == k) {
        const float qv = p.q[io + k] + inc;
        p.m[io + k] = macc[k];
        p.m_prev[io + k] = macc[k];  // update_prior (Models.py:47-48)
        p.q[io + k] = qv;
        p.sigma[io + k] = __fdiv_rn(1.0f, __fsqrt_rn(qv));
      }
    }
    loss_total += loss_acc;
    passes_total += passes;
    passes_max = passes > passes_max ? passes : passes_max;
  }
  if (lane == 0 && info) { info[0] = float(passes_max); info[1] = float(passes_total); info[2] = float(loss_total); info[3] = float(n); }
}

}  // namespace

int launch_fit_newton(agym_handle* h, const FitParams& fp, int max_passes, cudaStream_t s) {
  if (fp.K != 5) return set_error(h, AGYM_ERR_UNSUPPORTED, "AGYM_FIT_NEWTON: obs_embedding_size must be 4");
  const long long fits = (long long)fp.R * fp.A;
  const size_t smem = 4 * size_t(2 * fp.I + 2) * sizeof(int);
  if (smem > 200 * 1024) return set_error(h, AGYM_ERR_UNSUPPORTED, "AGYM_FIT_NEWTON: too many items");
  if (smem > 48 * 1024) {
    const cudaError_t e = cudaFuncSetAttribute(fit_newton_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return check_cuda(h, e, "fit_newton_kernel attribute");
  }
  fit_newton_kernel<5><<<unsigned((fits + 3) / 4), 128, smem, s>>>(fp, max_passes > 0 ? max_passes : kNewtonDefaultPasses);
  h->launches += 1;
  return check_cuda(h, cudaGetLastError(), "fit_newton_kernel");
}

}  // namespace agym
