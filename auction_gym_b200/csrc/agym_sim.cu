// K1-K5 fused: one auction opportunity per lane group (reference src/Auction.py:28-74).
//
// A lane group (G = 8/16/32 lanes, picked from the catalog width) owns one opportunity:
//   context draw (Auction.py:33,36) -> participants (Auction.py:42) -> for each participant slot:
//   item scores over lanes (Agent.py:29-42, BidderAllocation.py:67-68,81-82, Models.py:28-33),
//   shuffle arg-max with lowest-index tie-break, MAP re-estimate of the chosen item, true CTR and
//   best expected value (Auction.py:52-53), bid (Bidder.py:34-35,171-179) -> running top-2 with
//   lowest-slot tie-break (AuctionAllocation.py:18-35) -> Bernoulli click (Auction.py:65) ->
//   charge / set_price / revenue and every per-agent metric (Agent.py:70-118, main.py:131-148).
// The same body serves production mode (in-kernel Philox noise) and replay mode (host-drawn noise).
#include <cstdlib>
#include <type_traits>

#include "agym_round.cuh"

namespace agym {

#ifndef AGYM_SIM_MINB
#define AGYM_SIM_MINB 3
#endif
constexpr int kCatSmThreads = 768;  // one CTA per SM when the catalog is staged in shared memory: 24 warps, as 3 CTAs of 256 threads
template <typename Real, int G, int DMAX, bool kReplay, int DT, int DoT, bool kMulti, bool kCatSm = false>
__global__ void __launch_bounds__(kCatSm ? kCatSmThreads : 256, kCatSm ? 1 : (DT > 0 ? AGYM_SIM_MINB : 1))
sim_kernel(const SimParams p, const agym_replay_inputs in, const agym_round_log log) {
  using A_ = Arith<Real>;
  extern __shared__ __align__(16) unsigned char sim_smem[];
  uint32_t cat_sm = 0;
  if (kCatSm) {  // the static catalog tiles -> shared memory, once per CTA
    const int n16 = p.A * tiles_of(p.I) * (kCatTile / 16);
    const float4* __restrict__ src = reinterpret_cast<const float4*>(p.cat8);
    for (int j = threadIdx.x; j < n16; j += blockDim.x) reinterpret_cast<float4*>(sim_smem)[j] = src[j];
    __syncthreads();
    cat_sm = uint32_t(__cvta_generic_to_shared(sim_smem));
  }
  const int lane = threadIdx.x % G, group = threadIdx.x / G, ngroups = blockDim.x / G;
  const int chunks = int((p.T + p.chunk - 1) / p.chunk);
  const int rl = blockIdx.x / chunks, ck = blockIdx.x % chunks;
  const int run = p.run0 + rl;
  const long long tb = (long long)ck * p.chunk;
  const long long te = (tb + p.chunk < p.T) ? tb + p.chunk : p.T;
  const int iters = (p.chunk + ngroups - 1) / ngroups;
  const int D = DT > 0 ? DT : p.D, Do = DT > 0 ? DoT : p.Do, P = p.P, A = p.A;
  const PhiloxKey key = make_key(p.seed, uint32_t(p.run_offset + run));
  const bool has_log = log.agent || log.item || log.est || log.value || log.bid || log.true_ctr || log.best_ev || log.price || log.second ||
                       log.gamma || log.propensity || log.outcome || log.won || log.winner || log.ctx;

  for (int it = 0; it < iters; ++it) {
    const long long t_raw = tb + (long long)it * ngroups + group;
    const bool active = t_raw < te;
    const long long t = active ? t_raw : tb;
    const long long ta = p.round0 + t;             // round index inside the iteration
    const long long ri = (long long)rl * p.T + t;  // index into per-launch arrays
    const RoundCounter rc(ta, p.iter);

    // ---- context (Auction.py:33) ----
    Real ctx[DMAX];
    if (kReplay) {
#pragma unroll
      for (int d = 0; d < DMAX; ++d) ctx[d] = d < D ? Real(in.ctx[ri * D + d]) : Real(0);
    } else {
      draw_context<Real, G, DMAX>(ctx, D, Real(p.embedding_var), rc, key, lane);
    }

    // ---- participants (Auction.py:42): lane s holds the agent of slot s ----
    int my_agent = 0;
    if (kReplay) {
      if (lane < P) my_agent = in.parts[ri * P + lane];
    } else {
      my_agent = draw_participants<G>(P, A, rc, key, lane);
    }

    // ---- bids ----
    Real best = A_::neg_inf(), second = A_::neg_inf();
    int wslot = 0;
    int r_item = 0;
    Real r_est = 0, r_val = 0, r_bid = 0, r_true = 1, r_bev = 0;
    Real r_gamma = Real(CUDART_NAN), r_prop = Real(CUDART_NAN);

    for (int s = 0; s < P; ++s) {
      const int a = shfl_idx<G>(my_agent, s);
      const float* eps_slot = (kReplay && in.ts_eps) ? in.ts_eps + ((size_t)ri * P + s) * p.I * (Do + 1) : nullptr;
      const SlotEval<Real> ev = eval_slot<Real, G, DMAX, kReplay, DT, DoT, kCatSm>(p, run, a, s, ctx, rc, key, eps_slot, lane, cat_sm);
      Real gamma, prop;
      int eff;
      Real bid = shade_bid<Real>(p, run, a, s, ev.value, ev.est, kReplay,
                                 (kReplay && in.gamma_z) ? in.gamma_z[ri * P + s] : 0.0, rc, key, gamma, prop, eff);
      if (eff == AGYM_BID_SEARCH) {  // Bidder.py:180-196
        const double* gu = (kReplay && in.grid_u) ? in.grid_u + ((size_t)ri * P + s) * in.grid_n : nullptr;
        gamma = search_gamma<Real, G, kReplay>(p, run, a, s, bid, ev.est, ev.value, gu, kReplay ? in.grid_n : 0, rc, key, lane);
        prop = Real(1);
        bid = bid * gamma;
      }
      // running top-2, strict '>' keeps the lowest slot on ties (AuctionAllocation.py:19,33)
      if (bid > best) { second = best; best = bid; wslot = s; }
      else if (bid > second) { second = bid; }
      if (lane == s) {
        r_item = ev.item; r_est = ev.est; r_val = ev.value; r_bid = bid; r_true = ev.true_sel; r_bev = ev.best_ev;
        r_gamma = gamma; r_prop = prop;
      }
    }

    // ---- resolution + click (AuctionAllocation.py:18-35, Auction.py:60-74) ----
    // per participant (lane < P): did it win a slot, what was it charged, what does its log record say afterwards
    const bool valid = P >= 2;  // P == 1: the reference's price array is empty and nobody is charged
    const int S = kMulti ? p.max_slots : 1;  // rows of the winner log per round (kMulti: a separate instantiation, the single-slot hot path pays nothing)
    bool won_me = false, clk_me = false;
    Real paid_me = 0, sec_me = 0, logged_price = 0;
    int top_slot = wslot, my_rank = 0, n_charged = valid ? 1 : 0;
    if (S <= 1) {
      const Real price = valid ? (p.mechanism == AGYM_FIRST_PRICE ? best : second) : Real(0);
      const Real tw = shfl_idx<G>(r_true, wslot);
      Real u;
      if (kReplay) u = Real(in.u[ri]);
      else u = sizeof(Real) == 8 ? Real(click_uniform_d(rc, key)) : Real(click_uniform_f(rc, key));
      won_me = valid && lane == wslot;
      clk_me = won_me && (u < tw);
      paid_me = price;
      sec_me = valid ? second : Real(0);
      logged_price = price;
      my_rank = won_me ? 0 : 1;
    } else {
      // Several slots (Auction.py:30,60-74; AuctionAllocation.py:19-23,33-35): slot k goes to the k-th highest bid (equal
      // bids keep slot order), price_k = sorted[k] (first price) or sorted[k + 1] (second price), second_k = sorted[k + 1].
      // The reference zips winners, prices and second_prices, so only min(num_slots, P - 1) slots are charged.  Every
      // slot's set_price overwrites the logged price of all the others: afterwards every participant's record holds the
      // LAST slot's price, while utilities and revenue were charged slot by slot.
      const int ns = kReplay ? (in.num_slots ? in.num_slots[ri] : 1) : draw_num_slots(S, rc, key);
      n_charged = min(ns, P - 1);
      const Real my_bid = lane < P ? r_bid : A_::neg_inf();
      for (int j2 = 0; j2 < P; ++j2) {
        const Real bj = shfl_idx<G>(my_bid, j2);
        my_rank += (bj > my_bid) || (bj == my_bid && j2 < lane);
      }
      const int shift = p.mechanism == AGYM_FIRST_PRICE ? 0 : 1;
      for (int j2 = 0; j2 < P; ++j2) {  // sorted[q] = the bid of the participant ranked q
        const Real bj = shfl_idx<G>(my_bid, j2);
        const int rj = shfl_idx<G>(my_rank, j2);
        if (rj == my_rank + shift) paid_me = bj;
        if (rj == my_rank + 1) sec_me = bj;
        if (rj == n_charged - 1 + shift) logged_price = bj;
        if (rj == 0) top_slot = j2;
      }
      if (n_charged < 1) logged_price = Real(0);
      won_me = lane < P && my_rank < n_charged;
      if (won_me) {
        Real u;
        if (kReplay) u = Real(in.u[ri * S + my_rank]);
        else u = sizeof(Real) == 8 ? (Real(click_word_slot(rc, key, my_rank)) + Real(0.5)) * Real(1.0 / 4294967296.0)
                                   : Real(u32_to_unit(click_word_slot(rc, key, my_rank)));
        clk_me = u < r_true;
      } else {
        sec_me = Real(0);
      }
    }
    // the winner of the first slot: what the single-slot winner record is made of
    const bool click = shfl_idx<G>(int(clk_me), top_slot) != 0;
    const int w_agent = shfl_idx<G>(my_agent, top_slot);
    const int w_item = shfl_idx<G>(r_item, top_slot);

    if (active) {
      // ---- charge / set_price and metric sums (Agent.py:70-118) ----
      const long long tl = ta + p.log_base;  // row in the bid log (retained records sit in front)
      if (lane < P) {
        const bool won = won_me;
        const Real tv = r_true * r_val;
        double* __restrict__ ac = p.acc + ((size_t)run * A + my_agent) * kNumMetrics;
        const Real de = r_true - r_est;
        const double t_alloc = double(r_bev - tv), t_estim = double(r_est * r_val - tv), t_sq = double(de * de);
        double t_over = 0.0, t_under = 0.0, t_bias = 0.0;
        if (won) {
          const Real got = clk_me ? r_val : Real(0);
          t_over = double(logged_price - sec_me);
          t_bias = double(r_est / r_true);
          atomicAdd(ac + AGYM_M_NET, double(got - paid_me));
          atomicAdd(ac + AGYM_M_GROSS, double(got));
          atomicAdd(ac + AGYM_M_OVERBID_REGRET, t_over);
          atomicAdd(ac + AGYM_M_BIAS, t_bias);
          atomicAdd(ac + AGYM_M_NWON, 1.0);
          if (S > 1) atomicAdd(p.revenue + run, double(paid_me));  // Auction.py:74, once per charged slot
        } else if (logged_price < tv) {  // nobody charged (P == 1): the logged price stays 0 (Impression.py), the getter still counts it
          t_under = double(logged_price - r_bid);
          atomicAdd(ac + AGYM_M_UNDERBID_REGRET, t_under);
        }
        atomicAdd(ac + AGYM_M_ALLOC_REGRET, t_alloc);
        atomicAdd(ac + AGYM_M_ESTIM_REGRET, t_estim);
        atomicAdd(ac + AGYM_M_SQERR, t_sq);
        atomicAdd(ac + AGYM_M_NPART, 1.0);
        atomicAdd(ac + AGYM_M_BEST_EV, double(r_bev));
        if (r_gamma == r_gamma) atomicAdd(ac + AGYM_M_GAMMA, double(r_gamma));
        // per-record summands, kept so that retained records can be re-summed next iteration (Agent.py:124-129)
        if (p.terms != nullptr && tl < p.bid_Tcap) {
          double2* __restrict__ tr = reinterpret_cast<double2*>(p.terms + (((size_t)run * p.bid_Tcap + tl) * P + lane) * AGYM_TERM_ROW);
          tr[0] = make_double2(t_alloc, t_estim);
          tr[1] = make_double2(t_over, t_under);
          tr[2] = make_double2(t_sq, t_bias);
          tr[3] = make_double2(r_gamma == r_gamma ? double(r_gamma) : 0.0, double(r_bev));
        }
      }
      if (S <= 1 && lane == 0 && valid) atomicAdd(p.revenue + run, double(logged_price));  // Auction.py:74

      // ---- winner records for the allocator fit (Agent.py:81-91: won rows only), S rows per round ----
      if (p.fit_ctx != nullptr) {
        if (S <= 1) {
          if (tl < p.Tcap) {
            const size_t fi = (size_t)run * p.Tcap + tl;
#pragma unroll
            for (int k = 0; k < DMAX; ++k)
              if (k < Do && lane == (k % G)) p.fit_ctx[fi * Do + k] = float(ctx[k]);
            if (lane == 0) p.fit_meta[fi] = valid ? pack_meta(w_agent, w_item, click) : 0u;
          }
        } else {
          const long long row0 = p.log_base + ta * S;
          if (row0 + S <= p.Tcap) {
            const size_t f0 = (size_t)run * p.Tcap + row0;
            if (won_me) {  // the winner of slot my_rank writes that slot's row
#pragma unroll
              for (int k = 0; k < DMAX; ++k)
                if (k < Do) p.fit_ctx[(f0 + my_rank) * Do + k] = float(ctx[k]);
              p.fit_meta[f0 + my_rank] = pack_meta(my_agent, r_item, clk_me);
            }
            if (lane < S && lane >= n_charged) p.fit_meta[f0 + lane] = 0u;  // slots nobody was charged for
          }
        }
      }

      // ---- bid records for the bidder fits (Agent.py:81-94: bidder.update sees every row) ----
      if (p.bid_rows != nullptr && tl < p.bid_Tcap && lane < P) {
        const bool won = won_me;
        const size_t bi = ((size_t)run * p.bid_Tcap + tl) * P + lane;
        float* __restrict__ row = p.bid_rows + bi * AGYM_BID_ROW;
        row[0] = float(r_est); row[1] = float(r_val); row[2] = float(r_gamma); row[3] = float(r_prop); row[4] = float(logged_price);
        p.bid_meta[bi] = kBidValid | (won ? kBidWon : 0u) | ((won && clk_me) ? kBidClick : 0u) | (uint32_t(r_item) << 12) | uint32_t(my_agent);
      }

      // ---- detailed log (Impression.py:4-31); `has_log` spares the production launch the per-field null checks ----
      if (has_log && lane < P) {
        const bool won = won_me;
        const size_t li = (size_t)ri * P + lane;
        if (log.agent) log.agent[li] = my_agent;
        if (log.item) log.item[li] = r_item;
        if (log.est) log.est[li] = double(r_est);
        if (log.value) log.value[li] = double(r_val);
        if (log.bid) log.bid[li] = double(r_bid);
        if (log.true_ctr) log.true_ctr[li] = double(r_true);
        if (log.best_ev) log.best_ev[li] = double(r_bev);
        if (log.price) log.price[li] = double(logged_price);
        if (log.second) log.second[li] = won ? double(sec_me) : 0.0;
        if (log.gamma) log.gamma[li] = double(r_gamma);
        if (log.propensity) log.propensity[li] = double(r_prop);
        if (log.outcome) log.outcome[li] = (won && clk_me) ? 1 : 0;
        if (log.won) log.won[li] = won ? 1 : 0;
      }
      if (has_log && lane == 0) {
        if (log.winner) log.winner[ri] = top_slot;
        if (log.ctx) {
#pragma unroll
          for (int d = 0; d < DMAX; ++d)
            if (d < D) log.ctx[ri * D + d] = double(ctx[d]);
        }
      }
    }
  }
}

__global__ void refresh_sigma_kernel(const float* __restrict__ q, float* __restrict__ sigma, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) sigma[i] = __fdiv_rn(1.0f, __fsqrt_rn(q[i]));  // Models.py:31  1.0/torch.sqrt(q)
}

// Standard shape (K = 5): the tiled copy of {m, 1 / q} the production round loop reads (layout: SimParams::pk), 1 / q =
// sigma^2 (Models.py:31).  One thread per item slot of every (run, agent), padding slots of the last tile included.
__global__ void pack_state_kernel(const float* __restrict__ m, const float* __restrict__ q, unsigned char* __restrict__ pk, int I, int NT,
                                  size_t n_slots) {
  const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_slots) return;
  const size_t ra = g / (size_t)(NT * kTile);       // (run, agent) pair
  const int slot = int(g - ra * (size_t)(NT * kTile));
  const int t = slot / kTile, pos = slot % kTile;
  float mv[5] = {0.f, 0.f, 0.f, 0.f, 0.f}, v[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  if (slot < I) {
    const float* __restrict__ mi = m + (ra * I + slot) * 5;
    const float* __restrict__ qi = q + (ra * I + slot) * 5;
#pragma unroll
    for (int k = 0; k < 5; ++k) { mv[k] = mi[k]; v[k] = __fdiv_rn(1.0f, qi[k]); }
  }
  unsigned char* tile = pk + (ra * NT + t) * (size_t)kPkTile;
  *reinterpret_cast<float4*>(tile + pos * 16) = make_float4(mv[0], mv[1], mv[2], mv[3]);
  *reinterpret_cast<float4*>(tile + 128 + pos * 16) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float2*>(tile + 256 + pos * 8) = make_float2(mv[4], v[4]);
}

int launch_pack_state(agym_handle* h, cudaStream_t s) {
  if (!h->d_pk || !h->m || !h->q) return AGYM_OK;
  const int NT = tiles_of(h->shape.I);
  const size_t n = (size_t)h->shape.R * h->shape.A * NT * kTile;
  pack_state_kernel<<<unsigned((n + 255) / 256), 256, 0, s>>>(h->m, h->q, h->d_pk, h->shape.I, NT, n);
  h->launches += 1;
  h->pk_valid = true;
  return check_cuda(h, cudaGetLastError(), "pack_state_kernel");
}

int launch_refresh_sigma(agym_handle* h, cudaStream_t s) {
  if (!h->q || !h->sigma) return set_error(h, AGYM_ERR_STATE, "agym_refresh_sigma: allocator state not bound");
  const size_t n = (size_t)h->shape.R * h->shape.A * h->shape.I * h->K;
  refresh_sigma_kernel<<<unsigned((n + 255) / 256), 256, 0, s>>>(h->q, h->sigma, n);
  h->launches += 1;
  const int rc = check_cuda(h, cudaGetLastError(), "refresh_sigma");
  return rc ? rc : launch_pack_state(h, s);
}

template <typename Real, int G, int DMAX>
static int launch_g(agym_handle* h, const SimParams& p, const agym_replay_inputs* in, const agym_round_log* log, cudaStream_t s) {
  const bool std_shape = DMAX == 8 && p.D == 5 && p.Do == 4;  // every shipped config and the bench shape
  const bool multi = p.max_slots > 1;
  // production, float, standard shape, packed state current, catalog small enough: stage the catalog tiles in shared memory
  const size_t cat_bytes = (size_t)p.A * tiles_of(p.I) * kCatTile;
  // -- OPT-IN (option "sim_cat_smem" 1), measured on B200 at the bench shape and not adopted: 8-lane groups 4.55 ms against 4.47 ms
  // from global memory, 4-lane groups 4.52 against 4.06 (the catalog's lines are L1 hits anyway; one CTA of 768 threads per SM
  // fills the tail of the grid worse than three of 256)
  const bool cat_sm = std::is_same<Real, float>::value && G <= 8 && std_shape && !in && !multi && p.pk && p.cat8 && cat_bytes <= 160 * 1024 &&
                      h->has_option("sim_cat_smem") && h->option("sim_cat_smem", 0) != 0;
  const int threads = cat_sm ? kCatSmThreads : 256;
  const int ngroups = threads / G;
  SimParams q = p;
  // rounds per CTA: 16 per lane group, fewer when the launch is small so the grid still fills the SMs
  long long chunk = (long long)ngroups * 16;
  while (chunk > ngroups && ((p.T + chunk - 1) / chunk) * p.n_runs < (cat_sm ? 2LL : 4LL) * h->num_sms) chunk /= 2;
  q.chunk = int(chunk);
  const long long chunks = (p.T + chunk - 1) / chunk;
  const long long grid = chunks * p.n_runs;
  if (grid <= 0) return AGYM_OK;
  if (grid > 0x7fffffffLL) return set_error(h, AGYM_ERR_INVALID, "launch too large: split T");
  agym_round_log lg = {};
  if (log) lg = *log;
  agym_replay_inputs ri = {};
  h->launches += 1;
  if (cat_sm) {
    if constexpr (std::is_same<Real, float>::value && G <= 8 && DMAX == 8) {
      auto kern = sim_kernel<float, G, 8, false, 5, 4, false, true>;
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(cat_bytes));
      if (e != cudaSuccess) return check_cuda(h, e, "sim_kernel attribute");
      kern<<<unsigned(grid), threads, cat_bytes, s>>>(q, ri, lg);
      return check_cuda(h, cudaGetLastError(), "sim_kernel launch");
    }
  }
#define AGYM_SIM(REPLAY, MULTI)                                                                                                   \
  do {                                                                                                                            \
    if (std_shape) sim_kernel<Real, G, DMAX, REPLAY, (DMAX == 8 ? 5 : 0), (DMAX == 8 ? 4 : 0), MULTI><<<unsigned(grid), threads, 0, s>>>(q, ri, lg); \
    else sim_kernel<Real, G, DMAX, REPLAY, 0, 0, MULTI><<<unsigned(grid), threads, 0, s>>>(q, ri, lg);                             \
  } while (0)
  if (in) {
    ri = *in;
    if (multi) AGYM_SIM(true, true); else AGYM_SIM(true, false);
  } else {
    if (multi) AGYM_SIM(false, true); else AGYM_SIM(false, false);
  }
#undef AGYM_SIM
  return check_cuda(h, cudaGetLastError(), "sim_kernel launch");
}

#ifndef AGYM_SIM_G
#define AGYM_SIM_G 8
#endif
int sim_group_width(const agym_handle* h, int P, int DMAX) {
  // 4 lanes per opportunity where the packed production path exists (float, standard shape, learnt allocators): everything
  // outside the item loop is paid per warp instruction, i.e. per 8 opportunities instead of 4 (B200, bench shape: 4.47 -> 4.06 ms);
  // the generic loop with its scalar loads prefers 8 (Oracle allocators: 3.56 ms against 4.34 with 4).  A property of the handle's
  // configuration, never of a launch's size.
  int G = (h->shape.precision == AGYM_FP32 && h->d_pk != nullptr) ? 4 : AGYM_SIM_G;
  if (h->has_option("sim_g")) { const int v = int(h->option("sim_g", AGYM_SIM_G)); if (v == 4 || v == 8 || v == 16 || v == 32) G = v; }
  while (G < P) G *= 2;
  if (DMAX / 4 > G) G = 32;
  return G;
}

template <typename Real, int DMAX>
static int launch_d(agym_handle* h, const SimParams& p, const agym_replay_inputs* in, const agym_round_log* log, cudaStream_t s) {
  // Lane-group width.  Everything outside the item loop (context, participants, resolution, click, accumulators: more than
  // half of the instructions at 64 items) is issued once per warp instruction whatever G is, so narrow groups -- more
  // opportunities per warp, more items per lane -- amortise it: B200, bench shape, G = 32 / 16 / 8 -> 10.6 / 7.7 / 7.2 ms
  // (Oracle allocators 6.1 / 4.7 / 3.6 ms).  The width is the same for every launch size (the Thompson noise of an item
  // is addressed through its lane's position, so a launch-size-dependent width would break "chunked calls == one call").
  const int G = sim_group_width(h, p.P, DMAX);
  switch (G) {
    case 4: return launch_g<Real, 4, DMAX>(h, p, in, log, s);
    case 8: return launch_g<Real, 8, DMAX>(h, p, in, log, s);
    case 16: return launch_g<Real, 16, DMAX>(h, p, in, log, s);
    default: return launch_g<Real, 32, DMAX>(h, p, in, log, s);
  }
}

int launch_simulate(agym_handle* h, const SimParams& p, const agym_replay_inputs* in, const agym_round_log* log, cudaStream_t s) {
  if (p.P > kMaxP) return set_error(h, AGYM_ERR_UNSUPPORTED, "num_participants_per_round > 32 is not supported yet");
  if (p.D > 32) return set_error(h, AGYM_ERR_UNSUPPORTED, "embedding_size > 32 is not supported yet");
  const bool f64 = h->shape.precision == AGYM_FP64;
  if (p.D <= 8) return f64 ? launch_d<double, 8>(h, p, in, log, s) : launch_d<float, 8>(h, p, in, log, s);
  return f64 ? launch_d<double, 32>(h, p, in, log, s) : launch_d<float, 32>(h, p, in, log, s);
}

}  // namespace agym
