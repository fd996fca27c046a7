// Device building blocks of one auction opportunity, shared by the fused kernel (agym_sim.cu) and the
// staged kernels (agym_staged.cu) so both produce the same numbers from the same Philox counters.
#pragma once

#include <limits.h>
#include <math_constants.h>

#include "agym_common.cuh"

namespace agym {

template <typename Real>
__device__ __forceinline__ Real pick4(const float4& v, int j) {
  return Real(j == 0 ? v.x : j == 1 ? v.y : j == 2 ? v.z : v.w);
}

struct RoundCounter {
  uint32_t c0, c1;
  long long round;
  int iter;
  __device__ __forceinline__ RoundCounter(long long round_in_iter, int iter_)
      : c0(uint32_t(round_in_iter)), c1(uint32_t(iter_) ^ (uint32_t(round_in_iter >> 32) << 20)), round(round_in_iter), iter(iter_) {}
};

// Auction.py:33 -- D normals scaled by embedding_var (used as the std), spread over the group:
// lane j draws Philox block j = normals 4j..4j+3.
template <typename Real, int G, int DMAX>
__device__ __forceinline__ void draw_context(Real (&ctx)[DMAX], int D, Real std, RoundCounter rc, PhiloxKey key, int lane) {
  const float4 nrm = philox_normal4(rc.c0, rc.c1, kPurposeCtx << 16, uint32_t(lane), key);
#pragma unroll
  for (int d = 0; d < DMAX; ++d) {
    const float v = shfl_idx<G>(pick4<float>(nrm, d & 3), (d >> 2) % G);
    ctx[d] = d < D ? Real(v) * std : Real(0);
  }
}

// Per-thread variant of the same draw (staged K1): context component d.
__device__ __forceinline__ float context_component(int d, RoundCounter rc, PhiloxKey key) {
  const float4 nrm = philox_normal4(rc.c0, rc.c1, kPurposeCtx << 16, uint32_t(d >> 2), key);
  return pick4<float>(nrm, d & 3);
}

// Auction.py:42 -- ordered uniform sample of P agents out of A without replacement.  Slot j takes the
// r_j-th smallest agent not chosen yet, r_j uniform on [0, A-j) from Philox block j.  Lane s ends up
// holding slot s's agent.  (numpy's tail-shuffle produces the same distribution, not the same stream.)
__device__ __forceinline__ int participant_draw(int j, int A, RoundCounter rc, PhiloxKey key) {
  const uint4 w = philox4x32_10(rc.c0, rc.c1, kPurposePart << 16, uint32_t(j), key);
  return int(__umulhi(w.x, uint32_t(A - j > 0 ? A - j : 1)));
}

template <int G>
__device__ __forceinline__ int draw_participants(int P, int A, RoundCounter rc, PhiloxKey key, int lane) {
  const int rj = participant_draw(lane, A, rc, key);
  int mine = 0x7fffffff;
  const unsigned gmask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << ((threadIdx.x & 31) / G * G);
  for (int j = 0; j < P; ++j) {
    const int x = shfl_idx<G>(rj, j);
    int xp = x;
    bool changed;
    do {  // x-th unchosen agent: fixed point of xp = x + #{chosen <= xp}
      const unsigned b = __ballot_sync(0xffffffffu, lane < j && mine <= xp) & gmask;
      const int nx = x + __popc(b);
      changed = nx != xp;
      xp = nx;
    } while (__any_sync(0xffffffffu, changed));
    if (lane == j) mine = xp;
  }
  return lane < P ? mine : 0;
}

// Per-thread variant (staged K1): writes the P agents to out[0..P).
__device__ __forceinline__ void draw_participants_thread(int P, int A, RoundCounter rc, PhiloxKey key, int* out) {
  for (int j = 0; j < P; ++j) {
    const int x = participant_draw(j, A, rc, key);
    int xp = x;
    while (true) {
      int cnt = 0;
      for (int i = 0; i < j; ++i) cnt += out[i] <= xp;
      if (x + cnt == xp) break;
      xp = x + cnt;
    }
    out[j] = xp;
  }
}

template <typename Real, bool kReplay>
constexpr bool kPackedOk = false;
template <>
constexpr bool kPackedOk<float, false> = true;

template <typename Real>
struct SlotEval {
  int item;
  Real est, true_sel, value, best_ev;
};

// ---- production path of the standard shape (embedding_size 5, obs_embedding_size 4, float): packed 128-bit loads and
// single-instruction MUFU approximations.  Never used in replay mode or FP64, i.e. nowhere parity is claimed. ----
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sqrt_approx(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigmoid_approx(float z) { return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * z)); }
__device__ __forceinline__ float2 box_muller_approx(uint32_t a, uint32_t b) {
  const float r = sqrt_approx(-1.3862943611198906f * lg2_approx(u32_to_unit_open0(a)));  // sqrt(-2 ln u), u in (0, 1]
  float sn, cs;
  __sincosf(6.283185307179586f * u32_to_unit(b), &sn, &cs);
  return make_float2(r * cs, r * sn);
}
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
// true CTR of item i for context x (Auction.py:52); the item loop and the chosen item's record share this exact sequence,
// so the allocation regret best - chosen can never come out negative
__device__ __forceinline__ float true_ctr_packed(const float4 ea, const float4 eb, const float (&x)[5]) {
  float z = fmaf(x[0], ea.x, eb.y);
  z = fmaf(x[1], ea.y, z);
  z = fmaf(x[2], ea.z, z);
  z = fmaf(x[3], ea.w, z);
  z = fmaf(x[4], eb.x, z);
  return sigmoid_approx(z);
}

#ifndef AGYM_SIM_BRANCHFREE
#define AGYM_SIM_BRANCHFREE 0  // measured on B200: 1 (no branch around an item) spills under the 80-register cap: 4.68 vs 4.39 ms; 128 registers / 2 CTAs per SM: 5.06 ms
#endif
// Item scores of one participant over the lanes of its group: four items per lane and pass (one Philox block = two
// Box-Muller pairs = their four Thompson normals, as in the generic loop: the noise of an item does not depend on the path).
// kCatSm: the catalog tiles were staged in shared memory by the CTA (cat_sm = their shared-window address): they are static and
// shared by every run, 128 KB at 64 x 64 items, and two of the five loads of an item.
template <int G, bool kCatSm>
__device__ __forceinline__ void eval_items_packed(const SimParams& p, int run, int a, int s, const float (&x)[5], bool ts, RoundCounter rc,
                                                  PhiloxKey key, int lane, uint32_t cat_sm, float& bscore, int& bi, float& btv) {
  const int nI = p.n_items[a], NT = tiles_of(p.I);
  const unsigned char* __restrict__ cat = p.cat8 + (size_t)a * NT * kCatTile;
  const unsigned char* __restrict__ pk = p.pk + ((size_t)run * p.A + a) * NT * kPkTile;
  const float xx0 = x[0] * x[0], xx1 = x[1] * x[1], xx2 = x[2] * x[2], xx3 = x[3] * x[3];
  for (int i0 = lane; i0 < nI; i0 += 4 * G) {
    float nz[4] = {0.f, 0.f, 0.f, 0.f};
    if (ts) {
      const uint4 b = philox4x32_10(rc.c0, rc.c1, (kPurposeTS << 16) | uint32_t(s), uint32_t(i0), key);
      const float2 n01 = box_muller_approx(b.x, b.y), n23 = box_muller_approx(b.z, b.w);
      nz[0] = n01.x; nz[1] = n01.y; nz[2] = n23.x; nz[3] = n23.y;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i_raw = i0 + j * G;
#if AGYM_SIM_BRANCHFREE
      // no branch around an item: the loads of all four items can be in flight together; an item past the end re-reads item i0
      const bool ok = i_raw < nI;
      const int i = ok ? i_raw : i0;
      {
#else
      const bool ok = true;
      const int i = i_raw;
      if (i < nI) {
#endif
        // tile i / 8, position i % 8: the group's lanes read 128 (64) contiguous bytes per field
        const unsigned char* ct = cat + (i >> 3) * kCatTile + (i & 7) * 16;
        const unsigned char* pt = pk + (i >> 3) * kPkTile;
        float4 ea, eb;
        if (kCatSm) {
          const uint32_t cs = cat_sm + uint32_t((a * NT + (i >> 3)) * kCatTile + (i & 7) * 16);
          ea = lds128(cs);
          eb = lds128(cs + 128);
        } else {
          ea = *reinterpret_cast<const float4*>(ct);
          eb = *reinterpret_cast<const float4*>(ct + 128);
        }
        const float4 ma = *reinterpret_cast<const float4*>(pt + (i & 7) * 16), va = *reinterpret_cast<const float4*>(pt + 128 + (i & 7) * 16);
        const float2 mb = *reinterpret_cast<const float2*>(pt + 256 + (i & 7) * 8);
        const float tv = ok ? true_ctr_packed(ea, eb, x) * eb.z : -INFINITY;
        btv = fmaxf(btv, tv);
        float zl = fmaf(x[0], ma.x, mb.x);
        zl = fmaf(x[1], ma.y, zl);
        zl = fmaf(x[2], ma.z, zl);
        zl = fmaf(x[3], ma.w, zl);
        // logit-space Thompson draw: sum_k (m_k + eps_k sigma_k) x_k ~ N(m.x, sum_k sigma_k^2 x_k^2) (Models.py:31)
        float var = fmaf(xx0, va.x, mb.y);
        var = fmaf(xx1, va.y, var);
        var = fmaf(xx2, va.z, var);
        var = fmaf(xx3, va.w, var);
        zl = fmaf(sqrt_approx(var), nz[j], zl);
        const float score = sigmoid_approx(zl) * eb.z;  // Agent.py:33
        if (ok && score > bscore) { bscore = score; bi = i; }
      }
    }
  }
}

// Agent.select_item + the true-CTR bookkeeping of Auction.py:52-53 for the agent in slot s.
// Lanes of the group stride over the agent's items; the result is uniform across the group.
// DT / DoT > 0 fix embedding_size / obs_embedding_size at compile time (the shipped configs' 5 / 4), 0 = run time.
template <typename Real, int G, int DMAX, bool kReplay, int DT = 0, int DoT = 0, bool kCatSm = false>
__device__ __forceinline__ SlotEval<Real> eval_slot(const SimParams& p, int run, int a, int s, const Real (&ctx)[DMAX],
                                                    RoundCounter rc, PhiloxKey key, const float* __restrict__ ts_eps_slot,
                                                    int lane, uint32_t cat_sm = 0) {
  using A_ = Arith<Real>;
  const int D = DT > 0 ? DT : p.D, Do = DT > 0 ? DoT : p.Do, K = Do + 1, I = p.I;
  const int nI = p.n_items[a];
  const int akind = p.alloc_kind[a];
  const Real* __restrict__ Ea = Catalog<Real>::E(p) + (size_t)a * I * (D + 1);
  const Real* __restrict__ Va = Catalog<Real>::V(p) + (size_t)a * I;
  const size_t soff = ((size_t)run * p.A + a) * I * K;
  const float* __restrict__ ma = p.m + soff;
  const float* __restrict__ sa = p.sigma + soff;

  Real bscore = A_::neg_inf(), btv = A_::neg_inf();
  int bi = INT_MAX;
  uint4 ts_bits = make_uint4(0u, 0u, 0u, 0u);
  float2 ts_pair = make_float2(0.f, 0.f);
  // production, float, standard shape, learnt allocator, packed copies current: the 128-bit path
  const bool packed = kPackedOk<Real, kReplay> && D == 5 && Do == 4 && akind != AGYM_ALLOC_ORACLE && p.pk != nullptr && p.cat8 != nullptr;
  float px[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  if (packed) {
#pragma unroll
    for (int d = 0; d < 5; ++d) px[d] = float(ctx[d < DMAX ? d : 0]);
    float fs = -INFINITY, ft = -INFINITY;
    eval_items_packed<G, kCatSm>(p, run, a, s, px, akind == AGYM_ALLOC_TS, rc, key, lane, cat_sm, fs, bi, ft);
    bscore = Real(fs);
    btv = Real(ft);
  }
  for (int i = lane; i < (packed ? 0 : nI); i += G) {
    const Real* __restrict__ e = Ea + (size_t)i * (D + 1);
    Real z = 0;
#pragma unroll
    for (int d = 0; d < DMAX; ++d)
      if (d < D) z = fma(ctx[d], e[d], z);
    z += e[D];
    const Real tc = A_::sigmoid(z);  // Auction.py:52
    const Real v = Va[i];
    const Real tv = tc * v;
    btv = tv > btv ? tv : btv;
    Real score = tv;  // OracleAllocator: the estimate is the true CTR (BidderAllocation.py:81-82)
    if (akind != AGYM_ALLOC_ORACLE) {
      const float* __restrict__ mi = ma + (size_t)i * K;
      float zl = 0.0f;
      if (akind == AGYM_ALLOC_TS && !kReplay) {
        // Production mode samples in logit space: with independent eps_k ~ N(0, 1) on the weights (Models.py:31),
        // sum_k (m_k + eps_k sigma_k) x_k  ~  N(m.x, sum_k (sigma_k x_k)^2) exactly, so ONE normal per item gives the
        // reference's distribution of sampled CTRs with a fifth of the normals.  (Replay mode below keeps the
        // weight-space form: there the host supplies eps_k and parity is value for value.)
        const float* __restrict__ si = sa + (size_t)i * K;
        float var = 0.0f;
#pragma unroll
        for (int k = 0; k < DMAX + 1; ++k) {
          if (k < K) {
            const float x = k < Do ? float(ctx[k < DMAX ? k : 0]) : 1.0f;
            zl = fmaf(mi[k], x, zl);
            const float sx = si[k] * x;
            var = fmaf(sx, sx, var);
          }
        }
        // one Philox block serves four consecutive items of this lane, one Box-Muller pair two of them
        const int t = (i - lane) / G;
        if ((t & 3) == 0) ts_bits = philox4x32_10(rc.c0, rc.c1, (kPurposeTS << 16) | uint32_t(s), uint32_t(i), key);
        if ((t & 1) == 0) ts_pair = (t & 2) ? box_muller(ts_bits.z, ts_bits.w) : box_muller(ts_bits.x, ts_bits.y);
        zl = fmaf(sqrtf(var), (t & 1) ? ts_pair.y : ts_pair.x, zl);
      } else if (akind == AGYM_ALLOC_TS) {
        const float* __restrict__ si = sa + (size_t)i * K;
#pragma unroll
        for (int kb = 0; kb < (DMAX + 4) / 4; ++kb) {
          if (kb * 4 < K) {
            float4 nrm = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!kReplay) nrm = philox_normal4(rc.c0, rc.c1, (kPurposeTS << 16) | uint32_t(s), uint32_t(i * ((DMAX + 4) / 4) + kb), key);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int k = kb * 4 + j;
              if (k < K) {
                const float eps = kReplay ? ts_eps_slot[(size_t)i * K + k] : pick4<float>(nrm, j);
                const float w = A_::ts_weight(mi[k], eps, si[k]);  // Models.py:31
                const float x = k < Do ? float(ctx[k < DMAX ? k : 0]) : 1.0f;
                zl = A_::mac(w, x, zl);
              }
            }
          }
        }
      } else {
#pragma unroll
        for (int k = 0; k < DMAX + 1; ++k)
          if (k < K) zl = A_::mac(mi[k], k < Do ? float(ctx[k < DMAX ? k : 0]) : 1.0f, zl);
      }
      score = Real(A_::sigmoid32(zl)) * v;  // Agent.py:33
    }
    if (score > bscore) { bscore = score; bi = i; }
  }
  // arg-max over the group, lowest item index on ties (np.argmax, Agent.py:35); max of true value
#pragma unroll
  for (int off = G / 2; off > 0; off >>= 1) {
    const Real os = shfl_xor<G>(bscore, off);
    const int oi = shfl_xor<G>(bi, off);
    const Real ot = shfl_xor<G>(btv, off);
    if (os > bscore || (os == bscore && oi < bi)) { bscore = os; bi = oi; }
    btv = ot > btv ? ot : btv;
  }
  if (bi == INT_MAX) bi = 0;
  if (packed) {  // chosen item through the same loads and the same true-CTR sequence as the loop
    const int NT = tiles_of(I);
    const unsigned char* ct = p.cat8 + ((size_t)a * NT + (bi >> 3)) * kCatTile + (bi & 7) * 16;
    const unsigned char* pt = p.pk + (((size_t)run * p.A + a) * NT + (bi >> 3)) * kPkTile;
    float4 ea, eb;
    if (kCatSm) {
      const uint32_t cs = cat_sm + uint32_t((a * NT + (bi >> 3)) * kCatTile + (bi & 7) * 16);
      ea = lds128(cs);
      eb = lds128(cs + 128);
    } else {
      ea = *reinterpret_cast<const float4*>(ct);
      eb = *reinterpret_cast<const float4*>(ct + 128);
    }
    const float4 ma = *reinterpret_cast<const float4*>(pt + (bi & 7) * 16);
    const float2 mb = *reinterpret_cast<const float2*>(pt + 256 + (bi & 7) * 8);
    float zl = fmaf(px[0], ma.x, mb.x);
    zl = fmaf(px[1], ma.y, zl);
    zl = fmaf(px[2], ma.z, zl);
    zl = fmaf(px[3], ma.w, zl);
    SlotEval<Real> r;
    r.item = bi;
    r.true_sel = Real(true_ctr_packed(ea, eb, px));
    r.est = Real(sigmoid_approx(zl));  // MAP estimate of the chosen item (Agent.py:38-42)
    r.value = Real(eb.z);
    r.best_ev = btv;
    return r;
  }
  // chosen item: true CTR and the estimate that is logged and bid on (Agent.py:38-42)
  const Real* __restrict__ e = Ea + (size_t)bi * (D + 1);
  Real z = 0;
#pragma unroll
  for (int d = 0; d < DMAX; ++d)
    if (d < D) z = fma(ctx[d], e[d], z);
  z += e[D];
  SlotEval<Real> r;
  r.item = bi;
  r.true_sel = A_::sigmoid(z);
  r.est = r.true_sel;
  if (akind != AGYM_ALLOC_ORACLE) {
    const float* __restrict__ mi = ma + (size_t)bi * K;
    float zl = 0.0f;
#pragma unroll
    for (int k = 0; k < DMAX + 1; ++k)
      if (k < K) zl = A_::mac(mi[k], k < Do ? float(ctx[k < DMAX ? k : 0]) : 1.0f, zl);
    r.est = Real(A_::sigmoid32(zl));
  }
  r.value = Va[bi];
  r.best_ev = btv;
  return r;
}

// Bidder.bid (Bidder.py:34-35,47-58,171-179,348-356,455-463).  gamma / prop stay NaN for truthful bidders.
// `eff` returns the bid-time behaviour; AGYM_BID_SEARCH is left to search_gamma (it needs the whole lane group).
template <typename Real>
__device__ __forceinline__ Real shade_bid(const SimParams& p, int run, int a, int s, Real value, Real est, bool replay,
                                          double gamma_z_replay, RoundCounter rc, PhiloxKey key, Real& gamma, Real& prop,
                                          int& eff) {
  Real bid = value * est;
  gamma = Real(CUDART_NAN);
  prop = Real(CUDART_NAN);
  const int bkind = p.bidder_kind[a];
  eff = bkind;
  if (bkind == AGYM_BID_TRUTHFUL) return bid;
  const double* __restrict__ bd = p.bidder_d + ((size_t)run * p.A + a) * AGYM_BIDDER_D;
  const Real prev = Real(bd[0]), sg = Real(bd[1]);
  eff = (bkind >= AGYM_BID_SEARCH && bd[2] == 0.0) ? AGYM_BID_GAUSS : bkind;
  if (eff == AGYM_BID_GAUSS || eff == AGYM_BID_GAUSS_CLIP) {
    Real zg;
    if (replay) zg = Real(gamma_z_replay);
    else zg = Real(philox_normal4(rc.c0, rc.c1, (kPurposeGamma << 16) | uint32_t(s), 0u, key).x);
    gamma = prev + sg * zg;  // Bidder.py:51,177,354,461
    if (eff == AGYM_BID_GAUSS_CLIP) {
      gamma = gamma < Real(0) ? Real(0) : (gamma > Real(1) ? Real(1) : gamma);  // Bidder.py:52-55
    } else {
      const Real q_ = (prev - gamma) / sg;
      prop = exp(-(q_ * q_) / Real(2)) / (sg * Real(2.5066282746310002));  // Bidder.py:178
    }
    bid = bid * gamma;
  } else if (eff == AGYM_BID_BANDIT || eff == AGYM_BID_POLICY) {
    // BidShadingContextualBandit.forward / BidShadingPolicy.forward (Models.py:82-90,146-155): float32 throughout
    using A_ = Arith<Real>;
    const float* __restrict__ th = p.bidder_w + ((size_t)run * p.A + a) * AGYM_BIDDER_W + 4;
    const float x0 = float(est), x1 = float(value);
    const float h0 = A_::mac(x1, th[1], A_::mac(x0, th[0], 0.0f)) + th[4];
    const float h1 = A_::mac(x1, th[3], A_::mac(x0, th[2], 0.0f)) + th[5];
    const float s0 = h0 > 20.f ? h0 : log1pf(expf(h0)), s1 = h1 > 20.f ? h1 : log1pf(expf(h1));
    const float amu = A_::mac(s1, th[7], A_::mac(s0, th[6], 0.0f)) + th[8];
    const float asg = A_::mac(s1, th[10], A_::mac(s0, th[9], 0.0f)) + th[11];
    const float mu = amu > 20.f ? amu : log1pf(expf(amu));
    const float sg = (asg > 20.f ? asg : log1pf(expf(asg))) + 1e-2f;
    float eps;
    if (replay) eps = float(gamma_z_replay);
    else eps = philox_normal4(rc.c0, rc.c1, (kPurposeGamma << 16) | uint32_t(s), 0u, key).x;
    const float raw = __fadd_rn(mu, __fmul_rn(eps, sg));                 // rsample: loc + eps * scale
    const float dv = raw - mu;
    const float logp = -(dv * dv) / (2.0f * (sg * sg)) - logf(sg) - 0.9189385332046727f;  // Normal.log_prob
    prop = Real(expf(logp));
    const float gcl = fminf(fmaxf(raw, 0.0f), 1.0f);                     // torch.clip(sampled_value, 0, 1)
    gamma = Real(gcl);
    bid = bid * gamma;
  }
  return bid;
}

// ValueLearningBidder 'search' (Bidder.py:180-196): 128 gammas ~ U(0.1, 1); P(win | CTR, value, gamma) from the win-rate
// model (Models.py:61-62, float32 on float32 features); the gamma with the largest estimated utility W * (V - gamma V) is
// bid.  The reference sorts the grid and takes the first arg-max, i.e. the smallest gamma among equal utilities.
constexpr int kSearchGrid = 128;

template <typename Real>
__device__ __forceinline__ void search_point(Real u, Real bid0, float est32, float val32, float w0, float w1, float w2, float b,
                                             Real& best_u, Real& best_g) {
  using A_ = Arith<Real>;
  const Real g = Real(0.1) + (Real(1.0) - Real(0.1)) * u;  // rng.uniform(0.1, 1.0)
  float z = A_::mac(est32, w0, 0.0f);
  z = A_::mac(val32, w1, z);
  z = A_::mac(float(g), w2, z);
  z = A_::mac(b, 1.0f, z);
  const Real util = Real(A_::sigmoid32(z)) * (bid0 - bid0 * g);  // Bidder.py:192-194
  if (util > best_u || (util == best_u && g < best_g)) { best_u = util; best_g = g; }
}

template <typename Real, int G, bool kReplay>
__device__ __forceinline__ Real search_gamma(const SimParams& p, int run, int a, int s, Real bid0, Real est, Real value,
                                             const double* __restrict__ grid_u_slot, int grid_n, RoundCounter rc, PhiloxKey key,
                                             int lane) {
  const float* __restrict__ w = p.bidder_w + ((size_t)run * p.A + a) * AGYM_BIDDER_W;
  const float w0 = w[0], w1 = w[1], w2 = w[2], b = w[3];
  const float est32 = float(est), val32 = float(value);
  Real bu = Arith<Real>::neg_inf(), bg = Real(2);
  if (kReplay) {
    for (int j = lane; j < grid_n; j += G) search_point<Real>(Real(grid_u_slot[j]), bid0, est32, val32, w0, w1, w2, b, bu, bg);
  } else {
    for (int blk = lane; blk < kSearchGrid / 4; blk += G) {
      const uint4 r = philox4x32_10(rc.c0, rc.c1, (kPurposeGrid << 16) | uint32_t(s), uint32_t(blk), key);
      const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const Real u = sizeof(Real) == 8 ? Real((double(rr[j]) + 0.5) * (1.0 / 4294967296.0)) : Real(u32_to_unit(rr[j]));
        search_point<Real>(u, bid0, est32, val32, w0, w1, w2, b, bu, bg);
      }
    }
  }
#pragma unroll
  for (int off = G / 2; off > 0; off >>= 1) {
    const Real ou = shfl_xor<G>(bu, off), og = shfl_xor<G>(bg, off);
    if (ou > bu || (ou == bu && og < bg)) { bu = ou; bg = og; }
  }
  return bg;
}

// Auction.py:65 -- the click uniform of round t is word (t & 3) of Philox block (t >> 2), so the staged
// resolution kernel can serve four consecutive rounds from one block.
__device__ __forceinline__ uint4 click_block(long long round, int iter, PhiloxKey key) {
  return philox4x32_10(uint32_t(round >> 2), uint32_t(iter) ^ (uint32_t(round >> 34) << 20), kPurposeClick << 16, 0u, key);
}
__device__ __forceinline__ uint32_t click_word(RoundCounter rc, PhiloxKey key) {
  const uint4 w = click_block(rc.round, rc.iter, key);
  const int j = int(rc.round & 3);
  return j == 0 ? w.x : j == 1 ? w.y : j == 2 ? w.z : w.w;
}
// several slots per round (Auction.py:30,65): slot k > 0 takes its click uniform from its own Philox block, and the
// number of slots of the round is uniform in [1, max_slots]
__device__ __forceinline__ uint32_t click_word_slot(RoundCounter rc, PhiloxKey key, int slot) {
  if (slot == 0) return click_word(rc, key);
  return philox4x32_10(rc.c0, rc.c1, kPurposeClick << 16, uint32_t(slot), key).x;
}
__device__ __forceinline__ int draw_num_slots(int max_slots, RoundCounter rc, PhiloxKey key) {
  return 1 + int(philox4x32_10(rc.c0, rc.c1, kPurposeSlots << 16, 0u, key).x % uint32_t(max_slots));
}
__device__ __forceinline__ float click_uniform_f(RoundCounter rc, PhiloxKey key) { return u32_to_unit(click_word(rc, key)); }
__device__ __forceinline__ double click_uniform_d(RoundCounter rc, PhiloxKey key) {
  return (double(click_word(rc, key)) + 0.5) * (1.0 / 4294967296.0);
}

}  // namespace agym
