// Staged kernels K1-K4: the same arithmetic as the fused kernel (agym_sim.cu) with the intermediates
// materialised in HBM as float / uint8 SoA arrays [N = R*T][P].  They exist so each stage can be
// tested and profiled in isolation against its HBM roofline:
//   K1 contexts + participants   writes 4D + P            bytes / opportunity
//   K2 allocation                reads 4D + P, writes 17P  (item 1 + est 4 + true 4 + best_ev 4 + value 4)
//   K3 bids                      reads 8P + P, writes 4P (+8P shaded)
//   K4 resolution + click        reads 13P, writes 10      = 36 B at P = 2
// All four draw from the same Philox counters as the fused kernel, so K1->K2->K3->K4 reproduces the
// fused FP32 results exactly.
#include <cstdlib>

#include "agym_round.cuh"

namespace agym {

// ------------------------------------------------------------------------------------------------
// K1: one thread per opportunity (Auction.py:33,42)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k1_kernel(const SimParams p, float* __restrict__ ctx, uint8_t* __restrict__ parts, long long N) {
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int run = int(n / p.T);
  const long long t = n - (long long)run * p.T;
  const PhiloxKey key = make_key(p.seed, uint32_t(p.run_offset + run));
  const RoundCounter rc(p.round0 + t, p.iter);
  const float sd = float(p.embedding_var);
  for (int d0 = 0; d0 < p.D; d0 += 4) {
    const float4 nrm = philox_normal4(rc.c0, rc.c1, kPurposeCtx << 16, uint32_t(d0 >> 2), key);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (d0 + j < p.D) ctx[n * p.D + d0 + j] = pick4<float>(nrm, j) * sd;
  }
  int chosen[kMaxP];
  draw_participants_thread(p.P, p.A, rc, key, chosen);
  for (int s = 0; s < p.P; ++s) parts[n * p.P + s] = uint8_t(chosen[s]);
}

int launch_k1(agym_handle* h, const SimParams& p, float* ctx, uint8_t* parts, cudaStream_t s) {
  if (p.P > kMaxP) return set_error(h, AGYM_ERR_UNSUPPORTED, "K1: P > 32");
  const long long N = (long long)p.R * p.T;
  k1_kernel<<<unsigned((N + 255) / 256), 256, 0, s>>>(p, ctx, parts, N);
  h->launches += 1;
  return check_cuda(h, cudaGetLastError(), "k1_kernel");
}

// ------------------------------------------------------------------------------------------------
// K2: one lane group per opportunity (Agent.py:29-42, Auction.py:52-53)
// ------------------------------------------------------------------------------------------------
template <int G, int DMAX>
__global__ void __launch_bounds__(256) k2_kernel(const SimParams p, const float* __restrict__ ctx_in, const uint8_t* __restrict__ parts,
                                                 uint8_t* __restrict__ item, float* __restrict__ est, float* __restrict__ true_ctr,
                                                 float* __restrict__ best_ev, float* __restrict__ value, long long N) {
  const int lane = threadIdx.x % G;
  const long long n_raw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
  const bool active = n_raw < N;
  const long long n = active ? n_raw : N - 1;
  const int run = int(n / p.T);
  const long long t = n - (long long)run * p.T;
  const PhiloxKey key = make_key(p.seed, uint32_t(p.run_offset + run));
  const RoundCounter rc(p.round0 + t, p.iter);
  float ctx[DMAX];
#pragma unroll
  for (int d = 0; d < DMAX; ++d) ctx[d] = d < p.D ? ctx_in[n * p.D + d] : 0.0f;
  for (int s = 0; s < p.P; ++s) {
    const int a = parts[n * p.P + s];
    const SlotEval<float> ev = eval_slot<float, G, DMAX, false>(p, run, a, s, ctx, rc, key, nullptr, lane);
    if (active && lane == 0) {
      const long long o = n * p.P + s;
      item[o] = uint8_t(ev.item);
      est[o] = ev.est;
      true_ctr[o] = ev.true_sel;
      best_ev[o] = ev.best_ev;
      value[o] = ev.value;
      if (p.acc) {
        double* __restrict__ ac = p.acc + ((size_t)run * p.A + a) * kNumMetrics;
        const float tv = ev.true_sel * ev.value;
        const float de = ev.true_sel - ev.est;
        atomicAdd(ac + AGYM_M_ALLOC_REGRET, double(ev.best_ev - tv));
        atomicAdd(ac + AGYM_M_ESTIM_REGRET, double(ev.est * ev.value - tv));
        atomicAdd(ac + AGYM_M_SQERR, double(de * de));
        atomicAdd(ac + AGYM_M_NPART, 1.0);
        atomicAdd(ac + AGYM_M_BEST_EV, double(ev.best_ev));
      }
    }
  }
}

template <int DMAX>
static int launch_k2_d(agym_handle* h, const SimParams& p, const float* ctx, const uint8_t* parts, uint8_t* item, float* est,
                       float* true_ctr, float* best_ev, float* value, cudaStream_t s) {
  const long long N = (long long)p.R * p.T;
  const int G = sim_group_width(h, p.P, DMAX);  // the fused kernel's lane-group width: same Thompson noise addressing
  const long long threads = N * G;
  const unsigned grid = unsigned((threads + 255) / 256);
  if (G == 4) k2_kernel<4, DMAX><<<grid, 256, 0, s>>>(p, ctx, parts, item, est, true_ctr, best_ev, value, N);
  else if (G == 8) k2_kernel<8, DMAX><<<grid, 256, 0, s>>>(p, ctx, parts, item, est, true_ctr, best_ev, value, N);
  else if (G == 16) k2_kernel<16, DMAX><<<grid, 256, 0, s>>>(p, ctx, parts, item, est, true_ctr, best_ev, value, N);
  else k2_kernel<32, DMAX><<<grid, 256, 0, s>>>(p, ctx, parts, item, est, true_ctr, best_ev, value, N);
  h->launches += 1;
  return check_cuda(h, cudaGetLastError(), "k2_kernel");
}

int launch_k2(agym_handle* h, const SimParams& p, const float* ctx, const uint8_t* parts, uint8_t* item, float* est,
              float* true_ctr, float* best_ev, float* value, cudaStream_t s) {
  if (p.D > 32) return set_error(h, AGYM_ERR_UNSUPPORTED, "K2: embedding_size > 32");
  if (p.D <= 8) return launch_k2_d<8>(h, p, ctx, parts, item, est, true_ctr, best_ev, value, s);
  return launch_k2_d<32>(h, p, ctx, parts, item, est, true_ctr, best_ev, value, s);
}

// ------------------------------------------------------------------------------------------------
// K3: one thread per (opportunity, slot) (Bidder.py:34-35,171-179)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k3_kernel(const SimParams p, const uint8_t* __restrict__ parts, const float* __restrict__ est,
                                                 const float* __restrict__ value, float* __restrict__ bid, float* __restrict__ gamma,
                                                 float* __restrict__ prop, long long NP) {
  const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= NP) return;
  const long long n = o / p.P;
  const int s = int(o - n * p.P);
  const int run = int(n / p.T);
  const long long t = n - (long long)run * p.T;
  const PhiloxKey key = make_key(p.seed, uint32_t(p.run_offset + run));
  const RoundCounter rc(p.round0 + t, p.iter);
  const int a = parts[o];
  float g, pr;
  int eff;
  float b = shade_bid<float>(p, run, a, s, value[o], est[o], false, 0.0, rc, key, g, pr, eff);
  if (eff == AGYM_BID_SEARCH) {  // one thread walks the whole grid here (the fused kernel spreads it over the lane group)
    g = search_gamma<float, 1, false>(p, run, a, s, b, est[o], value[o], nullptr, 0, rc, key, 0);
    pr = 1.0f;
    b *= g;
  }
  bid[o] = b;
  if (gamma) gamma[o] = g;
  if (prop) prop[o] = pr;
  if (p.acc && g == g) atomicAdd(p.acc + ((size_t)run * p.A + a) * kNumMetrics + AGYM_M_GAMMA, double(g));
}

int launch_k3(agym_handle* h, const SimParams& p, const uint8_t* parts, const float* est, const float* value, float* bid,
              float* gamma, float* propensity, cudaStream_t s) {
  const long long NP = (long long)p.R * p.T * p.P;
  k3_kernel<<<unsigned((NP + 255) / 256), 256, 0, s>>>(p, parts, est, value, bid, gamma, propensity, NP);
  h->launches += 1;
  return check_cuda(h, cudaGetLastError(), "k3_kernel");
}

// ------------------------------------------------------------------------------------------------
// K4 + K5: resolution, click, charge (AuctionAllocation.py:18-35, Auction.py:60-74, Agent.py:70-77)
// ------------------------------------------------------------------------------------------------
struct Resolved {
  int wslot;
  float price, second;
  bool click;
};

__device__ __forceinline__ void k4_accumulate(const SimParams& p, int run, int P, const float* b, const float* c, const float* v,
                                              const int* ag, const Resolved& r, bool valid) {
  for (int s = 0; s < P; ++s) {
    double* __restrict__ ac = p.acc + ((size_t)run * p.A + ag[s]) * kNumMetrics;
    const float tv = c[s] * v[s];
    if (valid && s == r.wslot) {
      const float got = r.click ? v[s] : 0.0f;
      atomicAdd(ac + AGYM_M_NET, double(got - r.price));
      atomicAdd(ac + AGYM_M_GROSS, double(got));
      atomicAdd(ac + AGYM_M_OVERBID_REGRET, double(r.price - r.second));
      atomicAdd(ac + AGYM_M_NWON, 1.0);
    } else if (r.price < tv) {
      atomicAdd(ac + AGYM_M_UNDERBID_REGRET, double(r.price - b[s]));
    }
  }
}

// P == 2 fast path: each thread resolves 4 consecutive opportunities with 128-bit loads / stores and one
// Philox block for the four click uniforms.
__global__ void __launch_bounds__(256) k4_kernel_p2(const SimParams p, const float4* __restrict__ bid, const float4* __restrict__ ctr,
                                                    const float4* __restrict__ val, const uint2* __restrict__ parts,
                                                    uint32_t* __restrict__ winner, float4* __restrict__ price, float4* __restrict__ second,
                                                    uint32_t* __restrict__ outcome, long long N4, int accumulate) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= N4) return;
  const long long n0 = q * 4;
  const int run = int(n0 / p.T);
  const long long t0 = n0 - (long long)run * p.T;  // multiple of 4 (host guarantees T % 4 == 0)
  const PhiloxKey key = make_key(p.seed, uint32_t(p.run_offset + run));
  const long long ta = p.round0 + t0;
  const uint4 w = click_block(ta, p.iter, key);
  const float4 b0 = __ldg(bid + 2 * q), b1 = __ldg(bid + 2 * q + 1);
  const float4 c0 = __ldg(ctr + 2 * q), c1 = __ldg(ctr + 2 * q + 1);
  const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
  const float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
  const uint32_t uu[4] = {w.x, w.y, w.z, w.w};
  float pr[4], se[4];
  uint32_t wpack = 0, opack = 0;
  const bool first = p.mechanism == AGYM_FIRST_PRICE;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float x = bb[2 * j], y = bb[2 * j + 1];
    const int ws = y > x ? 1 : 0;  // lowest slot on ties
    const float hi = ws ? y : x, lo = ws ? x : y;
    se[j] = lo;
    pr[j] = first ? hi : lo;
    const bool click = u32_to_unit(uu[j]) < cc[2 * j + ws];
    wpack |= uint32_t(ws) << (8 * j);
    opack |= uint32_t(click) << (8 * j);
  }
  winner[q] = wpack;
  outcome[q] = opack;
  price[q] = make_float4(pr[0], pr[1], pr[2], pr[3]);
  second[q] = make_float4(se[0], se[1], se[2], se[3]);
  if (accumulate) {
    const float4 v0 = __ldg(val + 2 * q), v1 = __ldg(val + 2 * q + 1);
    const float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    const uint2 pa = __ldg(parts + q);
    float rev = 0.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t pw = j < 2 ? pa.x : pa.y;
      const int ag[2] = {int((pw >> (16 * (j & 1))) & 0xFF), int((pw >> (16 * (j & 1) + 8)) & 0xFF)};
      Resolved r{int((wpack >> (8 * j)) & 1), pr[j], se[j], ((opack >> (8 * j)) & 1) != 0};
      k4_accumulate(p, run, 2, bb + 2 * j, cc + 2 * j, vv + 2 * j, ag, r, true);
      rev += pr[j];
    }
    // revenue: one atomic per warp when the warp sits inside one run (Auction.py:74)
    const unsigned full = __activemask();
    const int run_lo = __shfl_sync(full, run, __ffs(full) - 1);
    if (full == 0xffffffffu && __all_sync(full, run == run_lo)) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) rev += __shfl_xor_sync(full, rev, off);
      if ((threadIdx.x & 31) == 0) atomicAdd(p.revenue + run, double(rev));
    } else {
      atomicAdd(p.revenue + run, double(rev));
    }
  }
}

// P == 2 with accumulation (K4 + K5): the same resolution, plus charge / set_price bookkeeping (Agent.py:70-77,
// 104-112, Auction.py:74).  A CTA walks kK4ChunksPerCta chunks of kK4Chunk consecutive opportunities of ONE run (thread =
// 4 opportunities per chunk, 128-bit loads / stores).  The CTA keeps one histogram [A] x {net, gross, overbid,
// underbid} of FIXED-POINT sums (2^-32, two 32-bit words) plus a win counter in shared memory and charges it with
// native fire-and-forget shared-memory integer atomics (k4_add_fixed): integer addition commutes, so the result is
// exact to 2^-32 per term and bit-reproducible whatever the order, and lanes that charge the same agent need no
// software arbitration.
// At the end the CTA stores per-agent doubles to its own slot of
// scratch [R][CTAs per run][A * 5 + 1] (the last value is the CTA's revenue); the last CTA of a run to finish (one
// counter per run) adds the run's slots in order into the FP64 accumulators.  Extra HBM traffic: under 1 %.
// Measured and rejected on B200 (512 runs x 10 000 rounds, 184 MB algorithmic, L2 flushed): one red.global.add.v4.f32
// per opportunity (round 1) 88 us = 0.32 of the HBM peak, bound by the L2 atomic units; float read-modify-write with
// __match_any_sync ranks, one rank per round, 67 - 72 us (650 warp instructions per 128 opportunities, half of the issue
// slots idle on the shared-memory round trips); redux.sync over each group's mask 240 us (a reduction per distinct mask
// serialises); 64-bit fixed point as atom.shared.add.u64 (a CAS spin on this architecture) or as two 32-bit adds with a
// carry taken from the returned old value 49 - 58 us (the returned value stalls the warp); shared-memory float atomics
// are CAS loops as well.
constexpr int kK4Chunk = 1024;      // opportunities per CTA and chunk
constexpr int kK4Vals = 5;          // per-agent sums kept by the staged kernel: net, gross, overbid, underbid, wins
constexpr int kK4MaxA = 256;        // uint8 agent ids
// v = C * 2^-12 + F * 2^-32 with integers C = rint(v * 2^12) and F = rint((v - C * 2^-12) * 2^32), |F| <= 2^19: two
// fire-and-forget 32-bit atomics per value (no returned value to wait for, no carry).  A CTA charges at most
// kK4ChunksPerCta * kK4Chunk = 2048 opportunities to one agent, so |v| < 64 cannot overflow either word; larger values
// (no shipped catalog has them: V ~ LogNormal(0.1, 0.2)) go to a float cell through the CAS loop.
// Layout [value][C, F, overflow][A]: the same word of consecutive agents sits in consecutive banks.
constexpr int kK4Words = 3;
__device__ __forceinline__ void k4_add_fixed(int* __restrict__ hist, int A, int value, int agent, float v) {
  int* cell = hist + (kK4Words * value) * A + agent;
  if (fabsf(v) < 64.0f) {
    const float c = rintf(v * 4096.0f);
    const float r = fmaf(c, -1.0f / 4096.0f, v);  // exact
    atomicAdd(cell, int(c));
    atomicAdd(cell + A, __float2int_rn(r * 4294967296.0f));
  } else {
    atomicAdd(reinterpret_cast<float*>(cell + 2 * A), v);
  }
}
__device__ __forceinline__ double k4_read_fixed(const int* __restrict__ hist, int A, int value, int agent) {
  const int* cell = hist + (kK4Words * value) * A + agent;
  return double(cell[0]) * (1.0 / 4096.0) + double(cell[A]) * (1.0 / 4294967296.0) + double(__int_as_float(cell[2 * A]));
}

template <int kK4ChunksPerCta, int kMinBlocks>  // chunks a CTA walks with one histogram (zeroed and folded once)
__global__ void __launch_bounds__(256, kMinBlocks) k4_kernel_p2_acc(const SimParams p, const float4* __restrict__ bid, const float4* __restrict__ ctr,
                                                           const float4* __restrict__ val, const uint2* __restrict__ parts,
                                                           uint32_t* __restrict__ winner, float4* __restrict__ price,
                                                           float4* __restrict__ second, uint32_t* __restrict__ outcome, int ppr,
                                                           double* __restrict__ scratch, int* __restrict__ done) {
  extern __shared__ __align__(16) int k4_hist[];  // [4 values][C, F, overflow][A] fixed point, then int [A] wins, then float [8] revenue per warp
  __shared__ int s_last;
  static_assert(kK4ChunksPerCta * kK4Chunk <= 2048, "k4_add_fixed: overflow bound");
  const int A = p.A, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int run = blockIdx.x / ppr, part = blockIdx.x - run * ppr;
  int* __restrict__ wins = k4_hist + (size_t)A * 4 * kK4Words;
  float* __restrict__ wrev = reinterpret_cast<float*>(wins + A);
  for (int i = threadIdx.x; i < A * (4 * kK4Words + 1); i += 256) k4_hist[i] = 0;  // one histogram per CTA: the atomics arbitrate between warps too
  __syncthreads();
  const PhiloxKey key = make_key(p.seed, uint32_t(p.run_offset + run));
  const bool first = p.mechanism == AGYM_FIRST_PRICE;
  float rev = 0.0f;
  for (int c = 0; c < kK4ChunksPerCta; ++c) {
    const long long tc = ((long long)part * kK4ChunksPerCta + c) * kK4Chunk;
    if (tc >= p.T) break;  // uniform across the CTA
    const long long t0 = tc + 4 * threadIdx.x;
    if (t0 >= p.T) continue;  // the run's last chunk is partial
    const long long q = ((long long)run * p.T + t0) >> 2;
    // streamed once: evict-first loads and stores
    const float4 b0 = __ldcs(bid + 2 * q), b1 = __ldcs(bid + 2 * q + 1);
    const float4 c0 = __ldcs(ctr + 2 * q), c1 = __ldcs(ctr + 2 * q + 1);
    const float4 v0 = __ldcs(val + 2 * q), v1 = __ldcs(val + 2 * q + 1);
    const uint2 pa = __ldcs(parts + q);
    const uint4 w = click_block(p.round0 + t0, p.iter, key);
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    const float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
    const float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    const uint32_t uu[4] = {w.x, w.y, w.z, w.w};
    float pr[4], se[4];
    uint32_t wpack = 0, opack = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float x = bb[2 * j], y = bb[2 * j + 1];
      const int ws = y > x ? 1 : 0;  // lowest slot on ties
      const float hi = ws ? y : x, lo = ws ? x : y;
      se[j] = lo;
      pr[j] = first ? hi : lo;
      const float c_w = ws ? cc[2 * j + 1] : cc[2 * j], c_l = ws ? cc[2 * j] : cc[2 * j + 1];  // selects: no dynamically indexed arrays
      const float v_w = ws ? vv[2 * j + 1] : vv[2 * j], v_l = ws ? vv[2 * j] : vv[2 * j + 1];
      const bool click = u32_to_unit(uu[j]) < c_w;
      wpack |= uint32_t(ws) << (8 * j);
      opack |= uint32_t(click) << (8 * j);
      const uint32_t pw = (j < 2 ? pa.x : pa.y) >> (16 * (j & 1));
      const int ag0 = int(pw & 0xFF), ag1 = int((pw >> 8) & 0xFF);
      const int ag_w = ws ? ag1 : ag0, ag_l = ws ? ag0 : ag1;
      const float got = click ? v_w : 0.0f;
      k4_add_fixed(k4_hist, A, 0, ag_w, got - pr[j]);                  // net utility      (Agent.py:72-73)
      if (click) k4_add_fixed(k4_hist, A, 1, ag_w, got);               // gross utility    (Agent.py:74)
      if (pr[j] != lo) k4_add_fixed(k4_hist, A, 2, ag_w, pr[j] - lo);  // overbid regret   (Agent.py:104-106): first price only
      atomicAdd(wins + ag_w, 1);
      if (pr[j] < c_l * v_l && pr[j] != lo) k4_add_fixed(k4_hist, A, 3, ag_l, pr[j] - lo);  // underbid regret; the loser's bid is `lo`
      rev += pr[j];
    }
    __stcs(winner + q, wpack);
    __stcs(outcome + q, opack);
    __stcs(price + q, make_float4(pr[0], pr[1], pr[2], pr[3]));
    __stcs(second + q, make_float4(se[0], se[1], se[2], se[3]));
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) rev += __shfl_xor_sync(0xffffffffu, rev, off);
  if (lane == 0) wrev[warp] = rev;
  __syncthreads();
  const int stride = A * kK4Vals + 1;
  double* __restrict__ out = scratch + ((size_t)run * ppr + part) * stride;
  for (int i = threadIdx.x; i < A * 4; i += 256) {
    const int v = i / A, a = i - v * A;
    out[a * kK4Vals + v] = k4_read_fixed(k4_hist, A, v, a);
  }
  for (int a = threadIdx.x; a < A; a += 256) out[a * kK4Vals + 4] = double(wins[a]);
  if (threadIdx.x == 0) {
    float sum = 0.f;
    for (int wi = 0; wi < 8; ++wi) sum += wrev[wi];
    out[A * kK4Vals] = double(sum);
  }
  // the run's last CTA to finish adds the run's partial sums, in part order, into the FP64 accumulators
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(done + run, 1) == ppr - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const double* __restrict__ src = scratch + (size_t)run * ppr * stride;
  for (int i = threadIdx.x; i <= A * kK4Vals; i += 256) {
    double sum = 0.0;
    for (int c = 0; c < ppr; ++c) sum += __ldcg(src + (size_t)c * stride + i);
    if (i == A * kK4Vals) { p.revenue[run] += sum; continue; }
    const int a = i / kK4Vals, k = i - a * kK4Vals;
    const int col = k == 0 ? AGYM_M_NET : k == 1 ? AGYM_M_GROSS : k == 2 ? AGYM_M_OVERBID_REGRET : k == 3 ? AGYM_M_UNDERBID_REGRET : AGYM_M_NWON;
    p.acc[((size_t)run * A + a) * kNumMetrics + col] += sum;
  }
  if (threadIdx.x == 0) done[run] = 0;  // ready for the next launch
}

// general P: one thread per opportunity, running top-2
__global__ void __launch_bounds__(256) k4_kernel_any(const SimParams p, const float* __restrict__ bid, const float* __restrict__ ctr,
                                                     const float* __restrict__ val, const uint8_t* __restrict__ parts,
                                                     uint8_t* __restrict__ winner, float* __restrict__ price, float* __restrict__ second,
                                                     uint8_t* __restrict__ outcome, long long N, int accumulate) {
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int P = p.P;
  const int run = int(n / p.T);
  const long long t = n - (long long)run * p.T;
  const PhiloxKey key = make_key(p.seed, uint32_t(p.run_offset + run));
  const long long ta = p.round0 + t;
  const uint32_t uw = click_word(RoundCounter(ta, p.iter), key);
  float best = -INFINITY, sec = -INFINITY;
  int ws = 0;
  for (int s = 0; s < P; ++s) {
    const float b = bid[n * P + s];
    if (b > best) { sec = best; best = b; ws = s; }
    else if (b > sec) { sec = b; }
  }
  const bool valid = P >= 2;
  Resolved r;
  r.wslot = ws;
  r.second = valid ? sec : 0.0f;
  r.price = valid ? (p.mechanism == AGYM_FIRST_PRICE ? best : sec) : 0.0f;
  r.click = valid && (u32_to_unit(uw) < ctr[n * P + ws]);
  winner[n] = uint8_t(ws);
  price[n] = r.price;
  second[n] = r.second;
  outcome[n] = r.click ? 1 : 0;
  if (accumulate) {
    float b[kMaxP], c[kMaxP], v[kMaxP];
    int ag[kMaxP];
    for (int s = 0; s < P; ++s) { b[s] = bid[n * P + s]; c[s] = ctr[n * P + s]; v[s] = val[n * P + s]; ag[s] = parts[n * P + s]; }
    k4_accumulate(p, run, P, b, c, v, ag, r, valid);
    if (valid) atomicAdd(p.revenue + run, double(r.price));
  }
}

// ------------------------------------------------------------------------------------------------
// Allocator.estimate_CTR for ONE context: every item of one (run, agent), in the reference's precision mix
// (OracleAllocator: float64, BidderAllocation.py:81-82; PyTorchLogisticRegressionAllocator: the context cast to float32,
// weights m or m + eps / sqrt(q) with eps ~ N(0, 1) per weight, BidderAllocation.py:67-68, Models.py:28-33).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64) estimate_kernel(const SimParams p, int run, int a, const double* __restrict__ ctx, int sample,
                                                      const float* __restrict__ eps, double* __restrict__ out) {
  using A_ = Arith<double>;
  const int nI = p.n_items[a], K = p.Do + 1;
  const PhiloxKey key = make_key(p.seed, uint32_t(p.run_offset + run));
  for (int i = threadIdx.x; i < nI; i += blockDim.x) {
    if (p.alloc_kind[a] == AGYM_ALLOC_ORACLE) {
      const double* __restrict__ e = p.E64 + ((size_t)a * p.I + i) * (p.D + 1);
      double z = 0.0;
      for (int d = 0; d <= p.D; ++d) z = fma(ctx[d], e[d], z);
      out[i] = A_::sigmoid(z);
      continue;
    }
    const size_t o = (((size_t)run * p.A + a) * p.I + i) * K;
    float zl = 0.0f;
    for (int kb = 0; kb * 4 < K; ++kb) {
      float4 nrm = make_float4(0.f, 0.f, 0.f, 0.f);
      if (sample && !eps) nrm = philox_normal4(uint32_t(p.round0), uint32_t(p.iter) ^ (uint32_t(p.round0 >> 32) << 20), kPurposeTS << 16, uint32_t(i * 16 + kb), key);
      for (int j = 0; j < 4 && kb * 4 + j < K; ++j) {
        const int k = kb * 4 + j;
        float w = p.m[o + k];
        if (sample) w = A_::ts_weight(w, eps ? eps[(size_t)i * K + k] : pick4<float>(nrm, j), p.sigma[o + k]);  // Models.py:31
        zl = A_::mac(w, float(ctx[k]), zl);
      }
    }
    out[i] = double(A_::sigmoid32(zl));
  }
}

int launch_estimate(agym_handle* h, const SimParams& p, int run, int a, const double* ctx, int sample, const float* eps, double* out, cudaStream_t s) {
  estimate_kernel<<<1, 64, 0, s>>>(p, run, a, ctx, sample, eps, out);
  h->launches += 1;
  return check_cuda(h, cudaGetLastError(), "estimate_kernel");
}

int launch_k4(agym_handle* h, const SimParams& p, const float* bid, const float* true_ctr, const float* value,
              const uint8_t* parts, uint8_t* winner, float* price, float* second, uint8_t* outcome, int accumulate,
              cudaStream_t s) {
  if (p.P > kMaxP) return set_error(h, AGYM_ERR_UNSUPPORTED, "K4: P > 32");
  const long long N = (long long)p.R * p.T;
  const bool aligned = ((uintptr_t)bid % 16 == 0) && ((uintptr_t)true_ctr % 16 == 0) && ((uintptr_t)value % 16 == 0) &&
                       ((uintptr_t)parts % 8 == 0) && ((uintptr_t)winner % 4 == 0) && ((uintptr_t)price % 16 == 0) &&
                       ((uintptr_t)second % 16 == 0) && ((uintptr_t)outcome % 4 == 0);
  if (p.P == 2 && p.T % 4 == 0 && p.round0 % 4 == 0 && aligned) {
    const long long N4 = N / 4;
    if (accumulate && p.A <= kK4MaxA) {
      const int chunks = int((p.T + kK4Chunk - 1) / kK4Chunk);
      constexpr int cpc = 2;  // B200, 512 runs: 1 -> 55 us, 2 -> 49 us, 5 -> 51 us (with the 64-bit variant); 5 / 6 / 4 CTAs per SM -> 49 / 51 / 51 us
      const int ppr = (chunks + cpc - 1) / cpc;  // CTAs (= partial-sum slots) per run
      const size_t nfl = (size_t)p.R * ppr * (p.A * kK4Vals + 1);
      const size_t nsc = nfl * sizeof(double) + (size_t)p.R * sizeof(int);
      if (h->k4_scratch == nullptr || h->k4_scratch_bytes != nsc) {  // first use: allocate + clear the counters (configuration-time, synchronous)
        cudaFree(h->k4_scratch);
        h->k4_scratch = nullptr;
        cudaError_t e = cudaMalloc(&h->k4_scratch, nsc);
        if (e == cudaSuccess) e = cudaMemset(h->k4_scratch, 0, nsc);
        if (e != cudaSuccess) return check_cuda(h, e, "k4 scratch");
        h->k4_scratch_bytes = nsc;
      }
      const size_t smem = (size_t)p.A * (4 * kK4Words + 1) * sizeof(int) + 8 * sizeof(float);
      k4_kernel_p2_acc<cpc, 5><<<unsigned(p.R) * unsigned(ppr), 256, smem, s>>>(
          p, (const float4*)bid, (const float4*)true_ctr, (const float4*)value, (const uint2*)parts, (uint32_t*)winner, (float4*)price,
          (float4*)second, (uint32_t*)outcome, ppr, reinterpret_cast<double*>(h->k4_scratch),
          reinterpret_cast<int*>(reinterpret_cast<double*>(h->k4_scratch) + nfl));
      h->launches += 1;
      return check_cuda(h, cudaGetLastError(), "k4_kernel_p2_acc");
    }
    k4_kernel_p2<<<unsigned((N4 + 255) / 256), 256, 0, s>>>(p, (const float4*)bid, (const float4*)true_ctr, (const float4*)value,
                                                            (const uint2*)parts, (uint32_t*)winner, (float4*)price, (float4*)second,
                                                            (uint32_t*)outcome, N4, accumulate);
  } else {
    k4_kernel_any<<<unsigned((N + 255) / 256), 256, 0, s>>>(p, bid, true_ctr, value, parts, winner, price, second, outcome, N, accumulate);
  }
  h->launches += 1;
  return check_cuda(h, cudaGetLastError(), "k4_kernel");
}

}  // namespace agym
