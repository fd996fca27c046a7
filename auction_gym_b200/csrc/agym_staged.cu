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
  int G = 8;  // the fused kernel's lane-group width (agym_sim.cu launch_d): same Thompson noise addressing
  if (h->has_option("sim_g")) { const int v = int(h->option("sim_g", 8)); if (v == 8 || v == 16 || v == 32) G = v; }
  while (G < p.P) G *= 2;
  if (DMAX / 4 > G) G = 32;
  const long long threads = N * G;
  const unsigned grid = unsigned((threads + 255) / 256);
  if (G == 8) k2_kernel<8, DMAX><<<grid, 256, 0, s>>>(p, ctx, parts, item, est, true_ctr, best_ev, value, N);
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
// 104-112, Auction.py:74).  The four winner-side sums of one opportunity {net, gross, overbid, wins} go out as ONE
// 128-bit vector reduction (red.global.add.v4.f32, sm_90+) into a float scratch block [R][A][kK4Buckets][8] that is
// bucketed by CTA (short float partial sums, less same-address contention); k4_fold_kernel adds the scratch into the
// FP64 accumulators and clears it.  Shared-memory float atomics (a CAS loop on this architecture) and one FP64 atomic
// per metric were both measured ~4x slower.
constexpr int kK4Buckets = 8;
__global__ void __launch_bounds__(256) k4_kernel_p2_acc(const SimParams p, const float4* __restrict__ bid, const float4* __restrict__ ctr,
                                                        const float4* __restrict__ val, const uint2* __restrict__ parts,
                                                        uint32_t* __restrict__ winner, float4* __restrict__ price,
                                                        float4* __restrict__ second, uint32_t* __restrict__ outcome, long long N4,
                                                        float* __restrict__ scratch) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= N4) return;
  const long long n0 = q * 4;
  const int run = int(n0 / p.T);
  const long long t0 = n0 - (long long)run * p.T;
  const PhiloxKey key = make_key(p.seed, uint32_t(p.run_offset + run));
  const uint4 w = click_block(p.round0 + t0, p.iter, key);
  const float4 b0 = __ldg(bid + 2 * q), b1 = __ldg(bid + 2 * q + 1);
  const float4 c0 = __ldg(ctr + 2 * q), c1 = __ldg(ctr + 2 * q + 1);
  const float4 v0 = __ldg(val + 2 * q), v1 = __ldg(val + 2 * q + 1);
  const uint2 pa = __ldg(parts + q);
  const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
  const float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
  const float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
  const uint32_t uu[4] = {w.x, w.y, w.z, w.w};
  float pr[4], se[4], rev = 0.0f;
  uint32_t wpack = 0, opack = 0;
  const bool first = p.mechanism == AGYM_FIRST_PRICE;
  float* __restrict__ mine = scratch + ((size_t)run * p.A * kK4Buckets + (blockIdx.x % kK4Buckets)) * 8;
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
    const uint32_t pw = (j < 2 ? pa.x : pa.y) >> (16 * (j & 1));
    const int ag_w = int((pw >> (8 * ws)) & 0xFF), ag_l = int((pw >> (8 * (1 - ws))) & 0xFF);
    const float got = click ? vv[2 * j + ws] : 0.0f;
    atomicAdd(reinterpret_cast<float4*>(mine + (size_t)ag_w * kK4Buckets * 8), make_float4(got - pr[j], got, pr[j] - lo, 1.0f));
    const float tv_l = cc[2 * j + 1 - ws] * vv[2 * j + 1 - ws];
    if (pr[j] < tv_l && pr[j] != lo) atomicAdd(mine + (size_t)ag_l * kK4Buckets * 8 + 4, pr[j] - lo);  // the loser's bid is `lo`
    rev += pr[j];
  }
  winner[q] = wpack;
  outcome[q] = opack;
  price[q] = make_float4(pr[0], pr[1], pr[2], pr[3]);
  second[q] = make_float4(se[0], se[1], se[2], se[3]);
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

// scratch [R*A][kK4Buckets][8] -> acc [R*A][12] (FP64), scratch cleared; one thread per (run, agent)
__global__ void __launch_bounds__(256) k4_fold_kernel(float* __restrict__ scratch, double* __restrict__ acc, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s[5] = {0, 0, 0, 0, 0};
  float4* sc = reinterpret_cast<float4*>(scratch + (size_t)i * kK4Buckets * 8);
#pragma unroll
  for (int b = 0; b < kK4Buckets; ++b) {
    const float4 u = sc[2 * b], v = sc[2 * b + 1];
    s[0] += u.x; s[1] += u.y; s[2] += u.z; s[3] += u.w; s[4] += v.x;
    sc[2 * b] = make_float4(0.f, 0.f, 0.f, 0.f);
    sc[2 * b + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  double* a = acc + (size_t)i * kNumMetrics;
  a[AGYM_M_NET] += s[0]; a[AGYM_M_GROSS] += s[1]; a[AGYM_M_OVERBID_REGRET] += s[2]; a[AGYM_M_NWON] += s[3];
  a[AGYM_M_UNDERBID_REGRET] += s[4];
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
    if (accumulate) {
      const size_t nsc = (size_t)p.R * p.A * kK4Buckets * 8;
      if (h->k4_scratch == nullptr) {  // first use: allocate + clear (configuration-time cost, synchronous)
        cudaError_t e = cudaMalloc(&h->k4_scratch, nsc * sizeof(float));
        if (e == cudaSuccess) e = cudaMemset(h->k4_scratch, 0, nsc * sizeof(float));
        if (e != cudaSuccess) return check_cuda(h, e, "k4 scratch");
      }
      k4_kernel_p2_acc<<<unsigned((N4 + 255) / 256), 256, 0, s>>>(p, (const float4*)bid, (const float4*)true_ctr, (const float4*)value,
                                                                 (const uint2*)parts, (uint32_t*)winner, (float4*)price,
                                                                 (float4*)second, (uint32_t*)outcome, N4, h->k4_scratch);
      k4_fold_kernel<<<unsigned((p.R * p.A + 255) / 256), 256, 0, s>>>(h->k4_scratch, p.acc, p.R * p.A);
      h->launches += 2;
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
