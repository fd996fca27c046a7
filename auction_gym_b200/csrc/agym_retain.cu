// Log retention across iterations: Agent.clear_logs with memory > 0 (reference src/Agent.py:124-129) and the
// bidders' clear_logs (src/Bidder.py:149-153,327-333,433-439,617-623).
//
// The reference keeps `self.logs[-memory:]` of every agent; everything that reads the logs afterwards -- the
// metric getters (Agent.py:96-118), allocator.update on the won rows and bidder.update on all rows
// (Agent.py:79-94) -- then sees the kept records followed by the next iteration's.  Here the kept records are
// moved to the head of the winner log and of the bid log (rows [0, B), B = sum(memory), agent a owns rows
// [off[a], off[a] + memory[a]), oldest first, slot 0 only; unused rows are invalid), the round loop appends at
// row B, and the bucketing kernels of both fits -- stable by agent -- deliver retained-then-new rows without
// knowing about retention.  The metric accumulators restart from the sums over the kept records.
//
// One warp per (run, agent).  In-place compaction is safe: rows are visited in time order, a kept record's new row
// index never exceeds its old one, every 32-row batch is read completely before any of it is written, and other
// agents' regions are never written by this warp.  Not a hot path (it moves <= memory records per agent per
// iteration), so it is written for obviousness.
#include "agym_common.cuh"

namespace agym {

struct RetainParams {
  int R, A, P, Do;
  long long bid_Tcap, fit_Tcap;  // capacities of the two logs (rows per run)
  long long Tn;                  // rows to scan: retained region + rounds recorded this iteration
  const int* memory;             // [A]
  const int* mem_off;            // [A]
  float* bid_rows;
  uint32_t* bid_meta;
  double* terms;
  float* fit_ctx;      // nullable
  uint32_t* fit_meta;  // nullable
  double* acc;
};

constexpr int kRetainMaxDo = 32;

__global__ void __launch_bounds__(32) retain_kernel(const RetainParams p) {
  const int a = blockIdx.x, run = blockIdx.y, lane = threadIdx.x;
  const int P = p.P, Do = p.Do;
  const int mem = p.memory[a], off = p.mem_off[a];
  uint32_t* __restrict__ meta = p.bid_meta + (size_t)run * p.bid_Tcap * P;
  float* __restrict__ rows = p.bid_rows + (size_t)run * p.bid_Tcap * P * AGYM_BID_ROW;
  double* __restrict__ terms = p.terms + (size_t)run * p.bid_Tcap * P * AGYM_TERM_ROW;
  float* __restrict__ fctx = p.fit_ctx ? p.fit_ctx + (size_t)run * p.fit_Tcap * Do : nullptr;
  uint32_t* __restrict__ fmeta = p.fit_meta ? p.fit_meta + (size_t)run * p.fit_Tcap : nullptr;
  const long long NR = p.Tn * P;
  const unsigned lt = (1u << lane) - 1u;

  int cnt = 0;
  if (mem > 0) {
    for (long long base = 0; base < NR; base += 32) {
      const long long j = base + lane;
      const uint32_t mt = j < NR ? meta[j] : 0u;
      cnt += __popc(__ballot_sync(0xffffffffu, (mt & kBidValid) && int(mt & 0xFFFu) == a));
    }
  }
  const int keep = cnt < mem ? cnt : mem, skip = cnt - keep;

  double sum[AGYM_TERM_ROW];
#pragma unroll
  for (int k = 0; k < AGYM_TERM_ROW; ++k) sum[k] = 0.0;
  int nwon = 0, seen = 0;
  for (long long base = 0; base < NR && seen < cnt; base += 32) {
    const long long j = base + lane;
    const uint32_t mt = j < NR ? meta[j] : 0u;
    const bool mine = (mt & kBidValid) && int(mt & 0xFFFu) == a;
    const unsigned b = __ballot_sync(0xffffffffu, mine);
    const int rank = seen + __popc(b & lt);
    seen += __popc(b);
    const bool take = mine && rank >= skip;
    float row[AGYM_BID_ROW];
    double tm[AGYM_TERM_ROW];
    float cx[kRetainMaxDo];
    if (take) {
#pragma unroll
      for (int k = 0; k < AGYM_BID_ROW; ++k) row[k] = rows[(size_t)j * AGYM_BID_ROW + k];
#pragma unroll
      for (int k = 0; k < AGYM_TERM_ROW; ++k) tm[k] = terms[(size_t)j * AGYM_TERM_ROW + k];
      if (fctx)
        for (int k = 0; k < Do; ++k) cx[k] = fctx[(size_t)(j / P) * Do + k];
    }
    __syncwarp();
    if (take) {
      const size_t d = size_t(off + rank - skip);  // retained row, slot 0
      const bool won = (mt & kBidWon) != 0;
#pragma unroll
      for (int k = 0; k < AGYM_BID_ROW; ++k) rows[d * P * AGYM_BID_ROW + k] = row[k];
#pragma unroll
      for (int k = 0; k < AGYM_TERM_ROW; ++k) { terms[d * P * AGYM_TERM_ROW + k] = tm[k]; sum[k] += tm[k]; }
      meta[d * P] = mt;
      if (fctx) {
        for (int k = 0; k < Do; ++k) fctx[d * Do + k] = cx[k];
        fmeta[d] = won ? pack_meta(a, int((mt >> 12) & 0xFFFu), (mt & kBidClick) != 0) : 0u;
      }
      nwon += won;
    }
    __syncwarp();
  }
  // rows of this agent's region that hold nothing
  for (int j = keep + lane; j < mem; j += 32) {
    meta[size_t(off + j) * P] = 0u;
    if (fmeta) fmeta[off + j] = 0u;
  }
  // the getters' sums over the kept records, in a fixed order (lane partials in time order, then a shuffle tree)
#pragma unroll
  for (int k = 0; k < AGYM_TERM_ROW; ++k)
    for (int o = 16; o > 0; o >>= 1) sum[k] += __shfl_xor_sync(0xffffffffu, sum[k], o);
  for (int o = 16; o > 0; o >>= 1) nwon += __shfl_xor_sync(0xffffffffu, nwon, o);
  if (lane == 0) {
    double* __restrict__ ac = p.acc + ((size_t)run * p.A + a) * kNumMetrics;
    ac[AGYM_M_ALLOC_REGRET] = sum[0]; ac[AGYM_M_ESTIM_REGRET] = sum[1];
    ac[AGYM_M_OVERBID_REGRET] = sum[2]; ac[AGYM_M_UNDERBID_REGRET] = sum[3];
    ac[AGYM_M_SQERR] = sum[4]; ac[AGYM_M_BIAS] = sum[5];
    ac[AGYM_M_GAMMA] = sum[6]; ac[AGYM_M_BEST_EV] = sum[7];
    ac[AGYM_M_NPART] = double(keep); ac[AGYM_M_NWON] = double(nwon);
  }
}

int launch_retain_logs(agym_handle* h, cudaStream_t s) {
  const agym_shape& sh = h->shape;
  if (sh.Do > kRetainMaxDo) return set_error(h, AGYM_ERR_UNSUPPORTED, "agym_retain_logs: obs_embedding_size > 32");
  RetainParams rp{};
  rp.R = sh.R; rp.A = sh.A; rp.P = sh.P; rp.Do = sh.Do;
  rp.bid_Tcap = h->bid_Tcap; rp.fit_Tcap = h->Tcap;
  rp.Tn = h->log_base + h->rounds_in_iter;
  rp.memory = h->d_memory; rp.mem_off = h->d_mem_off;
  rp.bid_rows = h->bid_rows; rp.bid_meta = h->bid_meta; rp.terms = h->terms;
  rp.fit_ctx = h->fit_ctx; rp.fit_meta = h->fit_meta;
  rp.acc = h->acc;
  if (sh.R > 65535) return set_error(h, AGYM_ERR_UNSUPPORTED, "agym_retain_logs: more than 65535 resident runs");
  retain_kernel<<<dim3(unsigned(sh.A), unsigned(sh.R)), 32, 0, s>>>(rp);
  h->launches += 1;
  return check_cuda(h, cudaGetLastError(), "retain_kernel");
}

// invalidate the retained rows of both logs (all slots), e.g. when retention is (re)configured or dropped
int clear_retained_rows(agym_handle* h, cudaStream_t s) {
  if (h->log_base <= 0) return AGYM_OK;
  const agym_shape& sh = h->shape;
  cudaError_t e = cudaMemset2DAsync(h->bid_meta, (size_t)h->bid_Tcap * sh.P * sizeof(uint32_t), 0,
                                    (size_t)h->log_base * sh.P * sizeof(uint32_t), sh.R, s);
  if (e == cudaSuccess && h->fit_meta)
    e = cudaMemset2DAsync(h->fit_meta, (size_t)h->Tcap * sizeof(uint32_t), 0, (size_t)h->log_base * sizeof(uint32_t), sh.R, s);
  return check_cuda(h, e, "clear retained rows");
}

}  // namespace agym
