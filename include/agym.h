/*
 * agym.h -- C ABI of the B200-native AuctionGym round-loop engine (libagym.so).
 *
 * The reference (soopark0221/auction-gym) is pure Python and has no FFI of its own
 * (SURVEY.md section 8b): its boundary is the Python class surface
 *   Auction.simulate_opportunity()            reference src/Auction.py:28-74
 *   Agent.select_item / bid / charge / update  reference src/Agent.py:29-94
 *   *.allocate / estimate_CTR / Bidder.bid     reference src/AuctionAllocation.py:18-35,
 *                                              src/BidderAllocation.py:29-82, src/Bidder.py:28-208
 * The entry points below are what those Python classes bind (through ctypes, see
 * auction_gym_b200/_lib.py and INTEGRATION.md).  Conventions:
 *   - plain C types only; no torch / C++ types cross the boundary;
 *   - every function returns 0 on success and a negative agym_status otherwise; the message is
 *     available from agym_last_error(); nothing throws across the ABI;
 *   - "device pointer" arguments are BORROWED for the lifetime stated; the caller (torch) owns the
 *     memory.  "host pointer" arguments are small configuration arrays that are copied;
 *   - all device work is enqueued on the cudaStream_t passed as `void* stream`; no hidden syncs
 *     except in agym_create / agym_set_* (configuration time);
 *   - one caller thread per handle (the reference is single-threaded, src/main.py:112-155).
 *
 * Shapes:  R runs resident on this device, A agents, I = max items per agent (padded), D =
 * embedding_size, Do = obs_embedding_size, K = Do+1, P = participants per round, T rounds.
 */
#ifndef AGYM_H
#define AGYM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AGYM_ABI_VERSION 2

typedef struct agym_handle agym_handle;

enum agym_status {
  AGYM_OK = 0,
  AGYM_ERR_INVALID = -1,     /* bad argument / unsupported shape */
  AGYM_ERR_CUDA = -2,        /* CUDA runtime error (message has the cudaError string) */
  AGYM_ERR_STATE = -3,       /* a required buffer was not bound */
  AGYM_ERR_UNSUPPORTED = -4  /* feature not built yet (out-of-scope table) */
};

/* AuctionAllocation.py:10,25 */
enum agym_mechanism { AGYM_SECOND_PRICE = 0, AGYM_FIRST_PRICE = 1 };

/* BidderAllocation.py:71 (Oracle), :21 with thompson_sampling True / False */
enum agym_alloc_kind { AGYM_ALLOC_ORACLE = 0, AGYM_ALLOC_TS = 1, AGYM_ALLOC_MAP = 2 };

/* Bid-time behaviour (Bidder.py).  Kinds >= AGYM_BID_SEARCH fall back to AGYM_BID_GAUSS while the
 * per-(run, agent) `initialised` flag in bidder_d is 0 (Bidder.py:174,351,458). */
enum agym_bidder_kind {
  AGYM_BID_TRUTHFUL = 0,    /* TruthfulBidder.bid            Bidder.py:34-35  */
  AGYM_BID_GAUSS = 1,       /* gamma ~ N(prev, sigma), unclipped   Bidder.py:177,354,461 */
  AGYM_BID_GAUSS_CLIP = 2,  /* EmpiricalShadedBidder.bid     Bidder.py:47-58  */
  AGYM_BID_SEARCH = 3,      /* ValueLearningBidder 'search'  Bidder.py:180-196 */
  AGYM_BID_BANDIT = 4,      /* PolicyLearning / DoublyRobust Bidder.py:357-362,464-470 */
  AGYM_BID_POLICY = 5       /* ValueLearningBidder 'policy'  Bidder.py:198-203 */
};

/* Arithmetic of the decision path.
 *   FP32: contexts, CTRs, bids and prices in float (production throughput mode)
 *   FP64: the reference's own mix -- float64 everywhere except the learnt CTR estimate, which is
 *         float32 on a float32 copy of the observed context (BidderAllocation.py:67-68). */
enum agym_precision { AGYM_FP32 = 0, AGYM_FP64 = 1 };

typedef struct agym_shape {
  int32_t R, A, I, D, Do, P;
  int32_t mechanism;   /* agym_mechanism */
  int32_t precision;   /* agym_precision */
  int32_t run_offset;  /* global index of this device's first run: RNG keys use run_offset + r, so a
                          sharded job reproduces the single-device job (main.py:186) */
  int32_t max_slots;   /* slots per auction round are uniform in [1, max_slots] (Auction.py:30); 0 or 1 = the reference driver's
                          single slot (main.py:36-37).  With several slots the winner log holds max_slots rows per round. */
  double embedding_var; /* used as the STD of the context normal, as Auction.py:33 does */
} agym_shape;

/* Columns of the per-(run, agent) accumulator block (Agent.py:70-118, main.py:131-148). */
enum agym_metric {
  AGYM_M_NET = 0, AGYM_M_GROSS, AGYM_M_ALLOC_REGRET, AGYM_M_ESTIM_REGRET, AGYM_M_OVERBID_REGRET,
  AGYM_M_UNDERBID_REGRET, AGYM_M_SQERR, AGYM_M_BIAS, AGYM_M_NPART, AGYM_M_NWON, AGYM_M_BEST_EV,
  AGYM_M_GAMMA, AGYM_NUM_METRICS
};

/* Per-(run, agent) bidder state. bidder_d: doubles {prev_gamma, gamma_sigma, initialised, reserved};
 * bidder_w: floats {winrate w[3], b (Models.py:56), policy W1[4], b1[2], w_mu[2], b_mu, w_sigma[2],
 * b_sigma (Models.py:97-101)}. */
#define AGYM_BIDDER_D 4
#define AGYM_BIDDER_W 16

/* Detailed per-(round, slot) log = the SoA form of ImpressionOpportunity (Impression.py:4-31).
 * Device pointers, each may be NULL (field not wanted).  Layout [n_runs][T][P] unless noted. */
typedef struct agym_round_log {
  int32_t* agent;     /* participating agent per slot (Auction.py:42) */
  int32_t* item;      /* Agent.py:35 */
  double* est;        /* estimated_CTR logged by Agent.bid (Agent.py:57) */
  double* value;
  double* bid;
  double* true_ctr;   /* Auction.py:53 */
  double* best_ev;    /* Auction.py:53 */
  double* price;      /* Agent.py:70-77: winner and losers both log the price */
  double* second;     /* winner only */
  double* gamma;      /* NaN for truthful bidders */
  double* propensity; /* NaN where the reference keeps none */
  uint8_t* outcome;
  uint8_t* won;
  int32_t* winner;    /* [n_runs][T] winning slot */
  double* ctx;        /* [n_runs][T][D] sampled context without the trailing 1 (production mode) */
} agym_round_log;

/* Host-drawn noise for replay mode (SURVEY.md section 8c "replay seams").  Device pointers. */
typedef struct agym_replay_inputs {
  const double* ctx;      /* [n_runs][T][D]  rng.normal(0, embedding_var, D)      Auction.py:33 */
  const int32_t* parts;   /* [n_runs][T][P]  rng.choice(A, P, replace=False)      Auction.py:42 */
  const float* ts_eps;    /* [n_runs][T][P][I][K] standard normals for Models.py:31, or NULL */
  const double* gamma_z;  /* [n_runs][T][P] standard normals for Bidder.py:177,354,461, or NULL */
  const double* grid_u;   /* [n_runs][T][P][grid_n] uniforms for Bidder.py:185, or NULL */
  const double* u;        /* [n_runs][T][max_slots] click uniforms, one per slot: outcome = (u < p)   Auction.py:65 */
  int32_t grid_n;
  int32_t reserved;
  const int32_t* num_slots; /* [n_runs][T] rng.integers(1, max_slots + 1) (Auction.py:30), or NULL = 1 slot; only read when max_slots > 1 */
} agym_replay_inputs;

/* every entry point below is exported even when the library is built with -fvisibility=hidden */
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

int agym_abi_version(void);
const char* agym_last_error(const agym_handle* h); /* h may be NULL: error of the last failed create */

/* ---- lifetime / configuration (host pointers, synchronous) ---- */
int agym_create(const agym_shape* shape, int device, agym_handle** out);
int agym_destroy(agym_handle* h);
/* per-agent static configuration: n_items[A] (<= I), alloc_kind[A], bidder_kind[A]  (main.py:77-95) */
int agym_set_agents(agym_handle* h, const int32_t* n_items, const int32_t* alloc_kind, const int32_t* bidder_kind);
/* catalog: E [A][I][D+1] item embeddings with the intercept column, V [A][I] item values (main.py:60-72) */
int agym_set_catalog(agym_handle* h, const double* E, const double* V);

/* ---- device state owned by the caller (device pointers, borrowed until re-bound / destroy) ---- */
/* learnt allocator state: m, q, m_prev, sigma = 1/sqrt(q)   each [R][A][I][K] float  (Models.py:21-24) */
int agym_bind_allocator_state(agym_handle* h, float* m, float* q, float* m_prev, float* sigma);
/* Call after the HOST wrote m or q: sigma <- 1/sqrt(q) (Models.py:31), and the library's packed copy of {m, 1/q} that the
 * production round loop reads with 128-bit loads is rebuilt (agym_update_allocators does both itself). */
int agym_refresh_sigma(agym_handle* h, void* stream);
/* bidder state: bidder_d [R][A][AGYM_BIDDER_D] double, bidder_w [R][A][AGYM_BIDDER_W] float */
int agym_bind_bidder_state(agym_handle* h, double* bidder_d, float* bidder_w);
/* accumulators: acc [R][A][AGYM_NUM_METRICS] double, revenue [R] double  (Agent.py:20-21, Auction.py:16) */
int agym_bind_metrics(agym_handle* h, double* acc, double* revenue);
/* winner records that feed the allocator fit: fit_ctx [R][Tcap][Do] float, fit_meta [R][Tcap] uint32
 * (bit 31 valid, bit 30 click, bits 12..23 agent, bits 0..11 item)   (Agent.py:81-91) */
int agym_bind_fit_log(agym_handle* h, float* fit_ctx, uint32_t* fit_meta, int64_t Tcap);
size_t agym_workspace_bytes(const agym_handle* h, int64_t Tcap);
int agym_bind_workspace(agym_handle* h, void* ws, size_t bytes);

/* ---- the round loop (K1-K5 fused) ---- */
/* T rounds for every resident run with in-kernel Philox4x32-10 noise keyed (seed, run, iter, round).
 * Appends at round index `rounds_in_iteration`; `log` may be NULL.  Replaces the loop over
 * Auction.simulate_opportunity at src/main.py:116-117. */
int agym_simulate_rounds(agym_handle* h, uint64_t seed, int32_t iter, int64_t T, const agym_round_log* log, void* stream);
/* Same arithmetic, noise served from host-drawn arrays, for runs [run0, run0 + n_runs). */
int agym_replay_rounds(agym_handle* h, int32_t run0, int32_t n_runs, int64_t T, const agym_replay_inputs* in,
                       const agym_round_log* log, void* stream);
int64_t agym_rounds_in_iteration(const agym_handle* h);
/* For hosts that write fit_ctx / fit_meta themselves (tests): declare how many rounds of the bound logs are
 * filled (not counting the retained rows at their head). */
int agym_set_rounds_in_iteration(agym_handle* h, int64_t n);
/* Agent.clear_utility / clear_logs + Auction.clear_revenue (Agent.py:120-129, Auction.py:76): zero the
 * accumulators and rewind the logs; with log retention configured, also drops every retained record
 * (the reference's memory == 0 behaviour). */
int agym_clear_iteration(agym_handle* h, void* stream);

/* ---- log retention across iterations (Agent(memory=...), Agent.py:124-129; Bidder.clear_logs, Bidder.py:149-153,
 * 327-333,433-439,617-623) ----
 * memory [A] (host): how many of its most recent records agent a keeps when an iteration ends (0 = none).
 * The records live in the bound logs themselves: the first B = sum(memory) rows of the winner log and of the
 * bid log are reserved ("retained rows", agent-major, oldest first, slot 0 only), the round loop appends behind
 * them, and agym_update_allocators / agym_update_bidders read retained + new rows in the reference's order.
 * Needs the bid log (agym_bind_bid_log) with bid_Tcap > B, the winner log too when some allocator learns, and
 * terms [R][bid_Tcap][P][AGYM_TERM_ROW] double (device): the per-record summands of the metric getters
 * (Agent.py:96-118, main.py:142-148) {allocation regret, estimation regret, overbid, underbid, squared CTR error,
 * est/true on won rows, gamma, best expected value}, written by the round loop.
 * The caller hands over bid_meta / fit_meta with their first B rows zeroed (no retained records yet), or holding
 * records retained earlier (when it moves the logs to larger buffers: rebind the logs, then call this again with the
 * new terms buffer).  memory == NULL switches retention off.  sum(memory) may only change between iterations. */
#define AGYM_TERM_ROW 8
int agym_set_log_retention(agym_handle* h, const int32_t* memory, double* terms);
/* B = sum(memory): rows reserved at the head of the logs (0 when retention is off). */
int64_t agym_retained_capacity(const agym_handle* h);
/* Agent.clear_logs for every agent: keep each agent's last memory[a] records (retained + this iteration's, in
 * time order), rewind the logs behind them, and set the ten log-derived accumulators of every (run, agent) to the
 * sums over the kept records, which is what the reference's getters return from the shortened self.logs.
 * Net / gross utility and revenue are NOT touched (Agent.clear_utility and Auction.clear_revenue are separate calls). */
int agym_retain_logs(agym_handle* h, void* stream);

/* Allocator.estimate_CTR for ONE context (BidderAllocation.py:67-68, 81-82; Models.py:28-33): the estimated CTR of every
 * item of `agent` in resident run `run`.  context (HOST, D+1 doubles for an OracleAllocator agent -- the true context --,
 * Do+1 for a learnt one; the trailing 1 included, Auction.py:33-49).  sample != 0: Thompson draw m + eps / sqrt(q) per weight,
 * eps (HOST, nullable, [I][Do+1] float) supplied by the caller (replay) or drawn from Philox keyed by (seed, run, iter, query).
 * out (HOST, I doubles; float32 values for learnt agents, as the reference returns).  Synchronises the stream: like the
 * reference's method it returns a host array. */
int agym_estimate_ctr(agym_handle* h, int32_t run, int32_t agent, const double* context, int32_t sample, const float* eps, uint64_t seed,
                      int32_t iter, int64_t query, double* out, void* stream);

/* ---- K8: the job's one collective (src/main.py:186-222 keeps a row per run, agent and iteration, so gather, not reduce) ----
 * NCCL is bound at run time (dlopen of libnccl.so.2); a single-GPU caller never needs it.
 * agym_nccl_unique_id: rank 0 fills 128 bytes (ncclUniqueId) and ships them to the other ranks by its own means.
 * agym_comm_init: every rank, same id; creates this handle's communicator on its device (collective call).
 * agym_gather_metrics_nccl: all-gather of the bound accumulators [R][A][AGYM_NUM_METRICS] and revenue [R] into
 *   recv_acc [world][R][A][AGYM_NUM_METRICS] and recv_revenue [world][R] (device pointers) on `stream`; every rank holds
 *   the same R (pad the last shard). */
int agym_nccl_unique_id(char* out128);
int agym_comm_init(agym_handle* h, const char* id128, int32_t rank, int32_t world);
int agym_gather_metrics_nccl(agym_handle* h, double* recv_acc, double* recv_revenue, void* stream);
/* The same collective for a block the caller kept: `count` doubles at `send` (e.g. the metric blocks of all the iterations
 * of a job, [N][R][A][AGYM_NUM_METRICS] -- main.py:186-222 assembles its per-run rows once, when the runs have finished)
 * are all-gathered into recv [world][count] on `stream`. */
int agym_gather_block_nccl(agym_handle* h, const double* send, double* recv, int64_t count, void* stream);

/* Number of kernels of this library launched through this handle so far (bench.py's gpu_launches is a difference of
 * two readings; the reference has no counterpart: it launches nothing). */
uint64_t agym_launch_count(const agym_handle* h);

/* Kernel-selection overrides for tests and experiments (the library never reads the environment): "fit_warp" 0 = CTA
 * fit kernels only, "fit_dense" 0/1, "fit_nt" threads per CTA, "fit_ncap" rows staged per fit as a multiple of the mean,
 * "fit_heavy" whole-warp threshold of the CTA kernel, "sim_g" lane-group width of the round loop (4, 8, 16, 32), "sim_cat_smem" 0 = catalog read from global memory,
 * "bidfit_wide" 0/1.  They choose between implementations of the same function (the reference has one:
 * BidderAllocation.py:29-65, Auction.py:28-74, Bidder.py:210-615); unknown names are AGYM_ERR_INVALID. */
int agym_set_option(agym_handle* h, const char* name, double value);

/* ---- per-iteration model updates ---- */
enum agym_fit_mode {
  AGYM_FIT_ADAM_REF = 0,  /* IEEE divide / sqrt, accurate expf / logf: the operations torch's CPU kernels perform */
  AGYM_FIT_ADAM_FAST = 1, /* same state machine with MUFU approximations (~2 ulp); sparse regime only, else == REF */
  AGYM_FIT_NEWTON = 2     /* OPT-IN, a different algorithm: damped Newton per item to the optimum of the reference's objective
                           * (Models.py:39-41) instead of the reference's Adam trajectory, which stops short of it; the Laplace
                           * update and update_prior (Models.py:43-48) are the reference's.  obs_embedding_size 4 only.  Never a
                           * parity claim: max_epochs = objective evaluations per item (0 = 50); fit_info = {max passes over
                           * the items, total passes, objective, rows}. */
};
/* PyTorchLogisticRegressionAllocator.update for every (run, learnt agent) on the winner records of
 * this iteration (BidderAllocation.py:29-65, Models.py:35-48).  fit_info (device, nullable)
 * [R][A][4] float: {stop_epoch or -1, epochs run, final loss, rows}. */
int agym_update_allocators(agym_handle* h, int32_t fit_mode, int32_t max_epochs, float* fit_info, void* stream);

/* Per-(round, slot) bid records that feed the bidder fits (Agent.py:81-94: bidder.update sees ALL rows):
 * bid_rows [R][Tcap][P][AGYM_BID_ROW] float {estimated CTR, value, gamma, propensity, price},
 * bid_meta [R][Tcap][P] uint32 (bit 31 valid, bit 30 won, bit 29 click, bits 12..23 item, bits 0..11 agent).
 * Written by the round loop whenever a log is bound and some agent has a shaded bidder. */
#define AGYM_BID_ROW 5
int agym_bind_bid_log(agym_handle* h, float* bid_rows, uint32_t* bid_meta, int64_t Tcap);
size_t agym_bidder_workspace_bytes(const agym_handle* h, int64_t Tcap);
int agym_bind_bidder_workspace(agym_handle* h, void* ws, size_t bytes);
/* What bidder.update does for each agent (Bidder.py).  agym_set_agents derives VL_SEARCH / VL_POLICY from the bid kinds
 * AGYM_BID_SEARCH / AGYM_BID_POLICY; agents that bid with AGYM_BID_BANDIT must be told which bidder they are. */
enum agym_bidder_fit {
  AGYM_BFIT_NONE = 0,
  AGYM_BFIT_VL_SEARCH = 1,     /* ValueLearningBidder('search'): win-rate fit                      Bidder.py:210-260 */
  AGYM_BFIT_VL_POLICY = 2,     /* ValueLearningBidder('policy'): win-rate fit + policy fit         Bidder.py:278-316 */
  AGYM_BFIT_PL_REINFORCE = 3,  /* PolicyLearningBidder(loss=...)                                   Bidder.py:369-431 */
  AGYM_BFIT_PL_OFFPOLICY = 4,
  AGYM_BFIT_PL_TRPO = 5,
  AGYM_BFIT_PL_PPO = 6,
  AGYM_BFIT_DR = 7,            /* DoublyRobustBidder                                               Bidder.py:477-615 */
  AGYM_BFIT_EMPIRICAL = 8      /* EmpiricalShadedBidder: bucketised search for prev_gamma          Bidder.py:60-125  */
};
int agym_set_bidder_fits(agym_handle* h, const int32_t* fit_kind /* [A], host */);
/* Agent.update -> bidder.update for every (run, agent) whose bidder learns: win-rate fit, initialise_policy on the
 * first update, then the bidder's policy loss; sets the per-(run, agent) `initialised` flag (ValueLearning bidders that
 * won nothing clear it instead, Bidder.py:213-216).  (seed, iter) key the rsample noise of the stochastic losses.
 * fit_info (device, nullable) [R][A][3][4] float: per stage {win-rate, initialise_policy, policy} the tuple
 * {stop_epoch or -1, epochs run, final loss (NaN if any loss was NaN, cf. Bidder.py:412-419), rows}. */
int agym_update_bidders(agym_handle* h, uint64_t seed, int32_t iter, int32_t max_epochs, float* fit_info, void* stream);

/* ---- staged kernels (intermediates in HBM; used for roofline evidence and isolation tests) ---- */
/* K1  Auction.py:33,42: contexts [N][D] float and participants [N][P] uint8 for N = n_runs*T opportunities */
int agym_k1_contexts(agym_handle* h, uint64_t seed, int32_t iter, int64_t T, float* ctx, uint8_t* parts, void* stream);
/* K2  Agent.py:29-42 + Auction.py:52-53: per participant item (uint8), est, true_ctr, best_ev, value (float) */
int agym_k2_allocate(agym_handle* h, uint64_t seed, int32_t iter, int64_t T, const float* ctx, const uint8_t* parts,
                     uint8_t* item, float* est, float* true_ctr, float* best_ev, float* value, void* stream);
/* K3  Bidder.py:34-35,171-179: bids (float) and, for shaded bidders, gamma / propensity (nullable) */
int agym_k3_bids(agym_handle* h, uint64_t seed, int32_t iter, int64_t T, const uint8_t* parts, const float* est,
                 const float* value, float* bid, float* gamma, float* propensity, void* stream);
/* K4+K5  AuctionAllocation.py:18-35 + Auction.py:65-74 + Agent.py:70-77: segmented top-2 over the P bids
 * of each opportunity, lowest-slot tie-break, FP/SP price, Bernoulli click, accumulation.
 * Reads bid, true_ctr, value [N][P] float + parts [N][P] uint8; writes winner [N] uint8, price, second [N]
 * float, outcome [N] uint8.  `accumulate` != 0 also adds the winner-side metrics and revenue. */
int agym_k4_resolve(agym_handle* h, uint64_t seed, int32_t iter, int64_t T, const float* bid, const float* true_ctr,
                    const float* value, const uint8_t* parts, uint8_t* winner, float* price, float* second,
                    uint8_t* outcome, int32_t accumulate, void* stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif

#ifdef __cplusplus
}
#endif
#endif /* AGYM_H */
