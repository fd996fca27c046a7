"""GPU: the reference-named Python surface (Auction / Agent / parse_config / instantiate_* / CSV driver) on the engine."""
import json
import os

import numpy as np
import pytest

from tests.conftest import ROOT

pytestmark = pytest.mark.gpu


def _need_gpu():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def _small_config(tmp_path, name, **over):
    cfg = json.load(open(os.path.join(ROOT, "config", name + ".json")))
    cfg.update(over)
    cfg["output_dir"] = str(tmp_path / "out") + "/"
    p = tmp_path / f"{name}.json"
    json.dump(cfg, open(p, "w"))
    return str(p)


def test_one_round_at_a_time_like_the_notebooks(tmp_path):
    _need_gpu()
    import auction_gym_b200 as ag
    from auction_gym_b200 import _lib

    path = _small_config(tmp_path, "SP_Oracle")
    rng, config, agent_configs, E, V, num_runs, max_slots, D, var, Do = ag.parse_config(path)
    agents = ag.instantiate_agents(rng, agent_configs, V, E)
    auction, num_iter, rounds_per_iter, output_dir = ag.instantiate_auction(rng, config, E, V, agents, max_slots, D, var, Do, precision=_lib.FP64)
    for _ in range(40):
        auction.simulate_opportunity()  # notebook cell 4 / main.py:117
    logs = [ag_.logs for ag_ in auction.agents]
    assert sum(len(l) for l in logs) == 40 * config["num_participants_per_round"]
    won = [o for l in logs for o in l if o.won]
    assert len(won) == 40
    assert auction.revenue == pytest.approx(sum(o.price for o in won), rel=1e-12)
    for agent in auction.agents:
        l = agent.logs
        assert agent.net_utility == pytest.approx(sum(o.value * o.outcome - o.price for o in l if o.won), rel=1e-12, abs=1e-12)
        assert agent.gross_utility == pytest.approx(sum(o.value * o.outcome for o in l if o.won), rel=1e-12, abs=1e-12)
        assert agent.get_allocation_regret() == pytest.approx(sum(o.best_expected_value - o.true_CTR * o.value for o in l), abs=1e-12)
        assert agent.get_overbid_regret() == 0.0  # second price
        for o in l:
            assert o.context.shape == (D + 1,) and o.context[-1] == 1.0   # Oracle agents see the true context
            assert o.estimated_CTR == pytest.approx(o.true_CTR, rel=1e-12) and o.bid == pytest.approx(o.value * o.estimated_CTR, rel=1e-12)
            # the chosen item is the arg-max of true CTR x value for this agent (Agent.py:33-35)
            ctr = 1 / (1 + np.exp(-(E[agent.name] @ o.context)))
            assert o.item == int(np.argmax(ctr * V[agent.name])) and o.best_expected_value == pytest.approx(np.max(ctr * V[agent.name]), rel=1e-12)
        agent.update(iteration=0)
        agent.clear_utility()
        agent.clear_logs()
        assert agent.net_utility == 0.0 and agent.logs == []
    auction.clear_revenue()
    assert auction.revenue == 0.0 and auction.iteration == 1
    auction.simulate_opportunity()
    assert sum(len(a.logs) for a in auction.agents) == config["num_participants_per_round"]
    auction.engine.close()


def test_batched_experiment_writes_the_reference_csvs(tmp_path):
    _need_gpu()
    import pandas as pd

    import auction_gym_b200 as ag
    from oracle import auction_oracle as ao

    path = _small_config(tmp_path, "SP_Oracle", num_runs=16, num_iter=3, rounds_per_iter=2000)
    result = ag.run_experiment(path)
    assert result["metrics"].shape == (16, 3, 6, 10) and result["revenue"].shape == (16, 3)
    out = ag.write_csvs(result)
    res = pd.read_csv(os.path.join(out, "results_2000_rounds_3_iters_16_runs_4_emb_of_5.csv"))
    rev = res[res["Measure Name"] == "Auction Revenue"]["Measure"].values
    np.testing.assert_allclose(np.sort(rev), np.sort(result["revenue"].ravel()))
    # distribution-level check against the oracle with numpy-drawn noise on the same catalog
    rng, config, agent_configs, E, V, *_ = ag.parse_config(path)
    names = [ac["name"] for ac in agent_configs]
    case = {"A": 6, "I": 12, "D": 5, "Do": 4, "P": 2, "mechanism": ao.MECH_SECOND, "embedding_var": 1.0, "n_items": np.full(6, 12),
            "E": np.stack([E[n] for n in names]), "V": np.stack([V[n] for n in names]), "alloc_kind": np.zeros(6, int),
            "bidder_kind": np.zeros(6, int), "bidder_f": np.zeros((6, 4))}
    r2 = np.random.default_rng(5)
    ref_rev = []
    for _ in range(24):
        nz = ao.draw_replay_inputs(r2, 2000, 6, 2, 5)
        ref_rev.append(ao.simulate_rounds(case, nz["ctx"], nz["parts"], nz["u"])[1]["revenue"])
    ref_rev, got = np.asarray(ref_rev), result["revenue"].ravel()
    se = np.sqrt(ref_rev.var(ddof=1) / len(ref_rev) + got.var(ddof=1) / len(got))
    assert abs(ref_rev.mean() - got.mean()) < 5 * se
    # second price + truthful: no overbid / underbid regret
    assert np.all(result["metrics"][..., 4] == 0) and np.all(result["metrics"][..., 5] == 0)


def test_learning_config_improves_over_iterations(tmp_path):
    _need_gpu()
    import auction_gym_b200 as ag

    path = _small_config(tmp_path, "SP_Truthful_TS", num_runs=8, num_iter=4, rounds_per_iter=4000)
    result = ag.run_experiment(path)
    m = result["metrics"]  # [R, N, A, 10]
    rmse = m[..., 6].mean(axis=(0, 2))
    welfare = m[..., 1].sum(axis=2).mean(axis=0)
    assert rmse[-1] < 0.5 * rmse[0], rmse         # the CTR model learns (reference: CTR RMSE curve falls)
    assert welfare[-1] > welfare[0], welfare       # welfare rises as allocation improves
    assert np.isfinite(m[..., :7]).all()


@pytest.mark.parametrize("over", [{}, {"memory": 3000}, {"max_slots": 2}])
def test_opt_in_newton_mode_through_the_driver(tmp_path, over):
    """`run_experiment(fit_mode="newton")` / `main.py --fit-mode newton`: the opt-in regularised Newton allocator fit (NOT the
    reference's algorithm, csrc/agym_fit_newton.cu) learns on the shipped config -- also on retained logs (`memory`) and with
    several slots per round, where its rows are simply more rows -- and the default stays the reference's Adam trajectory."""
    _need_gpu()
    import auction_gym_b200 as ag

    cfg = json.load(open(os.path.join(ROOT, "config", "SP_Truthful_TS.json")))
    if "memory" in over:
        for ac in cfg["agents"]:
            ac["memory"] = over["memory"]
    if "max_slots" in over:
        cfg["max_slots"] = over["max_slots"]
    cfg.update(num_runs=8, num_iter=4, rounds_per_iter=3000, output_dir=str(tmp_path) + "/o/")
    path = str(tmp_path / "cfg.json")
    json.dump(cfg, open(path, "w"))
    nw = ag.run_experiment(path, fit_mode="newton")["metrics"]
    assert np.isfinite(nw[..., :7]).all()
    rmse = nw[..., 6].mean(axis=(0, 2))
    welfare = nw[..., 1].sum(axis=2).mean(axis=0)
    assert rmse[-1] < 0.8 * rmse[0] and welfare[-1] > 1.5 * welfare[0], (rmse, welfare)  # with `memory` the RMSE getter also covers kept records
    if not over:
        ref = ag.run_experiment(path)["metrics"]  # default fit mode: same first iteration (no fit yet), a different trajectory after it
        np.testing.assert_allclose(nw[:, 0], ref[:, 0], rtol=1e-12, atol=1e-12)
        assert not np.allclose(nw[:, 2], ref[:, 2])


def test_memory_config_retains_logs_across_iterations(tmp_path):
    """`memory` in an agent's config (main.py:87, Agent.py:124-129): the getters sum over kept + new records, so the
    allocation regret reported per iteration (a sum of non-negative terms) covers several iterations' rounds, and the
    allocator is fitted on more rows."""
    _need_gpu()
    import auction_gym_b200 as ag

    res = {}
    for mem in (0, 100000):
        cfg = json.load(open(os.path.join(ROOT, "config", "SP_Truthful_TS.json")))
        if mem:
            for ac in cfg["agents"]:
                ac["memory"] = mem
        cfg.update(num_runs=4, num_iter=4, rounds_per_iter=1500, output_dir=str(tmp_path) + f"/m{mem}/")
        path = str(tmp_path / f"mem{mem}.json")
        json.dump(cfg, open(path, "w"))
        res[mem] = ag.run_experiment(path)["metrics"]  # [R, N, A, 10]
        assert np.isfinite(res[mem][..., :7]).all()
    reg0, reg1 = res[0][..., 2].mean(axis=(0, 2)), res[100000][..., 2].mean(axis=(0, 2))  # allocation regret per iteration
    assert abs(reg1[0] / reg0[0] - 1) < 0.2, (reg0, reg1)   # first iteration: nothing retained yet
    assert reg1[-1] > 2.0 * reg0[-1], (reg0, reg1)          # fourth iteration: sums over four iterations' records
    # utilities restart every iteration regardless of memory (Agent.py:120-122): same scale in both runs (the run with
    # memory fits its allocators on more rows, so its welfare may be higher, not four times higher)
    w0, w1 = res[0][..., 1].sum(axis=2).mean(axis=0), res[100000][..., 1].sum(axis=2).mean(axis=0)
    assert np.all(w1 / w0 > 0.8) and np.all(w1 / w0 < 1.8), (w0, w1)


def test_empirical_shaded_bidder_config_runs(tmp_path):
    """EmpiricalShadedBidder (Bidder.py:38-153) end to end: gamma is clipped to [0, 1] and prev_gamma moves after updates."""
    _need_gpu()
    import auction_gym_b200 as ag

    cfg = json.load(open(os.path.join(ROOT, "config", "FP_DM_Oracle.json")))
    cfg["agents"][0]["bidder"] = {"type": "EmpiricalShadedBidder", "kwargs": {"gamma_sigma": 0.1, "init_gamma": 0.9}}
    cfg.update(num_runs=4, num_iter=3, rounds_per_iter=3000, output_dir=str(tmp_path) + "/")
    path = str(tmp_path / "emp.json")
    json.dump(cfg, open(path, "w"))
    result = ag.run_experiment(path)
    gamma = result["metrics"][..., 9]  # [R, N, A]
    assert np.isfinite(result["metrics"][..., :7]).all() and (gamma >= 0).all() and (gamma <= 1).all()
    assert abs(gamma[:, 0].mean() - 0.9) < 0.02              # clipped N(0.9, 0.1) has a mean slightly below 0.9
    assert np.abs(gamma[:, 1:] - 0.9).mean() > 0.01          # prev_gamma moved away from its initial value


@pytest.mark.parametrize("cfg", ["FP_IPS_TS", "FP_DR_TS", "FP_DM_TS"])
def test_policy_learning_configs_run_and_shade(tmp_path, cfg):
    """config/FP_IPS_TS.json, FP_DR_TS.json, FP_DM_TS.json (BASELINE.json configs[3]): TS allocation + a learnt Gaussian
    shading policy.  After the first update the policy drives the bids: gammas stay in [0, 1], every metric stays finite,
    and the bidders no longer bid their full value."""
    _need_gpu()
    import auction_gym_b200 as ag

    path = _small_config(tmp_path, cfg, num_runs=4, num_iter=3, rounds_per_iter=2000)
    result = ag.run_experiment(path)
    m = result["metrics"]  # [R, N, A, 10]
    assert np.isfinite(m[..., :7]).all() and np.isfinite(m[..., 9]).all()
    gamma = m[..., 9].mean(axis=(0, 2))
    assert abs(gamma[0] - 1.0) < 0.01, gamma               # iteration 0: gamma ~ N(1, 0.02)
    assert 0.0 <= gamma[1] <= 1.0 and 0.0 <= gamma[2] <= 1.0, gamma
    assert gamma[2] < 0.999, gamma                           # a fitted policy samples below 1 and is clipped at 1


def test_first_price_value_learning_config_learns_to_shade(tmp_path):
    """config/FP_DM_Oracle.json (BASELINE.json configs[2]): after the first win-rate fit the bidders search the gamma grid
    and shade their bids; first-price revenue drops and bidder surplus rises, as in the reference's figures."""
    _need_gpu()
    import auction_gym_b200 as ag

    path = _small_config(tmp_path, "FP_DM_Oracle", num_runs=8, num_iter=3, rounds_per_iter=3000)
    result = ag.run_experiment(path)
    m = result["metrics"]  # [R, N, A, 10]
    gamma = m[..., 9].mean(axis=(0, 2))
    surplus = m[..., 0].sum(axis=2).mean(axis=0)
    revenue = result["revenue"].mean(axis=0)
    assert abs(gamma[0] - 1.0) < 0.01, gamma            # iteration 0: gamma ~ N(1, 0.02)  (Bidder.py:177)
    assert gamma[1] < 0.97 and gamma[2] < 0.97, gamma   # afterwards: searched gammas in [0.1, 1]
    assert surplus[1] > surplus[0] and revenue[1] < revenue[0], (surplus, revenue)
    assert np.isfinite(m).all()


def _same_metrics(got, want, err_msg=""):
    """Two executions of the same job: every draw, decision, logged row and fitted parameter is bit-identical, but the per-
    iteration metric sums are FP64 `red.global.add`s whose arrival order differs from launch to launch, so a sum may differ
    in its last bits (observed: 1 ulp in 1 of 720 values).  1e-12 is four orders below anything the CSVs print."""
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12, err_msg=err_msg)


@pytest.mark.parametrize("cfg", ["FP_DR_TS", "SP_Truthful_TS"])
def test_shards_reproduce_the_single_device_job(tmp_path, cfg, monkeypatch):
    """The same seed gives the same per-run metrics whichever rank owns a run (world = 1, 2 and num_runs): the Philox key AND
    every initial model draw (allocator m, win-rate and policy weights) are keyed by the global run index."""
    _need_gpu()
    from auction_gym_b200 import driver

    path = _small_config(tmp_path, cfg, num_runs=4, num_iter=2, rounds_per_iter=400)
    monkeypatch.setattr(driver, "gather_runs", lambda local, world: local)  # no process group here: shards are compared directly
    whole = driver.run_experiment(path, rank=0, world=1)
    for world in (2, 4):
        parts = [driver.run_experiment(path, rank=r, world=world) for r in range(world)]
        _same_metrics(np.concatenate([p["metrics"] for p in parts]), whole["metrics"], err_msg=f"{cfg} world {world}")
        _same_metrics(np.concatenate([p["revenue"] for p in parts]), whole["revenue"], err_msg=f"{cfg} world {world}")


def _surface_from_case(case):
    """Reference-named objects (Agent / allocator / bidder / Auction) over a golden case's catalog and learnt state."""
    import auction_gym_b200 as ag
    from auction_gym_b200 import _lib
    from oracle import auction_oracle as ao

    rng = np.random.default_rng(0)
    A, I, D, Do = int(case["A"]), int(case["I"]), int(case["D"]), int(case["Do"])
    agents, a2i, a2v = [], {}, {}
    for a in range(A):
        n = int(case["n_items"][a])
        name = f"agent {a}"
        a2i[name], a2v[name] = case["E"][a, :n], case["V"][a, :n]
        alloc = ag.OracleAllocator(rng) if case["alloc_kind"][a] == ao.ALLOC_ORACLE else \
            ag.PyTorchLogisticRegressionAllocator(rng, Do, n, thompson_sampling=bool(case["alloc_kind"][a] == ao.ALLOC_TS))
        agents.append(ag.Agent(rng, name, n, a2v[name], alloc, ag.TruthfulBidder(rng)))
        if isinstance(alloc, ag.OracleAllocator):
            alloc.update_item_embeddings(a2i[name])
    mech = ag.SecondPrice() if case["mechanism"] == ao.MECH_SECOND else ag.FirstPrice()
    auction = ag.Auction(rng, mech, agents, a2i, a2v, 1, D, float(case["embedding_var"]), Do, int(case["P"]), precision=_lib.FP64, seed=3)
    auction._build()
    if auction.engine.any_learnt:
        auction.engine.set_allocator_state(case["m"][None], case["q"][None])
    return auction


@pytest.mark.parametrize("name", ["rounds_sp_ts", "rounds_sp_oracle", "rounds_sp_ts_64x64"])
def test_per_context_methods_replay_the_reference(name):
    """Allocator.estimate_CTR / Agent.select_item / Agent.bid for ONE context (BidderAllocation.py:67-68,81-82; Agent.py:29-68),
    fed the golden rounds' contexts and Thompson noise: the chosen item is the reference's, the MAP estimate and the truthful
    bid match to float32 rounding."""
    _need_gpu()
    from tests.conftest import load_golden

    case, inp, ref, met = load_golden(name)
    auction = _surface_from_case(case)
    D, Do = int(case["D"]), int(case["Do"])
    T = min(len(inp["ctx"]), 48)
    for t in range(T):
        true_ctx = np.concatenate([inp["ctx"][t, :D], [1.0]])
        obs_ctx = np.concatenate([inp["ctx"][t, :Do], [1.0]])
        for s, a in enumerate(inp["parts"][t]):
            agent = auction.agents[int(a)]
            oracle = type(agent.allocator).__name__ == "OracleAllocator"
            ctx = true_ctx if oracle else obs_ctx
            eps = None if oracle else inp["ts_eps"][t, s]
            item, est = agent.select_item(ctx, _eps=eps)
            assert item == int(ref["item"][t, s]), (t, s)
            np.testing.assert_allclose(est, ref["est"][t, s], rtol=1e-12 if oracle else 2e-6, err_msg=str((t, s)))
            assert np.asarray(est).dtype == (np.float64 if oracle else np.float32)  # the reference's mix
            bid, item2 = agent.bid(ctx, _eps=eps)
            assert item2 == item
            np.testing.assert_allclose(bid, ref["bid"][t, s], rtol=1e-12 if oracle else 2e-6)
            if not oracle:  # MAP estimates are deterministic; sampled ones differ between two draws
                e1, e2 = agent.allocator.estimate_CTR(ctx, sample=False), agent.allocator.estimate_CTR(ctx, sample=False)
                assert np.array_equal(e1, e2) and e1.shape == (agent.num_items,)
                s1, s2 = agent.allocator.estimate_CTR(ctx), agent.allocator.estimate_CTR(ctx)
                assert bool(agent.allocator.thompson_sampling) == (not np.array_equal(s1, s2))
    auction.engine.close()


def test_per_context_shaded_bid_follows_the_bidder_state(tmp_path):
    """Bidder.bid for one opportunity (Bidder.py:171-179): before the first fit gamma ~ N(prev_gamma, gamma_sigma), bid = gamma x value x CTR."""
    _need_gpu()
    import auction_gym_b200 as ag

    path = _small_config(tmp_path, "FP_DM_Oracle")
    rng, config, agent_configs, E, V, num_runs, max_slots, D, var, Do = ag.parse_config(path)
    agents = ag.instantiate_agents(rng, agent_configs, V, E)
    auction, *_ = ag.instantiate_auction(rng, config, E, V, agents, max_slots, D, var, Do)
    shaded = [a for a in auction.agents if not a.bidder.truthful][0]
    gam = []
    for _ in range(64):
        b = shaded.bidder.bid(1.25, None, 0.2)
        gam.append(b / (1.25 * 0.2))
    gam = np.array(gam)
    assert abs(gam.mean() - shaded.bidder.prev_gamma) < 5 * shaded.bidder.gamma_sigma / 8 and 0.3 * shaded.bidder.gamma_sigma < gam.std() < 2 * shaded.bidder.gamma_sigma
    auction.engine.close()


@pytest.mark.parametrize("cfg,over", [("SP_Truthful_TS", {}), ("FP_DR_TS", {}), ("SP_Truthful_TS", {"memory": 300})])
def test_resume_from_checkpoint_equals_the_uninterrupted_job(tmp_path, cfg, over, monkeypatch):
    """Per-iteration checkpoint of the learnt state: stop after two of four iterations, resume in a
    fresh process-equivalent (new engine), and get the uninterrupted job's metrics bit for bit."""
    _need_gpu()
    from auction_gym_b200 import driver

    cfgd = json.load(open(os.path.join(ROOT, "config", cfg + ".json")))
    if "memory" in over:
        for a in cfgd["agents"]:
            a["memory"] = over["memory"]
    cfgd.update(num_runs=3, num_iter=4, rounds_per_iter=600, output_dir=str(tmp_path / "out") + "/")
    path = str(tmp_path / "cfg.json")
    json.dump(cfgd, open(path, "w"))
    ck0 = str(tmp_path / "ck_whole")
    whole = driver.run_experiment(path, checkpoint_dir=ck0)
    ck = str(tmp_path / "ck")
    part = driver.run_experiment(path, checkpoint_dir=ck, stop_after=2)
    _same_metrics(part["metrics"], whole["metrics"][:, :2])
    rest = driver.run_experiment(path, checkpoint_dir=ck, resume=True)
    _same_metrics(rest["metrics"], whole["metrics"])
    _same_metrics(rest["revenue"], whole["revenue"])
    # the learnt state (allocator m / q / prev_iter_m, bidder state, retained log rows) after the last iteration: bit for bit
    a, b = (dict(np.load(os.path.join(d, "state_rank0_of_1.npz"))) for d in (ck0, ck))
    assert set(a) == set(b)
    for k in ("shape", "iteration", "seed", "run_offset", "m", "q", "m_prev", "bidder_d", "bidder_w", "log_fit_meta", "log_bid_meta"):
        if k in a:
            np.testing.assert_array_equal(a[k], b[k], err_msg=k)
    if "log_fit_ctx" in a:  # retained winner rows (fields of rows nobody wrote are whatever the allocation held)
        valid = (a["log_fit_meta"].view(np.uint32) >> 31) == 1
        np.testing.assert_array_equal(a["log_fit_ctx"][valid], b["log_fit_ctx"][valid])


def test_metric_gather_through_the_abi_single_rank():
    """K8 through the C ABI (agym_nccl_unique_id / agym_comm_init / agym_gather_metrics_nccl) with a one-rank communicator:
    the gathered block is this rank's block.  (Two and more ranks: bench.py under torchrun, profiles/r2_bench_n2_line.json.)"""
    _need_gpu()
    import torch
    from tests.conftest import load_golden
    from tests import gpu_util as gu
    from auction_gym_b200 import _lib

    case, *_ = load_golden("rounds_sp_oracle")
    eng = gu.engine_from_case(case, R=3, precision=_lib.FP32)
    eng.simulate(1, 0, 500)
    eng.comm_init(0, 1)
    acc_g, rev_g = eng.gather_metrics()
    torch.cuda.synchronize()
    assert acc_g.shape == (1, 3, eng.A, _lib.NUM_METRICS) and torch.equal(acc_g[0], eng.acc) and torch.equal(rev_g[0], eng.revenue)
    # a block the caller kept (all the iterations of a job): agym_gather_block_nccl
    hist = torch.stack([eng.acc, 2 * eng.acc])
    g = eng.gather_block(hist)
    torch.cuda.synchronize()
    assert g.shape == (1, 2, 3, eng.A, _lib.NUM_METRICS) and torch.equal(g[0], hist)
    eng.close()
