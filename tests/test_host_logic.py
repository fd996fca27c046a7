"""CPU: host-side mirror of the reference interface -- config parsing, class table, mechanisms, run sharding,
CSV schema.  No device work."""
import json
import os
import sys

import numpy as np
import pytest

import auction_gym_b200 as ag
from auction_gym_b200 import driver
from oracle import auction_oracle as ao
from tests.conftest import ROOT


def test_parse_config_reproduces_the_reference_catalog():
    # known answers: the unmodified reference's parse_config on its own config/SP_Oracle.json, seed 0 (numpy 2.3 PCG64)
    rng, config, agent_configs, E, V, num_runs, max_slots, D, var, Do = ag.parse_config(os.path.join(ROOT, "config", "SP_Oracle.json"))
    assert (num_runs, max_slots, D, var, Do) == (3, 1, 5, 1.0, 4)
    assert [ac["name"] for ac in agent_configs] == [f"Truthful Oracle {k}" for k in range(1, 7)]
    Es, Vs = np.stack([E[k] for k in E]), np.stack([V[k] for k in V])
    assert Es.shape == (6, 12, 6) and Vs.shape == (6, 12)
    assert float(Es.sum()) == pytest.approx(-265.7772298549855, rel=1e-14)
    assert float(Vs.sum()) == pytest.approx(80.7858057403378, rel=1e-14)
    assert float(Es[0, 0, 0]) == 0.1257302210933933 and float(Es[5, 11, 5]) == -3.652348965689021 and float(Vs[3, 7]) == 1.6883650018900054
    _, _, ac, E, V, *_ = ag.parse_config(os.path.join(ROOT, "config", "FP_DR_TS.json"))
    assert list(E) == ["DR 1", "DR 2", "DR 3"]
    assert float(np.stack(list(E.values())).sum()) == pytest.approx(-122.00579069546444, rel=1e-14)


@pytest.mark.parametrize("cfg", ["SP_Oracle", "SP_Truthful_TS", "FP_DM_Oracle", "FP_DM_TS", "FP_DR_TS", "FP_IPS_TS"])
def test_every_shipped_config_instantiates(cfg):
    rng, config, agent_configs, E, V, *_ = ag.parse_config(os.path.join(ROOT, "config", cfg + ".json"))
    agents = ag.instantiate_agents(rng, agent_configs, V, E)
    assert len(agents) == len(agent_configs)
    a0 = agents[0]
    assert type(a0.allocator).__name__ == agent_configs[0]["allocator"]["type"]
    assert type(a0.bidder).__name__ == agent_configs[0]["bidder"]["type"]
    assert a0.bidder.truthful == (agent_configs[0]["bidder"]["type"] == "TruthfulBidder")
    if cfg == "FP_DM_Oracle":
        assert a0.bidder.inference == "search" and a0.bidder.kind == ag._lib.BID_SEARCH and a0.bidder.gamma_sigma == 0.02
        assert a0.allocator.item_embeddings is E[a0.name]
    if cfg == "FP_IPS_TS":
        assert a0.bidder.loss == "PPO" and a0.allocator.thompson_sampling and a0.allocator.response_model.m.shape == (12, 5)
    if cfg == "FP_DM_TS":
        assert a0.bidder.kind == ag._lib.BID_POLICY


def test_kwargs_decoding_and_unknown_types():
    assert driver.parse_kwargs({"inference": "\"search\"", "gamma_sigma": 0.02, "flag": True}) == {"inference": "search", "gamma_sigma": 0.02, "flag": True}
    with pytest.raises(ValueError, match="unknown type"):
        driver.build("NoSuchBidder", None, {})
    assert ag.LogisticTSAllocator is ag.PyTorchLogisticRegressionAllocator  # stale name in src/main.py:16 / BASELINE.json


def test_mechanisms_match_the_oracle_rule():
    rng = np.random.default_rng(0)
    for P in (1, 2, 3, 7):
        bids = rng.random((50, P))
        bids[::5] = np.round(bids[::5], 1)  # force ties
        for mech, code in ((ag.FirstPrice(), ao.MECH_FIRST), (ag.SecondPrice(), ao.MECH_SECOND)):
            assert mech.code == code
            w, price, second, valid = ao.resolve(bids, code)
            for t in range(len(bids)):
                winners, prices, seconds = mech.allocate(bids[t], 1)
                assert winners[0] == w[t]
                if P == 1:
                    assert (len(prices) == 0) if code == ao.MECH_SECOND else (len(seconds) == 0)  # AuctionAllocation.py:22,34
                else:
                    assert prices[0] == price[t] and seconds[0] == second[t]


def test_shard_runs_is_a_partition():
    for R in (1, 3, 8, 10, 4096):
        for world in (1, 2, 3, 8):
            spans = [ag.shard_runs(R, world, r) for r in range(world)]
            assert sum(c for _, c in spans) == R
            nxt = 0
            for first, count in spans:
                assert first == nxt and count >= 0
                nxt += count
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def test_csv_schema(tmp_path):
    import pandas as pd

    R, N, A = 2, 3, 3
    rng = np.random.default_rng(0)
    cfg = json.load(open(os.path.join(ROOT, "config", "FP_DR_TS.json")))
    cfg.update(num_runs=R, num_iter=N)
    names = ["DR 1", "DR 2", "DR 3"]
    result = {"config": cfg, "agent_names": names, "metrics": rng.random((R, N, A, len(driver.MEASURES))), "revenue": rng.random((R, N)),
              "truthful": [False] * A}
    out = driver.write_csvs(result, str(tmp_path))
    sfx = "10000_rounds_3_iters_2_runs_4_emb_of_5"  # main.py:270 naming
    files = sorted(os.listdir(out))
    assert files == sorted(f"{p}_{sfx}.csv" for p in ("net_utility", "gross_utility", "overbid_regret", "underbid_regret", "results"))
    net = pd.read_csv(f"{out}/net_utility_{sfx}.csv")
    assert list(net.columns) == ["Run", "Agent", "Iteration", "Net Utility"]
    assert net[["Agent", "Run", "Iteration"]].values.tolist() == sorted(net[["Agent", "Run", "Iteration"]].values.tolist())  # main.py:270
    ob = pd.read_csv(f"{out}/overbid_regret_{sfx}.csv")
    assert list(ob.columns) == ["Run", "Agent", "Iteration", "Overbid Regret"]
    assert ob["Run"].tolist() == sorted(ob["Run"].tolist())  # unsorted in the reference = run-major insertion order (main.py:228-237)
    res = pd.read_csv(f"{out}/results_{sfx}.csv")
    assert list(res.columns) == ["Run", "Iteration", "Measure", "Measure Name"]
    assert res["Measure Name"].unique().tolist() == ["Auction Revenue", "Social Surplus", "Social Welfare"]
    surplus = res[res["Measure Name"] == "Social Surplus"].sort_values(["Run", "Iteration"])["Measure"].values
    np.testing.assert_allclose(surplus, result["metrics"][..., 0].sum(axis=2).ravel())


def test_bare_name_modules_like_the_reference():
    sys.path.insert(0, os.path.join(ROOT, "auction_gym_b200", "src"))
    try:
        import importlib

        for mod, names in (("Agent", ["Agent"]), ("Auction", ["Auction"]), ("AuctionAllocation", ["FirstPrice", "SecondPrice"]),
                           ("Bidder", ["TruthfulBidder", "ValueLearningBidder", "PolicyLearningBidder", "DoublyRobustBidder", "EmpiricalShadedBidder"]),
                           ("BidderAllocation", ["OracleAllocator", "PyTorchLogisticRegressionAllocator"]),
                           ("Impression", ["ImpressionOpportunity"]), ("Models", ["sigmoid"]),
                           ("main", ["parse_config", "instantiate_agents", "instantiate_auction"])):
            m = importlib.import_module(mod)
            for n in names:
                assert hasattr(m, n), (mod, n)
    finally:
        sys.path.pop(0)
        for mod in ("Agent", "Auction", "AuctionAllocation", "Bidder", "BidderAllocation", "Impression", "Models", "main"):
            sys.modules.pop(mod, None)


def test_philox_restatement_known_answers():
    """Random123's published known-answer vectors for philox4x32-10 (kat_vectors): the numpy restatement of the engine's
    counter-based generator (csrc/agym_common.cuh) reproduces them; the device side is compared with the restatement in
    tests/test_gpu_production.py."""
    from oracle import philox_oracle as ph

    z, f = np.uint32(0), np.uint32(0xFFFFFFFF)
    assert [int(x) for x in ph.philox4x32_10(z, z, z, z, (z, z))] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert [int(x) for x in ph.philox4x32_10(f, f, f, f, (f, f))] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    c = [np.uint32(v) for v in (0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344)]
    k = (np.uint32(0xA4093822), np.uint32(0x299F31D0))
    assert [int(x) for x in ph.philox4x32_10(*c, k)] == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm: the unmodified reference from oracle/_ref on the host cores, or the numpy port
    when that copy is absent) needs no GPU; its single JSON line must carry the keys the driver reads, on the same metric /
    unit / workload / iterations as the GPU arm."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "auction opportunities/sec" and d["unit"] == "opportunities/s"
    assert d["higher_is_better"] is True and d["scaling"] == "strong" and d["vs_baseline"] is None and d["n_gpus"] == 1
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0
    have_ref = os.path.isfile(os.path.join(root, "oracle", "_ref", "src", "Auction.py")) or os.path.isfile("/root/reference/src/Auction.py")
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["config"]["iterations"] == [0, 1] and d["config"]["runs"] == 4096
    # iteration 0 of the trajectory: m ~ N(0, 1), q = 1 -- the long fits (the GPU arm's fit_epochs_mean of iteration 0 is ~8 100)
    assert 5000 < d["fit_epochs_mean"] < 12000 and d["cpu_baseline"]["fits_timed"] >= 8
    if have_ref:
        assert d["port"]["kind"] == "port" and abs(d["port"]["fit_epochs"] / d["fit_epochs_mean"] - 1) < 0.05
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_bench_reads_its_profile_numbers_from_the_committed_captures(tmp_path, monkeypatch):
    """bench.py prints no constants: the `from_profile` counters come from profiles/r*_ncu_full.txt at run time, from the newest
    capture of the kernel -- of the SAME GRID when there is one (only then is `traffic` reported) -- with ncu's units honoured."""
    import bench

    prof = tmp_path / "profiles"
    prof.mkdir()
    (prof / "r9_x_ncu_full.txt").write_text(
        "-" * 20 + "\n"
        "Kernel Name    void k4_kernel_p2_acc<2, 5>(SimParams)\n"
        "launch__grid_size    2560\n"
        "launch__registers_per_thread    48 register/thread\n"
        "dram__bytes_read.sum    100.5 Mbyte\n"
        "dram__bytes_write.sum    50 Mbyte\n"
        "smsp__issue_active.avg.pct_of_peak_sustained_active    58.6 %\n"
        + "-" * 20 + "\n"
        "Kernel Name    void k4_kernel_p2_acc<2, 5>(SimParams)\n"
        "launch__grid_size    20480\n"
        "dram__bytes_read.sum    1.09 Gbyte\n"
        "dram__bytes_write.sum    433 Mbyte\n")
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    small = bench.profile_numbers(r"k4_kernel_p2_acc", 2560)
    assert small["traffic_matches_this_grid"] and abs(small["dram_bytes"] - 150.5e6) < 1 and small["registers_per_thread"] == 48
    assert small["issue_slots_active_pct"] == 58.6 and small["source"].endswith("r9_x_ncu_full.txt")
    big = bench.profile_numbers(r"k4_kernel_p2_acc", 20480)
    assert big["traffic_matches_this_grid"] and abs(big["dram_bytes"] - 1.523e9) < 1e3
    other = bench.profile_numbers(r"k4_kernel_p2_acc", 777)  # no capture of that grid: the newest one, flagged
    assert other["grid"] == 20480 and not other["traffic_matches_this_grid"]
    assert bench.profile_numbers(r"no_such_kernel", None) is None
