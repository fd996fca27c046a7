"""Experiment driver: the reference's ``main.py`` (src/main.py) re-designed around the batched engine.

Kept verbatim from the reference: the JSON schema of ``config/*.json`` (CONFIG.md:9-33, including the
double-quoted string kwargs the reference needs for its ``eval``), the helper signatures
``parse_config`` / ``instantiate_agents`` / ``instantiate_auction`` with the same return tuples
(main.py:74,109), the catalog draw order (main.py:60-72: all embeddings, then all values, then all
intercepts, from one PCG64 generator seeded with ``random_seed``) and the five CSV files with their names
and columns (main.py:270-289,328-345).  Different by design: all ``num_runs`` runs execute at once on the
device (sharded by run across ranks when launched under torchrun), and the per-iteration metrics are read
from the engine's accumulators instead of Python lists of log records.
"""
from __future__ import annotations

import argparse
import ast
import json
import os
from copy import deepcopy

import numpy as np

from . import _lib
from .agent import Agent
from .allocators import Allocator, LogisticTSAllocator, OracleAllocator, PyTorchLogisticRegressionAllocator  # noqa: F401
from .auction import Auction
from .bidders import (Bidder, DoublyRobustBidder, EmpiricalShadedBidder, PolicyLearningBidder, TruthfulBidder,  # noqa: F401
                      ValueLearningBidder)
from .mechanisms import AllocationMechanism, FirstPrice, SecondPrice  # noqa: F401

# what the reference resolves with eval() on the strings in the config (main.py:85-86,100)
CLASS_TABLE = {c.__name__: c for c in (FirstPrice, SecondPrice, OracleAllocator, PyTorchLogisticRegressionAllocator,
                                       TruthfulBidder, EmpiricalShadedBidder, ValueLearningBidder, PolicyLearningBidder,
                                       DoublyRobustBidder)}
CLASS_TABLE["LogisticTSAllocator"] = PyTorchLogisticRegressionAllocator

MEASURES = ("Net Utility", "Gross Utility", "Allocation Regret", "Estimation Regret", "Overbid Regret", "Underbid Regret",
            "CTR RMSE", "CTR Bias", "Mean Expected Value for Top Ad", "Shading Factors")


def parse_kwargs(kwargs):
    """The reference builds the source text ``,k=v,...`` for eval (main.py:19-21); here the values are decoded
    instead: ``"\\"search\\""`` -> ``"search"``, numbers and booleans stay as they are."""
    out = {}
    for key, value in kwargs.items():
        if isinstance(value, str):
            try:
                value = ast.literal_eval(value)
            except (ValueError, SyntaxError):
                pass
        out[key] = value
    return out


def build(type_name, rng, kwargs=None):
    try:
        cls = CLASS_TABLE[type_name]
    except KeyError:
        raise ValueError(f"unknown type {type_name!r} in config (known: {sorted(CLASS_TABLE)})") from None
    if issubclass(cls, AllocationMechanism):
        return cls()
    return cls(rng=rng, **parse_kwargs(kwargs or {}))


def parse_config(path):
    """main.py:24-74 -- same return tuple, same catalog for the same seed."""
    with open(path) as f:
        config = json.load(f)
    rng = np.random.default_rng(config["random_seed"])
    np.random.seed(config["random_seed"])
    num_runs = config["num_runs"] if "num_runs" in config.keys() else 1
    max_slots = 1
    embedding_size = config["embedding_size"]
    embedding_var = config["embedding_var"]
    obs_embedding_size = config["obs_embedding_size"]
    agent_configs = []
    num_agents = 0
    for agent_config in config["agents"]:
        if "num_copies" in agent_config.keys():
            for _ in range(1, agent_config["num_copies"] + 1):
                copy = deepcopy(agent_config)
                copy["name"] += f" {num_agents + 1}"
                agent_configs.append(copy)
                num_agents += 1
        else:
            agent_configs.append(agent_config)
            num_agents += 1
    agents2items = {ac["name"]: rng.normal(0.0, embedding_var, size=(ac["num_items"], embedding_size)) for ac in agent_configs}
    agents2item_values = {ac["name"]: rng.lognormal(0.1, 0.2, ac["num_items"]) for ac in agent_configs}
    for agent, items in agents2items.items():
        agents2items[agent] = np.hstack((items, -3.0 - 1.0 * rng.random((items.shape[0], 1))))
    return rng, config, agent_configs, agents2items, agents2item_values, num_runs, max_slots, embedding_size, embedding_var, obs_embedding_size


def instantiate_agents(rng, agent_configs, agents2item_values, agents2items):
    """main.py:77-95."""
    agents = [Agent(rng=rng, name=ac["name"], num_items=ac["num_items"], item_values=agents2item_values[ac["name"]],
                    allocator=build(ac["allocator"]["type"], rng, ac["allocator"]["kwargs"]),
                    bidder=build(ac["bidder"]["type"], rng, ac["bidder"]["kwargs"]),
                    memory=(0 if "memory" not in ac.keys() else ac["memory"]))
              for ac in agent_configs]
    for agent in agents:
        if isinstance(agent.allocator, OracleAllocator):
            agent.allocator.update_item_embeddings(agents2items[agent.name])
    return agents


def instantiate_auction(rng, config, agents2items, agents2item_values, agents, max_slots, embedding_size, embedding_var,
                        obs_embedding_size, **engine_kwargs):
    """main.py:98-109 -- same 4-tuple; ``engine_kwargs`` (num_runs, run_offset, device, precision, seed) are new."""
    return (Auction(rng, build(config["allocation"], rng), agents, agents2items, agents2item_values, max_slots, embedding_size,
                    embedding_var, obs_embedding_size, config["num_participants_per_round"], **engine_kwargs),
            config["num_iter"], config["rounds_per_iter"], config["output_dir"])


# ------------------------------------------------------------------------------------------------
# batched run loop + sharding
# ------------------------------------------------------------------------------------------------
def shard_runs(num_runs, world, rank):
    """Contiguous split of the runs over ranks (runs are independent, main.py:186): (first run, count)."""
    base, extra = divmod(int(num_runs), int(world))
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def iteration_block(auction):
    """The ten per-agent numbers main.py:131-148 records after each iteration, [R, A, 10], plus revenue [R]."""
    acc = auction.engine.acc.cpu().numpy()
    rev = auction.engine.revenue.cpu().numpy().copy()
    with np.errstate(invalid="ignore", divide="ignore"):
        blk = np.stack([
            acc[..., _lib.M_NET], acc[..., _lib.M_GROSS], acc[..., _lib.M_ALLOC_REGRET], acc[..., _lib.M_ESTIM_REGRET],
            acc[..., _lib.M_OVERBID_REGRET], acc[..., _lib.M_UNDERBID_REGRET],
            np.sqrt(acc[..., _lib.M_SQERR] / acc[..., _lib.M_NPART]), acc[..., _lib.M_BIAS] / acc[..., _lib.M_NWON],
            acc[..., _lib.M_BEST_EV] / acc[..., _lib.M_NPART], acc[..., _lib.M_GAMMA] / acc[..., _lib.M_NPART]], axis=-1)
    return blk, rev


def simulation_run(auction, num_iter, rounds_per_iter, verbose=False, checkpoint=None, resume=False):
    """main.py:112-155 for every resident run at once.  Returns metrics [R, N, A, 10] and revenue [R, N].
    ``checkpoint`` (a path): the learnt state and the metrics so far are written after every iteration; ``resume`` continues
    from that file, and the result equals an uninterrupted job bit for bit (same seed, same Philox counters)."""
    blocks, revs = [], []
    first = 0
    if checkpoint and resume and os.path.isfile(checkpoint):
        d = auction.load_checkpoint(checkpoint)
        first = auction.iteration
        blocks = [d["metrics"][:, i] for i in range(first)]
        revs = [d["revenue_so_far"][:, i] for i in range(first)]
    for i in range(first, num_iter):
        auction.simulate_rounds(rounds_per_iter)
        auction._update_models()  # agent.update() of every agent (main.py:128-129)
        blk, rev = iteration_block(auction)
        if verbose:
            print(f"==== ITERATION {i} ====  mean revenue {rev.mean():.3f}  mean welfare {blk[..., 1].sum(axis=1).mean():.3f}")
        blocks.append(blk)
        revs.append(rev)
        auction.end_iteration()  # clear_utility / clear_logs / clear_revenue (main.py:151-155)
        if checkpoint:
            auction.save_checkpoint(checkpoint, metrics=np.stack(blocks, axis=1), revenue_so_far=np.stack(revs, axis=1))
    return np.stack(blocks, axis=1), np.stack(revs, axis=1)


def gather_runs(local, world):
    """All ranks' run blocks in rank order (torch.distributed all_gather; NCCL on GPUs, gloo in the CPU tests).
    The only collective of the job: a few MB of metrics at the end."""
    if world == 1:
        return local
    import torch
    import torch.distributed as dist

    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.as_tensor(local).to(dev)
    n = torch.tensor([t.shape[0]], device=dev)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    mx = int(max(c.item() for c in counts))
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
    pad[:t.shape[0]] = t
    parts = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return np.concatenate([p[:int(c.item())].cpu().numpy() for p, c in zip(parts, counts)], axis=0)


FIT_MODES = {"adam_ref": _lib.FIT_ADAM_REF, "adam_fast": _lib.FIT_ADAM_FAST, "newton": _lib.FIT_NEWTON}


def run_experiment(config_path, device=0, rank=0, world=1, precision=None, verbose=False, checkpoint_dir=None, resume=False, stop_after=None,
                   fit_mode="adam_ref"):
    """Parse the config, simulate this rank's share of the runs, gather.  Returns a dict with the reference's
    run -> agent -> per-iteration structure flattened into arrays (see ``write_csvs``)."""
    rng, config, agent_configs, agents2items, agents2item_values, num_runs, max_slots, embedding_size, embedding_var, \
        obs_embedding_size = parse_config(config_path)
    first, count = shard_runs(num_runs, world, rank)
    metrics = np.zeros((0, config["num_iter"], len(agent_configs), len(MEASURES)))
    revenue = np.zeros((0, config["num_iter"]))
    if count > 0:
        agents = instantiate_agents(rng, agent_configs, agents2item_values, agents2items)
        auction, num_iter, rounds_per_iter, output_dir = instantiate_auction(
            rng, config, agents2items, agents2item_values, agents, max_slots, embedding_size, embedding_var, obs_embedding_size,
            num_runs=count, run_offset=first, device=device, precision=precision, seed=config["random_seed"],
            rounds_capacity=config["rounds_per_iter"], per_run_init=True)
        auction.fit_mode = FIT_MODES[fit_mode]  # "newton": opt-in, NOT the reference's algorithm (csrc/agym_fit_newton.cu)
        ckpt = None
        if checkpoint_dir:
            os.makedirs(checkpoint_dir, exist_ok=True)
            ckpt = os.path.join(checkpoint_dir, f"state_rank{rank}_of_{world}.npz")
        metrics, revenue = simulation_run(auction, num_iter if stop_after is None else min(num_iter, stop_after), rounds_per_iter,
                                          verbose=verbose and rank == 0, checkpoint=ckpt, resume=resume)
        auction.engine.close()
    metrics = gather_runs(metrics, world)
    revenue = gather_runs(revenue, world)
    return {"config": config, "agent_names": [ac["name"] for ac in agent_configs], "metrics": metrics, "revenue": revenue,
            "truthful": [ac["bidder"]["type"] == "TruthfulBidder" for ac in agent_configs]}


# ------------------------------------------------------------------------------------------------
# CSV output (main.py:228-237,270-289,296-303,328-345)
# ------------------------------------------------------------------------------------------------
def file_suffix(config):
    return (f"{config['rounds_per_iter']}_rounds_{config['num_iter']}_iters_{config.get('num_runs', 1)}_runs_"
            f"{config['obs_embedding_size']}_emb_of_{config['embedding_size']}")


def measure_per_agent2df(result, measure_name):
    """main.py:228-237: rows ordered run -> agent -> iteration, columns [Run, Agent, Iteration, <measure>]."""
    import pandas as pd

    m = result["metrics"][..., MEASURES.index(measure_name)]  # [R, N, A]
    R, N, A = m.shape
    rows = {"Run": [], "Agent": [], "Iteration": [], measure_name: []}
    for run in range(R):
        for a, name in enumerate(result["agent_names"]):
            if measure_name == "Shading Factors" and result["truthful"][a]:
                continue  # main.py:144: truthful bidders log no gamma
            rows["Run"] += [run] * N
            rows["Agent"] += [name] * N
            rows["Iteration"] += list(range(N))
            rows[measure_name] += list(m[run, :, a])
    return pd.DataFrame(rows)


def write_csvs(result, output_dir=None):
    """The five CSV files of the reference, same names and columns."""
    import pandas as pd

    config = result["config"]
    output_dir = output_dir or config["output_dir"]
    os.makedirs(output_dir, exist_ok=True)
    sfx = file_suffix(config)
    net = measure_per_agent2df(result, "Net Utility").sort_values(["Agent", "Run", "Iteration"])
    net.to_csv(f"{output_dir}/net_utility_{sfx}.csv", index=False)
    gross = measure_per_agent2df(result, "Gross Utility").sort_values(["Agent", "Run", "Iteration"])
    gross.to_csv(f"{output_dir}/gross_utility_{sfx}.csv", index=False)
    measure_per_agent2df(result, "Overbid Regret").to_csv(f"{output_dir}/overbid_regret_{sfx}.csv", index=False)
    measure_per_agent2df(result, "Underbid Regret").to_csv(f"{output_dir}/underbid_regret_{sfx}.csv", index=False)
    R, N = result["revenue"].shape
    rev = pd.DataFrame({"Run": np.repeat(np.arange(R), N), "Iteration": np.tile(np.arange(N), R), "Measure": result["revenue"].ravel()})
    surplus = net.groupby(["Run", "Iteration"])["Net Utility"].sum().reset_index().rename(columns={"Net Utility": "Measure"})
    welfare = gross.groupby(["Run", "Iteration"])["Gross Utility"].sum().reset_index().rename(columns={"Gross Utility": "Measure"})
    rev["Measure Name"], surplus["Measure Name"], welfare["Measure Name"] = "Auction Revenue", "Social Surplus", "Social Welfare"
    pd.concat((rev, surplus, welfare)).to_csv(f"{output_dir}/results_{sfx}.csv", index=False)
    return output_dir


def main(argv=None):
    """``python main.py <config.json>`` (main.py:157-161); under torchrun the runs are sharded over the ranks."""
    parser = argparse.ArgumentParser()
    parser.add_argument("config", type=str, help="Path to experiment configuration file")
    parser.add_argument("--output-dir", default=None)
    parser.add_argument("--precision", default="fp32", choices=["fp32", "fp64"])
    parser.add_argument("--checkpoint-dir", default=None, help="write the learnt state after every iteration (one file per rank)")
    parser.add_argument("--resume", action="store_true", help="continue from --checkpoint-dir")
    parser.add_argument("--fit-mode", default="adam_ref", choices=sorted(FIT_MODES),
                        help="allocator fit: adam_ref = the reference's Adam trajectory (default); newton = opt-in regularised Newton solve, a different algorithm")
    args = parser.parse_args(argv)
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    result = run_experiment(args.config, device=local, rank=rank, world=world,
                            precision=_lib.FP64 if args.precision == "fp64" else _lib.FP32, verbose=True,
                            checkpoint_dir=args.checkpoint_dir, resume=args.resume, fit_mode=args.fit_mode)
    if rank == 0:
        out = write_csvs(result, args.output_dir)
        print(f"wrote CSVs to {out}")
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


if __name__ == "__main__":
    main()
