"""Worker of the multi-rank GPU tests (launched by torch.distributed.run): every rank evaluates its shard through
mcportfolio.dist, rank 0 also runs the whole job alone and compares.  Backend 'gloo' puts all ranks on cuda:0 (the host-side merge
code; kernels of different ranks never wait on each other), 'nccl' gives every rank its own GPU and uses libmcp's communicator."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [HERE, ROOT, os.path.join(ROOT, "monte-carlo-portfolio_b200")]
import numpy as np
import torch
import torch.distributed as dist


def main():
    backend, out_path = sys.argv[1], sys.argv[2]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = int(os.environ["LOCAL_RANK"]) if backend == "nccl" else 0
    torch.cuda.set_device(dev)
    if backend == "nccl":
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    else:
        dist.init_process_group("gloo")
    import mcportfolio as mcp
    from mcportfolio import dist as mdist
    from conftest import synthetic_inputs
    mu, sigma = synthetic_inputs(16)
    P, M = 3_000_001, 200_003
    res = {}
    r = mdist.simulate_portfolios_sharded(mu, sigma, P, risk_free=0.03, seed=4, return_arrays=False, device=dev)
    res["picks"] = [r.max_sharpe["global_index"], r.target_risk["global_index"], r.max_sharpe["sharpe"], r.target_risk["risk"]]
    res["n_acc"] = r.extra["n_accepted_global"]
    w = r.max_sharpe["weights"]
    p = mdist.simulate_paths_sharded(mu, sigma, w, M, 32, seed=4, device=dev, return_terminal=False)
    res["stats"] = {str(a): list(v) for a, v in p["stats"].items()}
    mu2, sigma2 = synthetic_inputs(64)
    e = mdist.frontier_envelope_sharded(mu2, sigma2, 100_000, 32, risk_free=0.03, seed=4, device=dev)
    res["env_idx"] = e.extra["envelope"]["best_index"].tolist()
    res["env_ret"] = e.extra["envelope"]["best_return"].tolist()
    res["env_pick"] = e.target_risk["global_index"]
    # bounded run: skipped portfolios, merged counts
    lo, hi = np.full(16, 0.01), np.full(16, 0.2)
    b = mdist.simulate_portfolios_sharded(mu, sigma, 50_000, risk_free=0.03, seed=5, min_weights=lo, max_weights=hi, return_arrays=False, device=dev)
    res["bounded"] = [b.max_sharpe["global_index"], b.extra["n_accepted_global"]]
    gathered = [None] * world
    dist.all_gather_object(gathered, res)
    ok = all(g == gathered[0] for g in gathered)
    if rank == 0:
        a = mcp.simulate_portfolios(mu, sigma, P, risk_free=0.03, seed=4, return_arrays=False, device=dev)
        single = {"picks": [a.max_sharpe["global_index"], a.target_risk["global_index"], a.max_sharpe["sharpe"], a.target_risk["risk"]], "n_acc": a.n_accepted}
        pa = mcp.simulate_paths(mu, sigma, w, M, 32, seed=4, device=dev, return_terminal=False)
        single["stats"] = {str(al): list(v) for al, v in pa["stats"].items()}
        ea = mcp.frontier_envelope(mu2, sigma2, 100_000, 32, risk_free=0.03, seed=4, device=dev)
        single["env_idx"] = ea.extra["envelope"]["best_index"].tolist()
        single["env_ret"] = ea.extra["envelope"]["best_return"].tolist()
        single["env_pick"] = ea.target_risk["global_index"]
        ba = mcp.simulate_portfolios(mu, sigma, 50_000, risk_free=0.03, seed=5, min_weights=lo, max_weights=hi, return_arrays=False, device=dev)
        single["bounded"] = [ba.max_sharpe["global_index"], ba.n_accepted]
        with open(out_path, "w") as fh:
            json.dump({"ranks_agree": ok, "sharded": gathered[0], "single": single, "world": world, "backend": backend,
                       "comm": list(mcp.get_engine(dev).comm_info())}, fh)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
