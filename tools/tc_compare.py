"""Development check: tcgen05 large-N sweep vs the SIMT kernel (MCP_LARGE_TC=0) on the same seeds.
usage: python tools/tc_compare.py [n] [P]   (spawns itself twice; prints rates and the result differences)"""
import json, os, subprocess, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "monte-carlo-portfolio_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))


def child(n, P):
    import numpy as np
    import mcportfolio as mcp
    from conftest import synthetic_inputs
    mu, sigma = synthetic_inputs(n, seed=1)
    out = {}
    small = mcp.simulate_portfolios(mu, sigma, 5000, risk_free=0.03, seed=5, first_index=123456789012, dtype="float32")
    out["risks"] = np.asarray(small.risks, dtype=np.float64).tolist()
    out["returns"] = np.asarray(small.returns, dtype=np.float64).tolist()
    out["wsum"] = float(np.abs(small.weights.sum(1) - 1).max())
    out["w0"] = np.asarray(small.weights[0][:8], dtype=np.float64).tolist()
    out["pick"] = [int(small.max_sharpe["index"]), float(small.max_sharpe["sharpe"]), int(small.target_risk["index"])]
    out["consistent"] = bool(small.max_sharpe["sharpe"] == float(small.sharpes[small.max_sharpe["index"]])
                             and np.array_equal(small.max_sharpe["weights"], small.weights[small.max_sharpe["index"]]))
    for _ in range(2):
        mcp.simulate_portfolios(mu, sigma, P // 10, risk_free=0.03, seed=1, dtype="float32", return_arrays=False)
    t0 = time.perf_counter()
    r = mcp.simulate_portfolios(mu, sigma, P, risk_free=0.03, seed=1, dtype="float32", return_arrays=False)
    dt = time.perf_counter() - t0
    out["rate"] = P / dt
    out["big_pick"] = [int(r.max_sharpe["global_index"]), float(r.max_sharpe["sharpe"]), int(r.target_risk["global_index"]), float(r.target_risk["risk"])]
    print("RESULT" + json.dumps(out))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(int(sys.argv[2]), int(sys.argv[3]))
        sys.exit(0)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    P = int(sys.argv[2]) if len(sys.argv) > 2 else 20_000_000
    import numpy as np
    res = {}
    for tag, env in (("tc", "1"), ("simt", "0")):
        e = dict(os.environ, MCP_LARGE_TC=env)
        p = subprocess.run([sys.executable, __file__, "--child", str(n), str(P)], env=e, capture_output=True, text=True, timeout=600)
        line = [l for l in p.stdout.splitlines() if l.startswith("RESULT")]
        if not line:
            print(tag, "FAILED rc", p.returncode, p.stdout[-2000:], p.stderr[-3000:])
            sys.exit(1)
        res[tag] = json.loads(line[0][6:])
        print(f"{tag}: n={n} rate={res[tag]['rate']:.4g} pf/s  wsum_err={res[tag]['wsum']:.2e} consistent={res[tag]['consistent']} pick={res[tag]['pick']} big={res[tag]['big_pick']}")
    a, b = res["tc"], res["simt"]
    rr = np.abs(np.array(a["risks"]) / np.array(b["risks"]) - 1).max()
    rt = np.abs(np.array(a["returns"]) / np.array(b["returns"]) - 1).max()
    print(f"max rel diff tc vs simt: risk {rr:.3e} return {rt:.3e}  w0 diff {np.abs(np.array(a['w0']) - np.array(b['w0'])).max():.2e}  speedup {a['rate'] / b['rate']:.2f}x")
