"""Per-move host/device time of the CLI path at the reference's budget (n_sims = 500)."""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
import pomcp as cli
t0 = time.time()
agent = cli.main(["--env", "RockSample", "--solver", "POMCP", "--n_epochs", "100", "--n_sims", sys.argv[1] if len(sys.argv) > 1 else "500",
                  "--seed", "123", "--verbosity", "0"])
dt = time.time() - t0
er = agent.experiment_results
print("epochs 100: %.2f s wall, %.1f ms/epoch, search %.3f ms/move (%d sims total, %.3g sims/s in search), return %.2f +- %.2f"
      % (dt, 10 * dt, 1e3 * agent.total_search_seconds / max(1, agent.total_sims / int(sys.argv[1] if len(sys.argv) > 1 else 500)),
         agent.total_sims, agent.total_sims / agent.total_search_seconds, er.discounted_return.mean, er.discounted_return.std_err()))
