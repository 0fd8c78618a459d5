import sys, cProfile, pstats
sys.path.insert(0, '/root/repo')
import pomcp as cli
cli.main(["--env", "RockSample", "--solver", "POMCP", "--n_epochs", "5", "--n_sims", "500", "--seed", "1", "--verbosity", "0"])
pr = cProfile.Profile(); pr.enable()
cli.main(["--env", "RockSample", "--solver", "POMCP", "--n_epochs", "60", "--n_sims", "500", "--seed", "123", "--verbosity", "0"])
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
