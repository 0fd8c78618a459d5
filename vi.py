#!/usr/bin/env python
"""Value-iteration command line: the flags of the reference's vi.py (:13-35); --solver ValueIteration only (the
LinearAlphaNet / VI-Baseline branches need TensorFlow 0.12 / matplotlib and are out of scope).  Example:

    python vi.py --env Tiger --solver ValueIteration --planning_horizon 8 --n_epochs 1 --max_steps 10 --seed 123
"""
import argparse

import numpy as np


def build_parser():
    p = argparse.ArgumentParser(description="Set the run parameters.")
    p.add_argument("--env", type=str, help="Specify the env to solve {Tiger}")
    p.add_argument("--solver", type=str, help="Specify the solver to use {ValueIteration}")
    p.add_argument("--seed", default=1993, type=int, help="Specify the random seed for numpy.random")
    p.add_argument("--use_tf", dest="use_tf", action="store_true", help="Set if using TensorFlow")
    p.add_argument("--discount", default=0.95, type=float, help="Specify the discount factor (default=0.95)")
    p.add_argument("--n_epochs", default=1, type=int, help="Num of epochs of the experiment to conduct")
    p.add_argument("--max_steps", default=10, type=int, help="Max num of steps per trial/episode/trajectory/epoch")
    p.add_argument("--save", dest="save", action="store_true", help="Pickle the weights/alpha vectors")
    p.add_argument("--learning_rate", default=0.05, type=float)
    p.add_argument("--learning_rate_minimum", default=0.0025, type=float)
    p.add_argument("--learning_rate_decay", default=0.996, type=float)
    p.add_argument("--learning_rate_decay_step", default=50, type=int)
    p.add_argument("--beta", default=0.001, type=float, help="L2 regularization parameter")
    p.add_argument("--test", default=10, type=int, help="Evaluate the agent every `test` epochs")
    p.add_argument("--epsilon_start", default=0.02, type=float)
    p.add_argument("--epsilon_minimum", default=0.05, type=float)
    p.add_argument("--epsilon_decay", default=0.96, type=float)
    p.add_argument("--epsilon_decay_step", default=75, type=int)
    p.add_argument("--planning_horizon", default=5, type=int, help="Number of lookahead steps for value iteration")
    p.add_argument("--verbosity", default=2, type=int)
    p.set_defaults(use_tf=False, save=False)
    return p


def main(argv=None):
    args = vars(build_parser().parse_args(argv))
    from pomdpy_b200 import Agent, util
    from pomdpy_b200.tiger import TigerModel
    from pomdpy_b200.vi import ValueIteration
    util.VERBOSITY = args.pop("verbosity")
    np.random.seed(int(args["seed"]))
    if args["solver"] != "ValueIteration":
        raise ValueError("solver not supported")
    if args["env"] != "Tiger":
        print("Unknown env {}".format(args["env"]))
        return None
    agent = Agent(TigerModel(args), ValueIteration)
    agent.discounted_return()
    return agent


if __name__ == "__main__":
    main()
