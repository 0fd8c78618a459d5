#!/usr/bin/env python
"""POMCP command line: the reference's 22 flags (pomcp.py:13-43) verbatim, plus --map_file / --verbosity and the
generation schedule of the CUDA planner.  Example:

    python pomcp.py --env RockSample --solver POMCP --max_steps 200 --n_epochs 10 --n_sims 500 --seed 123
"""
import argparse
import os

import numpy as np


def build_parser():
    p = argparse.ArgumentParser(description="Set the run parameters.")
    p.add_argument("--env", type=str, help="Specify the env to solve {RockSample} (additive: Tiger)")
    p.add_argument("--solver", default="POMCP", type=str, help="Specify the solver to use {POMCP}")
    p.add_argument("--seed", default=1993, type=int, help="Specify the random seed for numpy.random")
    p.add_argument("--use_tf", dest="use_tf", action="store_true", help="Set if using TensorFlow")
    p.add_argument("--discount", default=0.95, type=float, help="Specify the discount factor (default=0.95)")
    p.add_argument("--n_epochs", default=100, type=int, help="Num of epochs of the experiment to conduct")
    p.add_argument("--max_steps", default=200, type=int, help="Max num of steps per trial/episode/trajectory/epoch")
    p.add_argument("--save", dest="save", action="store_true", help="Pickle the weights/alpha vectors")
    p.add_argument("--test", default=10, type=int, help="Evaluate the agent every `test` epochs")
    p.add_argument("--epsilon_start", default=0.5, type=float)
    p.add_argument("--epsilon_minimum", default=0.1, type=float)
    p.add_argument("--epsilon_decay", default=0.95, type=float)
    p.add_argument("--epsilon_decay_step", default=20, type=int)
    p.add_argument("--n_sims", default=500, type=int, help="For POMCP, this is the num of MC sims to do at each "
                   "belief node")
    p.add_argument("--timeout", default=3600, type=int, help="Max num of sec the experiment should run before timeout")
    p.add_argument("--preferred_actions", dest="preferred_actions", action="store_true",
                   help="For RockSample, specify whether smart actions should be used")
    p.add_argument("--ucb_coefficient", default=3.0, type=float, help="Coefficient for UCB algorithm used by MCTS")
    p.add_argument("--n_start_states", default=2000, type=int, help="Num of state particles to generate for root "
                   "belief node in MCTS")
    p.add_argument("--min_particle_count", default=1000, type=int, help="Lower bound on num of particles a belief "
                   "node can have in MCTS")
    p.add_argument("--max_particle_count", default=2000, type=int, help="Upper bound on num of particles a belief "
                   "node can have in MCTS")
    p.add_argument("--max_depth", default=100, type=int, help="Max depth for a DFS of the belief search tree in MCTS")
    p.add_argument("--action_selection_timeout", default=60, type=int, help="Max num of secs for action selection")
    # additive flags (not in the reference)
    p.add_argument("--preferred_rollouts", action="store_true",
                   help="with --preferred_actions: rollouts draw from the smart actions of their own history too")
    p.add_argument("--map_file", default=None, type=str, help="RockSample map name or path (default map-7-8)")
    p.add_argument("--verbosity", default=2, type=int, help="console verbosity 0..4 (reference prints at 3)")
    p.add_argument("--gen_w0", default=0, type=int, help="first generation width of the device search (0 = auto)")
    p.add_argument("--gen_growth", default=0, type=int, help="generation growth factor (0 = auto)")
    p.add_argument("--gen_wmax", default=0, type=int, help="max simulations in flight (0 = auto)")
    p.add_argument("--concurrent_epochs", default=0, type=int, help="run this many epochs at once on the GPU (0 = serial)")
    p.set_defaults(preferred_actions=False, use_tf=False, save=False)
    return p


def main(argv=None):
    args = vars(build_parser().parse_args(argv))
    from pomdpy_b200 import Agent, util
    from pomdpy_b200.rock_sample import RockModel
    from pomdpy_b200.solvers import POMCP
    util.VERBOSITY = args.pop("verbosity")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:                        # launched by torchrun: root-parallel search, one process per GPU
        import torch
        import torch.distributed as dist
        local = int(os.environ.get("LOCAL_RANK", "0"))
        if torch.cuda.is_available():
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group("gloo")
        if dist.get_rank() != 0:
            util.VERBOSITY = min(util.VERBOSITY, 0)
    np.random.seed(args["seed"])
    if args["solver"] != "POMCP":
        raise ValueError("solver not supported")
    if args["env"] in ("RockSample", "Tiger"):       # Tiger: additive, the reference's matrices under the CUDA planner
        lanes = args.pop("concurrent_epochs", 0)
        if args["env"] == "Tiger":
            from pomdpy_b200.tabular import tiger
            env = tiger(args)
        else:
            env = RockModel(args)
            if util.VERBOSITY >= 3:
                env.draw_env()
        if lanes and lanes > 1 and world == 1:
            from pomdpy_b200.batch import BatchedAgent
            from pomdpy_b200.util import console
            agent = BatchedAgent(env, lanes)
            mean, err = agent.run(args["n_epochs"])
            console(2, "agent", "epochs: " + str(args["n_epochs"]))
            console(2, "agent", "ave discounted return/step: " + str(mean) + " +- " + str(err))
            console(2, "agent", "ave time/epoch: " + str(agent.wall_seconds / max(1, args["n_epochs"])))
            return agent
        agent = Agent(env, POMCP)
        agent.discounted_return()
        return agent
    print("Unknown env {}".format(args["env"]))
    return None


if __name__ == "__main__":
    main()
