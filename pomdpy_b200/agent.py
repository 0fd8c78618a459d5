"""Experiment driver: epochs, environment stepping, return statistics (mirrors reference pomdpy/agent.py).

Only the POMCP branch is driven here (agent.py:132-213); the value-iteration branch lives in
pomdpy_b200.vi.  Printing follows the reference's summary lines so outputs diff cleanly, plus sims/s."""
import logging
import time

from .history import Histories, HistoryEntry
from .statistic import Statistic
from .util import console, print_divider

module = "agent"


class Results(object):
    def __init__(self):
        self.time = Statistic("Time")
        self.discounted_return = Statistic("discounted return")
        self.undiscounted_return = Statistic("undiscounted return")

    def update_reward_results(self, r, dr):
        self.undiscounted_return.add(r)
        self.discounted_return.add(dr)

    def reset_running_totals(self):
        self.time.running_total = 0.0
        self.discounted_return.running_total = 0.0
        self.undiscounted_return.running_total = 0.0

    def show(self, epoch):
        print_divider("large")
        print("\tEpoch #" + str(epoch) + " RESULTS")
        print_divider("large")
        for title, stat in (("discounted return statistics", self.discounted_return),
                            ("undiscounted return statistics", self.undiscounted_return), ("Time", self.time)):
            console(2, module, title)
            print_divider("medium")
            stat.show()
            print_divider("medium")


class Agent(object):
    def __init__(self, model, solver):
        self.logger = logging.getLogger("POMDPy.Solver")
        self.model = model
        self.results = Results()
        self.experiment_results = Results()
        self.histories = Histories()
        self.action_pool = self.model.create_action_pool()
        self.observation_pool = self.model.create_observation_pool(self)
        self.solver_factory = solver.reset
        self.total_sims = 0
        self.total_search_seconds = 0.0

    def discounted_return(self):
        if getattr(self.model, "solver", "POMCP") == "ValueIteration":
            from .vi import run_value_iteration_experiment
            run_value_iteration_experiment(self)
        else:
            self.multi_epoch()
        er = self.experiment_results
        print("\n")
        console(2, module, "epochs: " + str(self.model.n_epochs))
        console(2, module, "ave undiscounted return/step: " + str(er.undiscounted_return.mean) + " +- " +
                str(er.undiscounted_return.std_err()))
        console(2, module, "ave discounted return/step: " + str(er.discounted_return.mean) + " +- " +
                str(er.discounted_return.std_err()))
        console(2, module, "ave time/epoch: " + str(er.time.mean))
        if self.total_search_seconds > 0:
            console(2, module, "search throughput: %.4g sims/s (%d sims)" % (
                self.total_sims / self.total_search_seconds, self.total_sims))

    def multi_epoch(self):
        eps = getattr(self.model, "epsilon_start", 0.5)
        self.model.reset_for_epoch()
        for i in range(self.model.n_epochs):
            self.results = Results()
            eps = self.run_pomcp(i + 1, eps)
            self.model.reset_for_epoch()
            if self.experiment_results.time.running_total > getattr(self.model, "timeout", 3600):
                console(2, module, "Timed out after " + str(i) + " epochs in " +
                        str(self.experiment_results.time.running_total) + " seconds")
                break

    def run_pomcp(self, epoch, eps):
        """One episode (agent.py:150-213)."""
        epoch_start = time.time()
        solver = self.solver_factory(self)
        state = solver.belief_tree_index.sample_particle()          # the Monte-Carlo start state
        reward, discounted_reward, discount = 0, 0, 1.0
        for i in range(self.model.max_steps):
            start_time = time.time()
            action = solver.select_eps_greedy_action(eps, start_time)
            if eps > getattr(self.model, "epsilon_minimum", 0.1):
                eps *= getattr(self.model, "epsilon_decay", 0.95)
            step_result, is_legal = self.model.generate_step(state, action)
            reward += step_result.reward
            discounted_reward += discount * step_result.reward
            discount *= self.model.discount
            state = step_result.next_state
            self.display_step_result(i, step_result)
            if not step_result.is_terminal or not is_legal:
                solver.update(step_result)
            entry = solver.history.add_entry()
            HistoryEntry.update_history_entry(entry, step_result.reward, step_result.action,
                                              step_result.observation, step_result.next_state)
            if step_result.is_terminal or not is_legal:
                console(3, module, "Terminated after episode step " + str(i + 1))
                break
        self.total_sims += getattr(solver, "sims_done", 0)
        self.total_search_seconds += getattr(solver, "search_seconds", 0.0)
        self.results.time.add(time.time() - epoch_start)
        self.results.update_reward_results(reward, discounted_reward)
        from . import util
        if util.VERBOSITY >= 3:
            solver.history.show()
            self.results.show(epoch)
        console(3, module, "Total possible undiscounted return: " + str(self.model.get_max_undiscounted_return()))
        if util.VERBOSITY >= 3:
            print_divider("medium")
        er, r = self.experiment_results, self.results
        er.time.add(r.time.running_total)
        er.undiscounted_return.count += (r.undiscounted_return.count - 1)
        er.undiscounted_return.add(r.undiscounted_return.running_total)
        er.discounted_return.count += (r.discounted_return.count - 1)
        er.discounted_return.add(r.discounted_return.running_total)
        return eps

    @staticmethod
    def display_step_result(step_num, step_result):
        console(3, module, "Step Number = " + str(step_num))
        console(3, module, "Step Result.Action = " + step_result.action.to_string())
        console(3, module, "Step Result.Observation = " + step_result.observation.to_string())
        console(3, module, "Step Result.Reward = " + str(step_result.reward))
