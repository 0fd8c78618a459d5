"""Exact value iteration on the GPU: the drop-in for the reference's ValueIteration solver
(pomdpy/solvers/value_iteration.py:9-181, pomdpy/solvers/alpha_vector.py:1-11).

Same surface: ValueIteration.reset(agent), .value_iteration(t, o, r, horizon), .gamma (a set of AlphaVector),
.select_action(belief, vector_set), .prune(n_states).  The alpha-vector backup (the reference's own projection
formula) and the dominance prune run as CUDA kernels behind vi_solve_reference(); the prune is the
pointwise test that the reference's pairwise linear program computes (delta = 1e-10), applied in a fixed order so
that, unlike the reference (A.9), the result is deterministic."""
import ctypes as C
import time

import numpy as np

from . import native
from .solvers.solver import Solver
from .util import console, print_divider

module = "ValueIteration"


class AlphaVector(object):
    """An alpha vector and the action it belongs to (alpha_vector.py:1-11)."""

    def __init__(self, a, v):
        self.action = a
        self.v = v

    def copy(self):
        return AlphaVector(self.action, self.v)


REFERENCE_ALPHA_MODULE = "pomdpy.solvers.alpha_vector"      # where the reference's AlphaVector lives (alpha_vector.py:1-11)


def save_alpha_vectors(gamma, path):
    """Write a set of alpha vectors the way the reference's experiments/scripts/pickle_wrapper.py:22-27 does (protocol 2,
    fix_imports) and AS THE REFERENCE'S OWN CLASS: the pickle names `pomdpy.solvers.alpha_vector.AlphaVector`, a Python
    set of objects with attributes `action` and `v` (float64 ndarray) -- exactly the layout of the reference's
    experiments/pickle_jar/VI_planning_horizon_*.pkl, so its load_pkl (:29-34) and the scripts that consume those files
    (approximate_vi_eval.py:22-28) read it without this package installed."""
    import pickle
    import sys
    import types
    cls = type("AlphaVector", (object,), {"__module__": REFERENCE_ALPHA_MODULE,
                                          "__init__": lambda self, a, v: (setattr(self, "action", a), setattr(self, "v", v)) and None})
    objs = set()
    for av in gamma:
        objs.add(cls(int(av.action), np.asarray(av.v, dtype=np.float64).copy()))
    # pickle verifies that the class can be imported under its name: lend it a module of that name for the dump
    parts, added = REFERENCE_ALPHA_MODULE.split("."), []
    real = sys.modules.get(REFERENCE_ALPHA_MODULE)
    if real is not None and hasattr(real, "AlphaVector"):          # the reference itself is importable here: use its class
        objs = set(real.AlphaVector(o.action, o.v) for o in objs)
    else:
        for i in range(1, len(parts) + 1):
            name = ".".join(parts[:i])
            if name not in sys.modules:
                sys.modules[name] = types.ModuleType(name)
                added.append(name)
        sys.modules[REFERENCE_ALPHA_MODULE].AlphaVector = cls
    try:
        with open(path, "wb") as f:
            pickle.dump(objs, f, protocol=2, fix_imports=True)
    finally:
        for name in added:
            del sys.modules[name]
    return path


def load_alpha_vectors(path):
    """Read a pickle written by save_alpha_vectors or by the reference (pickle_wrapper.load_pkl semantics) into this
    package's AlphaVector objects, whether or not the reference is importable."""
    import pickle

    class _Unpickler(pickle.Unpickler):
        def find_class(self, module, name):
            if name == "AlphaVector":
                return AlphaVector
            return super().find_class(module, name)

    with open(path, "rb") as f:
        objs = _Unpickler(f, fix_imports=True, encoding="bytes").load()
    out = []
    for o in objs:
        d = {(k.decode() if isinstance(k, bytes) else k): v for k, v in o.__dict__.items()}
        out.append(AlphaVector(int(d["action"]), np.asarray(d["v"], dtype=np.float64)))
    return out


def _lib():
    L = native.load()
    if not getattr(L, "_vi_ready", False):
        P, dbl, i32 = C.POINTER, C.c_double, C.c_int32
        L.vi_solve_reference.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, P(dbl), P(dbl), P(dbl), dbl, C.c_int,
                                         C.c_int, P(dbl), P(i32), P(i32), C.c_char_p, C.c_int]
        L.vi_project_textbook_dev.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p, C.c_int]
        vp = C.c_void_p
        L.vi_point_based_backup_dev.argtypes = [C.c_int] * 6 + [vp] * 5 + [dbl] + [vp] * 7 + [C.c_int] + [vp] * 5 + \
                                               [C.c_char_p, C.c_int]
        L._vi_ready = True
    return L


def solve_reference(T, O, R, discount, horizon, capacity=4096, device=0):
    """(alpha [n][S], action [n]) after `horizon` backups with the reference's formula; fp64 on the device."""
    T = np.ascontiguousarray(T, dtype=np.float64)
    O = np.ascontiguousarray(O, dtype=np.float64)
    R = np.ascontiguousarray(R, dtype=np.float64)
    A, S, _ = T.shape
    Z = O.shape[2]
    alpha = np.zeros((capacity, S), dtype=np.float64)
    act = np.zeros(capacity, dtype=np.int32)
    n = C.c_int32()
    err = C.create_string_buffer(256)
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    rc = _lib().vi_solve_reference(int(device), S, A, Z, p(T, C.c_double), p(O, C.c_double), p(R, C.c_double),
                                   float(discount), int(horizon), int(capacity), p(alpha, C.c_double),
                                   p(act, C.c_int32), C.byref(n), err, 256)
    if rc != 0:
        raise native.PomcpError(err.value.decode() or "vi_solve_reference failed")
    return alpha[:n.value].copy(), act[:n.value].copy()


def project_textbook(T, O, gamma, device=0):
    """proj[a][s][o*G+g] = sum_s' T[a,s,s'] O[a,s',o] gamma[g,s'] on the fp64 tensor cores (torch CUDA tensors in,
    torch CUDA tensor out; |S| and |O|*|G| multiples of 64)."""
    import torch
    A, S, _ = T.shape
    Z, G = O.shape[2], gamma.shape[0]
    for t in (T, O, gamma):
        assert t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()
    scratch = torch.empty(S * Z * G, dtype=torch.float64, device=T.device)
    proj = torch.empty((A, S, Z * G), dtype=torch.float64, device=T.device)
    err = C.create_string_buffer(256)
    rc = _lib().vi_project_textbook_dev(int(device), S, A, Z, G, T.data_ptr(), O.data_ptr(), gamma.data_ptr(),
                                        scratch.data_ptr(), proj.data_ptr(),
                                        C.c_void_p(torch.cuda.current_stream().cuda_stream), err, 256)
    if rc != 0:
        raise native.PomcpError(err.value.decode() or "vi_project_textbook_dev failed")
    return proj


def point_based_backup(T, O, R, gamma, beliefs, discount, prune=True, device=0, keep_work=False):
    """One point-based backup of `gamma` at the belief points (torch CUDA float64 tensors; |S| and the number of points
    multiples of 128, |O|*|Gamma| a multiple of 64): the textbook projection and the belief scores on the fp64 tensor
    cores, one new alpha vector per belief point, pointwise-dominance prune.  Returns (alpha [n][S], action [n]) and,
    with keep_work, the dict of intermediate tensors (unpruned vectors, choices, values)."""
    import torch
    A, S, _ = T.shape
    Z, G, P = O.shape[2], gamma.shape[0], beliefs.shape[0]
    for t in (T, O, R, gamma, beliefs):
        assert t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()
    dev, N = T.device, Z * G
    f64 = lambda *shape: torch.empty(shape, dtype=torch.float64, device=dev)
    i32 = lambda *shape: torch.empty(shape, dtype=torch.int32, device=dev)
    w = dict(scratch=f64(S * N), proj=f64(A, S, N), score=f64(A, P, N), best=i32(A, P, Z), value=f64(A, P),
             alpha=f64(P, S), action=i32(P), pruned=f64(P, S), pruned_action=i32(P), keep=i32(P), n_out=i32(1))
    err = C.create_string_buffer(256)
    rc = _lib().vi_point_based_backup_dev(
        int(device), S, A, Z, G, P, T.data_ptr(), O.data_ptr(), R.data_ptr(), gamma.data_ptr(), beliefs.data_ptr(),
        float(discount), w["scratch"].data_ptr(), w["proj"].data_ptr(), w["score"].data_ptr(), w["best"].data_ptr(),
        w["value"].data_ptr(), w["alpha"].data_ptr(), w["action"].data_ptr(), int(bool(prune)), w["pruned"].data_ptr(),
        w["pruned_action"].data_ptr(), w["keep"].data_ptr(), w["n_out"].data_ptr(),
        C.c_void_p(torch.cuda.current_stream().cuda_stream), err, 256)
    if rc != 0:
        raise native.PomcpError(err.value.decode() or "vi_point_based_backup_dev failed")
    if prune:
        n = int(w["n_out"].item())
        out = (w["pruned"][:n], w["pruned_action"][:n])
    else:
        out = (w["alpha"], w["action"])
    return out + (w,) if keep_work else out


class ValueIteration(Solver):
    def __init__(self, agent):
        super(ValueIteration, self).__init__(agent)
        self.gamma = set()
        self.history = agent.histories.create_sequence()
        self.alpha = None
        self.actions = None

    @staticmethod
    def reset(agent):
        return ValueIteration(agent)

    def value_iteration(self, t, o, r, horizon):
        """Solve for `horizon` planning steps (value_iteration.py:24-73); fills self.gamma."""
        self.alpha, self.actions = solve_reference(t, o, r, self.model.discount, horizon)
        self.gamma = {AlphaVector(int(a), v.copy()) for a, v in zip(self.actions, self.alpha)}
        for k in range(horizon):
            console(3, module, "planning horizon {}...".format(k))
        return self.gamma

    def prune(self, n_states):
        """Already applied after every backup on the device; kept for API compatibility."""
        return self.gamma

    @staticmethod
    def select_action(belief, vector_set):
        """Arg-max of alpha . b (value_iteration.py:161-181)."""
        best, max_v = None, -np.inf
        for av in vector_set:
            v = float(np.dot(av.v, belief))
            if v > max_v:
                max_v, best = v, av
        if best is None:
            raise ValueError("Vector set should not be empty")
        return best.action, best


def run_value_iteration_experiment(agent):
    """Agent.discounted_return, ValueIteration branch (agent.py:37-45, 215-264)."""
    model = agent.model
    solver = agent.solver_factory(agent)
    t0 = time.time()
    solver.value_iteration(model.get_transition_matrix(), model.get_observation_matrix(), model.get_reward_matrix(),
                           model.planning_horizon)
    vectors = sorted(solver.gamma, key=lambda av: (av.action, tuple(av.v)))     # deterministic scan order
    b = model.get_initial_belief_state()
    reward, discounted_reward, discount = 0, 0, 1.0
    model.reset_for_epoch()
    for i in range(model.max_steps):
        action, _ = solver.select_action(b, vectors)
        step_result = model.generate_step(action)
        if not step_result.is_terminal:
            b = model.belief_update(b, action, step_result.observation)
        reward += step_result.reward
        discounted_reward += discount * step_result.reward
        discount *= model.discount
        agent.display_step_result(i, step_result)
        if step_result.is_terminal:
            console(3, module, "Terminated after episode step " + str(i + 1))
            break
    agent.results.time.add(time.time() - t0)
    agent.results.update_reward_results(reward, discounted_reward)
    er, r = agent.experiment_results, agent.results
    er.time.add(r.time.running_total)
    er.undiscounted_return.add(r.undiscounted_return.running_total)
    er.discounted_return.add(r.discounted_return.running_total)
    if getattr(model, "save", False):                               # agent.py:37-45 (--save)
        save_alpha_vectors(solver.gamma, "VI_planning_horizon_{}.pkl".format(model.planning_horizon))
    console(2, module, "alpha vectors: " + str(len(solver.gamma)))
    print_divider("medium")
    return solver
