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
        u32, u64 = C.c_uint32, C.c_uint64
        L.vi_split_tf32_dev.argtypes = [C.c_int, C.c_longlong, vp, vp, vp, vp, C.c_char_p, C.c_int]
        L.vi_gemm_tf32x3_dev.argtypes = [C.c_int] * 4 + [vp] * 5 + [C.c_int, vp, C.c_char_p, C.c_int]
        L.vi_project_textbook_tf32x3_dev.argtypes = [C.c_int] * 5 + [vp] * 6 + [C.c_int, vp, C.c_char_p, C.c_int]
        L.vi_point_based_select_dev.argtypes = [C.c_int] * 6 + [vp] * 2 + [dbl] + [vp] * 6 + [C.c_int] + [vp] * 5 + [C.c_char_p, C.c_int]
        L.vi_policy_rollout_dev.argtypes = [C.c_int] * 5 + [vp] * 6 + [u32, dbl, C.c_int, C.c_int, u64, vp, vp, C.c_char_p, C.c_int]
        L.vi_expand_beliefs_dev.argtypes = [C.c_int] * 5 + [vp] * 3 + [u32, u64, vp, C.c_char_p, C.c_int]
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


def split_tf32(x, device=0):
    """fp64 CUDA tensor -> (hi, lo) float32 tensors with hi = tf32(x), lo = tf32(x - hi): the operands of the tcgen05
    TF32 x 3 GEMM (include/vi_b200.h vi_split_tf32_dev)."""
    import torch
    assert x.is_cuda and x.dtype == torch.float64 and x.is_contiguous()
    hi = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    lo = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    err = C.create_string_buffer(256)
    rc = _lib().vi_split_tf32_dev(int(device), x.numel(), x.data_ptr(), hi.data_ptr(), lo.data_ptr(),
                                  C.c_void_p(torch.cuda.current_stream().cuda_stream), err, 256)
    if rc != 0:
        raise native.PomcpError(err.value.decode() or "vi_split_tf32_dev failed")
    return hi, lo


TC_VARIANTS = {"tile128": 0, "persistent": 1}      # include/vi_b200.h VI_TC_*


def gemm_tf32x3(a_split, bt_split, variant="tile128", device=0):
    """C [M x N] (fp64) = A [M x K] . Bt [N x K]^T on the tcgen05 tensor cores from split operands (split_tf32).
    variant: "tile128" (one CTA per 128 x 128 tile, 4 partial accumulators: most accurate) or "persistent" (one CTA per
    SM, 2 partial accumulators, a tile's epilogue under the next tile's MMAs: 3 % faster)."""
    import torch
    (ahi, alo), (bhi, blo) = a_split, bt_split
    M, K = ahi.shape
    N = bhi.shape[0]
    out = torch.empty((M, N), dtype=torch.float64, device=ahi.device)
    err = C.create_string_buffer(256)
    rc = _lib().vi_gemm_tf32x3_dev(int(device), M, N, K, ahi.data_ptr(), alo.data_ptr(), bhi.data_ptr(), blo.data_ptr(),
                                   out.data_ptr(), TC_VARIANTS[variant], C.c_void_p(torch.cuda.current_stream().cuda_stream), err, 256)
    if rc != 0:
        raise native.PomcpError(err.value.decode() or "vi_gemm_tf32x3_dev failed")
    return out


def project_textbook(T, O, gamma, device=0, precision="fp64", t_split=None, variant="tile128"):
    """proj[a][s][o*G+g] = sum_s' T[a,s,s'] O[a,s',o] gamma[g,s'] (torch CUDA tensors in, torch CUDA tensor out).
    precision "fp64": the fp64 tensor cores (DMMA), exact (|S| and |O|*|G| multiples of 64);
    precision "tf32x3": the tcgen05 tensor cores, operands split into TF32 pairs, fp32-grade (|S| a multiple of 128,
    |O|*|G| of 128; variant as in gemm_tf32x3); t_split = split_tf32(T) can be passed to reuse the
    split of T across backups."""
    import torch
    A, S, _ = T.shape
    Z, G = O.shape[2], gamma.shape[0]
    for t in (T, O, gamma):
        assert t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()
    proj = torch.empty((A, S, Z * G), dtype=torch.float64, device=T.device)
    err = C.create_string_buffer(256)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    if precision == "tf32x3":
        thi, tlo = t_split if t_split is not None else split_tf32(T, device)
        scratch = torch.empty(2 * S * Z * G, dtype=torch.float32, device=T.device)
        rc = _lib().vi_project_textbook_tf32x3_dev(int(device), S, A, Z, G, thi.data_ptr(), tlo.data_ptr(), O.data_ptr(),
                                                   gamma.data_ptr(), scratch.data_ptr(), proj.data_ptr(), TC_VARIANTS[variant],
                                                   stream, err, 256)
        if rc != 0:
            raise native.PomcpError(err.value.decode() or "vi_project_textbook_tf32x3_dev failed")
        return proj
    assert precision == "fp64"
    scratch = torch.empty(S * Z * G, dtype=torch.float64, device=T.device)
    rc = _lib().vi_project_textbook_dev(int(device), S, A, Z, G, T.data_ptr(), O.data_ptr(), gamma.data_ptr(),
                                        scratch.data_ptr(), proj.data_ptr(), stream, err, 256)
    if rc != 0:
        raise native.PomcpError(err.value.decode() or "vi_project_textbook_dev failed")
    return proj


def point_based_backup(T, O, R, gamma, beliefs, discount, prune=True, device=0, keep_work=False, precision="fp64",
                       t_split=None, variant="tile128"):
    """One point-based backup of `gamma` at the belief points (torch CUDA float64 tensors; |S| and the number of points
    multiples of 128, |O|*|Gamma| a multiple of 64): the textbook projection and the belief scores on the fp64 tensor
    cores, one new alpha vector per belief point, pointwise-dominance prune.  Returns (alpha [n][S], action [n]) and,
    with keep_work, the dict of intermediate tensors (unpruned vectors, choices, values).
    precision="tf32x3" takes the projection (95 % of the flops) on the tcgen05 tensor cores instead (fp32-grade,
    variant as in gemm_tf32x3; t_split = split_tf32(T) reuses the split of T); scores and the rest stay fp64."""
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
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    if precision == "tf32x3":
        w["proj"] = project_textbook(T, O, gamma, device, precision="tf32x3", t_split=t_split, variant=variant)
        rc = _lib().vi_point_based_select_dev(
            int(device), S, A, Z, G, P, R.data_ptr(), beliefs.data_ptr(), float(discount), w["proj"].data_ptr(),
            w["score"].data_ptr(), w["best"].data_ptr(), w["value"].data_ptr(), w["alpha"].data_ptr(),
            w["action"].data_ptr(), int(bool(prune)), w["pruned"].data_ptr(), w["pruned_action"].data_ptr(),
            w["keep"].data_ptr(), w["n_out"].data_ptr(), stream, err, 256)
    else:
        assert precision == "fp64"
        rc = _lib().vi_point_based_backup_dev(
            int(device), S, A, Z, G, P, T.data_ptr(), O.data_ptr(), R.data_ptr(), gamma.data_ptr(), beliefs.data_ptr(),
            float(discount), w["scratch"].data_ptr(), w["proj"].data_ptr(), w["score"].data_ptr(), w["best"].data_ptr(),
            w["value"].data_ptr(), w["alpha"].data_ptr(), w["action"].data_ptr(), int(bool(prune)), w["pruned"].data_ptr(),
            w["pruned_action"].data_ptr(), w["keep"].data_ptr(), w["n_out"].data_ptr(), stream, err, 256)
    if rc != 0:
        raise native.PomcpError(err.value.decode() or "vi_point_based_backup_dev failed")
    if prune:
        n = int(w["n_out"].item())
        out = (w["pruned"][:n], w["pruned_action"][:n])
    else:
        out = (w["alpha"], w["action"])
    return out + (w,) if keep_work else out


def policy_rollout(alpha, action, T, O, R, b0, discount, max_steps, n_epochs, seed=0, term_action_mask=0, device=0):
    """Agent.run_value_iteration's policy execution (agent.py:215-264) for n_epochs episodes in ONE kernel launch (a
    block per epoch; belief, arg-max over the alpha vectors, environment step and belief update all on the device).
    Torch CUDA float64 tensors (action: int32) in; returns a [n_epochs][4] CPU array
    {discounted return, undiscounted return, steps, last action}."""
    import torch
    A, S, _ = T.shape
    Z, G = O.shape[2], alpha.shape[0]
    for t in (alpha, T, O, R, b0):
        assert t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()
    assert action.is_cuda and action.dtype == torch.int32 and action.is_contiguous()
    out = torch.zeros((n_epochs, 4), dtype=torch.float64, device=T.device)
    err = C.create_string_buffer(256)
    rc = _lib().vi_policy_rollout_dev(int(device), S, A, Z, G, alpha.data_ptr(), action.data_ptr(), T.data_ptr(), O.data_ptr(),
                                      R.data_ptr(), b0.data_ptr(), int(term_action_mask), float(discount), int(max_steps),
                                      int(n_epochs), int(seed) & 0xFFFFFFFFFFFFFFFF, out.data_ptr(),
                                      C.c_void_p(torch.cuda.current_stream().cuda_stream), err, 256)
    if rc != 0:
        raise native.PomcpError(err.value.decode() or "vi_policy_rollout_dev failed")
    return out.cpu().numpy()


def expand_beliefs(beliefs, P, T, O, round_index, seed=0, device=0):
    """Belief-point expansion (stochastic simulation, random action): rows P..2P-1 of `beliefs` [>= 2P][S] are filled
    with one successor of each of the first P points."""
    import torch
    A, S, _ = T.shape
    assert beliefs.is_cuda and beliefs.dtype == torch.float64 and beliefs.is_contiguous() and beliefs.shape[0] >= 2 * P
    err = C.create_string_buffer(256)
    rc = _lib().vi_expand_beliefs_dev(int(device), S, A, O.shape[2], int(P), T.data_ptr(), O.data_ptr(), beliefs.data_ptr(),
                                      int(round_index), int(seed) & 0xFFFFFFFFFFFFFFFF,
                                      C.c_void_p(torch.cuda.current_stream().cuda_stream), err, 256)
    if rc != 0:
        raise native.PomcpError(err.value.decode() or "vi_expand_beliefs_dev failed")
    return beliefs


def pad_vector_set(gamma, Z):
    """The projection GEMM wants |O| * |Gamma| to be a multiple of 64: repeat the first vector (duplicates change
    neither the arg-max values nor, with lowest-index ties, the choices)."""
    import torch
    G = gamma.shape[0]
    need = 64 // np.gcd(64, Z)
    pad = (-G) % need
    return gamma if pad == 0 else torch.cat([gamma, gamma[:1].expand(pad, -1)], dim=0).contiguous()


def pbvi(T, O, R, beliefs0, discount, iterations, max_points=1024, expand_every=1, gamma0=None, seed=0, device=0,
         log=None):
    """Point-based value iteration, everything on the device: `iterations` point-based backups (vi_point_based_backup_dev:
    projection and belief scores on the fp64 tensor cores, one vector per belief point, dominance prune); the belief set
    starts as `beliefs0` [P0][S] (P0 a multiple of 128) and doubles by stochastic-simulation expansion
    (vi_expand_beliefs_dev) every `expand_every` backups until max_points.  The host only sequences the launches and
    reads the size of the pruned set (4 bytes) after each backup.  Returns (alpha [n][S], action [n], beliefs [P][S])."""
    import torch
    A, S, _ = T.shape
    Z = O.shape[2]
    P = beliefs0.shape[0]
    B = torch.zeros((max(max_points, P), S), dtype=torch.float64, device=T.device)
    B[:P] = beliefs0
    if gamma0 is None:                       # the usual lower bound: min_a,s R / (1 - discount) for every state
        gamma0 = torch.full((1, S), float(R.min().item()) / (1.0 - discount), dtype=torch.float64, device=T.device)
    gamma, act = gamma0.contiguous(), torch.zeros(gamma0.shape[0], dtype=torch.int32, device=T.device)
    for it in range(iterations):
        alpha, act = point_based_backup(T, O, R, pad_vector_set(gamma, Z), B[:P].contiguous(), discount)
        gamma = alpha.contiguous().clone()
        act = act.contiguous().clone()
        if log is not None:
            log.append(dict(iteration=it, points=P, vectors=int(gamma.shape[0])))
        if (it + 1) % expand_every == 0 and 2 * P <= max_points:
            expand_beliefs(B, P, T, O, it, seed=seed, device=device)
            P *= 2
    return gamma, act, B[:P]


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
    """Agent.discounted_return, ValueIteration branch (agent.py:37-45, 215-264): solve once, then n_epochs episodes of
    policy execution.  The episodes run on the device, all at once (vi_policy_rollout_dev): the belief, the arg-max over
    the alpha vectors, the environment step and the belief update never leave the GPU."""
    import torch
    model = agent.model
    solver = agent.solver_factory(agent)
    t0 = time.time()
    T, O, R = model.get_transition_matrix(), model.get_observation_matrix(), model.get_reward_matrix()
    solver.value_iteration(T, O, R, model.planning_horizon)
    n_epochs = max(1, int(getattr(model, "n_epochs", 1)))
    dev = torch.device("cuda", torch.cuda.current_device())
    as_dev = lambda x: torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).to(dev)
    term_mask = int(getattr(model, "terminal_action_mask", 0b110 if type(model).__name__ == "TigerModel" else 0))
    res = policy_rollout(as_dev(solver.alpha), torch.from_numpy(np.ascontiguousarray(solver.actions, dtype=np.int32)).to(dev),
                         as_dev(T), as_dev(O), as_dev(R), as_dev(model.get_initial_belief_state()), model.discount,
                         int(model.max_steps), n_epochs, seed=int(getattr(model, "seed", 0)), term_action_mask=term_mask)
    elapsed = time.time() - t0
    er = agent.experiment_results
    for e in range(n_epochs):
        agent.results.update_reward_results(float(res[e, 1]), float(res[e, 0]))
        er.undiscounted_return.add(float(res[e, 1]))
        er.discounted_return.add(float(res[e, 0]))
        er.time.add(elapsed / n_epochs)
        console(3, module, "epoch %d: %d steps, discounted return %.3f" % (e + 1, int(res[e, 2]), res[e, 0]))
    agent.results.time.add(elapsed)
    if getattr(model, "save", False):                               # agent.py:37-45 (--save)
        save_alpha_vectors(solver.gamma, "VI_planning_horizon_{}.pkl".format(model.planning_horizon))
    console(2, module, "alpha vectors: " + str(len(solver.gamma)))
    print_divider("medium")
    return solver
