"""Value iteration: the numpy oracle against the executed reference's golden sets (CPU), the CUDA kernels against
the oracle (GPU)."""
import json
import os

import numpy as np
import pytest

import vi_oracle
from helpers import GOLDEN


def _golden():
    with open(os.path.join(GOLDEN, "vi_tiger.json")) as f:
        return json.load(f)


def test_oracle_matches_reference_sets_and_envelopes():
    """Horizons 1-3 of the executed reference (value_iteration.py:24-142): identical upper envelope on a 201-point
    belief grid and identical canonically pruned set (3, 7, 13 vectors)."""
    g = _golden()
    T, O, R = (np.array(g[k]) for k in ("T", "O", "R"))
    grid = np.stack([np.linspace(0, 1, 201), 1 - np.linspace(0, 1, 201)], axis=1)
    for h, want_n in ((1, 3), (2, 7), (3, 13)):
        v, a = vi_oracle.value_iteration(T, O, R, g["discount"], h)
        cv, ca = vi_oracle.canonical_set(v, a)
        ref = g["horizons"][str(h)]
        assert len(cv) == want_n == len(ref["canonical"])
        assert np.allclose(cv, np.array(ref["canonical"]), rtol=1e-12, atol=1e-12)
        assert ca.tolist() == ref["canonical_actions"]
        assert np.abs(vi_oracle.envelope(v, grid) - np.array(ref["envelope"])).max() <= 1e-9


def test_oracle_horizon_8_size_and_lp_sanity():
    g = _golden()
    T, O, R = (np.array(g[k]) for k in ("T", "O", "R"))
    sizes = [len(vi_oracle.value_iteration(T, O, R, 0.95, h)[0]) for h in range(1, 9)]
    assert sizes == [3, 7, 13, 23, 41, 71, 125, 215]
    # the reference's own sanity script (test/vi_test.py:27-29): exactly (-40,-5) is dominated by (40,55)
    vecs = np.array([[-40., -5.], [48., -2.], [25., 25.], [40., 55.]])
    keep = vi_oracle.pointwise_prune(vecs, np.zeros(4, dtype=int))
    assert 0 not in keep and 3 in keep


@pytest.mark.gpu
@pytest.mark.timeout(300)
def test_gpu_tiger_bit_exact_vs_oracle():
    from pomdpy_b200 import vi
    g = _golden()
    T, O, R = (np.array(g[k]) for k in ("T", "O", "R"))
    for h in (1, 2, 3, 5, 8):
        ov, oa = vi_oracle.value_iteration(T, O, R, 0.95, h)
        gv, ga = vi.solve_reference(T, O, R, 0.95, h)
        assert gv.shape == ov.shape and np.array_equal(gv, ov) and np.array_equal(ga, oa), h
    grid = np.stack([np.linspace(0, 1, 201), 1 - np.linspace(0, 1, 201)], axis=1)
    gv, ga = vi.solve_reference(T, O, R, 0.95, 3)
    assert np.abs(vi_oracle.envelope(gv, grid) - np.array(g["horizons"]["3"]["envelope"])).max() <= 1e-9
    # relative 1e-6 bar of the north star, against the executed reference's canonical set
    cv, ca = vi_oracle.canonical_set(gv, ga)
    assert np.allclose(cv, np.array(g["horizons"]["3"]["canonical"]), rtol=1e-6, atol=0)


@pytest.mark.gpu
@pytest.mark.timeout(300)
def test_gpu_random_pomdp_bit_exact_vs_oracle():
    rng = np.random.default_rng(7)
    S, A, Z = 5, 3, 4
    T = rng.dirichlet(np.ones(S), size=(A, S))
    O = rng.dirichlet(np.ones(Z), size=(A, S))
    R = rng.uniform(-1, 1, size=(A, S))
    from pomdpy_b200 import vi
    for h in (1, 2, 3):
        ov, oa = vi_oracle.value_iteration(T, O, R, 0.9, h)
        gv, ga = vi.solve_reference(T, O, R, 0.9, h, capacity=8192)
        assert np.array_equal(gv, ov) and np.array_equal(ga, oa)


@pytest.mark.gpu
@pytest.mark.timeout(300)
def test_gpu_textbook_projection_dmma():
    """BASELINE config 5 at reduced size: proj = T_a diag(O_a[:,o]) Gamma on the fp64 tensor cores vs float64 einsum."""
    import torch
    from pomdpy_b200 import vi
    rng = np.random.default_rng(123)
    S, A, Z, G = 256, 3, 8, 16
    T = rng.dirichlet(np.ones(S), size=(A, S))
    O = rng.dirichlet(np.ones(Z), size=(A, S))
    gamma = rng.uniform(-1, 1, size=(G, S))
    want = vi_oracle.textbook_projection(gamma, T, O)                      # [a][o][g][s]
    proj = vi.project_textbook(torch.from_numpy(T).cuda(), torch.from_numpy(O).cuda(), torch.from_numpy(gamma).cuda())
    got = proj.cpu().numpy().reshape(A, S, Z, G).transpose(0, 2, 3, 1)      # [a][s][o*G+g] -> [a][o][g][s]
    assert np.abs(got - want).max() <= 1e-6 * np.abs(want).max()
    assert np.abs(got - want).max() <= 1e-12


@pytest.mark.gpu
@pytest.mark.timeout(300)
def test_gpu_vi_cli_runs(capsys):
    import vi as cli
    agent = cli.main(["--env", "Tiger", "--solver", "ValueIteration", "--planning_horizon", "8", "--n_epochs", "1",
                      "--max_steps", "10", "--seed", "123", "--verbosity", "1"])
    assert agent.experiment_results.discounted_return.count == 1


@pytest.mark.gpu
def test_gpu_point_based_backup_vs_oracle():
    """vi_point_based_backup_dev (DMMA projection + belief scores, choices, new vectors, prune) against the float64
    numpy restatement: values within 1e-9 relative, every choice optimal to 1e-9 and identical wherever the oracle's
    margin exceeds 1e-6, every vector the exact backup of its choices, the pruned set's envelope unchanged."""
    import torch
    from pomdpy_b200 import vi
    rng = np.random.default_rng(5)
    S, A, Z, G, P, disc = 256, 4, 8, 16, 128, 0.95
    T = rng.dirichlet(np.ones(S) * 0.3, size=(A, S))
    O = rng.dirichlet(np.ones(Z), size=(A, S))
    R = rng.uniform(-1, 1, size=(A, S))
    gamma = rng.uniform(-5, 5, size=(G, S))
    B = rng.dirichlet(np.ones(S) * 0.2, size=P)
    cu = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    alpha, act, w = vi.point_based_backup(cu(T), cu(O), cu(R), cu(gamma), cu(B), disc, prune=True, keep_work=True)
    torch.cuda.synchronize()
    o_alpha, o_act, o_best, o_value, o_score = vi_oracle.point_based_backup(T, O, R, gamma, B, disc)
    g_value, g_best = w["value"].cpu().numpy(), w["best"].cpu().numpy()
    g_alpha, g_act = w["alpha"].cpu().numpy(), w["action"].cpu().numpy()
    g_score = w["score"].cpu().numpy().reshape(A, P, Z, G)
    scale = np.abs(o_score).max()
    assert np.abs(g_score - o_score).max() <= 1e-9 * scale                       # DMMA scores == float64 einsum
    assert np.abs(g_value - o_value).max() <= 1e-9 * np.abs(o_value).max()
    # choices: optimal to rounding, and identical where the oracle's margin is clear
    chosen = np.take_along_axis(o_score, g_best[..., None].astype(np.int64), axis=3)[..., 0]
    assert (o_score.max(axis=3) - chosen).max() <= 1e-9 * scale
    srt = np.sort(o_score, axis=3)
    clear = (srt[..., -1] - srt[..., -2]) > 1e-6 * scale
    assert np.array_equal(g_best[clear], o_best[clear]) and clear.mean() > 0.9
    v_srt = np.sort(o_value, axis=0)
    clear_a = (v_srt[-1] - v_srt[-2]) > 1e-6 * np.abs(o_value).max()
    assert np.array_equal(g_act[clear_a], o_act[clear_a])
    # every new vector is the exact backup of the choices the device made
    proj = vi_oracle.textbook_projection(gamma, T, O)
    for p in range(P):
        a = int(g_act[p])
        want = R[a] + disc * proj[a, np.arange(Z), g_best[a, p], :].sum(axis=0)
        assert np.abs(g_alpha[p] - want).max() <= 1e-9 * np.abs(want).max()
    # value at the belief points never decreases through the prune; same envelope as the oracle's vectors
    alpha, act = alpha.cpu().numpy(), act.cpu().numpy()
    assert 1 <= len(alpha) <= P and set(act.tolist()) <= set(range(A))
    env_g, env_all, env_o = (B @ alpha.T).max(axis=1), (B @ g_alpha.T).max(axis=1), (B @ o_alpha.T).max(axis=1)
    assert np.abs(env_g - env_all).max() <= 1e-9 * np.abs(env_all).max()
    assert np.abs(env_g - env_o).max() <= 1e-9 * np.abs(env_o).max()
    # a backup never loses value at its own belief point: b_p . alpha_p equals the chosen action's value
    own = np.einsum("ps,ps->p", B, g_alpha)
    assert np.abs(own - g_value[g_act, np.arange(P)]).max() <= 1e-9 * np.abs(own).max()


def test_oracle_point_based_backup_is_the_one_step_lookahead():
    """The alpha-vector form of the point-based backup equals the Bayes-filter form of the Bellman backup:
    max_a [ b.R_a + discount * sum_o P(o | b, a) * V(b'_{a,o}) ],  V = max_g gamma_g . b'."""
    rng = np.random.default_rng(2)
    S, A, Z, G, P, disc = 7, 3, 4, 5, 9, 0.9
    T = rng.dirichlet(np.ones(S), size=(A, S))
    O = rng.dirichlet(np.ones(Z), size=(A, S))
    R = rng.uniform(-1, 1, size=(A, S))
    gamma = rng.uniform(-3, 3, size=(G, S))
    B = rng.dirichlet(np.ones(S), size=P)
    alpha, act, best, value, score = vi_oracle.point_based_backup(T, O, R, gamma, B, disc)
    for p in range(P):
        look = np.zeros(A)
        for a in range(A):
            joint = (B[p] @ T[a])[:, None] * O[a]                 # [s'][o] = P(s', o | b, a)
            po = joint.sum(axis=0)
            fut = sum(po[o] * (gamma @ (joint[:, o] / po[o])).max() for o in range(Z) if po[o] > 0)
            look[a] = B[p] @ R[a] + disc * fut
        assert np.allclose(look, value[:, p], rtol=1e-12, atol=1e-12)
        assert act[p] == int(np.argmax(look))
        assert np.isclose(alpha[p] @ B[p], look.max(), rtol=1e-12)


@pytest.mark.gpu
def test_policy_execution_on_the_device_equals_the_restatement():
    """vi_policy_rollout_dev (agent.py:215-264 for n_epochs episodes in one launch) against the numpy restatement with
    the same draws: Tiger with the horizon-4 vectors (episodes end on opening a door) and a random 96-state POMDP."""
    import torch
    from pomdpy_b200 import vi
    from pomdpy_b200.tiger.tiger_model import TigerModel as TM
    T, O, R = (np.asarray(x, dtype=np.float64) for x in (TM.get_transition_matrix(), TM.get_observation_matrix(), TM.get_reward_matrix()))
    alpha, act = vi.solve_reference(T, O, R, 0.95, 4)
    dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    b0 = np.array([0.5, 0.5])
    got = vi.policy_rollout(dev(alpha), dev(act.astype(np.int32)), dev(T), dev(O), dev(R), dev(b0), 0.95, 12, 40, seed=9,
                            term_action_mask=0b110)
    want = vi_oracle.policy_rollout(alpha, act, T, O, R, b0, 0.95, 12, 40, seed=9, term_action_mask=0b110)
    assert np.array_equal(got, want)
    assert set(got[:, 3].astype(int)) <= {0, 1, 2} and got[:, 2].min() >= 1
    rng = np.random.default_rng(4)
    S, A, Z, G = 96, 4, 5, 7
    T = rng.dirichlet(np.ones(S) * 0.3, size=(A, S))
    O = rng.dirichlet(np.ones(Z), size=(A, S))
    R = rng.uniform(-1, 1, size=(A, S))
    alpha = rng.uniform(-2, 2, size=(G, S))
    act = rng.integers(0, A, size=G).astype(np.int32)
    b0 = rng.dirichlet(np.ones(S))
    got = vi.policy_rollout(dev(alpha), dev(act), dev(T), dev(O), dev(R), dev(b0), 0.9, 15, 6, seed=3)
    want = vi_oracle.policy_rollout(alpha, act, T, O, R, b0, 0.9, 15, 6, seed=3)
    assert np.array_equal(got[:, 2:], want[:, 2:]) and np.allclose(got[:, :2], want[:, :2], rtol=1e-12, atol=1e-12)


@pytest.mark.gpu
def test_belief_expansion_and_the_pbvi_loop():
    """vi_expand_beliefs_dev against the restatement (same draws), and the device PBVI loop: the value at the belief
    points never decreases from one backup to the next (the initial vector is a lower bound), the belief set doubles
    up to max_points, every vector is a valid lower bound of the next iterate."""
    import torch
    from pomdpy_b200 import vi
    rng = np.random.default_rng(8)
    S, A, Z = 128, 4, 64
    T = rng.dirichlet(np.ones(S) * 0.2, size=(A, S))
    O = rng.dirichlet(np.ones(Z) * 0.5, size=(A, S))
    R = rng.uniform(-1, 1, size=(A, S))
    B0 = rng.dirichlet(np.ones(S) * 0.1, size=128)
    dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    Bd = torch.zeros((256, S), dtype=torch.float64, device="cuda")
    Bd[:128] = dev(B0)
    vi.expand_beliefs(Bd, 128, dev(T), dev(O), 3, seed=11)
    want = vi_oracle.expand_beliefs(np.vstack([B0, np.zeros((128, S))]), 128, T, O, 3, seed=11)
    got = Bd.cpu().numpy()
    assert np.array_equal(got[:128], B0) and np.allclose(got[128:], want[128:], rtol=1e-13, atol=1e-15)
    assert np.allclose(got[128:].sum(axis=1), 1.0)
    log = []
    gamma, act, B = vi.pbvi(dev(T), dev(O), dev(R), dev(B0), 0.9, iterations=5, max_points=512, log=log, seed=11)
    assert [e["points"] for e in log] == [128, 256, 512, 512, 512] and B.shape[0] == 512
    assert 1 <= gamma.shape[0] <= 512 and set(act.cpu().numpy().tolist()) <= set(range(A))
    # monotone improvement at the first 128 points: rerun with fewer iterations and compare the envelopes
    g3, _, _ = vi.pbvi(dev(T), dev(O), dev(R), dev(B0), 0.9, iterations=3, max_points=512, seed=11)
    env5 = (dev(B0) @ gamma.T).max(dim=1).values.cpu().numpy()
    env3 = (dev(B0) @ g3.T).max(dim=1).values.cpu().numpy()
    assert np.all(env5 >= env3 - 1e-9)
    lower = R.min() / (1 - 0.9)
    assert np.all(env3 >= lower - 1e-9)


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_gpu_textbook_projection_at_config5_state_count():
    """BASELINE config 5's |S| = 4096 (|O| = 64; two actions and 8 vectors keep the test short): the projection
    proj[a][s][o*G+g] = sum_s' T[a,s,s'] O[a,s',o] gamma[g,s'] on the device against float64 numpy on 48 sampled rows,
    1e-6 relative to the largest entry (the north star's bar for alpha vectors) -- measured error is ~1e-15."""
    import torch
    from pomdpy_b200 import vi
    rng = np.random.default_rng(123)
    S, A, Z, G = 4096, 2, 64, 8
    T = rng.dirichlet(np.ones(S), size=(A, S))
    O = rng.dirichlet(np.ones(Z), size=(A, S))
    gamma = rng.uniform(-1, 1, size=(G, S))
    proj = vi.project_textbook(*(torch.from_numpy(x).cuda() for x in (T, O, gamma)))
    torch.cuda.synchronize()
    rows = rng.integers(0, S, size=48)
    got = proj[:, rows, :].cpu().numpy()
    want = np.einsum("art,ato,gt->arog", T[:, rows, :], O, gamma).reshape(A, len(rows), Z * G)
    assert np.abs(got - want).max() <= 1e-6 * np.abs(want).max()
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()      # fp64 tensor cores: far inside the bar


@pytest.mark.gpu
def test_gpu_tf32x3_gemm_on_tcgen05():
    """The tcgen05 GEMM of the fast projection mode (vi_tcgen05.cu: TMA loads, kind::tf32 MMAs into TMEM, operands split
    into TF32 pairs, three products per fp64 product) against float64 matmul.  Tolerance: 1e-5 x (|A| |B|^T) per entry --
    fp32-grade; measured 3e-7 at K = 256, 2.7e-6 (tile 128, four partial accumulators) / 5.4e-6 (the persistent variant, two; 2048 x 2560 gives it more tiles than SMs, so both accumulator sets cycle) at
    K = 4096 on same-sign data, where the tensor core's fp32 accumulation is at its worst."""
    import torch
    from pomdpy_b200 import vi
    torch.manual_seed(5)
    for (M, N, K, tn) in ((256, 384, 512, "tile128"), (128, 512, 1024, "persistent"), (384, 256, 96, "tile128"),
                         (2048, 2560, 160, "persistent"), (128, 128, 32, "tile128")):
        a = torch.rand(M, K, dtype=torch.float64, device="cuda") - 0.3
        bt = torch.rand(N, K, dtype=torch.float64, device="cuda") - 0.6
        sa, sb = vi.split_tf32(a), vi.split_tf32(bt)
        assert float(((a - sa[0].double() - sa[1].double()).abs() / a.abs().clamp_min(1e-300)).max()) < 2.0 ** -21
        c = vi.gemm_tf32x3(sa, sb, variant=tn)
        ref = a @ bt.T
        assert float(((c - ref).abs() / (a.abs() @ bt.abs().T)).max()) < 1e-5, (M, N, K, tn)
    with pytest.raises(Exception):
        vi.gemm_tf32x3(vi.split_tf32(a[:100].contiguous()), sb)          # M not a multiple of 128


@pytest.mark.gpu
def test_gpu_tf32x3_projection_and_backup_follow_the_fp64_path():
    """precision="tf32x3": the projection within 5e-6 of the largest entry of the exact (fp64 DMMA) one -- measured 6e-7 at
    |S| = 256, 7e-6 at |S| = 4096, i.e. OUTSIDE the 1e-6 bar of the default path, which is why it is an option -- and a
    point-based backup built on it gives the same value at every belief point to 1e-5."""
    import torch
    from pomdpy_b200 import vi
    rng = np.random.default_rng(123)
    S, A, Z, G, P = 256, 3, 8, 16, 128
    T = rng.dirichlet(np.ones(S), size=(A, S))
    O = rng.dirichlet(np.ones(Z), size=(A, S))
    R = rng.uniform(-1, 1, size=(A, S))
    gamma = rng.uniform(-1, 1, size=(G, S))
    B = rng.dirichlet(np.ones(S) * 0.05, size=P)
    Td, Od, Rd, Gd, Bd = (torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in (T, O, R, gamma, B))
    exact = vi.project_textbook(Td, Od, Gd)
    t_split = vi.split_tf32(Td)
    fast = vi.project_textbook(Td, Od, Gd, precision="tf32x3", t_split=t_split)
    assert float((fast - exact).abs().max() / exact.abs().max()) < 5e-6
    a0, _ = vi.point_based_backup(Td, Od, Rd, Gd, Bd, 0.95, prune=False)
    a1, _ = vi.point_based_backup(Td, Od, Rd, Gd, Bd, 0.95, prune=False, precision="tf32x3", t_split=t_split)
    v0, v1 = (Bd * a0).sum(dim=1), (Bd * a1).sum(dim=1)
    assert float((v0 - v1).abs().max()) < 1e-5 * float(v0.abs().max())

