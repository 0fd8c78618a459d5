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
