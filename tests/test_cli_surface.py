"""The command-line surface is the reference's: every flag of pomcp.py (:13-43) and vi.py (:13-35) exists here with
the same type and default (tests/golden/cli_flags.json, extracted from the reference's sources); the extra flags are
additive.  Host-side bookkeeping classes behave like the reference's."""
import json
import math
import os

from helpers import GOLDEN


def _surface(parser):
    out = {}
    for a in parser._actions:
        for opt in a.option_strings:
            if opt.startswith("--") and opt != "--help":
                out[opt] = dict(default=a.default, type=getattr(a.type, "__name__", None),
                                store_true=a.__class__.__name__ == "_StoreTrueAction")
    return out


def _check(parser, golden, additive):
    mine = _surface(parser)
    for flag, g in golden.items():
        assert flag in mine, "missing reference flag %s" % flag
        m = mine[flag]
        if g["action"] == "store_true":
            assert m["store_true"] and m["default"] is False
        else:
            assert m["type"] == g["type"], flag
            assert m["default"] == g["default"], (flag, m["default"], g["default"])
    assert set(mine) - set(golden) == additive


def test_pomcp_cli_flags_match_reference():
    import pomcp as cli
    with open(os.path.join(GOLDEN, "cli_flags.json")) as f:
        g = json.load(f)
    assert len(g["pomcp"]) == 22
    _check(cli.build_parser(), g["pomcp"], {"--map_file", "--verbosity", "--gen_w0", "--gen_growth", "--gen_wmax", "--concurrent_epochs",
                                             "--preferred_rollouts"})


def test_vi_cli_flags_match_reference():
    import vi as cli
    with open(os.path.join(GOLDEN, "cli_flags.json")) as f:
        g = json.load(f)
    _check(cli.build_parser(), g["vi"], {"--verbosity"})


def test_statistic_running_moments():
    from pomdpy_b200.statistic import Statistic
    s = Statistic("x")
    vals = [3.0, -1.5, 10.0, 4.25, 0.0]
    for v in vals:
        s.add(v)
    mean = sum(vals) / len(vals)
    var = sum((v - mean) ** 2 for v in vals) / len(vals)
    assert abs(s.mean - mean) < 1e-12 and abs(s.variance - var) < 1e-12
    assert s.max == 10.0 and s.min == -1.5 and s.count == 5 and abs(s.std_err() - math.sqrt(var / 5)) < 1e-12


def test_maps_render_parse_and_errors(tmp_path):
    import pytest
    from pomdpy_b200.rock_sample import maps
    for name, (rows, cols, start, rocks) in maps.GEOMETRY.items():
        r, c, body = maps.load(name + ".txt")
        assert (r, c) == (rows, cols) and body[start[0]][start[1]] == "S"
        assert sum(line.count("o") for line in body) == len(rocks) and all(line.endswith("G") for line in body)
    bad = tmp_path / "bad.txt"
    bad.write_text("2 3\n.oG\n..\n")
    with pytest.raises(ValueError):
        maps.load(str(bad))
    bad.write_text("1 3\n.qG\n")
    with pytest.raises(ValueError):
        maps.load(str(bad))


def test_history_and_results_bookkeeping():
    from pomdpy_b200.history import Histories, HistoryEntry
    from pomdpy_b200.rock_sample import RockAction, RockObservation
    h = Histories()
    seq = h.create_sequence()
    e = seq.add_entry()
    HistoryEntry.update_history_entry(e, 10.0, RockAction(4), RockObservation(False, False), "s")
    assert seq.get_length() == 1 and e.reward == 10.0 and e.action.to_string() == "SAMPLE" and seq.get_states() == ["s"]
    assert RockAction(7).to_string() == "CHECK" and RockAction(7).rock_no == 2
    assert RockObservation().to_string() == "EMPTY" and RockObservation(True, False) == RockObservation(True, None)
    assert hash(RockObservation(True, False)) == 1 and RockObservation(False, False) == RockObservation(False, True)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the CUDA one) needs no GPU: one JSON line with
    the metric of BASELINE.json, impl = reference, the cpu_baseline description and an e2e object."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    with open(os.path.join(root, "BASELINE.json")) as f:
        base = json.load(f)
    assert line["impl"] == "reference" and line["unit"] == "sims/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("pomcp_sims_per_sec") and "sims/sec" in base["metric"]
    assert line["value"] > 1e4 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["gpu_launches"] == 0
