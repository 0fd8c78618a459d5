"""Record the reference's command-line surface (pomcp.py:13-43, vi.py:13-35) into tests/golden/cli_flags.json.
TEST INFRASTRUCTURE ONLY; needs /root/reference.  The scripts build their parsers under `if __name__ == '__main__'`,
so the argparse calls are extracted from the source with ast."""
import ast
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("POMDPY_REFERENCE", "/root/reference")
GOLDEN = os.path.join(os.path.dirname(os.path.dirname(HERE)), "tests", "golden")


def flags_of(path):
    tree = ast.parse(open(path).read())
    out, defaults = {}, {}
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute):
            if node.func.attr == "add_argument":
                name = node.args[0].value
                kw = {k.arg: k.value for k in node.keywords}
                d = dict(default=ast.literal_eval(kw["default"]) if "default" in kw else None,
                         type=kw["type"].id if "type" in kw else None,
                         action=ast.literal_eval(kw["action"]) if "action" in kw else None,
                         dest=ast.literal_eval(kw["dest"]) if "dest" in kw else None)
                out[name] = d
            elif node.func.attr == "set_defaults":
                for k in node.keywords:
                    defaults[k.arg] = ast.literal_eval(k.value)
    for name, d in out.items():
        key = d["dest"] or name.lstrip("-")
        if key in defaults:
            d["default"] = defaults[key]
    return out


def main():
    res = dict(generator="oracle/refharness/dump_cli_flags.py",
               pomcp=flags_of(os.path.join(REF, "pomcp.py")), vi=flags_of(os.path.join(REF, "vi.py")))
    with open(os.path.join(GOLDEN, "cli_flags.json"), "w") as f:
        json.dump(res, f, indent=1, sort_keys=True)
    print(len(res["pomcp"]), "pomcp flags,", len(res["vi"]), "vi flags")


if __name__ == "__main__":
    sys.exit(main())
