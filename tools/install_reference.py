"""Put an UNMODIFIED copy of the Python reference (pemami4911/POMDPy) under baseline/_ref so that it travels to the GPU
box (baseline/_ref is git-ignored, not gpurun-ignored) and `bench.py --impl reference` can time the reference's own
POMCP on the box's host beside the C port (`cpu_baseline_python`).

    python tools/install_reference.py          # in the build container, where /root/reference exists

Layout: baseline/_ref/ref = the reference tree (copied, never executed in place: Model.__init__ mkdirs next to the
package); baseline/_ref/shim = the two-module past/future shim the reference imports (the `future` package
is not installed here).  The only edit is the map selection in pomdpy/config/rock_sample_config.json (a config value the
reference reads; `map-15-15.txt` for BASELINE config 4).  `pip install --target` is not used: the reference has no
setup.py / build step."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "refharness"))

import ref_env  # noqa: E402

if __name__ == "__main__":
    work = os.path.join(ROOT, "baseline", "_ref")
    ref, shim = ref_env.prepare_copy(work, map_file=sys.argv[1] if len(sys.argv) > 1 else "map-15-15.txt")
    print("reference copied to", ref, "shims in", shim)
