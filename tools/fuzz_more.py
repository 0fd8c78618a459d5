"""One-off extended run of tests/test_gpu_fuzz.py over many more random configurations."""
import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo/oracle')
import test_gpu_fuzz as f
lo, hi = int(sys.argv[1]), int(sys.argv[2])
t0 = time.time(); n = 0
for case in range(lo, hi):
    f.test_fuzz_rocksample(case); n += 1
    if case % 2 == 0:
        f.test_fuzz_tabular(case); n += 1
print("fuzz ok: %d cases in %.1f s" % (n, time.time() - t0))
