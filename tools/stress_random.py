"""One-off stress: tests/test_gpu_parity.py::test_randomised_configs_vs_oracle over many more seeds.
    python tools/stress_random.py [first_seed] [last_seed]          (120 seeds took 21 s on a B200: 0 failures, round 1)"""
import os, sys
import pytest
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_parity import test_randomised_configs_vs_oracle as f
lo = int(sys.argv[1]) if len(sys.argv) > 1 else 12
hi = int(sys.argv[2]) if len(sys.argv) > 2 else 132
bad = 0
for seed in range(lo, hi):
    try:
        f(seed)
    except pytest.skip.Exception:
        print(seed, "skipped")
    except Exception as e:           # noqa: BLE001 -- report and continue
        bad += 1
        print(seed, "FAIL", str(e)[:300].replace("\n", " "))
print("failures", bad)
