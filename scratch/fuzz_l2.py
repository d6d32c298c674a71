"""Open-ended differential fuzz (l2): python scratch/fuzz_l2.py [seconds] [first seed].
Cases come from tests/fuzz_cases.py (seeded, so a failure is reproducible from the seed it prints);
tests/test_gpu_fuzz.py runs a fixed list of the same cases in the driver's GPU tier."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fuzz_cases

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
t0 = time.time(); n = bad = 0
while time.time() - t0 < budget:
    try:
        fuzz_cases.check_l2(seed)
    except AssertionError as e:
        bad += 1; print("MISMATCH", str(e)[:200], flush=True)
    except Exception as e:  # noqa: BLE001
        bad += 1; print("ERROR seed", seed, type(e).__name__, str(e)[:160], flush=True)
    n += 1; seed += 1
print(f"fuzz l2: {n} cases (seeds {seed - n}..{seed - 1}), {bad} bad, {time.time() - t0:.0f} s")
