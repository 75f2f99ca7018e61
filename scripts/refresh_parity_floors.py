"""gpurun_out/parity_errors.json (written by a `pytest -m gpu` session on the B200, see tests/helpers.py: hold) ->
tests/parity_floors.json (key -> measured maximum error; the tests then assert min(default, max(1e-5, 2 x measured))) and
profiles/r02_parity_errors.json (the full table: measured error, asserted bound, default bound, conditioning slack)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "parity_errors.json")
d = json.load(open(src))
floors = {k: v["max_err"] for k, v in sorted(d.items())}
json.dump(floors, open(os.path.join(ROOT, "tests", "parity_floors.json"), "w"), indent=0, sort_keys=True)
json.dump(d, open(os.path.join(ROOT, "profiles", "r02_parity_errors.json"), "w"), indent=1, sort_keys=True)
above = {k: v for k, v in floors.items() if v > 1e-5}
print(len(floors), "entries;", len(above), "above 1e-5:")
for k, v in sorted(above.items(), key=lambda kv: -kv[1]):
    print(f"  {v:.2e}  {k}  (slack {d[k]['slack']:.1e}, default {d[k]['default_tol']:.0e})")
