"""Writes reface_b200/param_manifest.json (state-dict key -> shape, init kind) from the oracle's spec.
The product never imports the oracle; tests/test_abi.py checks the manifest stays equal to the oracle spec."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import reface_oracle as O  # noqa: E402

spec = O.full_spec()
out = {k: [list(s), kind] for k, (s, kind) in spec.items()}
with open(os.path.join(ROOT, "reface_b200", "param_manifest.json"), "w") as f:
    json.dump(out, f, separators=(",", ":"))
print(len(out), "entries")
