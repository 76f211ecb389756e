#!/usr/bin/env python
"""Fixture generator (build container only; NOT a test): evaluates the REFERENCE's own reconstruction-filter plugins
through oracle/_ref/rfilter_vectors (oracle/ref_harness/rfilter_vectors.cpp) and writes tests/golden/rfilter_vectors.json,
which pins the oracle's filter evaluation (FilmSplat::eval) in tests/test_rfilters.py."""
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF_RT = os.path.join(ROOT, "oracle", "_ref")


def main():
    env = dict(os.environ, LD_LIBRARY_PATH=REF_RT, DTOF_REF_DIR=REF_RT)
    r = subprocess.run([os.path.join(REF_RT, "rfilter_vectors")], capture_output=True, text=True, env=env, check=True)
    data = json.loads(r.stdout)
    with open(os.path.join(ROOT, "tests", "golden", "rfilter_vectors.json"), "w") as f:
        json.dump(data, f, separators=(",", ":"))
    print({k: v["radius"] for k, v in data["filters"].items()}, len(data["x"]), "offsets")


if __name__ == "__main__":
    main()
