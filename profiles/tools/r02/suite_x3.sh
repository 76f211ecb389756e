#!/bin/bash
# the whole GPU suite three times in a row (flakiness check: float-atomics order, timing)
cd "$GRAFT_REPO_ROOT"
for i in 1 2 3; do python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -2; done
