set -u
cd "$GRAFT_REPO_ROOT"
python -c "import __graft_entry__ as g; g.smoke()"
python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -3
python bench.py --workload c1 --steps 5 --warmup 3 --no-cpu-baseline | cut -c1-200
