python -m pytest tests/test_mitsuba_plugin.py -x -q -m gpu 2>&1 | tail -6
