python -m pytest tests/test_tof.py -x -q -m gpu 2>&1 | tail -12
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
