python -m pytest tests/test_tof.py tests/test_wavefront.py -x -q -m gpu -s 2>&1 | tail -12
