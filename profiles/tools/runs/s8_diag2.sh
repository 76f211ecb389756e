python -m pytest tests/test_gpu_parity.py -x -q -k "c14_spot" 2>&1 | grep -E "assert|Error|differ|lanes" | head -12
