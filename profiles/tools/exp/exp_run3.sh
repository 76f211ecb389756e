source profiles/tools/exp/exp.sh
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
run c2_default "" c2 ""
run c2_fastmath libdtof_b200_fastmath.so c2 ""
run c4_fastmath libdtof_b200_fastmath.so c4 "--spp 512"
