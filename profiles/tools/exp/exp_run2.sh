source profiles/tools/exp/exp.sh
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in c1 c2 c3; do
  unset DTOF_MODE
  run ${w}_flat_bf "" $w ""
  run ${w}_flat_pretest libdtof_b200_pretest.so $w ""
  export DTOF_MODE=1
  run ${w}_bvh_leaf2 "" $w ""
done
unset DTOF_MODE
run c4_default "" c4 "--spp 512"
run c5_default "" c5 "--spp 128"
