source profiles/tools/exp/exp.sh
for n in 3 4 5 6; do
  lib=libdtof_b200_cta$n.so; [ $n = 3 ] && lib=""
  run c2_cta$n "$lib" c2 ""
  run c4_cta$n "$lib" c4 "--spp 512"
  run c5_cta$n "$lib" c5 "--spp 128"
done
export DTOF_MODE=1
run c2_bvh_cta4_leaf4 libdtof_b200_cta4.so c2 ""
export DTOF_MAX_LEAF=2
run c2_bvh_cta4_leaf2 libdtof_b200_cta4.so c2 ""
export DTOF_MAX_LEAF=1
run c2_bvh_cta4_leaf1 libdtof_b200_cta4.so c2 ""
unset DTOF_MODE
for l in 1 2 8; do export DTOF_MAX_LEAF=$l; run c4_cta4_leaf$l libdtof_b200_cta4.so c4 "--spp 512"; run c5_cta4_leaf$l libdtof_b200_cta4.so c5 "--spp 128"; done
