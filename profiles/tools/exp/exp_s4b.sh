source profiles/tools/exp/exp.sh
for w in c1 c2 c3; do run ${w}_pair "" $w ""; run ${w}_pair_fma exp/libdtof_fma.so $w ""; done
run c4_pair "" c4 "--spp 512"; run c4_pair_fma exp/libdtof_fma.so c4 "--spp 512"
run c5_pair "" c5 "--spp 128"; run c5_pair_fma exp/libdtof_fma.so c5 "--spp 128"
python profiles/tools/exp/lane_err.py 2>&1 | grep -v "^   " | cut -c1-200
