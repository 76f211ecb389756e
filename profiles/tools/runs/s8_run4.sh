source profiles/tools/exp/exp.sh
DTOF_LIB=$PWD/mitsuba3dopplertof_b200/libdtof_phased_2_1.so python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
for v in "" libdtof_phased_2_1.so libdtof_phased_1_1.so libdtof_phased_1_2.so; do
  run c2_${v:-base} "$v" c2 "--spp 256"
  run c1_${v:-base} "$v" c1 "--spp 256"
  run c4_${v:-base} "$v" c4 "--spp 256"
  DTOF_WAVEFRONT=0 run c5f_${v:-base} "$v" c5 "--spp 64"
done
