source profiles/tools/exp/exp.sh
for w in c1 c2; do run ${w}_default "" $w ""; run ${w}_fma exp/libdtof_fma.so $w ""; done
run c4_default "" c4 "--spp 512"; run c4_fma exp/libdtof_fma.so c4 "--spp 512"
run c5_default "" c5 "--spp 128"; run c5_fma exp/libdtof_fma.so c5 "--spp 128"
export DTOF_LIB=$PWD/mitsuba3dopplertof_b200/exp/libdtof_fma.so
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python profiles/tools/exp/lane_err.py 2>&1 | grep -v "^   " | cut -c1-200
unset DTOF_LIB
timeout 300 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 3 -c 1 -o gpurun_out/r01_c2_v6 -f python bench.py --workload c2 --spp 64 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c2_v6.log 2>&1
ls -la gpurun_out | tail -5
