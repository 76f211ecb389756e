run() { # name, workload, spp, env...
  name=$1; wl=$2; spp=$3; shift; shift; shift
  env "$@" python bench.py --workload $wl --spp $spp --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s8_$name.json 2> gpurun_out/s8_$name.err
  python -c "import json;d=json.load(open('gpurun_out/s8_$name.json'));print('$name', round(d['value'],1), round(d['ms_per_step'],2), d['gpu_launches'], d['roofline'].get('traversal_mode'))" || tail -5 gpurun_out/s8_$name.err
}
L=$PWD/mitsuba3dopplertof_b200
run c5_base c5 128
run c5_tr4 c5 128 DTOF_LIB=$L/libdtof_wftr4.so
run c5_tr4_g3 c5 128 DTOF_LIB=$L/libdtof_wftr4.so DTOF_WF_TRACE_GRID=3
run c5_tr6 c5 128 DTOF_LIB=$L/libdtof_wftr6.so
run c5_tr6_g3 c5 128 DTOF_LIB=$L/libdtof_wftr6.so DTOF_WF_TRACE_GRID=3
run c2wf_base c2 1024 DTOF_WAVEFRONT=1
run c2wf_tr4 c2 1024 DTOF_WAVEFRONT=1 DTOF_LIB=$L/libdtof_wftr4.so
run c2wf_tr6 c2 1024 DTOF_WAVEFRONT=1 DTOF_LIB=$L/libdtof_wftr6.so
# reference arm with the rebuilt oracle/_ref
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/s8_ref_arm.json 2> gpurun_out/s8_ref_arm.err; tail -c 400 gpurun_out/s8_ref_arm.err; cat gpurun_out/s8_ref_arm.json
python -m pytest tests/test_mitsuba_plugin.py -x -q -m gpu 2>&1 | tail -3
