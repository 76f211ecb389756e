python -m pytest tests/test_wavefront.py tests/test_large_scene.py -x -q -m gpu 2>&1 | tail -3
run() { # name, env...
  name=$1; shift
  env "$@" python bench.py --workload c5 --spp 128 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/s8_$name.json 2> gpurun_out/s8_$name.err
  python -c "import json;d=json.load(open('gpurun_out/s8_$name.json'));print('$name', round(d['value'],1), round(d['ms_per_step'],2), d['gpu_launches'])" || tail -5 gpurun_out/s8_$name.err
}
run wf_s2 DTOF_WF_STREAMS=2
run wf_s3 DTOF_WF_STREAMS=3
run wf_s4 DTOF_WF_STREAMS=4
run wf_s3_g3 DTOF_WF_STREAMS=3 DTOF_WF_TRACE_GRID=3
run wf_s4_g3 DTOF_WF_STREAMS=4 DTOF_WF_TRACE_GRID=3
run wf_s4_g2 DTOF_WF_STREAMS=4 DTOF_WF_TRACE_GRID=2
run wf_s2_shade4 DTOF_WF_STREAMS=2 DTOF_LIB=$PWD/mitsuba3dopplertof_b200/libdtof_shade4.so
run wf_s4_shade4 DTOF_WF_STREAMS=4 DTOF_LIB=$PWD/mitsuba3dopplertof_b200/libdtof_shade4.so
run wf_s4_b8M DTOF_WF_STREAMS=4 DTOF_WF_BATCH=8388608
