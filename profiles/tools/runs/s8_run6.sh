python -m pytest tests/test_wavefront.py tests/test_large_scene.py -x -q -m gpu 2>&1 | tail -5
run() { # name, env...
  name=$1; shift
  env "$@" python bench.py --workload c5 --spp 64 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/s8_$name.json 2> gpurun_out/s8_$name.err
  python -c "import json;d=json.load(open('gpurun_out/s8_$name.json'));print('$name', round(d['value'],1), round(d['ms_per_step'],2), d['gpu_launches'])" || tail -5 gpurun_out/s8_$name.err
}
run fused DTOF_WAVEFRONT=0
run wf_s1_g5 DTOF_WF_STREAMS=1 DTOF_WF_TRACE_GRID=5
run wf_s2_g5 DTOF_WF_STREAMS=2 DTOF_WF_TRACE_GRID=5
run wf_s2_g4 DTOF_WF_STREAMS=2 DTOF_WF_TRACE_GRID=4
run wf_s2_g3 DTOF_WF_STREAMS=2 DTOF_WF_TRACE_GRID=3
run wf_s2_g4_b4M DTOF_WF_STREAMS=2 DTOF_WF_TRACE_GRID=4 DTOF_WF_BATCH=4194304
run wf_s2_g4_b16M DTOF_WF_STREAMS=2 DTOF_WF_TRACE_GRID=4 DTOF_WF_BATCH=16777216
run wf_s2_g4_f28 DTOF_WF_STREAMS=2 DTOF_WF_TRACE_GRID=4 DTOF_WF_THRESHOLD=28
run wf_s2_g4_f20 DTOF_WF_STREAMS=2 DTOF_WF_TRACE_GRID=4 DTOF_WF_THRESHOLD=20
