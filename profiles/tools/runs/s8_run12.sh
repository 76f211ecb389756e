python -m pytest tests/test_wavefront.py tests/test_large_scene.py -x -q -m gpu 2>&1 | tail -3
run() { # name, workload, spp, env...
  name=$1; wl=$2; spp=$3; shift; shift; shift
  env "$@" python bench.py --workload $wl --spp $spp --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s8_$name.json 2> gpurun_out/s8_$name.err
  python -c "import json;d=json.load(open('gpurun_out/s8_$name.json'));print('$name', round(d['value'],1), round(d['ms_per_step'],2), d['gpu_launches'], d['roofline'].get('traversal_mode'))" || tail -5 gpurun_out/s8_$name.err
}
run c5_wf c5 128
run c5_wf_s2 c5 128 DTOF_WF_STREAMS=2
run c5_wf_g3 c5 128 DTOF_WF_TRACE_GRID=3
run c2_fused c2 1024 DTOF_WAVEFRONT=0
run c2_wf c2 1024 DTOF_WAVEFRONT=1
run c2_wf_g3 c2 1024 DTOF_WAVEFRONT=1 DTOF_WF_TRACE_GRID=3
run c2_wf_s2 c2 1024 DTOF_WAVEFRONT=1 DTOF_WF_STREAMS=2
run c1_wf c1 1024 DTOF_WAVEFRONT=1
run c4_wf c4 512 DTOF_WAVEFRONT=1
