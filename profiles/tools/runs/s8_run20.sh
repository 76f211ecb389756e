python -m pytest tests/test_wavefront.py tests/test_large_scene.py -x -q -m gpu 2>&1 | tail -2
run() { # name, workload, spp, env...
  name=$1; wl=$2; spp=$3; shift; shift; shift
  env "$@" python bench.py --workload $wl --spp $spp --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s8_$name.json 2> gpurun_out/s8_$name.err
  python -c "import json;d=json.load(open('gpurun_out/s8_$name.json'));print('$name', round(d['value'],1), d['config']['triangles'], d['roofline'].get('traversal_mode'), d['roofline'].get('pipeline'))" || tail -5 gpurun_out/s8_$name.err
}
run c5_nopost c5 128 DTOF_WF_POSTPONE=0
run c5_post c5 128
run c5_post_i12 c5 128 DTOF_WF_INNER=12
run c5_post_i20 c5 128 DTOF_WF_INNER=20
run c5_post_i24 c5 128 DTOF_WF_INNER=24
run c5_post_i24_d28 c5 128 DTOF_WF_INNER=24 DTOF_WF_DOUBLE=28
run m200_nopost c5 64 DTOF_BENCH_MESH_N=200 DTOF_WF_POSTPONE=0
run m200_post c5 64 DTOF_BENCH_MESH_N=200
run c2wf_nopost c2 1024 DTOF_WAVEFRONT=1 DTOF_WF_POSTPONE=0
run c2wf_post c2 1024 DTOF_WAVEFRONT=1
