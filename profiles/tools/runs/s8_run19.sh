python -m pytest tests/test_wavefront.py -x -q 2>&1 | tail -2
run() { # name, workload, spp, env...
  name=$1; wl=$2; spp=$3; shift; shift; shift
  env "$@" python bench.py --workload $wl --spp $spp --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s8_$name.json 2> gpurun_out/s8_$name.err
  python -c "import json;d=json.load(open('gpurun_out/s8_$name.json'));print('$name', round(d['value'],1), d['config']['triangles'], d['roofline'].get('traversal_mode'), d['roofline'].get('pipeline'))" || tail -5 gpurun_out/s8_$name.err
}
run c5_base c5 128
run c5_dbl24 c5 128 DTOF_WF_DOUBLE=24
run c5_dbl20 c5 128 DTOF_WF_DOUBLE=20
run c5_dbl28 c5 128 DTOF_WF_DOUBLE=28
run c5_leaf2 c5 128 DTOF_MAX_LEAF=2
run c5_leaf8 c5 128 DTOF_MAX_LEAF=8
run m200_base c5 64 DTOF_BENCH_MESH_N=200
run m200_dbl24 c5 64 DTOF_BENCH_MESH_N=200 DTOF_WF_DOUBLE=24
run m200_leaf2 c5 64 DTOF_BENCH_MESH_N=200 DTOF_MAX_LEAF=2
run m200_leaf8 c5 64 DTOF_BENCH_MESH_N=200 DTOF_MAX_LEAF=8
