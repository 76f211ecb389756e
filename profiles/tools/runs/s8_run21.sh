python -m pytest tests -x -q -m gpu 2>&1 | tail -2
run() { # name, workload, spp, env...
  name=$1; wl=$2; spp=$3; shift; shift; shift
  env "$@" python bench.py --workload $wl --spp $spp --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s8_$name.json 2> gpurun_out/s8_$name.err
  python -c "import json;d=json.load(open('gpurun_out/s8_$name.json'));print('$name', round(d['value'],1), d['config']['triangles'], d['roofline'].get('traversal_mode'), d['roofline'].get('pipeline'))" || tail -5 gpurun_out/s8_$name.err
}
run c5_nopf c5 128
run c5_pf c5 128 DTOF_WF_PREFETCH=1
run c5_nopf2 c5 128
run c5_pf2 c5 128 DTOF_WF_PREFETCH=1
run m200_nopf c5 64 DTOF_BENCH_MESH_N=200
run m200_pf c5 64 DTOF_BENCH_MESH_N=200 DTOF_WF_PREFETCH=1
