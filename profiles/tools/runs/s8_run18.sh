run() { # name, workload, spp, env...
  name=$1; wl=$2; spp=$3; shift; shift; shift
  env "$@" python bench.py --workload $wl --spp $spp --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s8_$name.json 2> gpurun_out/s8_$name.err
  python -c "import json;d=json.load(open('gpurun_out/s8_$name.json'));print('$name', round(d['value'],1), d['config']['triangles'], d['roofline'].get('traversal_mode'), d['roofline'].get('pipeline'))" || tail -5 gpurun_out/s8_$name.err
}
for n in 24 48 96 200 400; do
  run mesh${n}_fused c5 64 DTOF_BENCH_MESH_N=$n DTOF_WAVEFRONT=0
  run mesh${n}_wf c5 64 DTOF_BENCH_MESH_N=$n DTOF_WAVEFRONT=1
done
