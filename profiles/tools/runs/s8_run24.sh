python -m pytest tests -x -q -m gpu 2>&1 | tail -5
run() { # name, workload, spp, env...
  name=$1; wl=$2; spp=$3; shift; shift; shift
  env "$@" python bench.py --workload $wl --spp $spp --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s8_$name.json 2> gpurun_out/s8_$name.err
  python -c "import json;d=json.load(open('gpurun_out/s8_$name.json'));print('$name', round(d['value'],1), d['config']['triangles'], d['roofline'].get('traversal_mode'), d['roofline'].get('pipeline'))" || tail -5 gpurun_out/s8_$name.err
}
run c2_final c2 1024
run c5_final c5 128
