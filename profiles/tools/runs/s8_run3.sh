python -m pytest tests/test_wavefront.py -x -q 2>&1 | tail -5
run() { # name, env...
  name=$1; shift
  env "$@" python bench.py --workload c5 --spp 64 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/s8_$name.json 2> gpurun_out/s8_$name.err
  python -c "import json;d=json.load(open('gpurun_out/s8_$name.json'));print('$name', round(d['value'],1), round(d['ms_per_step'],2), d['gpu_launches'])" || tail -5 gpurun_out/s8_$name.err
}
run fused DTOF_WAVEFRONT=0
for inner in 8 12 16 20 24; do run wf_i${inner}_f24 DTOF_WAVEFRONT=1 DTOF_WF_INNER=$inner DTOF_WF_THRESHOLD=24; done
for f in 16 28 31; do run wf_i16_f$f DTOF_WAVEFRONT=1 DTOF_WF_INNER=16 DTOF_WF_THRESHOLD=$f; done
