python -m pytest tests/test_wavefront.py -x -q 2>&1 | tail -15
for wf in 0 1; do
  DTOF_WAVEFRONT=$wf python bench.py --workload c5 --spp 64 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s8_c5_wf$wf.json 2> gpurun_out/s8_c5_wf$wf.err
  python -c "import json;d=json.load(open('gpurun_out/s8_c5_wf$wf.json'));print('c5 wf=$wf', d['value'], d['ms_per_step'], d['gpu_launches'])" || tail -5 gpurun_out/s8_c5_wf$wf.err
done
