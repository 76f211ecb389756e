timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in c1 c2 c3 c4 c5; do
  timeout 300 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_${w}_v5.json
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 3 -c 1 -o gpurun_out/r01_c2_v5 -f python bench.py --workload c2 --spp 64 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c2_v5.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 3 -c 1 -o gpurun_out/r01_c5_v5 -f python bench.py --workload c5 --spp 16 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c5_v5.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r01_launches_c2_v5.csv python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launch_c2_v5.log 2>&1
python -c "
import json
for w in ['c1','c2','c3','c4','c5']:
    try:
        d=json.loads(open('gpurun_out/bench_%s_v5.json'%w).read().strip().split(chr(10))[-1]); r=d['roofline']; print(w, round(d['value'],1), round(d['e2e']['value'],1), r['traversal_mode'], round(r['bytes_per_sample']), round(r['frac'],3), round(r['issue']['frac'],3))
    except Exception as e: print(w,'ERR',e)
"
ls -la gpurun_out
