python -m pytest tests -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py --steps 3 --warmup 3 > gpurun_out/s8_final_bench.json 2> gpurun_out/s8_final_bench.err; wc -l gpurun_out/s8_final_bench.json; python -c "
import json; d=json.load(open('gpurun_out/s8_final_bench.json')); print(d['value'], d['e2e']['value'], d['cpu_baseline']['kind'], d['cpu_baseline']['value'], d['gpu_launches'], d['scaling'])"
python bench.py --impl reference --steps 1 --warmup 0 | cut -c1-160
