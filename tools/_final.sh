mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q) 2>&1 | tail -3
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 200 python bench.py > gpurun_out/bench_v17.json 2> gpurun_out/bench_v17.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_v17.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], {k:round(v['ms_per_step'],2) for k,v in d['kernel_classes'].items()}, d['roofline']['frac'])"
