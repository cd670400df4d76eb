mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_path.py -x -q -m gpu -k "bit_packed or eval_sweep or fused_iou" 2>&1 | tail -3
timeout 600 python -m anyref_b200.eval_sweep --images 256 --n-seg 2 --batch 16 > gpurun_out/sweep_n1_256b.json 2> gpurun_out/sweep_n1_256b.err; echo "sweep rc=$?"; cat gpurun_out/sweep_n1_256b.json | cut -c1-700
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/f9_bench.log 2> gpurun_out/f9_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open("gpurun_out/f9_bench.log").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["parity"], d["cpu_baseline"])
PY
tail -3 gpurun_out/f9_bench.err
