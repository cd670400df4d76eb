mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "residual_ln" 2>&1 | tail -3
SAM_GEMM_XLOAD=tma timeout 200 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "residual_ln" 2>&1 | tail -3
SAM_GEMM_XLOAD=lsu timeout 200 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "residual_ln" 2>&1 | tail -3
for i in 1 2; do
echo "--- tma"; SAM_GEMM_XLOAD=tma timeout 300 python tools/gpu_bench_gemm.py 100 proj,proj_ln,lin2,lin2_ln 2>&1 | tee -a gpurun_out/f7_gemm.log
echo "--- lsu"; SAM_GEMM_XLOAD=lsu timeout 300 python tools/gpu_bench_gemm.py 100 proj,proj_ln,lin2,lin2_ln 2>&1 | tee -a gpurun_out/f7_gemm.log
done
for v in auto tma lsu auto; do
if [ $v = auto ]; then unset SAM_GEMM_XLOAD; else export SAM_GEMM_XLOAD=$v; fi
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/f7_bench_$v.log 2> gpurun_out/f7_bench_$v.err; echo "bench $v rc=$?"
python - $v <<'PY'
import json,sys
n=sys.argv[1]
d=json.loads(open(f"gpurun_out/f7_bench_{n}.log").read().strip().splitlines()[-1])
print(n, round(d["value"],2), round(d["ms_per_step"],2), round(d["e2e"]["value"],2), d["clocks"]["sm_mhz"], {k:round(v["ms_per_step"],2) for k,v in d["kernel_classes"].items()}, round(d["roofline"]["achieved"],1))
PY
done
