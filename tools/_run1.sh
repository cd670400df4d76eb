mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/f8_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/f8_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
SAM_GEMM_DEBUG=1 timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/f8_bench.log 2> gpurun_out/f8_bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/f8_bench.log; grep -i "co-resident" gpurun_out/f8_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/f8_bench_ref.log 2>&1; echo "ref rc=$?"; cut -c1-300 gpurun_out/f8_bench_ref.log
