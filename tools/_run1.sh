mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_path.py -x -q -m gpu -k "eval_sweep or host_pipeline" 2>&1 | tail -5
timeout 600 python -m anyref_b200.eval_sweep --images 128 --n-seg 2 --batch 16 > gpurun_out/sweep_n1_128.json 2> gpurun_out/sweep_n1_128.err; echo "sweep rc=$?"; cat gpurun_out/sweep_n1_128.json; tail -3 gpurun_out/sweep_n1_128.err
