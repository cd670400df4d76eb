mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_path.py -x -q -m gpu -k "decoder or batched or whole_path or prompt or predictor or goldens" 2>&1 | tail -3
timeout 120 python tools/gpu_time_decoder.py 16 2>&1 | tail -1
timeout 120 python tools/gpu_time_decoder.py 32 2>&1 | tail -1
timeout 120 python tools/gpu_time_decoder.py 1 2>&1 | tail -1
