mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "window" 2>&1 | tail -3
for i in 1 2; do
echo "--- new"; timeout 200 python tools/gpu_check_attn.py bench 2>&1 | grep attn_
echo "--- prev"; ANYREF_SAM_LIB=$PWD/anyref_b200/libanyref_sam_prev.so timeout 200 python tools/gpu_check_attn.py bench 2>&1 | grep attn_
done
