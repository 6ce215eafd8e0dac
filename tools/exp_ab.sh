( timeout 400 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py -x -q ) 2>&1 | tail -3
bash tools/gpu_sweep.sh ab MCQ_DIRECT_EPI=1 MCQ_DIRECT_EPI=0 MCQ_DIRECT_EPI=1 MCQ_DIRECT_EPI=0
