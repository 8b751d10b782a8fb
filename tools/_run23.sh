timeout 900 python -m pytest tests/test_gpu_lk_dense.py tests/test_gpu_pipeline.py -q -x 2>&1 | tail -4
