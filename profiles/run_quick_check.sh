# quick GPU validation: the new tests, then the whole GPU suite
timeout 600 python -m pytest tests/test_gpu_dropout_sites.py tests/test_gpu_ops.py tests/test_gpu_shells.py -q -x 2>&1 | tail -15
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
