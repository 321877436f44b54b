timeout 200 python -m pytest tests/test_gpu_philox.py tests/test_gpu_ensemble.py -q -x 2>&1 | tail -2
timeout 200 python scripts/bench_configs.py 2>&1 | grep -E '"c[0-9]|esteps_per_s|seconds|objectives_per_s'
