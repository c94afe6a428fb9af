timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_sampler.py -q -m gpu -x -k "two_streams or qp_ or sampler or estimator or batch_sizes" 2>&1 | tail -4
