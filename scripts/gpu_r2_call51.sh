#!/bin/bash
timeout 400 python -m pytest tests/test_gpu_attn.py tests/test_gpu_engine.py tests/test_gpu_gemm.py -m gpu -q -x 2>&1 | tail -3
