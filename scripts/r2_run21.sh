#!/bin/bash
timeout 600 python -m pytest tests/test_modules_gpu.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -25 | cut -c1-220
