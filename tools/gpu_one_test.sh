#!/bin/bash
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "reference_api" 2>&1 | tail -40
