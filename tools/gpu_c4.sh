#!/bin/bash
# Six-point path on the GPU box: parity tests, then config C4 timing (pairs given as $1, default 20000).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sixpt.py -m gpu -q -x 2>&1 | tail -4
timeout 300 python tools/c4_probe.py ${1:-20000}
