#!/bin/bash
# Verification run used at the end of a work session (through gpurun): the whole -m gpu suite, smoke(), the per-method fill timings.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout=1500 > gpurun_out/pytest_verify.log 2>&1; tail -4 gpurun_out/pytest_verify.log
python __graft_entry__.py smoke 2>&1 | tail -1
python tools/run_fill.py 8192 0.70 idw,nn,cubic,kriging,bilinear 10; python tools/run_fill.py 8192 0.90 idw 10; python tools/run_fill.py 8192 0.97 idw 10
