#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout=1500 > gpurun_out/pytest_r2k.log 2>&1; tail -4 gpurun_out/pytest_r2k.log
python __graft_entry__.py smoke 2>&1 | tail -1
python tools/run_fill.py 8192 0.70 idw,nn,cubic,kriging,bilinear 10; python tools/run_fill.py 8192 0.90 idw 10; python tools/run_fill.py 8192 0.97 idw 10
python -m pytest tests/test_parity_gpu.py -m gpu -q -k "tiny" --timeout=600 > gpurun_out/plain_san.log 2>&1 && compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "tiny" --timeout=1500 > gpurun_out/memcheck_r2k.log 2>&1; echo "memcheck exit $?"; tail -6 gpurun_out/memcheck_r2k.log
