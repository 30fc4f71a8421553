#!/bin/bash
# kinematic obstacle (row f3, MeshCO on the barrier / TI path): its GPU tests
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_obstacle.py -m gpu -x -q 2>&1 | grep -v "^$" ) > gpurun_out/r02ad_pytest.log 2>&1; tail -40 gpurun_out/r02ad_pytest.log
