#!/bin/bash
# obstacle tests again (mean |p| of the swept build over the mesh's surface vertices only) + the CCD / deferred tests that share the code
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_obstacle.py tests/test_gpu_ccd.py tests/test_gpu_deferred.py tests/test_gpu_solve.py -m gpu -x -q 2>&1 | grep -v "^$" ) > gpurun_out/r02ah_pytest.log 2>&1; tail -12 gpurun_out/r02ah_pytest.log
