#!/bin/bash
# One gpurun call: A/B of edge_reserve / ring_cost on the bench workloads, then the GPU test-suite.  Logs go to gpurun_out/.
mkdir -p gpurun_out
echo "== reserve_ab"; timeout ${AB_TIMEOUT:-30} python profiles/reserve_ab.py > gpurun_out/r2_reserve_ab.txt 2>&1; echo "ab rc=$?"; cat gpurun_out/r2_reserve_ab.txt
echo "== pytest"; timeout ${PYTEST_TIMEOUT:-70} python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest.log
