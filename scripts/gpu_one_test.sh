#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest ${TESTS:-tests} -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/pytest_one.log
