#!/usr/bin/env bash
( timeout 900 python -m pytest tests -m gpu -q -rfs 2>&1 | tail -25 ) > gpurun_out/r2_pytest_only.log; tail -25 gpurun_out/r2_pytest_only.log
