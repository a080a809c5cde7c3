#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl
(timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -40) > gpurun_out/r2_tests.log 2>&1
tail -40 gpurun_out/r2_tests.log
