#!/bin/bash
# compute-sanitizer over every kernel of libgtc.so (small sizes); logs in gpurun_out/.  SURVEY.md section 5.
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $S --tool memcheck --error-exitcode 7 python scripts/sanitize_run.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -4 gpurun_out/sanitize_memcheck.log
timeout 900 $S --tool racecheck --error-exitcode 7 python scripts/sanitize_run.py --no-tma > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -4 gpurun_out/sanitize_racecheck.log
timeout 900 $S --tool initcheck --error-exitcode 7 python scripts/sanitize_run.py --no-tma > gpurun_out/sanitize_initcheck.log 2>&1; echo "initcheck rc=$?"
tail -4 gpurun_out/sanitize_initcheck.log
