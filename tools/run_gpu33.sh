#!/bin/bash
mkdir -p gpurun_out
timeout 110 compute-sanitizer --tool racecheck --racecheck-report all python tools/sanitize_small.py > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/racecheck.log; grep -c "Race reported\|hazard" gpurun_out/racecheck.log
