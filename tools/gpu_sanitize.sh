#!/bin/bash
mkdir -p gpurun_out
( for tool in memcheck racecheck synccheck; do
    echo "== compute-sanitizer --tool $tool python tools/sanitize.py"
    timeout 600 compute-sanitizer --tool $tool python tools/sanitize.py 2>&1 | grep -E "sanitize workload done|SUMMARY|hazard|Invalid|error" | head -20
  done ) | tee gpurun_out/sanitizer.log
