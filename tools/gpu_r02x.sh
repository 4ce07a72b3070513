#!/bin/bash
# round 2 (1 GPU): compute-sanitizer over every kernel family and mode, incl. the staged / TMA path, TAA and the band group
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
out=gpurun_out
timeout -k 10 700 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_modes.py > $out/compute_sanitizer_memcheck_r02.txt 2>&1; echo "memcheck rc=$?"
tail -4 $out/compute_sanitizer_memcheck_r02.txt | cut -c1-300
timeout -k 10 500 compute-sanitizer --tool racecheck --error-exitcode 3 python tools/sanitize_modes.py > $out/compute_sanitizer_racecheck_r02.txt 2>&1; echo "racecheck rc=$?"
tail -4 $out/compute_sanitizer_racecheck_r02.txt | cut -c1-300
