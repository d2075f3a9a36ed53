#!/bin/bash
# compute-sanitizer over the kernels added late in round 2: phase-parallel BPS (bps_par.cu), generic look-ahead trainer
K='phase_parallel or (bps_slicer and fast-par) or (bps_edge and fast-par) or (train_kernel_variants and gla)'
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_pipeline.py -x -q -m gpu -k "$K" > gpurun_out/ay_memcheck.log 2>&1; echo "memcheck rc $?"; tail -4 gpurun_out/ay_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_pipeline.py -x -q -m gpu -k "(phase_parallel and 40000) or (bps_slicer and fast-par and 64-64-16) or (train_kernel_variants and gla and 16-21)" > gpurun_out/ay_racecheck.log 2>&1; echo "racecheck rc $?"; tail -4 gpurun_out/ay_racecheck.log
