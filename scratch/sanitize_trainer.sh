#!/bin/bash
# compute-sanitizer over the trainer changes of the end of round 2: grid slicer returning level values, grid with empty
# cells (cross QAM), mddma compiled in, padded error rows
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_pipeline.py -x -q -m gpu -k "searched_alphabets or (train_kernel_variants and la8 and 64-45-2) or (train_kernel_variants and la32 and 16-21-2)" > gpurun_out/bn_memcheck.log 2>&1; echo "memcheck rc $?"; tail -3 gpurun_out/bn_memcheck.log
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_pipeline.py -x -q -m gpu -k "(searched_alphabets and qam32_low_snr) or (train_kernel_variants and la8 and 16-21-2)" > gpurun_out/bn_racecheck.log 2>&1; echo "racecheck rc $?"; tail -3 gpurun_out/bn_racecheck.log
