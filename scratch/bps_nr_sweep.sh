#!/bin/bash
# BPS rows per group (independent distance evaluations in flight per lane).  Build the variant HERE first
# (QB_NVCC_EXTRA="-DQB_BPS_NR=16" python -m qampy_b200.build), then run this on the box: times the C3 launch.
grep -A2 "bps_fast_kernelILi2ELi0" qampy_b200/lib/obj/bps_fast.ptxas.log | grep Used
python scratch/prof_bps.py 2368 2>&1 | grep -v Warn
