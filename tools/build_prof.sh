#!/bin/sh
# profiling build of the library with per-phase cycle counters (used by tools/phase_profile.py only)
cd "$(dirname "$0")/.." && mkdir -p cosim_b200/csrc/_build_prof && \
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -DCOSIM_PHASE_TIMING -shared -Xcompiler -fPIC \
  -o cosim_b200/csrc/_build_prof/libcosim_b200_prof.so cosim_b200/csrc/engine.cu cosim_b200/csrc/policy.cu
