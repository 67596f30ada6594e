// engine_w24.cu -- third instance of the per-env kernels: the fast path (no general constraint code) compiled for 768 threads per
// CTA, i.e. 24 env-warps of 80 registers instead of 20 of 96.  cosim_create() picks this set for models whose workspace lets 24
// env-warps share an SM (engine_setup.h, `wpb_cap`): flamingo_p_v3 runs 6 % faster with it, the other robots stay on engine.cu.
#define COSIM_GENERAL 0
#define COSIM_W24 1
#define COSIM_LB 768
#include "engine_kernels.cuh"
