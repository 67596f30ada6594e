// engine_w12.cu -- fourth instance of the per-env kernels: the fast path compiled for 384 threads per CTA, i.e. up to 12 env-warps of
// up to 168 registers.  Models whose workspace leaves room for at most 12 env-warps per SM (humanoid_p_v0: 11, w4_p_v2 on the
// stairs rasters: 12) cannot use the registers the 640-thread build gives up: +5.8 % (humanoid) / +4.6 % (w4) on B200, round 2.
#define COSIM_GENERAL 0
#define COSIM_W12 1
#define COSIM_LB 384
#include "engine_kernels.cuh"
