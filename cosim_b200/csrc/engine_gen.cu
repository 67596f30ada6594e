// engine_gen.cu -- second instance of the per-env kernels, compiled WITH the general constraint path (engine_general.h: condim
// 1 / 4 / 6, elliptic cone, PGS, impratio).  cosim_create() picks this set when the model asks for any of those options.
#define COSIM_GENERAL 1
#include "engine_kernels.cuh"
