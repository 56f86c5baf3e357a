// Third compilation of the LM program: grid-wide "wave" execution for batches of heavy areas (see lm_wave.cuh).
#define SRBA_WAVE 1
#include "lm_kernel.cu"
