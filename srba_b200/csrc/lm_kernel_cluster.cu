// The LM program of lm_kernel.cu compiled for thread-block CLUSTERS: 8 CTAs (8 SMs) share one large problem.
// Stage loops stride over the whole cluster, barriers are cluster barriers, reductions and the solver's pivot scratch are
// exchanged through distributed shared memory, every mutable array lives in the L2-resident global workspace.
#define SRBA_CLUSTER 8
#include "lm_kernel.cu"
