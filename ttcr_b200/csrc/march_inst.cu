// Translation unit of the fp32 first-order marching kernel (k_sweep_march, sweep_march.cuh): its 32 instances are the
// slowest part of the library to compile, so they build in parallel with grid.cu.
#define TTCR_B200_MARCH_DEFINE
#include "sweep_march.cuh"
