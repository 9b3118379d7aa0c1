// Translation unit of the four-nodes-per-thread marching kernel (k_sweep_march4, sweep_march4.cuh).
#define TTCR_B200_MARCH4_DEFINE
#include "sweep_march4.cuh"
