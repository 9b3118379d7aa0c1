// Translation unit of the marching kernel's WENO variant (k_sweep_march_weno, sweep_march_weno.cuh).
#define TTCR_B200_MARCHW_DEFINE
#include "sweep_march_weno.cuh"
