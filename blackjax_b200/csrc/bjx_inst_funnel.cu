// Kernel instantiations for target kind funnel (see bjx_launch.cuh).
#define BJX_INSTANTIATE_TK 1
#include "bjx_launch.cuh"
