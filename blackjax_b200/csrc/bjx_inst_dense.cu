// Kernel instantiations for target kind dense (see bjx_launch.cuh).
#define BJX_INSTANTIATE_TK 2
#include "bjx_launch.cuh"
