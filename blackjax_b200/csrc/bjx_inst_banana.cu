// Kernel instantiations for target kind banana (see bjx_launch.cuh).
#define BJX_INSTANTIATE_TK 3
#include "bjx_launch.cuh"
