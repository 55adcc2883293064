// Kernel instantiations for target kind diag (see bjx_launch.cuh).
#define BJX_INSTANTIATE_TK 0
#include "bjx_launch.cuh"
