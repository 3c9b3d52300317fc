// scan.cuh -- internal interface of the linear-quotient scan.
#pragma once
#include <cuda_runtime.h>

#include "field.cuh"

namespace kzg {
size_t scan_workspace_bytes(size_t n);
int scan_divide_linear(const Fr* d_a, size_t n, const Fr* d_x, const Fr* d_y, Fr* d_q, Fr* d_rem, void* d_ws,
                       cudaStream_t stream);
}  // namespace kzg
