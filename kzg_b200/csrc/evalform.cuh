#pragma once
#include <cuda_runtime.h>

#include "field.cuh"

namespace kzg {
size_t div_by_omega_workspace_bytes(uint32_t log_d);
int div_by_omega_run(const Fr* d_f, uint32_t log_d, uint32_t m, Fr* d_q, void* d_ws, cudaStream_t stream);
}  // namespace kzg
