#pragma once
#include <cuda_runtime.h>

#include "field.cuh"

namespace kzg {
size_t interp_workspace_bytes(size_t m);
int interpolate_run(const Fr* d_xs, const Fr* d_ys, size_t m, Fr* d_r, Fr* d_Z, int* d_flag, void* d_ws,
                    cudaStream_t stream);
size_t batch_div_workspace_bytes(size_t n, size_t m);
int batch_divide_run(const Fr* d_p, size_t n, const Fr* d_r, size_t lr, const Fr* d_Z, const Fr* d_xs, size_t m,
                     Fr* d_psi, size_t* psi_len, int* d_flag, void* d_ws, cudaStream_t stream,
                     cudaEvent_t r_ready = nullptr);
}  // namespace kzg
