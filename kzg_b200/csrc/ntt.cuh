// ntt.cuh -- internal NTT driver interface.
#pragma once
#include <cuda_runtime.h>

#include "field.cuh"

namespace kzg {
int ntt_run(Fr* d_data, Fr* d_tmp, uint32_t log_n, bool inverse, cudaStream_t stream, uint32_t batch = 1,
            const Fr* scale = nullptr);
struct NttScatter;
int ntt_columns_run(Fr* d_slab, uint32_t log_n, uint32_t log_n1, uint32_t log_c, uint32_t col_offset, bool inverse,
                    cudaStream_t stream, const NttScatter* scatter = nullptr);
int ntt_get_twiddles(uint32_t log_n, cudaStream_t stream, const Fr** out);
void ntt_free_twiddles();
}  // namespace kzg
