// ntt.cuh -- internal NTT driver interface.
#pragma once
#include <cuda_runtime.h>

#include "field.cuh"

namespace kzg {
// nonzero_prefix > 0: the caller guarantees d_data[i] == 0 for i >= nonzero_prefix; when the prefix fits inside the first
// pass's stride that pass is replaced by a direct evaluation (see ntt_sparse_first_pass_kernel)
int ntt_run(Fr* d_data, Fr* d_tmp, uint32_t log_n, bool inverse, cudaStream_t stream, uint32_t batch = 1,
            const Fr* scale = nullptr, uint32_t nonzero_prefix = 0);
// Fused four-step exchange: destination row matrices of every rank (peer[rank] is local, the others are CUDA-IPC
// mappings valid on this device); see ntt.cu / comm.cu.
constexpr int NTT_MAX_PEERS = 16;
struct NttScatter {
    Fr* peer[NTT_MAX_PEERS];   // rows buffer of every rank
    uint32_t log_rpr;          // log2(rows per rank) = log_n1 - log2(world)
    uint32_t rank, log_world, log_c;
};
int ntt_columns_run(Fr* d_slab, uint32_t log_n, uint32_t log_n1, uint32_t log_c, uint32_t col_offset, bool inverse,
                    cudaStream_t stream, const NttScatter* scatter = nullptr);
int ntt_get_twiddles(uint32_t log_n, cudaStream_t stream, const Fr** out);
}  // namespace kzg
