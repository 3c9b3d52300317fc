// batchinv.cuh -- one field inversion per thread BLOCK (Montgomery's trick across threads).
//
// Every thread of a block brings one non-zero value v_t (usually the product of its own small batch) and receives
// v_t^-1.  Prefix and suffix products across the block (log2 NT rounds through shared memory), ONE Kaliski inversion
// of the total by one thread, then v_t^-1 = total^-1 * prod_{u<t} v_u * prod_{u>t} v_u.  A binary inversion is ~16 k
// data-dependent (divergent) instructions for Fr; one per thread made it two thirds of pointwise_div_kernel and
// div_by_omega_kernel.  If any v_t is zero every result is zero (the callers' inputs are non-zero by construction).
#pragma once
#include "field.cuh"

namespace kzg {

template <int NT, class F>
KZG_DI F block_inverse(const F& v, F* sh /* 2 * NT + 1 elements */) {
    F* pre = sh;
    F* suf = sh + NT;
    const int t = threadIdx.x;
    pre[t] = v;
    suf[t] = v;
    __syncthreads();
    F p = v, s = v;
#pragma unroll 1
    for (int d = 1; d < NT; d <<= 1) {
        const bool hp = t >= d, hs = t + d < NT;
        F po, so;
        if (hp) po = pre[t - d];
        if (hs) so = suf[t + d];
        __syncthreads();
        if (hp) { p = p * po; pre[t] = p; }
        if (hs) { s = s * so; suf[t] = s; }
        __syncthreads();
    }
    if (t == NT - 1) sh[2 * NT] = p.inv();   // p = product of the whole block
    __syncthreads();
    F r = sh[2 * NT];
    if (t > 0) r = r * pre[t - 1];
    if (t + 1 < NT) r = r * suf[t + 1];
    return r;
}

}  // namespace kzg
