// comm.cu -- multi-GPU entry points of the C ABI: one process per GPU, NCCL over NVLink 5 / NVSwitch.
//
// north_star / SURVEY 8(e): "the MSM shards naturally by splitting points and scalars per GPU, with the per-GPU
// partial G1 sums reduced over NVLink via NCCL.  NTTs of 2^24 and above use the four-step split".  The reference has
// no multi-device code (its only parallelism is the crossbeam scope of parallel_fft, ft.rs:336-387); the calls these
// entry points serve are the same KZGProver::commit (coeff_form.rs:59-64) and EvaluationDomain::fft/ifft
// (ft.rs:111-140), issued by every rank of a job on its shard.
//
//  * kzgb_msm_g1_sharded*: rank g permanently holds the SRS shard gs[lo_g, hi_g) (its own kzgb_srs handle, window
//    table included) and passes the matching slice of the scalars.  Local Pippenger -> one raw XYZZ partial (192 B)
//    -> ncclAllGather of world x 192 B -> world-1 additions + affine conversion on every rank.  There is no
//    elliptic-curve ncclRedOp, and 192 bytes leave nothing to fuse or overlap.
//  * kzgb_ntt4_plan / kzgb_ntt_fr_sharded*: four-step NTT with the exchange FUSED into the column pass: every rank
//    owns two row matrices [N1/G][N2] (cudaMalloc) mapped into all other ranks through CUDA IPC; the column kernel's
//    epilogue stores element (k1, c) straight into the row matrix of the rank that owns row k1 (peer stores over
//    NVLink, already transposed), one stream-ordered barrier (a 4-byte ncclAllReduce), then the row pass.  The two
//    matrices alternate between calls, so one barrier per call suffices (a rank can only be storing into matrix k%2
//    of call k after every rank passed the barrier of call k-1, i.e. finished the row pass of call k-2).
//
// NCCL is bound at run time (dlopen of libnccl.so.2, the soname both the system library and PyTorch's bundled copy
// carry), so a process that has already loaded PyTorch shares its NCCL instead of loading a second one, and
// single-GPU users of libkzg_b200.so need no NCCL at all.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "ec.cuh"
#include "fr_util.cuh"
#include "msm.cuh"
#include "ntt.cuh"

namespace kzg {
namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

struct Comm {
    std::recursive_mutex mu;   // collectives must be issued in the same order on every rank: one at a time
    NcclApi api;
    bool ready = false;
    int rank = 0, world = 1;
    ncclComm_t comm = nullptr;
    int* d_token = nullptr;         // barrier token
    char* d_buf = nullptr;          // gather / combine scratch
    size_t buf_bytes = 0;
};
Comm g_comm;

int nccl_load() {
    NcclApi& a = g_comm.api;
    if (a.handle) return KZGB_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names)
        if ((h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!h) {
        set_error("multi-GPU entry points need NCCL: dlopen(libnccl.so.2) failed: %s", dlerror());
        return KZGB_ERR_INVALID_ARG;
    }
#define KZG_NCCL_SYM(field, sym)                                                     \
    *(void**)(&a.field) = dlsym(h, sym);                                             \
    if (!a.field) { set_error("NCCL symbol %s not found", sym); dlclose(h); return KZGB_ERR_INVALID_ARG; }
    KZG_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    KZG_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    KZG_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    KZG_NCCL_SYM(AllGather, "ncclAllGather")
    KZG_NCCL_SYM(AllReduce, "ncclAllReduce")
    KZG_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef KZG_NCCL_SYM
    a.handle = h;
    return KZGB_OK;
}

int nccl_fail(ncclResult_t r, const char* what, int line) {
    set_error("NCCL error %d (%s) at comm.cu:%d: %s", (int)r, g_comm.api.GetErrorString ? g_comm.api.GetErrorString(r) : "?",
              line, what);
    return KZGB_ERR_CUDA;
}
#define KZG_NCCL(expr)                                                 \
    do {                                                               \
        ncclResult_t _r = (expr);                                      \
        if (_r != ncclSuccess) return nccl_fail(_r, #expr, __LINE__);  \
    } while (0)

int comm_buf(size_t bytes, char** out) {
    Comm& c = g_comm;
    if (bytes > c.buf_bytes) {
        if (c.d_buf) {
            KZG_CUDA(cudaDeviceSynchronize());
            cudaFree(c.d_buf);
            c.d_buf = nullptr;
            c.buf_bytes = 0;
        }
        size_t want = bytes < ((size_t)1 << 16) ? ((size_t)1 << 16) : bytes;
        KZG_CUDA(cudaMalloc((void**)&c.d_buf, want));
        c.buf_bytes = want;
    }
    *out = c.d_buf;
    return KZGB_OK;
}

// stream-ordered barrier across the ranks: no host synchronisation
int comm_barrier(cudaStream_t st) {
    Comm& c = g_comm;
    if (c.world == 1) return KZGB_OK;
    KZG_NCCL(c.api.AllReduce(c.d_token, c.d_token, 1, ncclInt32, ncclSum, c.comm, st));
    return KZGB_OK;
}

size_t out_bytes_of(int mode) { return mode == 0 ? sizeof(kzgb_p1_affine) : (mode == 1 ? sizeof(kzgb_p1) : 192); }

// partials [items][192 B] on the device (this rank's) -> d_out[items] (out_mode layout) = sum over ranks
int gather_combine(const void* d_partials, uint32_t items, void* d_out, int out_mode, cudaStream_t st) {
    Comm& c = g_comm;
    const size_t pb = (size_t)items * 192;
    char* buf = nullptr;
    int rc = comm_buf((size_t)c.world * pb + pb + 256, &buf);
    if (rc) return rc;
    const void* gathered = d_partials;
    if (c.world > 1) {
        KZG_NCCL(c.api.AllGather(d_partials, buf, pb, ncclChar, c.comm, st));
        gathered = buf;
    }
    char* d_sum = buf + (size_t)c.world * pb;
    rc = msm_combine(gathered, (uint32_t)c.world, d_sum, st, items);
    if (rc) return rc;
    return msm_finish((const XYZZ*)d_sum, d_out, out_mode, st, items);
}

}  // namespace

void comm_shutdown() {
    Comm& c = g_comm;
    std::lock_guard<std::recursive_mutex> g(c.mu);
    if (c.ready && c.comm) c.api.CommDestroy(c.comm);
    if (c.d_token) cudaFree(c.d_token);
    if (c.d_buf) cudaFree(c.d_buf);
    c.comm = nullptr;
    c.d_token = nullptr;
    c.d_buf = nullptr;
    c.buf_bytes = 0;
    c.ready = false;
    c.rank = 0;
    c.world = 1;
}

// [R][C] -> [C][R] over 32-byte elements (two 16-byte halves), 32 x 32 tiles through shared memory
__global__ void __launch_bounds__(256) fr_transpose_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, uint32_t R,
                                                           uint32_t C) {
    __shared__ uint4 tile[2][32][33];
    const uint32_t c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const uint32_t tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (uint32_t j = ty; j < 32; j += 8) {
        const uint32_t r = r0 + j, c = c0 + tx;
        if (r < R && c < C) {
            const size_t e = ((size_t)r * C + c) * 2;
            tile[0][j][tx] = in[e];
            tile[1][j][tx] = in[e + 1];
        }
    }
    __syncthreads();
    for (uint32_t j = ty; j < 32; j += 8) {
        const uint32_t c = c0 + j, r = r0 + tx;
        if (r < R && c < C) {
            const size_t e = ((size_t)c * R + r) * 2;
            out[e] = tile[0][tx][j];
            out[e + 1] = tile[1][tx][j];
        }
    }
}

}  // namespace kzg

using namespace kzg;

// Plan of the fused four-step NTT for one (log_n, communicator): owns the IPC-shared row matrices.
struct kzgb_ntt4_plan {
    uint32_t log_n = 0, l1 = 0, l2 = 0, lc = 0;
    uint32_t world = 1, rank = 0;
    size_t nrows = 0, nbytes = 0;          // rows per rank, bytes of one row matrix = N/world elements
    void* local[2] = {nullptr, nullptr};
    void* peers[2][NTT_MAX_PEERS] = {};
    std::vector<void*> opened;
    uint64_t calls = 0;
    void* d_tmp = nullptr;                 // row-pass scratch (l2 > 10)
    // host-vector helpers
    void* d_slab = nullptr;
    void* d_t = nullptr;
    void* d_all = nullptr;
};

extern "C" {

KZGB_API int kzgb_comm_unique_id(unsigned char id[KZGB_COMM_ID_BYTES]) {
    static_assert(KZGB_COMM_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "id size");
    if (!id) { set_error("kzgb_comm_unique_id: null"); return KZGB_ERR_INVALID_ARG; }
    std::lock_guard<std::recursive_mutex> g(g_comm.mu);
    int rc = nccl_load();
    if (rc) return rc;
    ncclUniqueId u;
    KZG_NCCL(g_comm.api.GetUniqueId(&u));
    memcpy(id, u.internal, NCCL_UNIQUE_ID_BYTES);
    return KZGB_OK;
}

KZGB_API int kzgb_comm_init(int rank, int world, const unsigned char id[KZGB_COMM_ID_BYTES]) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    Comm& c = g_comm;
    std::lock_guard<std::recursive_mutex> g(c.mu);
    if (c.ready) {
        if (c.rank == rank && c.world == world) return KZGB_OK;
        set_error("kzgb_comm_init: already initialised as rank %d of %d", c.rank, c.world);
        return KZGB_ERR_INVALID_ARG;
    }
    if (world < 1 || rank < 0 || rank >= world || (world > 1 && !id) || world > NTT_MAX_PEERS) {
        set_error("kzgb_comm_init: bad rank/world %d/%d (world <= %d)", rank, world, NTT_MAX_PEERS);
        return KZGB_ERR_INVALID_ARG;
    }
    if (world > 1) {
        int rc = nccl_load();
        if (rc) return rc;
        ncclUniqueId u;
        memcpy(u.internal, id, NCCL_UNIQUE_ID_BYTES);
        KZG_NCCL(c.api.CommInitRank(&c.comm, world, u, rank));
    }
    KZG_CUDA(cudaMalloc((void**)&c.d_token, 256));
    KZG_CUDA(cudaMemset(c.d_token, 0, 256));
    c.rank = rank;
    c.world = world;
    c.ready = true;
    return KZGB_OK;
}

KZGB_API int kzgb_comm_destroy(void) {
    if (!ctx().initialised) return KZGB_OK;
    KZG_ENTER();
    cudaDeviceSynchronize();
    comm_shutdown();
    return KZGB_OK;
}

KZGB_API int kzgb_comm_rank(void) { return g_comm.rank; }
KZGB_API int kzgb_comm_world(void) { return g_comm.world; }

KZGB_API int kzgb_comm_barrier(void* stream) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    std::lock_guard<std::recursive_mutex> g(g_comm.mu);
    return comm_barrier((cudaStream_t)stream);
}

// ---- sharded MSM -------------------------------------------------------------------------------------------
KZGB_API int kzgb_msm_g1_sharded_dev(kzgb_srs* shard, const void* d_scalars, size_t m, void* d_out, int out_mode, void* stream) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    if (!shard || (!d_scalars && m) || !d_out || out_mode < 0 || out_mode > 2) {
        set_error("kzgb_msm_g1_sharded_dev: bad argument");
        return KZGB_ERR_INVALID_ARG;
    }
    std::lock_guard<std::recursive_mutex> g(g_comm.mu);
    KZG_LOCK_SRS(shard);
    cudaStream_t st = use_stream((cudaStream_t)stream);
    const XYZZ* res = nullptr;
    int rc = msm_run(shard, d_scalars, m, st, &res);
    if (rc) return rc;
    return gather_combine(res, 1, d_out, out_mode, st);
}

KZGB_API int kzgb_msm_g1_sharded(kzgb_srs* shard, const kzgb_fr* scalars, size_t m, kzgb_p1_affine* out) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    if (!shard || (!scalars && m) || !out) { set_error("kzgb_msm_g1_sharded: null argument"); return KZGB_ERR_INVALID_ARG; }
    if (m > shard->n) { set_error("kzgb_msm_g1_sharded: %zu scalars > %zu points in this shard", m, shard->n); return KZGB_ERR_INVALID_ARG; }
    std::lock_guard<std::recursive_mutex> g(g_comm.mu);
    KZG_LOCK_SRS(shard);
    cudaStream_t st = use_stream(lane().stream);
    const size_t bytes = m * sizeof(kzgb_fr);
    char* d = (char*)staging_reserve(bytes + 256);
    if (!d) return KZGB_ERR_CUDA;
    if (m) KZG_CUDA(upload_async(d + 256, scalars, bytes, st));
    int rc = kzgb_msm_g1_sharded_dev(shard, d + 256, m, d, 0, st);
    if (rc) return rc;
    KZG_CUDA(cudaMemcpyAsync(out, d, sizeof(kzgb_p1_affine), cudaMemcpyDeviceToHost, st));
    KZG_CUDA(cudaStreamSynchronize(st));
    return KZGB_OK;
}

// `count` sharded commits in one call: the local MSMs run through the pipelined batch (msm_batch_entry), the count
// partials travel in ONE all-gather (world x count x 192 B) and are combined side by side.
static int sharded_batch(kzgb_srs* shard, const void* const* scalars, const size_t* ms, size_t count, void* out, int out_mode,
                         bool scalars_dev, bool out_dev, cudaStream_t st) {
    if (count == 0) return KZGB_OK;
    // [partials count x 192][results count x 192] in the lane's scratch arena
    char* d = (char*)scratch_reserve(2 * count * 192 + 512);
    if (!d) return KZGB_ERR_CUDA;
    char* d_part = d;
    char* d_res = d + count * 192 + 256;
    int rc = msm_batch_entry(shard, scalars, ms, count, d_part, 2, (scalars_dev ? 1 : 0) | 2, st);
    if (rc) return rc;
    void* d_o = out_dev ? out : (void*)d_res;
    rc = gather_combine(d_part, (uint32_t)count, d_o, out_mode, st);
    if (rc) return rc;
    if (!out_dev) {
        KZG_CUDA(cudaMemcpyAsync(out, d_res, count * out_bytes_of(out_mode), cudaMemcpyDeviceToHost, st));
        KZG_CUDA(cudaStreamSynchronize(st));
    }
    return KZGB_OK;
}

KZGB_API int kzgb_msm_g1_sharded_batch(kzgb_srs* shard, const kzgb_fr* const* scalars, const size_t* ms, size_t count,
                                       kzgb_p1_affine* out) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    if (!shard || (count && (!scalars || !ms || !out))) { set_error("kzgb_msm_g1_sharded_batch: null argument"); return KZGB_ERR_INVALID_ARG; }
    for (size_t i = 0; i < count; i++)
        if (ms[i] && !scalars[i]) { set_error("kzgb_msm_g1_sharded_batch: item %zu is null", i); return KZGB_ERR_INVALID_ARG; }
    std::lock_guard<std::recursive_mutex> g(g_comm.mu);
    KZG_LOCK_SRS(shard);
    cudaStream_t st = use_stream(lane().stream);
    return sharded_batch(shard, (const void* const*)scalars, ms, count, out, 0, false, false, st);
}

KZGB_API int kzgb_msm_g1_sharded_batch_dev(kzgb_srs* shard, const void* const* d_scalars, const size_t* ms, size_t count,
                                           void* d_out, int out_mode, void* stream) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    if (!shard || out_mode < 0 || out_mode > 2 || (count && (!d_scalars || !ms || !d_out))) {
        set_error("kzgb_msm_g1_sharded_batch_dev: bad argument");
        return KZGB_ERR_INVALID_ARG;
    }
    for (size_t i = 0; i < count; i++)
        if (ms[i] && !d_scalars[i]) { set_error("kzgb_msm_g1_sharded_batch_dev: item %zu is null", i); return KZGB_ERR_INVALID_ARG; }
    std::lock_guard<std::recursive_mutex> g(g_comm.mu);
    KZG_LOCK_SRS(shard);
    cudaStream_t st = use_stream((cudaStream_t)stream);
    return sharded_batch(shard, d_scalars, ms, count, d_out, out_mode, true, true, st);
}

// ---- four-step NTT plan --------------------------------------------------------------------------------------
KZGB_API int kzgb_ntt4_plan_create(uint32_t log_n, uint32_t log_n1, kzgb_ntt4_plan** out) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    Comm& c = g_comm;
    std::lock_guard<std::recursive_mutex> g(c.mu);
    if (!out) { set_error("kzgb_ntt4_plan_create: null"); return KZGB_ERR_INVALID_ARG; }
    if (log_n >= 32) return KZGB_DEGREE_TOO_LARGE;   // ft.rs:67-69
    uint32_t lw = 0;
    while ((1u << lw) < (uint32_t)c.world) lw++;
    if ((1u << lw) != (uint32_t)c.world) { set_error("kzgb_ntt4_plan_create: world size %d is not a power of two", c.world); return KZGB_ERR_INVALID_ARG; }
    // N = N1 x N2, rank g holds C = N2/world columns; N1 <= 2^10 keeps the column transform in one shared-memory pass
    if (log_n1 == 0) log_n1 = log_n / 2 < 10 ? log_n / 2 : 10;
    if (log_n1 < lw) log_n1 = lw;
    if (log_n1 > 11 || log_n1 > log_n || log_n - log_n1 < lw) {
        set_error("kzgb_ntt4_plan_create: 2^%u is too small to split %d ways (N1 = 2^%u)", log_n, c.world, log_n1);
        return KZGB_ERR_INVALID_ARG;
    }
    kzgb_ntt4_plan* p = new kzgb_ntt4_plan();
    p->log_n = log_n;
    p->l1 = log_n1;
    p->l2 = log_n - log_n1;
    p->lc = p->l2 - lw;
    p->world = (uint32_t)c.world;
    p->rank = (uint32_t)c.rank;
    p->nrows = ((size_t)1 << p->l1) >> lw;
    p->nbytes = (p->nrows << p->l2) * 32;
    int rc = KZGB_OK;
    cudaError_t e = cudaSuccess;
    cudaIpcMemHandle_t mine[2];
    for (int k = 0; k < 2 && e == cudaSuccess; k++) {
        e = cudaMalloc(&p->local[k], p->nbytes);
        if (e == cudaSuccess && p->world > 1) e = cudaIpcGetMemHandle(&mine[k], p->local[k]);
    }
    if (e == cudaSuccess && p->l2 > 10) e = cudaMalloc(&p->d_tmp, p->nbytes);
    if (e != cudaSuccess) rc = cuda_fail(e, "ntt4 plan allocation", __FILE__, __LINE__);
    std::vector<cudaIpcMemHandle_t> all(2 * (size_t)p->world);
    if (!rc && p->world > 1) {   // exchange the IPC handles: all-gather of 2 x 64 bytes per rank
        char* buf = nullptr;
        const size_t hb = 2 * sizeof(cudaIpcMemHandle_t);
        rc = comm_buf(hb * (p->world + 1), &buf);
        cudaStream_t st = lane().stream;
        if (!rc && (e = cudaMemcpyAsync(buf, mine, hb, cudaMemcpyHostToDevice, st)) != cudaSuccess) rc = cuda_fail(e, "H2D handles", __FILE__, __LINE__);
        if (!rc) {
            ncclResult_t r = c.api.AllGather(buf, buf + hb, hb, ncclChar, c.comm, st);
            if (r != ncclSuccess) rc = nccl_fail(r, "all-gather of IPC handles", __LINE__);
        }
        if (!rc && ((e = cudaMemcpyAsync(all.data(), buf + hb, hb * p->world, cudaMemcpyDeviceToHost, st)) != cudaSuccess ||
                    (e = cudaStreamSynchronize(st)) != cudaSuccess))
            rc = cuda_fail(e, "D2H handles", __FILE__, __LINE__);
    }
    for (int k = 0; k < 2 && !rc; k++) {
        for (uint32_t r = 0; r < p->world && !rc; r++) {
            if (r == p->rank) { p->peers[k][r] = p->local[k]; continue; }
            void* q = nullptr;
            e = cudaIpcOpenMemHandle(&q, all[2 * r + k], cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) { rc = cuda_fail(e, "cudaIpcOpenMemHandle (peer row matrix)", __FILE__, __LINE__); break; }
            p->opened.push_back(q);
            p->peers[k][r] = q;
        }
    }
    if (rc) {
        for (void* q : p->opened) cudaIpcCloseMemHandle(q);
        for (int k = 0; k < 2; k++) if (p->local[k]) cudaFree(p->local[k]);
        if (p->d_tmp) cudaFree(p->d_tmp);
        delete p;
        return rc;
    }
    *out = p;
    return KZGB_OK;
}

KZGB_API int kzgb_ntt4_plan_destroy(kzgb_ntt4_plan* p) {
    if (!p) return KZGB_OK;
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    std::lock_guard<std::recursive_mutex> g(g_comm.mu);
    cudaStream_t st = lane().stream;
    // nobody may still be storing into a matrix that is about to go away
    KZG_CUDA(cudaDeviceSynchronize());
    int rc = comm_barrier(st);
    if (rc) return rc;
    KZG_CUDA(cudaStreamSynchronize(st));
    for (void* q : p->opened) cudaIpcCloseMemHandle(q);
    rc = comm_barrier(st);
    if (rc) return rc;
    KZG_CUDA(cudaStreamSynchronize(st));
    for (int k = 0; k < 2; k++) if (p->local[k]) cudaFree(p->local[k]);
    if (p->d_tmp) cudaFree(p->d_tmp);
    if (p->d_slab) cudaFree(p->d_slab);
    if (p->d_t) cudaFree(p->d_t);
    if (p->d_all) cudaFree(p->d_all);
    delete p;
    return KZGB_OK;
}

KZGB_API int kzgb_ntt4_plan_shape(const kzgb_ntt4_plan* p, uint32_t* log_n1, uint32_t* log_n2, uint32_t* log_c) {
    if (!p) { set_error("kzgb_ntt4_plan_shape: null"); return KZGB_ERR_INVALID_ARG; }
    if (log_n1) *log_n1 = p->l1;
    if (log_n2) *log_n2 = p->l2;
    if (log_c) *log_c = p->lc;
    return KZGB_OK;
}

// d_slab: this rank's [N1][C] column slab (slab[i1][c] = a[i1*N2 + rank*C + c]); only read.
// *d_rows_out: this rank's [N1/world][N2] block of the result, rows[r][k2] = X[(rank*N1/world + r) + N1*k2]
// (valid until the plan's next-but-one call).  Same values as the single-GPU transform / ft.rs:111-140.
KZGB_API int kzgb_ntt_fr_sharded_dev(kzgb_ntt4_plan* p, const void* d_slab, int inverse, void** d_rows_out, void* stream) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    if (!p || !d_slab || !d_rows_out) { set_error("kzgb_ntt_fr_sharded_dev: null argument"); return KZGB_ERR_INVALID_ARG; }
    std::lock_guard<std::recursive_mutex> g(g_comm.mu);
    cudaStream_t st = use_stream((cudaStream_t)stream);
    const int k = (int)(p->calls & 1);
    p->calls++;
    uint32_t lw = 0;
    while ((1u << lw) < p->world) lw++;
    NttScatter sc{};
    for (uint32_t r = 0; r < p->world; r++) sc.peer[r] = (Fr*)p->peers[k][r];
    sc.log_rpr = p->l1 - lw;
    sc.rank = p->rank;
    sc.log_world = lw;
    sc.log_c = p->lc;
    int rc = ntt_columns_run((Fr*)d_slab, p->log_n, p->l1, p->lc, p->rank << p->lc, inverse != 0, st, &sc);
    if (rc) return rc;
    rc = comm_barrier(st);
    if (rc) return rc;
    Fr scale;
    if (inverse) {
        FrHost mi = fr_host_inv(fr_host_from_u64(1ull << p->log_n));
        memcpy(scale.l, mi.l, 32);
    }
    rc = ntt_run((Fr*)p->local[k], (Fr*)p->d_tmp, p->l2, inverse != 0, st, (uint32_t)p->nrows, inverse ? &scale : nullptr);
    if (rc) return rc;
    *d_rows_out = p->local[k];
    return KZGB_OK;
}

// Host vectors, natural order in and out (the drop-in shape of EvaluationDomain::fft/ifft, ft.rs:111-140, issued by
// every rank of the job): `in` is the full 2^log_n vector (this rank reads only its columns), `out` the full-size
// result vector.  gather == 0: only this rank's elements of `out` are written (indices (rank*N1/world + r) + N1*k2);
// gather != 0: the blocks of all ranks are all-gathered over NCCL and every rank receives the complete vector.
KZGB_API int kzgb_ntt_fr_sharded(kzgb_ntt4_plan* p, const kzgb_fr* in, kzgb_fr* out, int inverse, int gather) {
    KZG_REQUIRE_INIT();
    KZG_ENTER();
    if (!p || !in || !out) { set_error("kzgb_ntt_fr_sharded: null argument"); return KZGB_ERR_INVALID_ARG; }
    Comm& c = g_comm;
    std::lock_guard<std::recursive_mutex> g(c.mu);
    cudaStream_t st = use_stream(lane().stream);
    const size_t n1 = (size_t)1 << p->l1, n2 = (size_t)1 << p->l2, C = (size_t)1 << p->lc;
    if (!p->d_slab) KZG_CUDA(cudaMalloc(&p->d_slab, p->nbytes));
    if (!p->d_t) KZG_CUDA(cudaMalloc(&p->d_t, p->nbytes));
    if (gather && p->world > 1 && !p->d_all) KZG_CUDA(cudaMalloc(&p->d_all, p->nbytes * p->world));
    KZG_CUDA(cudaMemcpy2DAsync(p->d_slab, C * 32, (const char*)in + (size_t)p->rank * C * 32, n2 * 32, C * 32, n1,
                               cudaMemcpyHostToDevice, st));
    void* rows = nullptr;
    int rc = kzgb_ntt_fr_sharded_dev(p, p->d_slab, inverse, &rows, st);
    if (rc) return rc;
    // rows[r][k2] -> t[k2][r]: runs of N1/world consecutive natural-order elements
    dim3 grid((unsigned)((n2 + 31) / 32), (unsigned)((p->nrows + 31) / 32));
    fr_transpose_kernel<<<grid, 256, 0, st>>>((const uint4*)rows, (uint4*)p->d_t, (uint32_t)p->nrows, (uint32_t)n2);
    KZG_LAUNCHED();
    const size_t run = p->nrows * 32;
    if (gather && p->world > 1) {
        KZG_NCCL(c.api.AllGather(p->d_t, p->d_all, p->nbytes, ncclChar, c.comm, st));
        for (uint32_t h = 0; h < p->world; h++)
            KZG_CUDA(cudaMemcpy2DAsync((char*)out + h * run, n1 * 32, (const char*)p->d_all + h * p->nbytes, run, run, n2,
                                       cudaMemcpyDeviceToHost, st));
    } else {
        KZG_CUDA(cudaMemcpy2DAsync((char*)out + p->rank * run, n1 * 32, p->d_t, run, run, n2, cudaMemcpyDeviceToHost, st));
    }
    KZG_CUDA(cudaStreamSynchronize(st));
    return KZGB_OK;
}

}  // extern "C"
