// multi_gpu_check.cpp -- C++ (no Python, no PyTorch) multi-process parity test of the multi-GPU C ABI:
// `world` processes, one per GPU, are started with the same id-file path; each initialises its GPU (device = rank),
// joins the communicator and runs
//   * kzgb_msm_g1_sharded / _batch on its shard of a 2^log_n-point SRS  == kzgb_msm_g1 on the whole SRS (rank 0),
//   * kzgb_ntt_fr_sharded (fused four-step)                             == kzgb_ntt_fr on one GPU (rank 0),
// byte for byte.  The sharded calls are issued from a second host thread, so the library's per-call device selection is
// exercised on every rank > 0 (the calling thread's current CUDA device is 0 there).
// Reference calls being served: KZGProver::commit (coeff_form.rs:59-64), EvaluationDomain::fft/ifft (ft.rs:111-140).
// Usage: multi_gpu_check <rank> <world> <id_file> [log_n_msm = 16] [log_n_ntt = 16]
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/kzg_b200.h"

static uint64_t sm_state;
static uint64_t splitmix() {
    uint64_t z = (sm_state += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
// any 256-bit pattern below 2^254 (< r) is the Montgomery image of some field element
static void fill_fr(std::vector<kzgb_fr>& v, uint64_t seed) {
    sm_state = seed;
    for (auto& e : v) {
        for (int i = 0; i < 4; i++) e.l[i] = splitmix();
        e.l[3] &= 0x3fffffffffffffffULL;
    }
}
#define CHECK(call)                                                                              \
    do {                                                                                         \
        int _st = (call);                                                                        \
        if (_st != KZGB_OK) {                                                                    \
            std::fprintf(stderr, "[rank %d] %s -> %d: %s\n", rank, #call, _st, kzgb_last_error_message()); \
            return 1;                                                                            \
        }                                                                                        \
    } while (0)

static int run(int rank, int world, const char* id_file, unsigned log_n, unsigned log_ntt) {
    CHECK(kzgb_init(rank));
    unsigned char id[KZGB_COMM_ID_BYTES];
    if (rank == 0) {
        CHECK(kzgb_comm_unique_id(id));
        std::string tmp = std::string(id_file) + ".tmp";
        FILE* f = std::fopen(tmp.c_str(), "wb");
        if (!f || std::fwrite(id, 1, sizeof(id), f) != sizeof(id)) return 1;
        std::fclose(f);
        std::rename(tmp.c_str(), id_file);
    } else {
        FILE* f = nullptr;
        for (int tries = 0; tries < 3000 && !(f = std::fopen(id_file, "rb")); tries++) usleep(10000);
        if (!f || std::fread(id, 1, sizeof(id), f) != sizeof(id)) { std::fprintf(stderr, "[rank %d] no id file\n", rank); return 1; }
        std::fclose(f);
    }
    CHECK(kzgb_comm_init(rank, world, id));

    // ---- sharded MSM ----
    const size_t n = (size_t)1 << log_n, per = n / world, lo = (size_t)rank * per;
    kzgb_fr secret = {{0x123456789abcdef1ULL, 0x0fedcba987654321ULL, 0x1111222233334444ULL, 0x0123456789abcdefULL}};
    kzgb_srs* shard = nullptr;
    CHECK(kzgb_srs_generate(0, &secret, n, lo, per, 0, &shard));
    std::vector<kzgb_fr> sc(n), sc2(n);
    fill_fr(sc, 42);
    fill_fr(sc2, 43);
    kzgb_p1_affine got{}, got_b[3]{};
    const kzgb_fr* items[3] = {sc.data() + lo, sc2.data() + lo, sc.data() + lo};
    const size_t ms[3] = {per, per - 5, 0};
    CHECK(kzgb_msm_g1_sharded(shard, sc.data() + lo, per, &got));
    CHECK(kzgb_msm_g1_sharded_batch(shard, items, ms, 3, got_b));
    int ok = 1;
    if (rank == 0) {
        kzgb_srs* full = nullptr;
        CHECK(kzgb_srs_generate(0, &secret, n, 0, n, 0, &full));
        kzgb_p1_affine want{}, want1{}, part{};
        CHECK(kzgb_msm_g1(full, sc.data(), n, &want));
        ok &= std::memcmp(&want, &got, sizeof(want)) == 0;
        ok &= std::memcmp(&want, &got_b[0], sizeof(want)) == 0;
        // item 1: every rank contributes its first per-5 scalars: zero the 5 last scalars of each shard in a copy
        std::vector<kzgb_fr> z(sc2);
        for (int r = 0; r < world; r++)
            for (size_t i = per - 5; i < per; i++) std::memset(&z[(size_t)r * per + i], 0, sizeof(kzgb_fr));
        CHECK(kzgb_msm_g1(full, z.data(), n, &want1));
        ok &= std::memcmp(&want1, &got_b[1], sizeof(want1)) == 0;
        std::memset(&part, 0, sizeof(part));
        ok &= std::memcmp(&part, &got_b[2], sizeof(part)) == 0;   // empty commit = identity (all-zero affine)
        std::printf("[rank 0] MSM 2^%u over %d GPUs: sharded == single-GPU: %s\n", log_n, world, ok ? "OK" : "MISMATCH");
        CHECK(kzgb_srs_destroy(full));
    }
    // ---- four-step NTT ----
    const size_t N = (size_t)1 << log_ntt;
    std::vector<kzgb_fr> a(N), X(N), back(N);
    fill_fr(a, 7);
    kzgb_ntt4_plan* plan = nullptr;
    CHECK(kzgb_ntt4_plan_create(log_ntt, 0, &plan));
    CHECK(kzgb_ntt_fr_sharded(plan, a.data(), X.data(), 0, 1));
    CHECK(kzgb_ntt_fr_sharded(plan, X.data(), back.data(), 1, 1));
    int ok_ntt = std::memcmp(back.data(), a.data(), N * sizeof(kzgb_fr)) == 0;   // ifft(fft(a)) == a (canonical inputs)
    // timing of the device-resident fused path (what bench.py reports), 10 calls
    if (rank == 0) {
        std::vector<kzgb_fr> ref(a);
        CHECK(kzgb_ntt_fr(ref.data(), log_ntt, 0));
        ok_ntt &= std::memcmp(ref.data(), X.data(), N * sizeof(kzgb_fr)) == 0;
        std::printf("[rank 0] NTT 2^%u over %d GPUs: four-step == single-GPU, ifft(fft) == id: %s\n", log_ntt, world,
                    ok_ntt ? "OK" : "MISMATCH");
    }
    // own-block-only variant writes exactly this rank's elements
    std::vector<kzgb_fr> own(N);
    std::memset(own.data(), 0xff, N * sizeof(kzgb_fr));
    CHECK(kzgb_ntt_fr_sharded(plan, a.data(), own.data(), 0, 0));
    uint32_t l1 = 0, l2 = 0, lc = 0;
    CHECK(kzgb_ntt4_plan_shape(plan, &l1, &l2, &lc));
    const size_t n1 = (size_t)1 << l1, rows = n1 / world;
    for (size_t i = 0; i < N && ok_ntt; i++) {
        const bool mine = (i % n1) / rows == (size_t)rank;
        static const kzgb_fr ones = {{~0ULL, ~0ULL, ~0ULL, ~0ULL}};
        if (mine) ok_ntt &= std::memcmp(&own[i], &X[i], sizeof(kzgb_fr)) == 0;
        else ok_ntt &= std::memcmp(&own[i], &ones, sizeof(kzgb_fr)) == 0;
    }
    if (!ok_ntt) std::fprintf(stderr, "[rank %d] NTT block check failed\n", rank);
    CHECK(kzgb_ntt4_plan_destroy(plan));
    CHECK(kzgb_srs_destroy(shard));
    CHECK(kzgb_comm_destroy());
    CHECK(kzgb_shutdown());
    return (ok && ok_ntt) ? 0 : 1;
}

int main(int argc, char** argv) {
    if (argc < 4) { std::fprintf(stderr, "usage: %s <rank> <world> <id_file> [log_n_msm] [log_n_ntt]\n", argv[0]); return 2; }
    const int rank = std::atoi(argv[1]), world = std::atoi(argv[2]);
    const unsigned log_n = argc > 4 ? (unsigned)std::atoi(argv[4]) : 16, log_ntt = argc > 5 ? (unsigned)std::atoi(argv[5]) : 16;
    int rc = 1;
    std::thread t([&] { rc = run(rank, world, argv[3], log_n, log_ntt); });   // NOT the main thread: see the header comment
    t.join();
    if (rank == 0) std::printf("multi_gpu_check: %s\n", rc == 0 ? "PASS" : "FAIL");
    return rc;
}
