// threads_check.cpp -- concurrent provers: 4 host threads call kzgb_msm_g1 on two SRS handles at the same time.
// The reference's provers are Send + Sync plain structs (coeff_form.rs:37-41) and `cargo test` / the `parallel`
// feature call in from worker threads; the library must (a) return the bytes of the sequential calls and (b) let
// calls on different handles overlap on the device (per-handle locks + lanes instead of one process-wide mutex).
// Prints the sequential and concurrent wall times; exit code 0 = all results bit-equal.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/kzg_b200.h"

static void fill_fr(std::vector<kzgb_fr>& v, uint64_t seed) {
    uint64_t s = seed;
    auto next = [&]() {
        uint64_t z = (s += 0x9e3779b97f4a7c15ULL);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
        return z ^ (z >> 31);
    };
    for (auto& e : v) {
        for (int i = 0; i < 4; i++) e.l[i] = next();
        e.l[3] &= 0x3fffffffffffffffULL;
    }
}

int main(int argc, char** argv) {
    const unsigned log_n = argc > 1 ? (unsigned)std::atoi(argv[1]) : 16;
    const int K = argc > 2 ? std::atoi(argv[2]) : 8;
    if (kzgb_init(0) != KZGB_OK) { std::printf("no B200 available: %s\n", kzgb_last_error_message()); return 2; }
    const size_t n = (size_t)1 << log_n;
    kzgb_fr sec[2] = {{{3, 5, 7, 11}}, {{13, 17, 19, 23}}};
    kzgb_srs* srs[2] = {nullptr, nullptr};
    for (int h = 0; h < 2; h++)
        if (kzgb_srs_generate(0, &sec[h], n, 0, n, 0, &srs[h]) != KZGB_OK) { std::printf("srs: %s\n", kzgb_last_error_message()); return 1; }
    const int T = 4;
    std::vector<std::vector<kzgb_fr>> sc(T, std::vector<kzgb_fr>(n));
    for (int t = 0; t < T; t++) fill_fr(sc[t], 100 + t);
    std::vector<kzgb_p1_affine> want(T), got((size_t)T * K);
    // sequential pass (also warms every arena)
    for (int t = 0; t < T; t++) kzgb_msm_g1(srs[t & 1], sc[t].data(), n, &want[t]);
    auto t0 = std::chrono::steady_clock::now();
    for (int k = 0; k < K; k++)
        for (int t = 0; t < T; t++) kzgb_msm_g1(srs[t & 1], sc[t].data(), n, &want[t]);
    const double seq_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    // concurrent pass: threads 0,2 share handle 0, threads 1,3 share handle 1
    int bad = 0;
    auto work = [&](int t) {
        for (int k = 0; k < K; k++)
            if (kzgb_msm_g1(srs[t & 1], sc[t].data(), n, &got[(size_t)t * K + k]) != KZGB_OK) __sync_fetch_and_add(&bad, 1);
    };
    for (int warm = 0; warm < 2; warm++) {   // first round warms the extra lanes' arenas
        t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++) th.emplace_back(work, t);
        for (auto& x : th) x.join();
    }
    const double par_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    for (int t = 0; t < T; t++)
        for (int k = 0; k < K; k++)
            if (std::memcmp(&got[(size_t)t * K + k], &want[t], sizeof(kzgb_p1_affine)) != 0) bad++;
    std::printf("threads_check: n=2^%u, %d threads x %d calls on 2 handles: sequential %.2f ms, concurrent %.2f ms "
                "(overlap %.2fx), mismatches %d\n", log_n, T, K, seq_ms, par_ms, seq_ms / par_ms, bad);
    kzgb_srs_destroy(srs[0]);
    kzgb_srs_destroy(srs[1]);
    kzgb_shutdown();
    std::printf("threads_check: %s\n", bad == 0 ? "PASS" : "FAIL");
    return bad == 0 ? 0 : 1;
}
