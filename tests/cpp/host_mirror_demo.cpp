// Compile check (and, on a GPU box, a runnable demo) of the C++ host mirror: commit + create_witness of a small
// polynomial, cross-checked through the library itself (witness of a point NOT on the polynomial must fail).
#include <cstdio>
#include <cstring>

#include "../../include/kzg_b200.hpp"

using namespace kzg_b200;

int main() {
    if (kzgb_init(0) != KZGB_OK) {
        std::printf("no B200 available: %s\n", kzgb_last_error_message());
        return 2;
    }
    // Montgomery images of small integers via the library's own to_mont test hook
    std::vector<Scalar> canon(8), mont(8);
    for (size_t i = 0; i < canon.size(); i++) canon[i] = Scalar{{i + 3, 0, 0, 0}};
    check(kzgb_test_fr_op(canon.data(), nullptr, mont.data(), canon.size(), 4), "to_mont");
    KZGParams params = KZGParams::setup(mont[7], 16);
    KZGProver prover(params);
    Polynomial p = Polynomial::make(std::vector<Scalar>(mont.begin(), mont.begin() + 6));
    G1Affine c = prover.commit(p);
    Scalar x = mont[1], y{}, q[8];
    check(kzgb_divide_linear(p.slice_coeffs(), p.num_coeffs(), &x, nullptr, q, &y), "eval");  // y = p(x)
    auto w = prover.create_witness(p, x, y);
    auto bad = prover.create_witness(p, x, mont[2]);
    std::printf("commit.x[0]=%016llx witness ok=%d bad rejected=%d\n", (unsigned long long)c.x.l[0], (int)w.ok(),
                (int)(!bad.ok() && bad.error == KZGError::PointNotOnPolynomial));
    // the pipelined batch entry points give the bytes of the single calls
    Polynomial p2 = Polynomial::make(std::vector<Scalar>(mont.begin() + 1, mont.begin() + 5));
    std::vector<G1Affine> many = prover.commit_many({&p, &p2, &p});
    G1Affine c2 = prover.commit(p2);
    bool batch_ok = std::memcmp(&many[0], &c, sizeof(c)) == 0 && std::memcmp(&many[1], &c2, sizeof(c2)) == 0 &&
                    std::memcmp(&many[2], &c, sizeof(c)) == 0;
    Scalar x2 = mont[3], y2{};
    check(kzgb_divide_linear(p.slice_coeffs(), p.num_coeffs(), &x2, nullptr, q, &y2), "eval");
    auto wm = prover.create_witness_many(p, {x, x2}, {y, y2});
    auto w2 = prover.create_witness(p, x2, y2);
    bool many_ok = wm.ok() && w2.ok() && std::memcmp(&(*wm.value)[0], &*w.value, sizeof(G1Affine)) == 0 &&
                   std::memcmp(&(*wm.value)[1], &*w2.value, sizeof(G1Affine)) == 0;
    auto wm_bad = prover.create_witness_many(p, {x, x2}, {y, mont[2]});
    std::printf("commit_many ok=%d create_witness_many ok=%d bad rejected=%d\n", (int)batch_ok, (int)many_ok, (int)!wm_bad.ok());
    kzgb_shutdown();
    return (w.ok() && !bad.ok() && batch_ok && many_ok && !wm_bad.ok()) ? 0 : 1;
}
