// Compile check (and, on a GPU box, a runnable demo) of the C++ host mirror: commit + create_witness of a small
// polynomial, cross-checked through the library itself (witness of a point NOT on the polynomial must fail).
#include <cstdio>

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
    kzgb_shutdown();
    return (w.ok() && !bad.ok()) ? 0 : 1;
}
