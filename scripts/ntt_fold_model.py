"""Index-level model of the multi-pass NTT of kzg_b200/csrc/ntt.cu with the inter-pass twiddles FOLDED into the
butterflies of the following pass (coset sub-transforms): every pass j >= 2 transforms its digit with the shifted
root g = omega_N^(S_j * kappa_j), kappa_j = k_1 + R_1 k_2 + ... (the output digits already fixed), so that stage s
(m = 2^s) of pass j uses the twiddle exponent
        e = j' * N/(2m) + kappa_j * N_j/(2m),      N_j = R_j * S_j,
and no separate multiplication by omega_N^(k_j * low) remains.  Checked here against the naive DFT over a small prime
field for all pass plans of small sizes (same loop structure as the kernel: bit-reversed rows, DIT stages, last pass
writes digit-reversed) before the CUDA kernel was changed."""
import itertools
import random

P = 257          # prime with 2^8 | P - 1
G = 3            # generator of F_257^*


def root(n):
    return pow(G, (P - 1) // n, P)


def naive(a, w):
    n = len(a)
    return [sum(a[j] * pow(w, j * k, P) for j in range(n)) % P for k in range(n)]


def brev(x, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


def ntt_folded(a, log_r, inverse=False):
    n_log = sum(log_r)
    N = 1 << n_log
    w = root(N)
    if inverse:
        w = pow(w, -1, P)
    tw = lambda e: pow(w, e % N, P)
    cur = list(a)
    npass = len(log_r)
    log_s = n_log
    for jp, lr in enumerate(log_r):
        R = 1 << lr
        log_s -= lr
        S = 1 << log_s
        log_nj = lr + log_s
        last = jp == npass - 1
        out = [0] * N
        for hi in range(N >> log_nj):
            for low in range(S):
                # kappa: output digits already fixed.  non-last pass: kappa = hi (valid for pass 1 (hi = 0) and pass 2 of 3);
                # last pass: kappa = k1 + R1 * k2 read from the digit positions
                if not last:
                    kappa = hi
                else:
                    if npass == 1:
                        kappa = 0
                    elif npass == 2:
                        kappa = hi                      # hi = k1
                    else:
                        r1, r2 = log_r[0], log_r[1]
                        k1, k2 = hi >> r2, hi & ((1 << r2) - 1)
                        kappa = k1 + (k2 << r1)
                base = (hi << log_nj) + low
                rows = [0] * R
                for d in range(R):
                    rows[brev(d, lr)] = cur[base + (d << log_s)]
                for s in range(lr):
                    m = 1 << s
                    for k in range(0, R, 2 * m):
                        for j in range(m):
                            e = (j << (n_log - s - 1)) + ((kappa << log_nj) >> (s + 1))
                            t = rows[k + j + m] * tw(e) % P
                            u = rows[k + j]
                            rows[k + j], rows[k + j + m] = (u + t) % P, (u - t) % P
                for kj in range(R):
                    if not last:
                        out[base + (kj << log_s)] = rows[kj]          # NO inter-pass twiddle
                    else:
                        # digit-reversed write: k1 + R1*k2 + R1*R2*k3
                        out[kappa + (kj << (n_log - lr))] = rows[kj]
        cur = out
    if inverse:
        ninv = pow(N, -1, P)
        cur = [v * ninv % P for v in cur]
    return cur


def main():
    random.seed(1)
    checked = 0
    for n_log in range(1, 9):
        N = 1 << n_log
        a = [random.randrange(P) for _ in range(N)]
        want = naive(a, root(N))
        for npass in (1, 2, 3):
            for plan in itertools.product(range(1, n_log + 1), repeat=npass):
                if sum(plan) != n_log:
                    continue
                got = ntt_folded(a, list(plan))
                assert got == want, (n_log, plan)
                assert ntt_folded(got, list(plan), inverse=True) == a, (n_log, plan, "inverse")
                checked += 1
    print("folded multi-pass NTT == naive DFT for", checked, "plans")


if __name__ == "__main__":
    main()
