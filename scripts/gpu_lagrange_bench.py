import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kzg_b200 as kzg
from kzg_b200 import _ffi
from oracle import kzg_oracle as O
_ffi.init(0)
S = O.fr_stream(69, 1)[0]
for lg in [int(x) for x in sys.argv[1:]] or [12, 16]:
    d = 1 << lg
    params = kzg.setup(S, d)
    t = time.time(); lag = kzg.compute_lagrange_basis(params); dt = time.time() - t
    exp = kzg.compute_lagrange_basis_known_secret(S, d)
    ok = bool((lag.read_affine() == exp.read_affine()).all())
    print("compute_lagrange_basis d=2^%d: %.3f s parity=%s" % (lg, dt, ok), flush=True)
