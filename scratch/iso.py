import sys, numpy as np
sys.path.insert(0, '.')
import __graft_entry__ as g
pkg = g.package()
sys.path.insert(0, 'tests')
from conftest import smooth_box
which = sys.argv[1]; n = int(sys.argv[2]); path = int(sys.argv[3]) if len(sys.argv) > 3 else 2
rng = np.random.default_rng(1)
d = (32,32,32) if which == '1' else (64,64,64)
base = [smooth_box(d, rng, dtype=np.float64, sym=bool(i%2)) for i in range(4)]
boxes = [base[i%4]*(1+0.001*i) for i in range(n)]
ctx = pkg.Context(0); ctx.set_path(path)
try:
    p = ctx.compress_batch(boxes, float(np.float32(0.999)))
    print(which, n, path, 'ok', sum(x.npairs for x in p))
except Exception as e:
    print(which, n, path, 'FAIL', str(e)[-60:])
