import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as onp
import ramba_b200 as np
m = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dt = onp.float64
U = np.fromfunction(lambda i, j, k: (i + 2 * j + 3 * k) % 64, (m, m, m), dtype=dt)
V = np.zeros((m, m, m), dtype=dt)
V[1:-1, 1:-1, 1:-1] = (U[:-2, 1:-1, 1:-1] + U[2:, 1:-1, 1:-1] + U[1:-1, :-2, 1:-1] + U[1:-1, 2:, 1:-1]
                       + U[1:-1, 1:-1, :-2] + U[1:-1, 1:-1, 2:] - 6.0 * U[1:-1, 1:-1, 1:-1])
np.sync()
for z0 in (0, m // 2 - 3, m - 8):
    sub = V[z0:z0 + 8].asarray()
    i, j, k = onp.meshgrid(onp.arange(z0 - 1, z0 + 9), onp.arange(m), onp.arange(m), indexing="ij")
    u = ((i + 2 * j + 3 * k) % 64).astype(dt)
    ref = onp.zeros((8, m, m), dtype=dt)
    ref[:, 1:-1, 1:-1] = (u[:-2, 1:-1, 1:-1] + u[2:, 1:-1, 1:-1] + u[1:-1, :-2, 1:-1] + u[1:-1, 2:, 1:-1] + u[1:-1, 1:-1, :-2] + u[1:-1, 1:-1, 2:] - 6.0 * u[1:-1, 1:-1, 1:-1])
    if z0 == 0: ref[0] = 0
    if z0 + 8 == m: ref[-1] = 0
    bad = onp.argwhere(sub != ref)
    print("z0", z0, "bad", len(bad), bad[:5].tolist(), [(float(sub[tuple(b)]), float(ref[tuple(b)])) for b in bad[:5]])
    uu = U[z0:z0+8].asarray()
    print("  U ok:", onp.array_equal(uu, u[1:-1] if z0 > 0 else u[1:-1]))
