import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import libmpc_b200 as L
from user_systems import OUTPUT_MAP_SRC, output_map_formulation
f = output_map_formulation()
sid = L.register_system(OUTPUT_MAP_SRC, "UserWithOutput")
rng = np.random.default_rng(5)
z = rng.standard_normal((3, f.nz)) * 0.6; z[:, -1] = 0.0
x0 = rng.uniform(-0.5, 0.5, (3, 2))
out = L.nlmpc_eval(sid, f.ph, f.ch, z, x0, f.params)
worst = 0.0
for b in range(3):
    fv, g = f.objective(z[b], x0[b]); ci, Ji = f.ineq_con(z[b], x0[b]); c, J = f.state_eq(z[b], x0[b])
    worst = max(worst, abs(out["f"][b] - fv) / max(1, abs(fv)), np.abs(out["cin"][b] - ci).max(), np.abs(out["grad"][b] - g).max() / max(1, abs(fv)) * 1e-6 / 5e-7,
                np.abs(out["Jin"][b] - Ji).max() * 1e-6 / 5e-7, np.abs(out["Jeq"][b] - J).max() * 1e-6 / 5e-7)
    assert abs(out["f"][b] - fv) <= 1e-12 * max(1, abs(fv)) and np.allclose(out["cin"][b], ci, rtol=1e-12, atol=1e-13)
    assert np.allclose(out["grad"][b], g, rtol=1e-6, atol=5e-7 * max(1.0, abs(fv))) and np.allclose(out["Jin"][b], Ji, rtol=1e-6, atol=5e-7)
print("OUTPUT MAP OK", worst)
