import sys, numpy as np
sys.path.insert(0, ".")
import libmpc_b200 as L
from libmpc_b200 import workloads as W
B, PH = int(sys.argv[1]), 20
x0, r = W.quadrotor_inputs(0, B)
yref = np.zeros((B, 12, PH)); yref[:, 2, :] = r[:, None]
c = W.build_quadrotor_controller(L, PH, B, 250)
c.set_engine(2, 0)
c.setReferences(yref, np.zeros((4, PH)), np.zeros((4, PH)))
for _ in range(2):
    out = c.optimize(x0, np.zeros((B, 4)))
print(out.iterations.mean())
