"""Phase split of the WINDOW kernel (build with -DUWBGO_WIN_TIMING, select with UWBGO_LIB)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from localization_b200 import Config, Solver, synthetic
s = Solver(0)
s.set_window_path(1 << 30)
for name, mk, it in (("uwb_only N=10", lambda: synthetic.uwb_only(1, 10, 4, seed=3), 10),
                     ("uwb_imu N=12 (C2)", lambda: synthetic.uwb_imu_lidar(1, 12, 4, v_max=3.0, antennas=0, lidar=False, seed=11), 10),
                     ("uwb_imu_lidar N=20", lambda: synthetic.uwb_imu_lidar(1, 20, 8, seed=2), 20),
                     ("uwb_twist N=15", lambda: synthetic.uwb_twist(1, 15, 8, seed=5), 12)):
    topo, batch, _ = mk()
    print(name, "E", topo.n_edges, flush=True)
    for _ in range(2):
        s.solve(topo, batch, Config(max_iterations=it))
