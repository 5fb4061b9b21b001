"""Small solves of every kernel path for compute-sanitizer runs."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from localization_b200 import Config, Solver, synthetic, Topology, Batch
from oracle import oracle
s = Solver(0)
cases = [("chain_ws", synthetic.uwb_only(70, 9, 4, seed=1), 4),
         ("imu_lidar", synthetic.uwb_imu_lidar(40, 6, 4, seed=2), 3),
         ("twist", synthetic.uwb_twist(40, 6, 4, seed=3), 3),
         ("pose_star", synthetic.uwb_pose(40, 9, 4, keyframe_len=3, seed=4), 3)]
topo, batch, _ = synthetic.uwb_only(40, 7, 4, seed=5)
e = np.stack([topo.edge_kind, topo.edge_a, topo.edge_b, topo.edge_ant, topo.edge_robust], 1)
topo2 = Topology.from_edges(7, 4, 0, np.concatenate([e, [[0, 0, 1, 0, 1]]]))
b2 = Batch(pose_t=batch.pose_t, anchors=batch.anchors, range_d=np.concatenate([batch.range_d, batch.range_d[:, :1]], 1),
           range_info=np.concatenate([batch.range_info, batch.range_info[:, :1]], 1))
cases.append(("fast_generic", (topo2, b2, None), 4))
for name, (topo, batch, _), it in cases:
    cfg = Config(max_iterations=it)
    got = s.solve(topo, batch, cfg)
    ref = oracle.solve(topo, batch, cfg)
    assert np.array_equal(got.pose_t, ref.pose_t) and np.array_equal(got.chi2, ref.chi2), name
    Hd, Ho, b, chi = s.linearize(topo, batch, cfg)
    print(name, "path", s.last_path, "ok")
x, ok = s.factor_solve(Hd, Ho, b, np.full(len(b), 0.3))
print("factor_solve ok", ok.all())
s.close()
