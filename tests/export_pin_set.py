"""Export a parity PIN SET: windows as g2o text files plus what this repo's oracle (and therefore its CUDA
kernels, which match it bit for bit) returns for them, so that someone with a real g2o build — the reference pins
g2o @ deafc01 with CHOLMOD, README.md:24-33 — can run the reference's own optimizer on the same inputs and close
the one open point of DESIGN.md section 2 ("parity unpinned").  TEST INFRASTRUCTURE: uses oracle/.

    python tests/export_pin_set.py --out /tmp/pin --windows 4

writes, per window, <shape>_<k>.g2o (localization_b200/tools/g2o_text.py: VERTEX_SE3:QUAT, FIX, EDGE_RANGE,
EDGE_RANGE_OFFSET, EDGE_SE3:QUAT, EDGE_SE3_PRIOR; '# ROBUST <edge>' comment lines mark the Cauchy edges) and
expected.json: final translations / rotations, {plain, robust, g2o-stale} chi2, final lambda, iterations, trials.
and <shape>_<k>.expected (the same numbers as plain "key value" lines, %.17g, for the C++ checker).
The g2o side is tools/pin_check/ (C++ + CMake): it loads each file with the reference's own edge types
compiled from the reference's sources, sets RobustKernelCauchy on the marked edges, builds BlockSolver_6_3 +
LinearSolverCholmod + OptimizationAlgorithmLevenberg as in localization.cpp:44-52, runs
initializeOptimization(); optimize(iterations); and compares: poses within 1e-6 m, chi2 within 1e-9 relative
(BASELINE.json north_star), equal iteration and trial counts."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from localization_b200 import Config, synthetic  # noqa: E402
from localization_b200.tools import g2o_text  # noqa: E402
from oracle import oracle  # noqa: E402

SHAPES = {
    "c1_uwb_only_n10": (lambda W: synthetic.uwb_only(W, 10, 4, seed=101), 10),
    "c3_uwb_only_n50": (lambda W: synthetic.uwb_only(W, 50, 8, seed=102), 10),
    "c2_uwb_imu_n12": (lambda W: synthetic.uwb_imu_lidar(W, 12, 4, antennas=0, lidar=False, seed=103), 10),
    "c4a_uwb_imu_lidar_n20": (lambda W: synthetic.uwb_imu_lidar(W, 20, 8, seed=104), 20),
    "c4b_uwb_twist_n15": (lambda W: synthetic.uwb_twist(W, 15, 8, seed=105), 12),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--windows", type=int, default=4)
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    expected = {}
    for name, (make, iters) in SHAPES.items():
        topo, batch, _ = make(a.windows)
        cfg = Config(max_iterations=iters)
        ref = oracle.solve(topo, batch, cfg, edge_chi2=True)
        for w in range(a.windows):
            f = f"{name}_{w}.g2o"
            g2o_text.write_window(os.path.join(a.out, f), topo, batch, w)
            expected[f] = {
                "iterations_max": iters,
                "pose_t": ref.pose_t[w].tolist(),
                "pose_R": None if ref.pose_R is None else ref.pose_R[w].reshape(-1, 9).tolist(),
                "chi2_plain": float(ref.chi2[w, 0]), "chi2_robust": float(ref.chi2[w, 1]),
                "chi2_g2o_stale": float(ref.chi2[w, 2]), "lambda": float(ref.chi2[w, 3]),
                "iterations": int(ref.status[w, 0]), "trials": int(ref.status[w, 1]), "flags": int(ref.status[w, 2]),
            }
            x = expected[f]
            N = topo.n_poses
            R = ref.pose_R[w].reshape(N, 9) if ref.pose_R is not None else [[1, 0, 0, 0, 1, 0, 0, 0, 1]] * N
            g = lambda v: "%.17g" % float(v)
            lines = [f"iterations_max {iters}", f"iterations {x['iterations']}", f"trials {x['trials']}",
                     f"flags {x['flags']}", f"chi2_plain {g(x['chi2_plain'])}", f"chi2_robust {g(x['chi2_robust'])}",
                     f"chi2_g2o_stale {g(x['chi2_g2o_stale'])}", f"lambda {g(x['lambda'])}"]
            for i in range(N):   # vertex ids as g2o_text.write_window assigns them: slot * 300 + self id (robot.cpp:43,94)
                lines.append(f"pose {i * 300 + 200} " + " ".join(g(v) for v in list(ref.pose_t[w, i]) + list(R[i])))
            if getattr(ref, "edge_chi2", None) is not None:
                lines.append("edge_chi2 " + " ".join(g(v) for v in ref.edge_chi2[w]))
                x["edge_chi2"] = ref.edge_chi2[w].tolist()
            with open(os.path.join(a.out, f[:-4] + ".expected"), "w") as fh:
                fh.write("\n".join(lines) + "\n")
    with open(os.path.join(a.out, "expected.json"), "w") as fh:
        json.dump(expected, fh, indent=1)
    print(f"{len(expected)} windows + expected.json in {a.out}")


if __name__ == "__main__":
    main()
