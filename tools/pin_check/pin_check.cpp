// pin_check — runs the REFERENCE's own optimizer stack on the windows of a pin set and compares with the
// answers this repo's oracle (and, bit for bit, its CUDA kernels) gave for them.
//
// Build it on a machine that has what the reference builds against: g2o at the commit the reference pins
// (README.md:26-33), Eigen3, SuiteSparse/CHOLMOD, and the reference's source tree (its two edge types are
// compiled from where they lie: src/types/types_edge_se3range*.cpp, registered with g2o's factory as
// EDGE_RANGE / EDGE_RANGE_OFFSET by their own G2O_REGISTER_TYPE lines).  See README.md in this directory.
//
//   python tests/export_pin_set.py --out pin --windows 4
//   pin_check pin            # every pin/*.g2o against pin/*.expected
//
// What it does per window is what Localization does (reference src/localization/localization.cpp):
//   :44-52   Solver = LinearSolverCholmod<BlockSolver_6_3::PoseMatrixType>, setBlockOrdering(false),
//            BlockSolver_6_3, OptimizationAlgorithmLevenberg
//   :608-627 RobustKernelCauchy (default delta) on the edges the pin set marks "# ROBUST <edge index>"
//   :168-170 initializeOptimization(); optimize(iteration_max);
//   :197     optimizer.chi2()  (the errors of the LAST trial, accepted or not)
// and then the comparisons BASELINE.json asks for: poses within 1e-6 m, chi2 within 1e-9 relative; equal
// iteration and trial counts show that the accept / reject history is the same.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <dirent.h>
#include <fstream>
#include <iostream>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <vector>

#include <g2o/core/block_solver.h>
#include <g2o/core/optimization_algorithm_levenberg.h>
#include <g2o/core/robust_kernel_impl.h>
#include <g2o/core/sparse_optimizer.h>
#include <g2o/solvers/cholmod/linear_solver_cholmod.h>
#include <g2o/types/slam3d/types_slam3d.h>

#include "types_edge_se3range.h"         // reference src/types/
#include "types_edge_se3range_offset.h"  // reference src/types/

typedef g2o::BlockSolver_6_3 SE3BlockSolver;                                     // localization.h:82
typedef g2o::LinearSolverCholmod<SE3BlockSolver::PoseMatrixType> Solver;         // localization.h:84

struct Expected {
    int iterations_max = 0, iterations = 0, trials = 0, flags = 0;
    double chi2_plain = 0, chi2_robust = 0, chi2_stale = 0, lambda = 0;
    std::map<int, std::vector<double>> pose;  // vertex id -> x y z r00 .. r22
    std::vector<double> edge_chi2;            // per edge, insertion order (optional)
};

static bool read_expected(const std::string &path, Expected &x)
{
    std::ifstream f(path);
    if (!f) return false;
    std::string line, key;
    while (std::getline(f, line)) {
        std::istringstream is(line);
        if (!(is >> key)) continue;
        if (key == "iterations_max") is >> x.iterations_max;
        else if (key == "iterations") is >> x.iterations;
        else if (key == "trials") is >> x.trials;
        else if (key == "flags") is >> x.flags;
        else if (key == "chi2_plain") is >> x.chi2_plain;
        else if (key == "chi2_robust") is >> x.chi2_robust;
        else if (key == "chi2_g2o_stale") is >> x.chi2_stale;
        else if (key == "lambda") is >> x.lambda;
        else if (key == "pose") {
            int id;
            is >> id;
            std::vector<double> v(12);
            for (double &d : v) is >> d;
            x.pose[id] = v;
        } else if (key == "edge_chi2") {
            double d;
            while (is >> d) x.edge_chi2.push_back(d);
        }
    }
    return x.iterations_max > 0 && !x.pose.empty();
}

// '# ROBUST <edge index>' lines (g2o's loader skips comments); index = position among the EDGE_* lines
static std::set<int> robust_edges(const std::string &path)
{
    std::set<int> r;
    std::ifstream f(path);
    std::string line, a, b;
    while (std::getline(f, line)) {
        std::istringstream is(line);
        int k;
        if ((is >> a >> b >> k) && a == "#" && b == "ROBUST") r.insert(k);
    }
    return r;
}

static double rel(double a, double b) { return std::fabs(a - b) / std::max(std::fabs(b), 1e-300); }

static int check_window(const std::string &g2o_file, const std::string &expected_file, bool verbose)
{
    Expected x;
    if (!read_expected(expected_file, x)) {
        std::printf("%s: cannot read %s\n", g2o_file.c_str(), expected_file.c_str());
        return 1;
    }
    g2o::SparseOptimizer optimizer;
    Solver *solver = new Solver();
    solver->setBlockOrdering(false);
    SE3BlockSolver *se3blockSolver = new SE3BlockSolver(solver);
    g2o::OptimizationAlgorithmLevenberg *lm = new g2o::OptimizationAlgorithmLevenberg(se3blockSolver);
    optimizer.setAlgorithm(lm);
    optimizer.setVerbose(verbose);
    if (!optimizer.load(g2o_file.c_str())) {  // PARAMS_SE3OFFSET 0 in the file = the zero_offset of localization.cpp:54-56
        std::printf("%s: g2o could not load it (are EDGE_RANGE / EDGE_RANGE_OFFSET registered?)\n", g2o_file.c_str());
        return 1;
    }
    // edges in insertion order = internal id order (what activeEdges() is sorted by)
    std::vector<g2o::OptimizableGraph::Edge *> edges;
    for (g2o::HyperGraph::Edge *e : optimizer.edges()) edges.push_back(static_cast<g2o::OptimizableGraph::Edge *>(e));
    std::sort(edges.begin(), edges.end(),
              [](g2o::OptimizableGraph::Edge *a, g2o::OptimizableGraph::Edge *b) { return a->internalId() < b->internalId(); });
    for (int k : robust_edges(g2o_file))
        if (k >= 0 && (size_t)k < edges.size()) edges[(size_t)k]->setRobustKernel(new g2o::RobustKernelCauchy());

    optimizer.initializeOptimization();
    const int iterations = optimizer.optimize(x.iterations_max);
    const double stale = optimizer.chi2();  // what publish() gates on
    std::vector<double> edge_stale;
    for (g2o::OptimizableGraph::Edge *e : edges) edge_stale.push_back(e->chi2());
    const double lambda = lm->currentLambda();
    const int trials = lm->levenbergIteration();
    optimizer.computeActiveErrors();
    const double plain = optimizer.activeChi2(), robust = optimizer.activeRobustChi2();

    double dt_max = 0, dR_max = 0;
    for (const auto &kv : x.pose) {
        const g2o::VertexSE3 *v = dynamic_cast<const g2o::VertexSE3 *>(optimizer.vertex(kv.first));
        if (!v) {
            std::printf("%s: vertex %d missing\n", g2o_file.c_str(), kv.first);
            return 1;
        }
        const Eigen::Isometry3d &T = v->estimate();
        for (int r = 0; r < 3; ++r) {
            dt_max = std::max(dt_max, std::fabs(T(r, 3) - kv.second[(size_t)r]));
            for (int c = 0; c < 3; ++c) dR_max = std::max(dR_max, std::fabs(T(r, c) - kv.second[(size_t)(3 + 3 * r + c)]));
        }
    }
    double de_max = 0;
    if (x.edge_chi2.size() == edge_stale.size())
        for (size_t k = 0; k < edge_stale.size(); ++k)
            de_max = std::max(de_max, std::fabs(edge_stale[k] - x.edge_chi2[k]) / std::max(1.0, std::fabs(x.edge_chi2[k])));
    const bool ok = dt_max < 1e-6 && dR_max < 1e-6 && rel(plain, x.chi2_plain) < 1e-9 && rel(robust, x.chi2_robust) < 1e-9 &&
                    rel(stale, x.chi2_stale) < 1e-9 && iterations == x.iterations && trials == x.trials && de_max < 1e-9;
    std::printf("%-34s %s  |dt| %.3e m  |dR| %.3e  chi2 rel: plain %.3e robust %.3e stale %.3e  edge chi2 %.3e  "
                "lambda %.6g (expected %.6g)  iterations %d/%d  trials %d/%d\n",
                g2o_file.substr(g2o_file.find_last_of('/') + 1).c_str(), ok ? "PASS" : "FAIL", dt_max, dR_max,
                rel(plain, x.chi2_plain), rel(robust, x.chi2_robust), rel(stale, x.chi2_stale), de_max, lambda, x.lambda,
                iterations, x.iterations, trials, x.trials);
    return ok ? 0 : 1;
}

int main(int argc, char **argv)
{
    if (argc < 2) {
        std::fprintf(stderr, "usage: pin_check <pin-set directory> [-v]\n");
        return 2;
    }
    const std::string dir = argv[1];
    const bool verbose = argc > 2 && std::string(argv[2]) == "-v";
    std::vector<std::string> files;
    if (DIR *d = opendir(dir.c_str())) {
        while (dirent *e = readdir(d)) {
            std::string n = e->d_name;
            if (n.size() > 4 && n.substr(n.size() - 4) == ".g2o") files.push_back(n);
        }
        closedir(d);
    }
    std::sort(files.begin(), files.end());
    if (files.empty()) {
        std::fprintf(stderr, "no .g2o files in %s\n", dir.c_str());
        return 2;
    }
    int failed = 0;
    for (const std::string &f : files)
        failed += check_window(dir + "/" + f, dir + "/" + f.substr(0, f.size() - 4) + ".expected", verbose);
    std::printf("%zu windows, %d failed\n", files.size(), failed);
    return failed ? 1 : 0;
}
