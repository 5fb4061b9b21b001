/*
 * host_localization.h — ROS-free drop-in for the reference's Localization class
 * (reference src/localization/localization.h:99-128, localization.cpp): same public methods with
 * the same meaning, plain-struct messages instead of ROS ConstPtrs, a Params struct instead of
 * the ROS parameter server (same keys, localization.cpp:58-159), and `solve()` re-pointed from
 * g2o (`initializeOptimization(); optimize(iteration_max)`, localization.cpp:168-170) to the C ABI
 * of include/uwbgo.h.  Nothing in here does solver arithmetic.
 *
 * Fleet runs many Localization instances in lockstep (many robots, Monte-Carlo replays,
 * parameter sweeps): their solve() calls are collected and issued as ONE uwbgo_solve_batch per
 * window structure.
 */
#ifndef UWBGO_HOST_LOCALIZATION_H
#define UWBGO_HOST_LOCALIZATION_H

#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/uwbgo.h"
#include "window_graph.h"

namespace uwbgo {
namespace host {

/* the solver behind solve(): uwbgo_solve_batch bound to a context, or (tests only) anything with
 * the same contract */
using SolveBackend = std::function<int(const uwbgo_topology *, const uwbgo_batch *, const uwbgo_config *,
                                       uwbgo_result *)>;
SolveBackend backend_from_ctx(uwbgo_ctx *ctx);

struct Params {
    /* optimizer/... (localization.cpp:58-69) */
    int maximum_iteration = 20;
    double minimum_optimize_error = 1000.0;
    bool verbose = false;
    /* robot/... (localization.cpp:72-79) */
    int trajectory_length = 10;
    double maximum_velocity = 1.0;
    double distance_outlier = 1.0;
    /* /uwb/nodesId, /uwb/nodesPos, /uwb/antennaOffset (localization.cpp:83-123): the LAST id is self */
    std::vector<int> nodesId;
    std::vector<double> nodesPos;
    std::vector<double> antennaOffset;
    /* log/filename_prefix (localization.cpp:126); empty = no log files */
    std::string filename_prefix;
    std::string filename_suffix; /* empty = the reference's _%Y_%b_%d_%H_%M_%S.txt */
    /* frame/... and publish_flag/... (localization.cpp:134-159) */
    std::string frame_target = "estimation", frame_source = "local_origin";
    bool publish_tf = false, publish_range = false, publish_pose = false, publish_twist = false,
         publish_lidar = false, publish_imu = false;
};

/* one window as the C ABI wants it, plus the way back to the vertices */
struct PackedWindow {
    std::vector<int32_t> kind, a, b, ant, ant_b, robust;
    int n_poses = 0, n_anchors = 0, n_antennas = 0, iteration_max = 0;
    std::vector<double> antenna_xyz; /* [n_antennas][3] */
    bool identity_rotations = true;
    std::vector<double> pose_t, pose_R, anchors, range_d, range_info, prior_Z, prior_info, se3_Z, se3_info;
    std::vector<int32_t> oplus;
    std::vector<VertexSE3 *> pose_vertices;
    std::vector<int32_t> key() const; /* structure key: windows with equal keys share a batch */
};

struct Published {
    PoseStamped realtime;  /* robots.at(self_id).current_pose()          localization.cpp:208 */
    PoseStamped optimized; /* path->poses[trajectory_length/2]          localization.cpp:220 */
    double error = 0.0;    /* optimizer.chi2()                           localization.cpp:197 */
};

class Fleet;

class Localization {
public:
    Localization(const Params &, SolveBackend backend);
    ~Localization();

    bool solve();   /* false: no solve ran (pack or solver error, counted in solver_errors()) */
    void publish();
    void solve_and_publish(); /* "solve(); publish();" of the callbacks; no publish after a failed solve */
    /* deferred mode: solve a window parked for Fleet::flush() now, on its own (see settle() in the .cpp);
     * every callback calls it before it touches the graph */
    void settle();
    int settled_alone() const { return n_settled_; }
    /* range edges built through the reference's edge classes (types_edge_se3range*.cpp) between vertices
     * of the window: from_age / to_age count from the oldest ring slot, to_anchor >= 0 names a fixed node
     * id instead.  off_from / off_to are antenna numbers: EdgeSE3Range gets setVertexOffset(0 / 1, ...),
     * EdgeSE3RangeOffset gets setParameterId(0 / 1, ...) on a table whose id k holds offsets[k-1] */
    bool addTypedRangeEdge(bool parameter_offsets, int from_age, int to_age, int to_anchor, double measurement,
                           double information, int off_from, int off_to, bool cauchy);
    void addRangeEdge(const UwbRange &);
    void addPoseEdge(const PoseWithCovarianceStamped &);
    void addLidarEdge(const PoseWithCovarianceStamped &);
    void addImuEdge(const Imu &);
    void addTwistEdge(const TwistWithCovarianceStamped &);
    void configCallback(bool publish_optimized_poses);
    void set_file();
    void set_file(std::vector<double> antennaOffset);

    /* what the ROS publishers / log files would have carried */
    const std::vector<Published> &published() const { return published_; }
    const std::vector<PoseStamped> &republished() const { return republished_; }
    int solves() const { return n_solves_; }
    int rejected_ranges() const { return n_rejected_; }
    int skipped_publishes() const { return n_skipped_; }
    int solver_errors() const { return n_errors_; }
    const std::string &last_error() const { return last_error_; }
    double last_chi2(int k) const { return last_chi2_[k]; }
    const int32_t *last_status() const { return last_status_; }
    const std::string &realtime_log() const { return realtime_filename; }
    const std::string &optimized_log() const { return optimized_filename; }
    Robot &self() { return robots.at(self_id); }

    /* the window solve() would hand to the solver right now */
    bool pack(PackedWindow &out, std::string &err);
    void unpack(const PackedWindow &w, const double *pose_t, const double *pose_R, const int32_t *oplus,
                const double *chi2, const int32_t *status);

private:
    friend class Fleet;
    bool solve_window(const PackedWindow &w);
    Edge make_range_edge(VertexSE3 *v1, VertexSE3 *v2, double distance, double covariance);
    Edge make_se3_edge_from_twist(VertexSE3 *v1, VertexSE3 *v2, const TwistWithCovarianceStamped &, double dt);
    void save_file(const PoseStamped &pose, const std::string &filename);
    void write_log_headers(const std::vector<double> *antennaOffset);

    Params prm;
    SolveBackend backend_;
    Graph optimizer;
    std::map<int, Robot> robots;
    std::vector<Isometry3d> offsets;
    int self_id = 0;
    int number_measurements = 0;
    int iteration_max, trajectory_length;
    double minimum_optimize_error, robot_max_velocity, distance_outlier;
    VertexSE3 *key_vertex = nullptr;
    bool flag_save_file = false;
    std::string realtime_filename, optimized_filename;

    Fleet *fleet_ = nullptr;         /* deferred mode */
    bool solve_pending_ = false, publish_pending_ = false;
    PackedWindow pending_;

    std::vector<Published> published_;
    std::vector<PoseStamped> republished_;
    int n_solves_ = 0, n_rejected_ = 0, n_skipped_ = 0, n_errors_ = 0, n_settled_ = 0;
    double last_chi2_[UWBGO_CHI2_STRIDE] = {0, 0, 0, 0};
    int32_t last_status_[UWBGO_STATUS_STRIDE] = {0, 0, 0, 0};
    std::string last_error_;
};

class Fleet {
public:
    explicit Fleet(SolveBackend backend) : backend_(std::move(backend)) {}
    Localization &add(const Params &);
    Localization &at(size_t i) { return *members_.at(i); }
    size_t size() const { return members_.size(); }
    /* run every pending solve() as one batch per window structure, then the pending publish()es */
    int flush();
    int64_t windows_solved() const { return windows_; }
    int batches() const { return batches_; }

private:
    SolveBackend backend_;
    std::vector<std::unique_ptr<Localization>> members_;
    int64_t windows_ = 0;
    int batches_ = 0;
};

const struct SensorType {
    unsigned char general = 0, pose = 1, range = 2, twist = 3, imu = 4;
} sensor_type;

}  // namespace host
}  // namespace uwbgo
#endif
