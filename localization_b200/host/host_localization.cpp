#include "host_localization.h"
#include "edge_types.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <set>

namespace uwbgo {
namespace host {

SolveBackend backend_from_ctx(uwbgo_ctx *ctx)
{
    return [ctx](const uwbgo_topology *t, const uwbgo_batch *b, const uwbgo_config *c, uwbgo_result *r) {
        return uwbgo_solve_batch(ctx, t, b, c, r);
    };
}

/* ---- small linear algebra on plain arrays (edge parameterisation only) --------------------- */
static bool invert6(const double *M, double *inv) /* Gauss-Jordan with partial pivoting */
{
    double a[6][12];
    for (int r = 0; r < 6; ++r)
        for (int c = 0; c < 6; ++c) {
            a[r][c] = M[6 * r + c];
            a[r][6 + c] = r == c ? 1.0 : 0.0;
        }
    for (int col = 0; col < 6; ++col) {
        int piv = col;
        for (int r = col + 1; r < 6; ++r)
            if (std::fabs(a[r][col]) > std::fabs(a[piv][col])) piv = r;
        if (a[piv][col] == 0.0) return false;
        if (piv != col)
            for (int c = 0; c < 12; ++c) std::swap(a[piv][c], a[col][c]);
        const double d = a[col][col];
        for (int c = 0; c < 12; ++c) a[col][c] /= d;
        for (int r = 0; r < 6; ++r) {
            if (r == col) continue;
            const double f = a[r][col];
            if (f == 0.0) continue;
            for (int c = 0; c < 12; ++c) a[r][c] -= f * a[col][c];
        }
    }
    for (int r = 0; r < 6; ++r)
        for (int c = 0; c < 6; ++c) inv[6 * r + c] = a[r][6 + c];
    return true;
}

/* Eigen Quaterniond(w,x,y,z).toRotationMatrix(), no normalisation (localization.cpp:509) */
static void quaternion_to_matrix(double w, double x, double y, double z, double *R)
{
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
    R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

/* tf::Quaternion::setRPY followed by tf::Matrix3x3::setRotation (localization.cpp:570-574) */
static void rpy_to_matrix(double roll, double pitch, double yaw, double *R)
{
    const double hy = yaw * 0.5, hp = pitch * 0.5, hr = roll * 0.5;
    const double cy = std::cos(hy), sy = std::sin(hy), cp = std::cos(hp), sp = std::sin(hp),
                 cr = std::cos(hr), sr = std::sin(hr);
    const double x = sr * cp * cy - cr * sp * sy, y = cr * sp * cy + sr * cp * sy,
                 z = cr * cp * sy - sr * sp * cy, w = cr * cp * cy + sr * sp * sy;
    const double d = x * x + y * y + z * z + w * w, s = 2.0 / d;
    const double xs = x * s, ys = y * s, zs = z * s;
    const double wx = w * xs, wy = w * ys, wz = w * zs, xx = x * xs, xy = x * ys, xz = x * zs,
                 yy = y * ys, yz = y * zs, zz = z * zs;
    R[0] = 1 - (yy + zz); R[1] = xy - wz;       R[2] = xz + wy;
    R[3] = xy + wz;       R[4] = 1 - (xx + zz); R[5] = yz - wx;
    R[6] = xz - wy;       R[7] = yz + wx;       R[8] = 1 - (xx + yy);
}

/* ---- construction -------------------------------------------------------------------------- */
Localization::Localization(const Params &p, SolveBackend backend) : prm(p), backend_(std::move(backend))
{
    iteration_max = p.maximum_iteration;
    minimum_optimize_error = p.minimum_optimize_error;
    trajectory_length = p.trajectory_length;
    robot_max_velocity = p.maximum_velocity;
    distance_outlier = p.distance_outlier;
    if (p.nodesId.empty() || p.nodesPos.size() != 3 * p.nodesId.size()) {
        last_error_ = "Can't get parameter nodesId / nodesPos from UWB";
        ++n_errors_;
        return;
    }
    self_id = p.nodesId.back(); /* localization.cpp:89: the last id is the moving robot */
    for (size_t i = 0; i < p.nodesId.size(); ++i) {
        const int id = p.nodesId[i];
        const bool moving = id == self_id;
        robots.emplace(id, Robot(id, !moving, moving ? trajectory_length : 1));
        Isometry3d pose = Isometry3d::Identity();
        pose.t[0] = p.nodesPos[3 * i];
        pose.t[1] = p.nodesPos[3 * i + 1];
        pose.t[2] = p.nodesPos[3 * i + 2];
        robots.at(id).init(optimizer, pose);
    }
    for (size_t i = 0; i + 2 < p.antennaOffset.size(); i += 3) {
        Isometry3d o = Isometry3d::Identity();
        o.t[0] = p.antennaOffset[i];
        o.t[1] = p.antennaOffset[i + 1];
        o.t[2] = p.antennaOffset[i + 2];
        offsets.push_back(o);
    }
    if (!p.filename_prefix.empty()) {
        if (!p.antennaOffset.empty()) set_file(p.antennaOffset);
        else set_file();
    }
}

Localization::~Localization()
{
    if (flag_save_file && robots.count(self_id)) { /* localization.cpp:705-716 */
        Path *path = robots.at(self_id).vertices2path();
        for (int i = trajectory_length / 2; i < trajectory_length; ++i) save_file(path->poses[i], optimized_filename);
    }
}

/* ---- the hot path boundary ------------------------------------------------------------------ */
std::vector<int32_t> PackedWindow::key() const
{
    std::vector<int32_t> k{n_poses, n_anchors, n_antennas, identity_rotations ? 1 : 0, (int32_t)kind.size()};
    for (size_t e = 0; e < kind.size(); ++e) {
        k.push_back(kind[e]); k.push_back(a[e]); k.push_back(b[e]); k.push_back(ant[e]); k.push_back(ant_b[e]);
        k.push_back(robust[e]);
    }
    /* one batch has ONE antenna table and ONE iteration limit (uwbgo_batch::ant_offsets, uwbgo_config):
     * members that differ in either go to different batches */
    k.push_back(iteration_max);
    for (double x : antenna_xyz) {
        int32_t bits[2];
        std::memcpy(bits, &x, sizeof bits);
        k.push_back(bits[0]); k.push_back(bits[1]);
    }
    return k;
}

bool Localization::pack(PackedWindow &w, std::string &err)
{
    w = PackedWindow();
    Robot &me = robots.at(self_id);
    /* g2o's active set: vertices with at least one edge.  Ring slots that were never overwritten
     * carry no edges; the active poses are the newest ones, contiguous in age. */
    std::set<const VertexSE3 *> used;
    for (const Edge &e : optimizer.edges()) {
        used.insert(e.from);
        if (e.to) used.insert(e.to);
    }
    std::map<const VertexSE3 *, int> pose_index, anchor_index;
    int first = -1;
    for (int age = 0; age < me.length(); ++age) {
        VertexSE3 *v = me.by_age(age);
        if (first < 0 && used.count(v)) first = age;
        if (first >= 0) {
            pose_index[v] = age - first;
            w.pose_vertices.push_back(v);
        }
    }
    if (first < 0) {
        err = "no active pose";
        return false;
    }
    w.n_poses = (int)w.pose_vertices.size();
    w.n_antennas = (int)offsets.size();
    w.iteration_max = iteration_max;
    for (const Isometry3d &o : offsets) w.antenna_xyz.insert(w.antenna_xyz.end(), o.t, o.t + 3);
    std::vector<bool> zero_offset(offsets.size());
    for (size_t k = 0; k < offsets.size(); ++k)
        zero_offset[k] = offsets[k].t[0] == 0 && offsets[k].t[1] == 0 && offsets[k].t[2] == 0;
    for (VertexSE3 *v : w.pose_vertices) {
        const Isometry3d &T = v->estimate();
        w.pose_t.insert(w.pose_t.end(), T.t, T.t + 3);
        w.pose_R.insert(w.pose_R.end(), T.R, T.R + 9);
        w.oplus.push_back(v->oplusCalls);
        if (!T.rotationIsIdentity()) w.identity_rotations = false;
    }
    auto anchor_of = [&](const VertexSE3 *v) {
        auto it = anchor_index.find(v);
        if (it != anchor_index.end()) return it->second;
        int idx = (int)anchor_index.size();
        anchor_index[v] = idx;
        const Isometry3d &T = v->estimate();
        w.anchors.insert(w.anchors.end(), T.t, T.t + 3);
        return idx;
    };
    auto push_Z = [](std::vector<double> &dst, const Isometry3d &Z) {
        dst.insert(dst.end(), Z.R, Z.R + 9);
        dst.insert(dst.end(), Z.t, Z.t + 3);
    };
    for (const Edge &e : optimizer.edges()) { /* insertion order */
        auto pa = pose_index.find(e.from);
        if (pa == pose_index.end()) {
            err = "edge whose vertex 0 is not a pose of the window";
            return false;
        }
        int kind, b = 0, ant = 0, ant_b = 0;
        if (e.kind == EdgeKind::Range) {
            if (e.antenna < 0 || (size_t)e.antenna > offsets.size() || e.antenna_b < 0 || (size_t)e.antenna_b > offsets.size()) {
                err = "range edge with an antenna number outside the antenna table";
                return false;
            }
            ant = e.antenna;
            ant_b = e.antenna_b;
            if (ant > 0 && zero_offset[(size_t)ant - 1]) ant = 0; /* a zero lever arm is the identity offset, bit for bit */
            if (ant_b > 0 && zero_offset[(size_t)ant_b - 1]) ant_b = 0;
            if (e.to->fixed()) {
                kind = UWBGO_EDGE_RANGE_ANCHOR;
                b = anchor_of(e.to);
            } else {
                kind = UWBGO_EDGE_RANGE_POSE;
                auto pb = pose_index.find(e.to);
                if (pb == pose_index.end() || pb->second <= pa->second) {
                    err = "range edge whose vertex 1 is not a newer pose of the window";
                    return false;
                }
                b = pb->second;
            }
            w.range_d.push_back(e.range);
            w.range_info.push_back(e.rangeInformation);
        } else if (e.kind == EdgeKind::Prior) {
            kind = UWBGO_EDGE_PRIOR;
            push_Z(w.prior_Z, e.measurement);
            w.prior_info.insert(w.prior_info.end(), e.information, e.information + 36);
        } else {
            kind = UWBGO_EDGE_SE3;
            auto pb = pose_index.find(e.to);
            if (pb == pose_index.end() || pb->second <= pa->second) {
                err = "EdgeSE3 whose vertex 1 is not a newer pose of the window";
                return false;
            }
            b = pb->second;
            push_Z(w.se3_Z, e.measurement);
            w.se3_info.insert(w.se3_info.end(), e.information, e.information + 36);
        }
        w.kind.push_back(kind);
        w.a.push_back(pa->second);
        w.b.push_back(b);
        w.ant.push_back(ant);
        w.ant_b.push_back(ant_b);
        w.robust.push_back(e.cauchy ? 1 : 0);
    }
    w.n_anchors = (int)anchor_index.size();
    return true;
}

void Localization::unpack(const PackedWindow &w, const double *pose_t, const double *pose_R, const int32_t *oplus,
                          const double *chi2, const int32_t *status)
{
    for (int i = 0; i < w.n_poses; ++i) {
        Isometry3d T = w.pose_vertices[i]->estimate();
        std::memcpy(T.t, pose_t + 3 * i, sizeof T.t);
        if (pose_R) std::memcpy(T.R, pose_R + 9 * i, sizeof T.R);
        w.pose_vertices[i]->setEstimate(T);
        if (oplus) w.pose_vertices[i]->oplusCalls = oplus[i];
    }
    std::memcpy(last_chi2_, chi2, sizeof last_chi2_);
    std::memcpy(last_status_, status, sizeof last_status_);
    ++n_solves_;
}

static void fill_abi(const PackedWindow &w, int64_t n_windows, uwbgo_topology &t, uwbgo_batch &b)
{
    const std::vector<double> &antenna_xyz = w.antenna_xyz;
    t = uwbgo_topology{w.n_poses, w.n_anchors, w.n_antennas, (int32_t)w.kind.size(), w.kind.data(), w.a.data(),
                       w.b.data(), w.ant.data(), w.robust.data(), w.ant_b.data()};
    b = uwbgo_batch{};
    b.n_windows = n_windows;
    b.pose_t = w.pose_t.data();
    b.pose_R = w.identity_rotations ? nullptr : w.pose_R.data();
    b.oplus_count = w.oplus.data();
    b.anchors = w.anchors.data();
    b.ant_offsets = antenna_xyz.empty() ? nullptr : antenna_xyz.data();
    b.range_d = w.range_d.data();
    b.range_info = w.range_info.data();
    b.prior_Z = w.prior_Z.data();
    b.prior_info = w.prior_info.data();
    b.se3_Z = w.se3_Z.data();
    b.se3_info = w.se3_info.data();
}

/* one window through the solver right now (the reference's synchronous optimize()) */
bool Localization::solve_window(const PackedWindow &w)
{
    uwbgo_topology topo;
    uwbgo_batch in;
    fill_abi(w, 1, topo, in);
    uwbgo_config cfg;
    uwbgo_config_default(&cfg);
    cfg.max_iterations = w.iteration_max;
    std::vector<double> pose_t(w.pose_t.size()), pose_R(w.pose_R.size());
    std::vector<int32_t> oplus(w.oplus.size());
    double chi2[UWBGO_CHI2_STRIDE];
    int32_t status[UWBGO_STATUS_STRIDE];
    uwbgo_result out{pose_t.data(), pose_R.data(), oplus.data(), chi2, status};
    int rc = backend_(&topo, &in, &cfg, &out);
    if (rc != UWBGO_OK) { /* the reference ignores optimize()'s return value; we at least count it */
        last_error_ = std::string("uwbgo_solve_batch failed: ") + uwbgo_last_error();
        ++n_errors_;
        return false;
    }
    unpack(w, pose_t.data(), pose_R.data(), oplus.data(), chi2, status);
    return true;
}

/* Deferred (fleet) mode parks a packed window -- vertex pointers and a snapshot of the estimates --
 * until Fleet::flush().  The reference solves synchronously, so the graph can never change under a
 * solve; here every callback that is about to touch the graph settles a parked window first (solved on
 * its own, outside the fleet's batch) so that neither a dangling vertex nor a stale snapshot can occur. */
void Localization::settle()
{
    if (!solve_pending_) return;
    solve_pending_ = false;
    ++n_settled_;
    const bool ok = solve_window(pending_);
    pending_ = PackedWindow();
    if (publish_pending_) {
        publish_pending_ = false;
        if (ok) publish(); else ++n_skipped_;
    }
}

/* replaces localization.cpp:164-192; false = no solve ran (counted in solver_errors()) */
bool Localization::solve()
{
    settle();
    std::string err;
    PackedWindow w;
    if (!pack(w, err)) {
        last_error_ = err;
        ++n_errors_;
        return false;
    }
    if (fleet_) { /* lockstep mode: Fleet::flush() issues the batch */
        pending_ = std::move(w);
        solve_pending_ = true;
        return true;
    }
    return solve_window(w);
}

/* what every callback does where the reference has "solve(); publish();": a publish() after a solve
 * that did not run would gate on the previous solve's chi2 and log an un-optimised estimate */
void Localization::solve_and_publish()
{
    if (solve()) publish(); else ++n_skipped_;
}

/* replaces localization.cpp:195-251 (ROS publishers / TF become records, log files stay) */
void Localization::publish()
{
    if (solve_pending_) {
        publish_pending_ = true;
        return;
    }
    const double error = last_chi2_[2]; /* optimizer.chi2(): errors of the last trial */
    if (!(error < minimum_optimize_error)) {
        ++n_skipped_;
        return;
    }
    Published out;
    out.error = error;
    out.realtime = robots.at(self_id).current_pose();
    out.realtime.header.frame_id = prm.frame_source;
    Path *path = robots.at(self_id).vertices2path();
    path->header.frame_id = prm.frame_source;
    out.optimized = path->poses[(size_t)trajectory_length / 2];
    published_.push_back(out);
    if (flag_save_file) {
        save_file(out.realtime, realtime_filename);
        save_file(out.optimized, optimized_filename);
    }
}

/* ---- graph builder callbacks ---------------------------------------------------------------- */
/* localization.cpp:608-627 */
Edge Localization::make_range_edge(VertexSE3 *v1, VertexSE3 *v2, double distance, double covariance)
{
    EdgeSE3Range edge;
    edge.vertices()[0] = v1;
    edge.vertices()[1] = v2;
    edge.setMeasurement(distance);
    edge.setInformation(1.0 / covariance); /* 1x1 covariance_matrix.inverse() */
    edge.setRobustKernel(new RobustKernelCauchy());
    return edge.asEdge(offsets);
}

/* the two range edge classes of src/types/ driven the way a g2o user drives them; see the header */
bool Localization::addTypedRangeEdge(bool parameter_offsets, int from_age, int to_age, int to_anchor,
                                     double measurement, double information, int off_from, int off_to, bool cauchy)
{
    settle();
    Robot &me = robots.at(self_id);
    auto fail = [&](const char *why) {
        last_error_ = why;
        ++n_errors_;
        return false;
    };
    if (from_age < 0 || from_age >= me.length()) return fail("typed range edge: vertex 0 is not a pose of the ring");
    VertexSE3 *v0 = me.by_age(from_age), *v1 = nullptr;
    if (to_anchor >= 0) {
        if (!robots.count(to_anchor) || !robots.at(to_anchor).is_static()) return fail("typed range edge: unknown anchor id");
        v1 = robots.at(to_anchor).last_vertex();
    } else {
        if (to_age <= from_age || to_age >= me.length()) return fail("typed range edge: vertex 1 is not a newer pose of the ring");
        v1 = me.by_age(to_age);
    }
    if (off_from < 0 || (size_t)off_from > offsets.size() || off_to < 0 || (size_t)off_to > offsets.size())
        return fail("typed range edge: antenna number outside the antenna table");
    try {
        if (parameter_offsets) { /* EDGE_RANGE_OFFSET: offsets by parameter id, id 0 = identity (types_edge_se3range_offset.cpp:61-79) */
            std::map<int, ParameterSE3Offset> params;
            params[0].setId(0);
            for (size_t k = 0; k < offsets.size(); ++k) {
                params[(int)k + 1].setId((int)k + 1);
                params[(int)k + 1].setOffset(offsets[k]);
            }
            EdgeSE3RangeOffset edge(&params);
            edge.vertices()[0] = v0;
            edge.vertices()[1] = v1;
            edge.setMeasurement(measurement);
            edge.setInformation(information);
            if (cauchy) edge.setRobustKernel(new RobustKernelCauchy());
            if (!edge.setParameterId(0, off_from) || !edge.setParameterId(1, off_to)) return fail("typed range edge: unknown parameter id");
            optimizer.addEdge(edge.asEdge(offsets));
        } else { /* EDGE_RANGE: offsets by value (types_edge_se3range.cpp:99-114) */
            EdgeSE3Range edge;
            edge.vertices()[0] = v0;
            edge.vertices()[1] = v1;
            edge.setMeasurement(measurement);
            edge.setInformation(information);
            if (cauchy) edge.setRobustKernel(new RobustKernelCauchy());
            if (off_from > 0) edge.setVertexOffset(0, offsets[(size_t)off_from - 1]);
            if (off_to > 0) edge.setVertexOffset(1, offsets[(size_t)off_to - 1]);
            optimizer.addEdge(edge.asEdge(offsets));
        }
    } catch (const std::invalid_argument &e) {
        return fail(e.what());
    }
    return true;
}

/* localization.cpp:560-605 */
Edge Localization::make_se3_edge_from_twist(VertexSE3 *v1, VertexSE3 *v2, const TwistWithCovarianceStamped &tw, double dt)
{
    Edge e;
    e.kind = EdgeKind::SE3;
    e.from = v1;
    e.to = v2;
    rpy_to_matrix(tw.twist.angular.x * dt, tw.twist.angular.y * dt, tw.twist.angular.z * dt, e.measurement.R);
    e.measurement.t[0] = tw.twist.linear.x * dt;
    e.measurement.t[1] = tw.twist.linear.y * dt;
    e.measurement.t[2] = tw.twist.linear.z * dt;
    double cov[36];
    for (int k = 0; k < 36; ++k) cov[k] = tw.covariance[(size_t)k] * dt * dt;
    if (!invert6(cov, e.information)) std::memset(e.information, 0, sizeof e.information);
    e.cauchy = true;
    return e;
}

/* localization.cpp:297-376 */
void Localization::addRangeEdge(const UwbRange &uwb)
{
    settle();
    ++number_measurements;
    if (!robots.count(uwb.requester_id) || !robots.count(uwb.responder_id)) {
        last_error_ = "range between unknown nodes";
        ++n_errors_;
        return;
    }
    Robot &requester = robots.at(uwb.requester_id);
    Robot &responder = robots.at(uwb.responder_id);
    const double *pr = requester.last_vertex()->estimate().t, *ps = responder.last_vertex()->estimate().t;
    const double dx = pr[0] - ps[0], dy = pr[1] - ps[1], dz = pr[2] - ps[2];
    const double distance_estimation = std::sqrt(dx * dx + dy * dy + dz * dz);
    /* outlier gate, active once the window has filled (localization.cpp:308); abs resolves to
     * std::abs(double) there (using namespace std + <cmath> through Eigen) */
    if (number_measurements > trajectory_length &&
        std::fabs(distance_estimation - (double)uwb.distance) > distance_outlier) {
        ++n_rejected_;
        return;
    }
    const double dt_requester = uwb.header.stamp.toSec() - requester.last_header().stamp.toSec();
    const double dt_responder = uwb.header.stamp.toSec() - responder.last_header().stamp.toSec();
    const double distance_cov = std::pow((double)uwb.distance_err, 2);
    const double cov_requester = std::pow(robot_max_velocity * dt_requester / 3, 2); /* 3 sigma principle */

    VertexSE3 *vertex_last_requester = requester.last_vertex();
    VertexSE3 *vertex_last_responder = responder.last_vertex();
    VertexSE3 *vertex_responder = responder.new_vertex(sensor_type.range, uwb.header, optimizer);
    const std::string frame_id = requester.last_header().frame_id;

    if (frame_id.find(uwb.header.frame_id) != std::string::npos || frame_id.find("none") != std::string::npos) {
        VertexSE3 *vertex_requester = requester.new_vertex(sensor_type.range, uwb.header, optimizer);
        Edge edge = make_range_edge(vertex_requester, vertex_responder, (double)uwb.distance, distance_cov);
        if (uwb.antenna > 0 && (size_t)uwb.antenna <= offsets.size()) edge.antenna = uwb.antenna;
        optimizer.addEdge(edge);
        optimizer.addEdge(make_range_edge(vertex_last_requester, vertex_requester, 0, cov_requester));
    } else {
        /* same pose, further range: one edge with the motion uncertainty folded in (decrease computation) */
        optimizer.addEdge(make_range_edge(vertex_last_requester, vertex_responder, (double)uwb.distance,
                                          distance_cov + cov_requester));
    }
    if (!responder.is_static()) {
        const double cov_responder = std::pow(robot_max_velocity * dt_responder / 3, 2);
        optimizer.addEdge(make_range_edge(vertex_last_responder, vertex_responder, 0, cov_responder));
    }
    if (prm.publish_range && number_measurements > trajectory_length) {
        solve_and_publish();
    }
}

/* localization.cpp:254-290: EdgeSE3 from the key vertex of the current keyframe to the new vertex */
void Localization::addPoseEdge(const PoseWithCovarianceStamped &pose_cov)
{
    settle();
    Robot &me = robots.at(self_id);
    if (pose_cov.header.frame_id != me.last_header(sensor_type.pose).frame_id) key_vertex = me.last_vertex(sensor_type.pose);
    VertexSE3 *fresh = me.new_vertex(sensor_type.pose, pose_cov.header, optimizer);
    Edge e;
    e.kind = EdgeKind::SE3;
    e.from = key_vertex;
    e.to = fresh;
    const Quaternion &q = pose_cov.pose.orientation;
    quaternion_to_matrix(q.w, q.x, q.y, q.z, e.measurement.R);
    e.measurement.t[0] = pose_cov.pose.position.x;
    e.measurement.t[1] = pose_cov.pose.position.y;
    e.measurement.t[2] = pose_cov.pose.position.z;
    if (!invert6(pose_cov.covariance.data(), e.information)) std::memset(e.information, 0, sizeof e.information);
    e.cauchy = true;
    optimizer.addEdge(e);
    if (prm.publish_pose) {
        solve_and_publish();
    }
}

/* localization.cpp:462-496 */
void Localization::addLidarEdge(const PoseWithCovarianceStamped &pose_cov)
{
    settle();
    Robot &me = robots.at(self_id);
    if (me.last_header().frame_id.find(pose_cov.header.frame_id) == std::string::npos) {
        me.append_last_header(pose_cov.header.frame_id);
        VertexSE3 *last = me.last_vertex(sensor_type.range);
        Isometry3d current_pose = last->estimate();
        current_pose.t[2] = pose_cov.pose.position.z;
        last->setEstimate(current_pose);
        Edge prior;
        prior.kind = EdgeKind::Prior;
        prior.from = last;
        prior.measurement = current_pose;
        prior.information[6 * 2 + 2] = 1 / 0.05;
        optimizer.addEdge(prior);
    }
    if (prm.publish_lidar) {
        solve_and_publish();
    }
}

/* localization.cpp:499-535 */
void Localization::addImuEdge(const Imu &imu)
{
    settle();
    Robot &me = robots.at(self_id);
    if (me.last_header().frame_id.find(imu.header.frame_id) == std::string::npos) {
        me.append_last_header(imu.header.frame_id);
        VertexSE3 *last = me.last_vertex(sensor_type.range);
        Isometry3d current_pose = Isometry3d::Identity();
        quaternion_to_matrix(imu.orientation.w, imu.orientation.x, imu.orientation.y, imu.orientation.z, current_pose.R);
        std::memcpy(current_pose.t, last->estimate().t, sizeof current_pose.t);
        last->setEstimate(current_pose);
        Edge prior;
        prior.kind = EdgeKind::Prior;
        prior.from = last;
        prior.measurement = current_pose;
        prior.information[6 * 3 + 3] = 1.0 / imu.orientation_covariance[0];
        prior.information[6 * 4 + 4] = 1.0 / imu.orientation_covariance[4];
        prior.information[6 * 5 + 5] = 1.0 / imu.orientation_covariance[8]; /* roll, pitch, yaw */
        optimizer.addEdge(prior);
    }
    if (prm.publish_imu) {
        solve_and_publish();
    }
}

/* localization.cpp:438-459 */
void Localization::addTwistEdge(const TwistWithCovarianceStamped &twist)
{
    settle();
    Robot &me = robots.at(self_id);
    const double dt = twist.header.stamp.toSec() - me.last_header().stamp.toSec();
    VertexSE3 *last = me.last_vertex();
    VertexSE3 *fresh = me.new_vertex(sensor_type.twist, twist.header, optimizer);
    optimizer.addEdge(make_se3_edge_from_twist(last, fresh, twist, dt));
    if (prm.publish_twist) {
        solve_and_publish();
    }
}

/* localization.cpp:539-557: dynamic_reconfigure callback re-publishes the newer half of the path */
void Localization::configCallback(bool publish_optimized_poses)
{
    if (!publish_optimized_poses) return;
    Path *path = robots.at(self_id).vertices2path();
    for (int i = trajectory_length / 2; i < trajectory_length; ++i) republished_.push_back(path->poses[(size_t)i]);
}

/* ---- TUM-format logs (localization.cpp:630-703) ---------------------------------------------- */
void Localization::save_file(const PoseStamped &pose, const std::string &filename)
{
    /* "%.9f" stamp, the rest with the stream's default formatting (6 significant digits = %g) */
    FILE *f = std::fopen(filename.c_str(), "a");
    if (!f) return;
    std::fprintf(f, "%.9f %g %g %g %g %g %g %g\n", pose.header.stamp.toSec(), pose.pose.position.x,
                 pose.pose.position.y, pose.pose.position.z, pose.pose.orientation.x, pose.pose.orientation.y,
                 pose.pose.orientation.z, pose.pose.orientation.w);
    std::fclose(f);
}

void Localization::write_log_headers(const std::vector<double> *antennaOffset)
{
    flag_save_file = true;
    std::string suffix = prm.filename_suffix;
    if (suffix.empty()) {
        char s[40];
        time_t now = time(nullptr);
        struct tm tim = *localtime(&now);
        strftime(s, sizeof s, "_%Y_%b_%d_%H_%M_%S.txt", &tim);
        suffix = s;
    }
    realtime_filename = prm.filename_prefix + "_realtime" + suffix;
    optimized_filename = prm.filename_prefix + "_optimized" + suffix;
    for (const std::string *name : {&realtime_filename, &optimized_filename}) {
        FILE *f = std::fopen(name->c_str(), "w");
        if (!f) continue;
        std::fprintf(f, "# iteration_max:%d\n# trajectory_length:%d\n# maximum_velocity:%g\n", iteration_max,
                     trajectory_length, robot_max_velocity);
        std::fclose(f);
    }
    if (antennaOffset && !antennaOffset->empty()) {
        /* the reference re-opens the optimized log truncating it, so only this line survives there */
        FILE *f = std::fopen(optimized_filename.c_str(), "w");
        if (f) {
            std::fprintf(f, "# antenna offsets: ");
            for (size_t i = 0; i + 1 < antennaOffset->size(); ++i) std::fprintf(f, "%g,", (*antennaOffset)[i]);
            std::fprintf(f, "%g\n", antennaOffset->back());
            std::fclose(f);
        }
    }
}
void Localization::set_file() { write_log_headers(nullptr); }
void Localization::set_file(std::vector<double> antennaOffset) { write_log_headers(&antennaOffset); }

/* ---- Fleet ---------------------------------------------------------------------------------- */
Localization &Fleet::add(const Params &p)
{
    members_.push_back(std::make_unique<Localization>(p, backend_));
    members_.back()->fleet_ = this;
    return *members_.back();
}

int Fleet::flush()
{
    std::map<std::vector<int32_t>, std::vector<Localization *>> groups;
    for (auto &m : members_)
        if (m->solve_pending_) groups[m->pending_.key()].push_back(m.get());
    int rc_all = UWBGO_OK;
    for (auto &g : groups) {
        std::vector<Localization *> &mem = g.second;
        const PackedWindow &w0 = mem[0]->pending_;
        const int64_t W = (int64_t)mem.size();
        PackedWindow cat = w0; /* structure of the group; per-window arrays are concatenated below */
        auto gather = [&](std::vector<double> PackedWindow::*f) {
            std::vector<double> &dst = cat.*f;
            dst.clear();
            for (Localization *m : mem) dst.insert(dst.end(), (m->pending_.*f).begin(), (m->pending_.*f).end());
        };
        gather(&PackedWindow::pose_t);
        gather(&PackedWindow::pose_R);
        gather(&PackedWindow::anchors);
        gather(&PackedWindow::range_d);
        gather(&PackedWindow::range_info);
        gather(&PackedWindow::prior_Z);
        gather(&PackedWindow::prior_info);
        gather(&PackedWindow::se3_Z);
        gather(&PackedWindow::se3_info);
        cat.oplus.clear();
        for (Localization *m : mem) cat.oplus.insert(cat.oplus.end(), m->pending_.oplus.begin(), m->pending_.oplus.end());
        uwbgo_topology topo; /* antenna table and iteration limit are part of the key: equal within the group */
        uwbgo_batch in;
        fill_abi(cat, W, topo, in);
        uwbgo_config cfg;
        uwbgo_config_default(&cfg);
        cfg.max_iterations = w0.iteration_max;
        const size_t N = (size_t)w0.n_poses;
        std::vector<double> pose_t((size_t)W * N * 3), pose_R((size_t)W * N * 9), chi2((size_t)W * UWBGO_CHI2_STRIDE);
        std::vector<int32_t> oplus((size_t)W * N), status((size_t)W * UWBGO_STATUS_STRIDE);
        uwbgo_result out{pose_t.data(), pose_R.data(), oplus.data(), chi2.data(), status.data()};
        int rc = backend_(&topo, &in, &cfg, &out);
        ++batches_;
        for (size_t k = 0; k < mem.size(); ++k) {
            Localization *m = mem[k];
            if (rc == UWBGO_OK) {
                m->unpack(m->pending_, pose_t.data() + k * N * 3, pose_R.data() + k * N * 9, oplus.data() + k * N,
                          chi2.data() + k * UWBGO_CHI2_STRIDE, status.data() + k * UWBGO_STATUS_STRIDE);
                ++windows_;
            } else {
                m->last_error_ = std::string("uwbgo_solve_batch failed: ") + uwbgo_last_error();
                ++m->n_errors_;
                rc_all = rc;
            }
            m->solve_pending_ = false;
            if (m->publish_pending_) {
                m->publish_pending_ = false;
                if (rc == UWBGO_OK) m->publish(); else ++m->n_skipped_;
            }
            m->pending_ = PackedWindow();
        }
    }
    return rc_all;
}

}  // namespace host
}  // namespace uwbgo
