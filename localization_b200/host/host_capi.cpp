/* host_capi.cpp — extern "C" surface of the host layer (include/uwbgo_host.h). */
#include <cstring>
#include <memory>

#include "../../include/uwbgo_host.h"
#include "host_localization.h"

using namespace uwbgo::host;

struct uwbgo_fleet {
    std::unique_ptr<Fleet> fleet;
};

static Header make_header(uint32_t seq, uint32_t sec, uint32_t nsec, const char *frame_id)
{
    Header h;
    h.seq = seq;
    h.stamp.sec = sec;
    h.stamp.nsec = nsec;
    h.frame_id = frame_id ? frame_id : "";
    return h;
}
static bool bad(const uwbgo_fleet *f, int member) { return !f || member < 0 || (size_t)member >= f->fleet->size(); }
static void stamp_pose(const PoseStamped &p, double *o)
{
    o[0] = p.header.stamp.toSec();
    o[1] = p.pose.position.x; o[2] = p.pose.position.y; o[3] = p.pose.position.z;
    o[4] = p.pose.orientation.x; o[5] = p.pose.orientation.y; o[6] = p.pose.orientation.z; o[7] = p.pose.orientation.w;
}

extern "C" {

uwbgo_fleet *uwbgo_fleet_create(uwbgo_ctx *ctx)
{
    if (!ctx) return nullptr;
    auto f = new uwbgo_fleet();
    f->fleet = std::make_unique<Fleet>(backend_from_ctx(ctx));
    return f;
}

uwbgo_fleet *uwbgo_fleet_create_with_solver(uwbgo_solve_fn fn, void *user)
{
    if (!fn) return nullptr;
    auto f = new uwbgo_fleet();
    f->fleet = std::make_unique<Fleet>(
        [fn, user](const uwbgo_topology *t, const uwbgo_batch *b, const uwbgo_config *c, uwbgo_result *r) {
            return fn(user, t, b, c, r);
        });
    return f;
}

void uwbgo_fleet_destroy(uwbgo_fleet *f) { delete f; }

int uwbgo_fleet_add(uwbgo_fleet *f, const uwbgo_loc_params *p)
{
    if (!f || !p || p->n_nodes < 1 || !p->nodes_id || !p->nodes_pos || p->trajectory_length < 1) return UWBGO_E_INVALID;
    Params q;
    q.trajectory_length = p->trajectory_length;
    q.maximum_iteration = p->maximum_iteration;
    q.maximum_velocity = p->maximum_velocity;
    q.distance_outlier = p->distance_outlier;
    q.minimum_optimize_error = p->minimum_optimize_error;
    q.nodesId.assign(p->nodes_id, p->nodes_id + p->n_nodes);
    q.nodesPos.assign(p->nodes_pos, p->nodes_pos + 3 * (size_t)p->n_nodes);
    if (p->n_antennas > 0 && p->antenna_offset) q.antennaOffset.assign(p->antenna_offset, p->antenna_offset + 3 * (size_t)p->n_antennas);
    q.publish_range = p->publish_range != 0;
    q.publish_pose = p->publish_pose != 0;
    q.publish_twist = p->publish_twist != 0;
    q.publish_lidar = p->publish_lidar != 0;
    q.publish_imu = p->publish_imu != 0;
    if (p->filename_prefix) q.filename_prefix = p->filename_prefix;
    if (p->filename_suffix) q.filename_suffix = p->filename_suffix;
    f->fleet->add(q);
    return (int)f->fleet->size() - 1;
}

int uwbgo_fleet_size(const uwbgo_fleet *f) { return f ? (int)f->fleet->size() : 0; }
int uwbgo_fleet_flush(uwbgo_fleet *f) { return f ? f->fleet->flush() : UWBGO_E_INVALID; }

int uwbgo_fleet_add_range(uwbgo_fleet *f, int member, uint32_t seq, uint32_t sec, uint32_t nsec, const char *frame_id,
                          int requester_id, int responder_id, float distance, float distance_err, int antenna)
{
    if (bad(f, member)) return UWBGO_E_INVALID;
    UwbRange m;
    m.header = make_header(seq, sec, nsec, frame_id);
    m.requester_id = requester_id;
    m.responder_id = responder_id;
    m.distance = distance;
    m.distance_err = distance_err;
    m.antenna = antenna;
    f->fleet->at((size_t)member).addRangeEdge(m);
    return UWBGO_OK;
}

int uwbgo_fleet_add_imu(uwbgo_fleet *f, int member, uint32_t seq, uint32_t sec, uint32_t nsec, const char *frame_id,
                        const double *q, const double *cov9)
{
    if (bad(f, member) || !q || !cov9) return UWBGO_E_INVALID;
    Imu m;
    m.header = make_header(seq, sec, nsec, frame_id);
    m.orientation = Quaternion{q[0], q[1], q[2], q[3]};
    std::memcpy(m.orientation_covariance.data(), cov9, 9 * sizeof(double));
    f->fleet->at((size_t)member).addImuEdge(m);
    return UWBGO_OK;
}

int uwbgo_fleet_add_lidar(uwbgo_fleet *f, int member, uint32_t seq, uint32_t sec, uint32_t nsec, const char *frame_id, double z)
{
    if (bad(f, member)) return UWBGO_E_INVALID;
    PoseWithCovarianceStamped m;
    m.header = make_header(seq, sec, nsec, frame_id);
    m.pose.position.z = z;
    f->fleet->at((size_t)member).addLidarEdge(m);
    return UWBGO_OK;
}

int uwbgo_fleet_add_twist(uwbgo_fleet *f, int member, uint32_t seq, uint32_t sec, uint32_t nsec, const char *frame_id,
                          const double *lin, const double *ang, const double *cov36)
{
    if (bad(f, member) || !lin || !ang || !cov36) return UWBGO_E_INVALID;
    TwistWithCovarianceStamped m;
    m.header = make_header(seq, sec, nsec, frame_id);
    m.twist.linear = Point{lin[0], lin[1], lin[2]};
    m.twist.angular = Point{ang[0], ang[1], ang[2]};
    std::memcpy(m.covariance.data(), cov36, 36 * sizeof(double));
    f->fleet->at((size_t)member).addTwistEdge(m);
    return UWBGO_OK;
}

int uwbgo_fleet_add_pose(uwbgo_fleet *f, int member, uint32_t seq, uint32_t sec, uint32_t nsec, const char *frame_id,
                         const double *pos, const double *q, const double *cov36)
{
    if (bad(f, member) || !pos || !q || !cov36) return UWBGO_E_INVALID;
    PoseWithCovarianceStamped m;
    m.header = make_header(seq, sec, nsec, frame_id);
    m.pose.position = Point{pos[0], pos[1], pos[2]};
    m.pose.orientation = Quaternion{q[0], q[1], q[2], q[3]};
    std::memcpy(m.covariance.data(), cov36, 36 * sizeof(double));
    f->fleet->at((size_t)member).addPoseEdge(m);
    return UWBGO_OK;
}

int uwbgo_fleet_add_range_each(uwbgo_fleet *f, uint32_t seq, uint32_t sec, uint32_t nsec, const char *frame_id,
                               int requester_id, int responder_id, const float *distance, const float *distance_err,
                               int antenna)
{
    if (!f || !distance || !distance_err) return UWBGO_E_INVALID;
    UwbRange m;
    m.header = make_header(seq, sec, nsec, frame_id);
    m.requester_id = requester_id;
    m.responder_id = responder_id;
    m.antenna = antenna;
    for (size_t i = 0; i < f->fleet->size(); ++i) {
        m.distance = distance[i];
        m.distance_err = distance_err[i];
        f->fleet->at(i).addRangeEdge(m);
    }
    return UWBGO_OK;
}

int uwbgo_fleet_add_imu_each(uwbgo_fleet *f, uint32_t seq, uint32_t sec, uint32_t nsec, const char *frame_id,
                             const double *q, const double *cov9)
{
    if (!f || !q || !cov9) return UWBGO_E_INVALID;
    Imu m;
    m.header = make_header(seq, sec, nsec, frame_id);
    std::memcpy(m.orientation_covariance.data(), cov9, 9 * sizeof(double));
    for (size_t i = 0; i < f->fleet->size(); ++i) {
        m.orientation = Quaternion{q[4 * i], q[4 * i + 1], q[4 * i + 2], q[4 * i + 3]};
        f->fleet->at(i).addImuEdge(m);
    }
    return UWBGO_OK;
}

int uwbgo_fleet_add_typed_range_edge(uwbgo_fleet *f, int member, int edge_class, int from_age, int to_age, int to_anchor,
                                     double measurement, double information, int off_from, int off_to, int cauchy)
{
    if (bad(f, member) || (edge_class != UWBGO_EDGE_CLASS_RANGE && edge_class != UWBGO_EDGE_CLASS_RANGE_OFFSET))
        return UWBGO_E_INVALID;
    return f->fleet->at((size_t)member).addTypedRangeEdge(edge_class == UWBGO_EDGE_CLASS_RANGE_OFFSET, from_age, to_age,
                                                          to_anchor, measurement, information, off_from, off_to,
                                                          cauchy != 0)
               ? UWBGO_OK
               : UWBGO_E_INVALID;
}

int uwbgo_fleet_solve(uwbgo_fleet *f, int member)
{
    if (bad(f, member)) return UWBGO_E_INVALID;
    f->fleet->at((size_t)member).solve_and_publish();
    return UWBGO_OK;
}

int64_t uwbgo_fleet_published_count(const uwbgo_fleet *f, int member)
{
    return bad(f, member) ? -1 : (int64_t)f->fleet->at((size_t)member).published().size();
}

int uwbgo_fleet_published(const uwbgo_fleet *f, int member, int64_t k, double *realtime8, double *optimized8, double *error)
{
    if (bad(f, member)) return UWBGO_E_INVALID;
    const auto &pub = f->fleet->at((size_t)member).published();
    if (k < 0 || (size_t)k >= pub.size()) return UWBGO_E_INVALID;
    if (realtime8) stamp_pose(pub[(size_t)k].realtime, realtime8);
    if (optimized8) stamp_pose(pub[(size_t)k].optimized, optimized8);
    if (error) *error = pub[(size_t)k].error;
    return UWBGO_OK;
}

int uwbgo_fleet_published_all(const uwbgo_fleet *f, int member, double *realtime8, double *optimized8, double *error)
{
    if (bad(f, member)) return UWBGO_E_INVALID;
    const auto &pub = f->fleet->at((size_t)member).published();
    for (size_t k = 0; k < pub.size(); ++k) {
        if (realtime8) stamp_pose(pub[k].realtime, realtime8 + 8 * k);
        if (optimized8) stamp_pose(pub[k].optimized, optimized8 + 8 * k);
        if (error) error[k] = pub[k].error;
    }
    return UWBGO_OK;
}

int uwbgo_fleet_stats(const uwbgo_fleet *f, int member, int64_t *s /* [7] */)
{
    if (bad(f, member) || !s) return UWBGO_E_INVALID;
    Localization &m = f->fleet->at((size_t)member);
    s[0] = m.solves();
    s[1] = m.rejected_ranges();
    s[2] = m.skipped_publishes();
    s[3] = m.solver_errors();
    s[4] = f->fleet->windows_solved();
    s[5] = f->fleet->batches();
    s[6] = m.settled_alone();
    return UWBGO_OK;
}

int uwbgo_fleet_last_solve(const uwbgo_fleet *f, int member, double *chi2_4, int32_t *status4)
{
    if (bad(f, member)) return UWBGO_E_INVALID;
    Localization &m = f->fleet->at((size_t)member);
    if (chi2_4)
        for (int k = 0; k < 4; ++k) chi2_4[k] = m.last_chi2(k);
    if (status4) std::memcpy(status4, m.last_status(), 4 * sizeof(int32_t));
    return UWBGO_OK;
}

const char *uwbgo_fleet_last_error(const uwbgo_fleet *f, int member)
{
    return bad(f, member) ? "bad fleet member" : f->fleet->at((size_t)member).last_error().c_str();
}

int uwbgo_fleet_window_poses(uwbgo_fleet *f, int member, double *pose_t, int capacity)
{
    if (bad(f, member)) return UWBGO_E_INVALID;
    PackedWindow w;
    std::string err;
    if (!f->fleet->at((size_t)member).pack(w, err)) return 0;
    if (pose_t && capacity >= w.n_poses) std::memcpy(pose_t, w.pose_t.data(), w.pose_t.size() * sizeof(double));
    return w.n_poses;
}

}  // extern "C"
