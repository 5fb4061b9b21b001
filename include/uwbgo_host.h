/*
 * uwbgo_host.h — C ABI of the ROS-free host layer (localization_b200/host/): the reference's
 * Localization / Robot classes (src/localization/localization.h:99-128, robot.h:60-99) with their
 * solve() re-pointed at include/uwbgo.h, exposed to the py3 evaluation tooling through ctypes.
 *
 * A fleet is a set of Localization instances advanced in lockstep (many robots, Monte-Carlo
 * replays, parameter sweeps).  Messages are fed to members with the *_add_* calls, which mirror the
 * reference's ROS callbacks (addRangeEdge localization.cpp:297, addImuEdge :499, addLidarEdge :462,
 * addTwistEdge :438, addPoseEdge :254); whenever a callback decides to solve
 * (publish_flag/..., localization.cpp:371,491,530,455,285) the window is queued, and
 * uwbgo_fleet_flush() solves all queued windows as one uwbgo_solve_batch per window structure and
 * then runs the queued publish() steps (chi2 gate, newest + mid-window pose, TUM log lines).
 *
 * Ordering rule of the deferred mode.  The reference solves inside the callback, so its graph never
 * changes under a solve.  A queued window holds the member's vertices and a snapshot of their
 * estimates; a member therefore never has more than one window queued, and a message (or a second
 * solve) that reaches a member whose window is still queued first SETTLES that window: it is solved at
 * once, on its own, outside the batch (counted in stats[6]).  Results are those of the reference's
 * synchronous order either way; to get one batch per step, call uwbgo_fleet_flush() before feeding a
 * member that has solved its next message.
 */
#ifndef UWBGO_HOST_H
#define UWBGO_HOST_H

#include "uwbgo.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct uwbgo_fleet uwbgo_fleet;

/* same contract as uwbgo_solve_batch with the context replaced by `user` */
typedef int (*uwbgo_solve_fn)(void *user, const uwbgo_topology *topo, const uwbgo_batch *in,
                              const uwbgo_config *cfg, uwbgo_result *out);

/* the ROS parameters Localization::Localization reads (localization.cpp:58-159) */
typedef struct uwbgo_loc_params {
    int32_t trajectory_length;       /* robot/trajectory_length            */
    int32_t maximum_iteration;       /* optimizer/maximum_iteration        */
    double  maximum_velocity;        /* robot/maximum_velocity             */
    double  distance_outlier;        /* robot/distance_outlier             */
    double  minimum_optimize_error;  /* optimizer/minimum_optimize_error   */
    int32_t n_nodes;                 /* /uwb/nodesId: anchors..., self LAST */
    int32_t n_antennas;              /* /uwb/antennaOffset / 3             */
    const int32_t *nodes_id;         /* [n_nodes]                          */
    const double  *nodes_pos;        /* [n_nodes][3]                       */
    const double  *antenna_offset;   /* [n_antennas][3]                    */
    int32_t publish_range, publish_pose, publish_twist, publish_lidar, publish_imu; /* publish_flag/... */
    int32_t reserved;
    const char *filename_prefix;     /* log/filename_prefix; NULL = no log files */
    const char *filename_suffix;     /* NULL = the reference's _%Y_%b_%d_%H_%M_%S.txt */
} uwbgo_loc_params;

/* solver = uwbgo_solve_batch on ctx (the product path) */
uwbgo_fleet *uwbgo_fleet_create(uwbgo_ctx *ctx);
/* solver = caller-supplied function (test infrastructure) */
uwbgo_fleet *uwbgo_fleet_create_with_solver(uwbgo_solve_fn fn, void *user);
void uwbgo_fleet_destroy(uwbgo_fleet *f);
/* returns the member index (>= 0) or a negative UWBGO_E_* code */
int  uwbgo_fleet_add(uwbgo_fleet *f, const uwbgo_loc_params *p);
int  uwbgo_fleet_size(const uwbgo_fleet *f);
int  uwbgo_fleet_flush(uwbgo_fleet *f);

/* one message to one member */
int uwbgo_fleet_add_range(uwbgo_fleet *f, int member, uint32_t seq, uint32_t sec, uint32_t nsec,
                          const char *frame_id, int requester_id, int responder_id, float distance,
                          float distance_err, int antenna);
int uwbgo_fleet_add_imu(uwbgo_fleet *f, int member, uint32_t seq, uint32_t sec, uint32_t nsec,
                        const char *frame_id, const double *orientation_xyzw, const double *orientation_cov9);
int uwbgo_fleet_add_lidar(uwbgo_fleet *f, int member, uint32_t seq, uint32_t sec, uint32_t nsec,
                          const char *frame_id, double z);
int uwbgo_fleet_add_twist(uwbgo_fleet *f, int member, uint32_t seq, uint32_t sec, uint32_t nsec,
                          const char *frame_id, const double *linear3, const double *angular3,
                          const double *cov36);
int uwbgo_fleet_add_pose(uwbgo_fleet *f, int member, uint32_t seq, uint32_t sec, uint32_t nsec,
                         const char *frame_id, const double *position3, const double *orientation_xyzw,
                         const double *cov36);
/* the same range message to every member, with per-member measured values (Monte-Carlo replay) */
int uwbgo_fleet_add_range_each(uwbgo_fleet *f, uint32_t seq, uint32_t sec, uint32_t nsec,
                               const char *frame_id, int requester_id, int responder_id,
                               const float *distance, const float *distance_err, int antenna);
int uwbgo_fleet_add_imu_each(uwbgo_fleet *f, uint32_t seq, uint32_t sec, uint32_t nsec,
                             const char *frame_id, const double *orientation_xyzw /* [M][4] */,
                             const double *orientation_cov9 /* [9], shared */);

/* A range edge built through the reference's edge classes between vertices the member already has:
 * EDGE_RANGE = EdgeSE3Range with setVertexOffset(0 / 1, ...) (src/types/types_edge_se3range.cpp:99-114),
 * EDGE_RANGE_OFFSET = EdgeSE3RangeOffset with setParameterId(0 / 1, ...) on a parameter table whose id k
 * is antenna k, id 0 the identity (types_edge_se3range_offset.cpp:61-79,126-149).  from_age / to_age
 * count ring slots from the oldest (to_age > from_age); to_anchor >= 0 names a fixed node id instead of
 * to_age.  off_from / off_to: antenna numbers, 0 = identity.  An offset the solve path cannot carry is
 * refused (UWBGO_E_INVALID, uwbgo_fleet_last_error), never dropped. */
#define UWBGO_EDGE_CLASS_RANGE        0
#define UWBGO_EDGE_CLASS_RANGE_OFFSET 1
int uwbgo_fleet_add_typed_range_edge(uwbgo_fleet *f, int member, int edge_class, int from_age, int to_age,
                                     int to_anchor, double measurement, double information, int off_from,
                                     int off_to, int cauchy);
/* solve() + publish() of one member now (queued for uwbgo_fleet_flush like a callback's solve) */
int uwbgo_fleet_solve(uwbgo_fleet *f, int member);

/* results of member `member` */
int64_t uwbgo_fleet_published_count(const uwbgo_fleet *f, int member);
/* k-th publish(): realtime / optimized = {stamp, x, y, z, qx, qy, qz, qw}; error = optimizer.chi2() */
int uwbgo_fleet_published(const uwbgo_fleet *f, int member, int64_t k, double *realtime8,
                          double *optimized8, double *error);
/* all publishes of a member at once: arrays [count][8], [count][8], [count] */
int uwbgo_fleet_published_all(const uwbgo_fleet *f, int member, double *realtime8, double *optimized8,
                              double *error);
/* stats[0..6] = solves, rejected ranges, skipped publishes, errors, total windows solved by the
 * fleet's batches, batches issued by the fleet, windows of this member settled on their own */
int uwbgo_fleet_stats(const uwbgo_fleet *f, int member, int64_t *stats7);
/* chi2[4] and status[4] of the member's last solve */
int uwbgo_fleet_last_solve(const uwbgo_fleet *f, int member, double *chi2_4, int32_t *status4);
const char *uwbgo_fleet_last_error(const uwbgo_fleet *f, int member);
/* window the member would solve right now: current poses oldest -> newest, [N][3]; returns N */
int uwbgo_fleet_window_poses(uwbgo_fleet *f, int member, double *pose_t, int capacity);

#ifdef __cplusplus
}
#endif
#endif /* UWBGO_HOST_H */
