#include "window_graph.h"

#include <cmath>

namespace uwbgo {
namespace host {

VertexSE3 *Graph::addVertex(std::unique_ptr<VertexSE3> v)
{
    owned_.push_back(std::move(v));
    return owned_.back().get();
}

void Graph::removeVertex(VertexSE3 *v)
{
    edges_.remove_if([v](const Edge &e) { return e.from == v || e.to == v; });
    owned_.remove_if([v](const std::unique_ptr<VertexSE3> &p) { return p.get() == v; });
}

Pose pose_to_msg(const Isometry3d &e)
{
    const double *R = e.R;
    double q[4]; /* x y z w */
    const double trace = R[0] + R[4] + R[8];
    if (trace > 0.0) {
        const double s = std::sqrt(trace + 1.0);
        q[3] = 0.5 * s;
        q[0] = (R[7] - R[5]) * (0.5 / s);
        q[1] = (R[2] - R[6]) * (0.5 / s);
        q[2] = (R[3] - R[1]) * (0.5 / s);
    } else {
        int i = R[4] > R[0] ? 1 : 0;
        if (R[8] > R[4 * i]) i = 2;
        const int j = (i + 1) % 3, k = (i + 2) % 3;
        const double s = std::sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
        q[i] = 0.5 * s;
        q[3] = (R[3 * k + j] - R[3 * j + k]) * (0.5 / s);
        q[j] = (R[3 * j + i] + R[3 * i + j]) * (0.5 / s);
        q[k] = (R[3 * k + i] + R[3 * i + k]) * (0.5 / s);
    }
    const double sign = q[3] < 0 ? -1.0 : 1.0;
    Pose m;
    m.position = Point{e.t[0], e.t[1], e.t[2]};
    m.orientation = Quaternion{sign * q[0], sign * q[1], sign * q[2], sign * q[3]};
    return m;
}

Robot::Robot(int ID, bool FLAG_STATIC, int trajectory_length)
    : node_id_(ID), anchored_(FLAG_STATIC), capacity_(trajectory_length)
{
}

void Robot::init(Graph &optimizer, Isometry3d vertex_init)
{
    ring_.assign((size_t)capacity_, Slot());
    path_.poses.assign((size_t)capacity_, PoseStamped());
    newest_ = 0;
    for (int slot = 0; slot < capacity_; ++slot) {
        auto v = std::make_unique<VertexSE3>();
        v->setId(node_id_ + slot * 300);
        v->setEstimate(vertex_init);
        v->setFixed(anchored_);
        ring_[slot].vertex = optimizer.addVertex(std::move(v));
    }
    ring_[0].stamp.frame_id = "none"; /* robot.cpp:57: lets the first range message open a vertex */
}

VertexSE3 *Robot::new_vertex(unsigned char type, const Header &new_header, Graph &optimizer)
{
    slot_of_.emplace(type, newest_);
    header_of_.emplace(type, new_header);
    if (anchored_) {
        ring_[newest_].stamp = new_header;
        return last_vertex(type);
    }
    auto fresh = std::make_unique<VertexSE3>();
    fresh->setEstimate(ring_[newest_].vertex->estimate()); /* start at the newest estimate */
    newest_ = (newest_ + 1) % capacity_;
    fresh->setId(newest_ * 300 + node_id_);
    optimizer.removeVertex(ring_[newest_].vertex);         /* evict the oldest, edges included */
    ring_[newest_].vertex = optimizer.addVertex(std::move(fresh));
    ring_[newest_].stamp = new_header;
    slot_of_.at(type) = newest_;
    header_of_.at(type) = new_header;
    return ring_[newest_].vertex;
}

VertexSE3 *Robot::last_vertex(unsigned char type)
{
    slot_of_.emplace(type, newest_);
    header_of_.emplace(type, ring_[newest_].stamp);
    return ring_.at((size_t)slot_of_[type]).vertex;
}
VertexSE3 *Robot::last_vertex() { return ring_.at((size_t)newest_).vertex; }
Header Robot::last_header(unsigned char type)
{
    header_of_.emplace(type, ring_[newest_].stamp);
    return header_of_.at(type);
}
Header Robot::last_header() { return ring_[newest_].stamp; }
void Robot::append_last_header(const std::string &frame_id) { ring_[newest_].stamp.frame_id += "-" + frame_id; }

Path *Robot::vertices2path()
{
    for (int age = 0; age < capacity_; ++age) {
        const Slot &s = ring_[(newest_ + 1 + age) % capacity_];
        path_.poses[age].pose = pose_to_msg(s.vertex->estimate());
        path_.poses[age].header = s.stamp;
    }
    path_.header = last_header();
    return &path_;
}

PoseStamped Robot::current_pose()
{
    PoseStamped p;
    p.header = last_header();
    p.pose = pose_to_msg(last_vertex()->estimate());
    return p;
}

}  // namespace host
}  // namespace uwbgo
