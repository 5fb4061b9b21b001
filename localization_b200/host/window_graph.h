/*
 * window_graph.h — host-side state the sliding window is packed from.
 *
 * Two things live here, both ROS-free and solver-free:
 *
 *  - Graph: the little of g2o's graph container API that the reference's Localization and Robot
 *    touch (reference src/localization/localization.cpp:44-56,263-281,481-486,588-602,610-624;
 *    src/localization/robot.cpp:41-54,88-106): SE(3) vertices with an estimate / fixed flag / id,
 *    edges kept in insertion order (g2o's internal-id order, the order it accumulates H and b),
 *    and removeVertex() which drops every edge incident to the vertex.
 *
 *  - Robot: drop-in for the reference's Robot (src/localization/robot.h:60-99): same constructor
 *    and method names and meanings.  A moving robot is a ring of `trajectory_length` vertices;
 *    new_vertex() advances the ring, evicts the oldest vertex (and with it its edges) and starts
 *    the new vertex at a copy of the newest estimate (robot.cpp:88-106); a static robot (anchor)
 *    is one fixed vertex.  Vertex ids are slot*300 + ID as in robot.cpp:43,94.
 */
#ifndef UWBGO_HOST_WINDOW_GRAPH_H
#define UWBGO_HOST_WINDOW_GRAPH_H

#include <cstdint>
#include <list>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "messages.h"

namespace uwbgo {
namespace host {

/* rigid transform, rotation row-major (what the reference holds in an Eigen::Isometry3d) */
struct Isometry3d {
    double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    double t[3] = {0, 0, 0};
    static Isometry3d Identity() { return Isometry3d(); }
    bool rotationIsIdentity() const
    {
        return R[0] == 1 && R[4] == 1 && R[8] == 1 && R[1] == 0 && R[2] == 0 && R[3] == 0 &&
               R[5] == 0 && R[6] == 0 && R[7] == 0;
    }
};

class VertexSE3 {
public:
    void setId(int id) { id_ = id; }
    int id() const { return id_; }
    void setFixed(bool f) { fixed_ = f; }
    bool fixed() const { return fixed_; }
    void setEstimate(const Isometry3d &e) { est_ = e; }
    const Isometry3d &estimate() const { return est_; }
    int32_t oplusCalls = 0; /* g2o VertexSE3::_numOplusCalls, survives across solves */

private:
    int id_ = 0;
    bool fixed_ = false;
    Isometry3d est_;
};

enum class EdgeKind { Range, Prior, SE3 };

struct Edge {
    EdgeKind kind = EdgeKind::Range;
    VertexSE3 *from = nullptr, *to = nullptr; /* vertices()[0], vertices()[1] (Prior: to unused) */
    bool cauchy = false;                      /* setRobustKernel(new RobustKernelCauchy()) */
    double range = 0.0, rangeInformation = 0.0;
    int antenna = 0;                          /* setVertexOffset(0, offsets[antenna-1]); 0: none */
    int antenna_b = 0;                        /* setVertexOffset(1, offsets[antenna_b-1]) / pidTo; 0: none */
    Isometry3d measurement;                   /* EdgeSE3 / EdgeSE3Prior */
    double information[36] = {0};             /* 6x6 row-major */
};

class Graph {
public:
    VertexSE3 *addVertex(std::unique_ptr<VertexSE3> v);
    void removeVertex(VertexSE3 *v);
    void addEdge(const Edge &e) { edges_.push_back(e); }
    const std::list<Edge> &edges() const { return edges_; }

private:
    std::list<std::unique_ptr<VertexSE3>> owned_;
    std::list<Edge> edges_;
};

class Robot {
public:
    Robot(int ID, bool FLAG_STATIC, int trajectory_length);
    void init(Graph &optimizer, Isometry3d vertex_init = Isometry3d::Identity());
    bool is_static() const { return anchored_; }
    VertexSE3 *new_vertex(unsigned char type, const Header &new_header, Graph &optimizer);
    VertexSE3 *last_vertex(unsigned char type);
    VertexSE3 *last_vertex();
    Header last_header(unsigned char type);
    Header last_header();
    void append_last_header(const std::string &frame_id);
    Path *vertices2path();
    PoseStamped current_pose();

    /* read-out for the window packer: age 0 = oldest slot, length()-1 = newest */
    int length() const { return (int)ring_.size(); }
    VertexSE3 *by_age(int age) const { return ring_[(newest_ + 1 + age) % ring_.size()].vertex; }

private:
    struct Slot {
        VertexSE3 *vertex = nullptr;
        Header stamp;
    };
    int node_id_;
    bool anchored_;
    int capacity_;
    int newest_ = 0;
    std::vector<Slot> ring_;
    /* slot and header each sensor type touched last; like the reference's two maps they are
     * filled on first touch (std::map::emplace never overwrites) and may be touched separately */
    std::map<unsigned char, int> slot_of_;
    std::map<unsigned char, Header> header_of_;
    Path path_;
};

/* tf::poseEigenToMsg: rotation matrix -> quaternion with w >= 0 */
Pose pose_to_msg(const Isometry3d &e);

}  // namespace host
}  // namespace uwbgo
#endif
