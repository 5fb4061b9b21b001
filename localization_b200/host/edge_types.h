/*
 * edge_types.h — host-side drop-ins for the reference's two g2o edge types
 *   g2o::EdgeSE3Range        (src/types/types_edge_se3range.h:45-74, .cpp:39-114)
 *   g2o::EdgeSE3RangeOffset  (src/types/types_edge_se3range_offset.h:45-78, .cpp:39-149)
 * with the same public methods: setMeasurement, setVertexOffset / setParameterId, vertices(),
 * setInformation, setRobustKernel, computeError, read / write (g2o text payload), initialEstimate,
 * initialEstimatePossible.  They are descriptors: asEdge() turns one into the plain Edge record
 * the window packer consumes; the batched solve never calls computeError() on the host (it exists
 * for inspection, initialEstimate and the outlier tooling, as in the reference).
 *
 * The solve path knows offsets as ANTENNA NUMBERS into the window's antenna table (translation-only
 * lever arms, /uwb/antennaOffset, localization.cpp:111-123).  asEdge() therefore resolves both
 * offsets of an edge -- offset[0] / offset[1], resp. the offsets of pidFrom / pidTo -- against that
 * table and THROWS std::invalid_argument for an offset the solve path cannot honour (a rotation, or a
 * lever arm that is not in the table): what computeError() evaluates on the host and what the GPU
 * optimises are the same residual, or the call fails.
 */
#ifndef UWBGO_HOST_EDGE_TYPES_H
#define UWBGO_HOST_EDGE_TYPES_H

#include <iosfwd>
#include <map>
#include <set>
#include <stdexcept>
#include <vector>

#include "window_graph.h"

namespace uwbgo {
namespace host {

struct RobustKernelCauchy {}; /* g2o::RobustKernelCauchy with its default delta = 1 */

/* g2o::ParameterSE3Offset: an id'd rigid offset shared by edges */
struct ParameterSE3Offset {
    int id_ = 0;
    Isometry3d offset_;
    void setId(int id) { id_ = id; }
    int id() const { return id_; }
    void setOffset(const Isometry3d &o) { offset_ = o; }
    const Isometry3d &offset() const { return offset_; }
};

Isometry3d compose(const Isometry3d &a, const Isometry3d &b);
Isometry3d inverse(const Isometry3d &a);
/* antenna number (1-based) of a translation-only offset in `antennas`, 0 for the identity; throws
 * std::invalid_argument when the offset rotates or its lever arm is not in the table */
int antenna_number(const Isometry3d &offset, const std::vector<Isometry3d> &antennas, const char *what);

class RangeEdgeBase {
public:
    using VertexSet = std::set<VertexSE3 *>;
    RangeEdgeBase() : vertices_(2, nullptr) {}
    virtual ~RangeEdgeBase() {}
    std::vector<VertexSE3 *> &vertices() { return vertices_; }
    const std::vector<VertexSE3 *> &vertices() const { return vertices_; }
    virtual void setMeasurement(const double &m) { measurement_ = m; }
    double measurement() const { return measurement_; }
    void setInformation(double info) { information_ = info; } /* 1x1 information matrix */
    double information() const { return information_; }
    void setRobustKernel(RobustKernelCauchy *k) { cauchy_ = k != nullptr; delete k; }
    bool robustKernel() const { return cauchy_; }
    virtual void computeError() = 0;
    double error() const { return error_; }
    double chi2() const { return error_ * information_ * error_; }
    virtual double initialEstimatePossible(const VertexSet &, VertexSE3 *) { return 1.; }
    virtual void initialEstimate(const VertexSet &from_, VertexSE3 *to_);

protected:
    std::vector<VertexSE3 *> vertices_;
    double measurement_ = 0.0, information_ = 1.0, error_ = 0.0;
    bool cauchy_ = false;
};

/* factory tag EDGE_RANGE; text payload "meas info" */
class EdgeSE3Range : public RangeEdgeBase {
public:
    bool read(std::istream &is);
    bool write(std::ostream &os) const;
    void computeError() override;
    void setVertexOffset(int vertex, Isometry3d &pose) { offset[(size_t)vertex] = pose; }
    std::vector<Isometry3d> offset = std::vector<Isometry3d>(2, Isometry3d::Identity());
    /* both offsets resolved against the window's antenna table (see the header comment) */
    Edge asEdge(const std::vector<Isometry3d> &antennas) const;
    static const char *tag() { return "EDGE_RANGE"; }
};

/* factory tag EDGE_RANGE_OFFSET; text payload "pidFrom pidTo meas info"; offsets come from
 * ParameterSE3Offset ids (g2o Parameter/Cache mechanism) */
class EdgeSE3RangeOffset : public RangeEdgeBase {
public:
    explicit EdgeSE3RangeOffset(const std::map<int, ParameterSE3Offset> *params = nullptr) : params_(params) {}
    bool read(std::istream &is);
    bool write(std::ostream &os) const;
    void computeError() override;
    bool setParameterId(int argNum, int paramId);
    int parameterId(int argNum) const { return pid_[argNum]; }
    /* the offsets of pidFrom / pidTo resolved against the window's antenna table; a parameter id that is
     * not registered throws (g2o's resolveCaches() would fail, types_edge_se3range_offset.cpp:134-149) */
    Edge asEdge(const std::vector<Isometry3d> &antennas) const;
    static const char *tag() { return "EDGE_RANGE_OFFSET"; }

private:
    const std::map<int, ParameterSE3Offset> *params_;
    int pid_[2] = {0, 0};
};

}  // namespace host
}  // namespace uwbgo
#endif
