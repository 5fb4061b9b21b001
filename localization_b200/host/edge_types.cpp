#include "edge_types.h"

#include <cmath>
#include <istream>
#include <ostream>
#include <string>

namespace uwbgo {
namespace host {

Isometry3d compose(const Isometry3d &a, const Isometry3d &b)
{
    Isometry3d o;
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c)
            o.R[3 * r + c] = a.R[3 * r] * b.R[c] + a.R[3 * r + 1] * b.R[3 + c] + a.R[3 * r + 2] * b.R[6 + c];
        o.t[r] = a.R[3 * r] * b.t[0] + a.R[3 * r + 1] * b.t[1] + a.R[3 * r + 2] * b.t[2] + a.t[r];
    }
    return o;
}

Isometry3d inverse(const Isometry3d &a)
{
    Isometry3d o;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) o.R[3 * r + c] = a.R[3 * c + r];
    for (int r = 0; r < 3; ++r) o.t[r] = -(o.R[3 * r] * a.t[0] + o.R[3 * r + 1] * a.t[1] + o.R[3 * r + 2] * a.t[2]);
    return o;
}

int antenna_number(const Isometry3d &offset, const std::vector<Isometry3d> &antennas, const char *what)
{
    if (!offset.rotationIsIdentity())
        throw std::invalid_argument(std::string(what) + ": the offset has a rotation; the solve path carries "
                                    "translation-only lever arms (antenna numbers)");
    if (offset.t[0] == 0 && offset.t[1] == 0 && offset.t[2] == 0) return 0;
    for (size_t k = 0; k < antennas.size(); ++k)
        if (antennas[k].t[0] == offset.t[0] && antennas[k].t[1] == offset.t[1] && antennas[k].t[2] == offset.t[2])
            return (int)k + 1;
    throw std::invalid_argument(std::string(what) + ": the lever arm is not in the window's antenna table");
}

static double range_between(const Isometry3d &p0, const Isometry3d &p1)
{
    const double dx = p0.t[0] - p1.t[0], dy = p0.t[1] - p1.t[1], dz = p0.t[2] - p1.t[2];
    return std::sqrt(dx * dx + dy * dy + dz * dz);
}

/* types_edge_se3range.cpp:67-97: slide the unknown end point along the line of sight until the
 * distance equals the measurement */
void RangeEdgeBase::initialEstimate(const VertexSet &from_, VertexSE3 * /*to_*/)
{
    VertexSE3 *v1 = vertices_[0], *v2 = vertices_[1];
    const bool forward = from_.count(v1) == 1;
    VertexSE3 *known = forward ? v1 : v2, *unknown = forward ? v2 : v1;
    Isometry3d delta = compose(inverse(known->estimate()), unknown->estimate());
    const double norm = std::sqrt(delta.t[0] * delta.t[0] + delta.t[1] * delta.t[1] + delta.t[2] * delta.t[2]);
    const double alpha = measurement_ / norm;
    for (double &c : delta.t) c *= alpha;
    unknown->setEstimate(compose(known->estimate(), delta));
}

bool EdgeSE3Range::read(std::istream &is)
{
    double meas;
    is >> meas;
    setMeasurement(meas);
    information_ = 1.0;
    is >> information_;
    return true;
}
bool EdgeSE3Range::write(std::ostream &os) const
{
    os << measurement_ << " " << information_;
    return os.good();
}
/* types_edge_se3range.cpp:105-114 */
void EdgeSE3Range::computeError()
{
    error_ = measurement_ - range_between(compose(vertices_[0]->estimate(), offset[0]),
                                          compose(vertices_[1]->estimate(), offset[1]));
}
Edge EdgeSE3Range::asEdge(const std::vector<Isometry3d> &antennas) const
{
    Edge e;
    e.kind = EdgeKind::Range;
    e.from = vertices_[0];
    e.to = vertices_[1];
    e.range = measurement_;
    e.rangeInformation = information_;
    e.cauchy = cauchy_;
    e.antenna = antenna_number(offset[0], antennas, "EdgeSE3Range offset[0]");
    e.antenna_b = antenna_number(offset[1], antennas, "EdgeSE3Range offset[1]");
    return e;
}

bool EdgeSE3RangeOffset::setParameterId(int argNum, int paramId)
{
    if (argNum < 0 || argNum > 1) return false;
    if (params_ && !params_->count(paramId)) return false;
    pid_[argNum] = paramId;
    return true;
}
bool EdgeSE3RangeOffset::read(std::istream &is)
{
    int pidFrom, pidTo;
    is >> pidFrom >> pidTo;
    if (!setParameterId(0, pidFrom)) return false;
    if (!setParameterId(1, pidTo)) return false;
    double meas;
    is >> meas;
    setMeasurement(meas);
    information_ = 1.0;
    is >> information_;
    return true;
}
bool EdgeSE3RangeOffset::write(std::ostream &os) const
{
    os << pid_[0] << " " << pid_[1] << " " << measurement_ << " " << information_;
    return os.good();
}
/* types_edge_se3range_offset.cpp:126-131: CacheSE3Offset::n2w() = estimate * offset */
void EdgeSE3RangeOffset::computeError()
{
    static const Isometry3d identity;
    auto off = [&](int k) -> const Isometry3d & {
        if (!params_) return identity;
        auto it = params_->find(pid_[k]);
        return it == params_->end() ? identity : it->second.offset();
    };
    error_ = measurement_ - range_between(compose(vertices_[0]->estimate(), off(0)),
                                          compose(vertices_[1]->estimate(), off(1)));
}
Edge EdgeSE3RangeOffset::asEdge(const std::vector<Isometry3d> &antennas) const
{
    auto off = [&](int k, const char *what) -> int {
        if (!params_) { /* no parameter table: only the identity id 0 the reference registers */
            if (pid_[k] != 0) throw std::invalid_argument(std::string(what) + ": parameter id without a parameter table");
            return 0;
        }
        auto it = params_->find(pid_[k]);
        if (it == params_->end()) throw std::invalid_argument(std::string(what) + ": unknown parameter id");
        return antenna_number(it->second.offset(), antennas, what);
    };
    Edge e;
    e.kind = EdgeKind::Range;
    e.from = vertices_[0];
    e.to = vertices_[1];
    e.range = measurement_;
    e.rangeInformation = information_;
    e.cauchy = cauchy_;
    e.antenna = off(0, "EdgeSE3RangeOffset pidFrom");
    e.antenna_b = off(1, "EdgeSE3RangeOffset pidTo");
    return e;
}

}  // namespace host
}  // namespace uwbgo
