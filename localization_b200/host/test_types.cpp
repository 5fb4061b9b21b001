/* Unit test of the host-side type API: exits 0 on success, prints the failing check otherwise. */
#include <cmath>
#include <cstdio>
#include <sstream>

#include "edge_types.h"
#include "window_graph.h"

using namespace uwbgo::host;

#define CHECK(c)                                              \
    do {                                                      \
        if (!(c)) {                                           \
            std::printf("FAILED %s:%d %s\n", __FILE__, __LINE__, #c); \
            return 1;                                         \
        }                                                     \
    } while (0)

int main()
{
    Graph g;
    VertexSE3 *a = g.addVertex(std::make_unique<VertexSE3>());
    VertexSE3 *b = g.addVertex(std::make_unique<VertexSE3>());
    Isometry3d Ta, Tb;
    Ta.t[0] = 1; Ta.t[1] = 2; Ta.t[2] = 3;
    /* 90 degrees about z */
    Ta.R[0] = 0; Ta.R[1] = -1; Ta.R[3] = 1; Ta.R[4] = 0;
    Tb.t[0] = 4; Tb.t[1] = 6; Tb.t[2] = 3;
    a->setEstimate(Ta);
    b->setEstimate(Tb);

    EdgeSE3Range e;
    e.vertices()[0] = a;
    e.vertices()[1] = b;
    e.setMeasurement(5.5);
    e.setInformation(4.0);
    e.computeError();
    CHECK(std::fabs(e.error() - 0.5) < 1e-15); /* |(3,4,0)| = 5 */
    CHECK(std::fabs(e.chi2() - 1.0) < 1e-15);
    Isometry3d off;
    off.t[0] = 0.2; /* lever arm along body x = world y after the rotation */
    e.setVertexOffset(0, off);
    e.computeError();
    CHECK(std::fabs(e.error() - (5.5 - std::sqrt(9.0 + 3.8 * 3.8))) < 1e-15);
    CHECK(e.initialEstimatePossible(RangeEdgeBase::VertexSet(), nullptr) == 1.0);

    /* text payload round trip: "meas info" */
    std::stringstream ss;
    CHECK(e.write(ss));
    CHECK(ss.str() == "5.5 4");
    EdgeSE3Range r;
    CHECK(r.read(ss));
    CHECK(r.measurement() == 5.5 && r.information() == 4.0);

    /* EDGE_RANGE_OFFSET: "pidFrom pidTo meas info", offsets by parameter id */
    std::map<int, ParameterSE3Offset> params;
    params[0].setId(0);
    params[1].setId(1);
    params[1].setOffset(off);
    EdgeSE3RangeOffset eo(&params);
    eo.vertices()[0] = a;
    eo.vertices()[1] = b;
    std::stringstream s2("1 0 5.5 4");
    CHECK(eo.read(s2));
    CHECK(eo.parameterId(0) == 1 && eo.parameterId(1) == 0);
    eo.computeError();
    CHECK(std::fabs(eo.error() - e.error()) < 1e-15);
    CHECK(!eo.setParameterId(0, 7)); /* unknown parameter id */
    std::stringstream s3;
    CHECK(eo.write(s3) && s3.str() == "1 0 5.5 4");
    /* asEdge resolves BOTH offsets against the antenna table, and refuses what the solve path cannot carry */
    {
        Isometry3d other;
        other.t[0] = 0.5;
        std::vector<Isometry3d> table{other, off};
        Edge pe = eo.asEdge(table), re = e.asEdge(table);
        CHECK(pe.antenna == 2 && pe.antenna_b == 0 && re.antenna == 2 && re.antenna_b == 0);
        e.setVertexOffset(1, other);
        re = e.asEdge(table);
        CHECK(re.antenna == 2 && re.antenna_b == 1);
        CHECK(eo.setParameterId(1, 1));
        pe = eo.asEdge(table);
        CHECK(pe.antenna == 2 && pe.antenna_b == 2);
        bool threw = false;
        try { e.asEdge(std::vector<Isometry3d>{off}); } catch (const std::invalid_argument &) { threw = true; }
        CHECK(threw); /* offset[1] is not in the table */
        Isometry3d rot = off;
        rot.R[0] = 0; rot.R[1] = -1; rot.R[3] = 1; rot.R[4] = 0;
        e.setVertexOffset(0, rot);
        threw = false;
        try { e.asEdge(table); } catch (const std::invalid_argument &) { threw = true; }
        CHECK(threw); /* a rotating offset */
        e.setVertexOffset(0, off);
        Isometry3d id = Isometry3d::Identity();
        e.setVertexOffset(1, id);
        CHECK(eo.setParameterId(1, 0));
    }

    /* initialEstimate: slide vertex 1 along the line of sight to the measured distance */
    EdgeSE3Range ie;
    ie.vertices()[0] = a;
    ie.vertices()[1] = b;
    ie.setMeasurement(10.0);
    RangeEdgeBase::VertexSet from{a};
    ie.initialEstimate(from, b);
    ie.computeError();
    CHECK(std::fabs(ie.error()) < 1e-12);
    CHECK(std::fabs(b->estimate().t[0] - (1 + 6)) < 1e-12 && std::fabs(b->estimate().t[1] - (2 + 8)) < 1e-12);

    /* Robot ring: ids slot*300 + ID, newest copies its predecessor, eviction drops edges */
    Graph g2;
    Robot rob(200, false, 3);
    Isometry3d start;
    start.t[2] = 0.87;
    rob.init(g2, start);
    CHECK(rob.last_header().frame_id == "none");
    Header h;
    h.frame_id = "uwb";
    VertexSE3 *v1 = rob.new_vertex(2, h, g2);
    CHECK(v1->id() == 1 * 300 + 200 && v1->estimate().t[2] == 0.87);
    Isometry3d moved = v1->estimate();
    moved.t[0] = 1.5;
    v1->setEstimate(moved);
    VertexSE3 *v2 = rob.new_vertex(2, h, g2);
    CHECK(v2->id() == 2 * 300 + 200 && v2->estimate().t[0] == 1.5);
    Edge tr;
    tr.from = v1;
    tr.to = v2;
    g2.addEdge(tr);
    CHECK(g2.edges().size() == 1);
    rob.new_vertex(2, h, g2);          /* slot 0 evicted (never had edges) */
    CHECK(g2.edges().size() == 1);
    rob.new_vertex(2, h, g2);          /* slot 1 = v1 evicted, its edge goes with it */
    CHECK(g2.edges().empty());
    CHECK(rob.vertices2path()->poses.size() == 3);
    Robot anchor(100, true, 1);
    anchor.init(g2, start);
    CHECK(anchor.is_static() && anchor.new_vertex(2, h, g2) == anchor.last_vertex() && anchor.last_vertex()->fixed());
    std::printf("host type API ok\n");
    return 0;
}
