// inner_rects.hpp -- axis-aligned rectangles that lie inside the road-corridor polygon.
//
// The kernels test "is this point inside the drivable area" (World::RemoveInvalidAgents,
// EvaluatorSafeDistDrivableArea) with an even-odd point-in-polygon scan.  A point inside one of
// these rectangles -- which keep a margin to every polygon edge that is far above FP32 rounding
// -- is inside the polygon, so the scan is skipped for it without changing any result.  Found
// once per road polygon on the host (double precision); road polygons whose vertices span too
// many distinct coordinates (curved roads) simply get no rectangle and always take the scan.
#pragma once

#include "bmcts_c.h"

namespace bmcts {

struct InnerRects {
  float rect[2][4];  // x0, x1, y0, y1
  int n;
  int pad_[3];
  // per lane-corridor segment: the interval [lo, hi] of local arc length on which the centre
  // line point keeps the margin to the road boundary (lo > hi: none).  An IDM agent, which
  // sits exactly on its centre line, needs no scan inside it.
  float arc_lo[BMCTS_MAX_CORRIDORS][BMCTS_MAX_CORRIDOR_PTS];
  float arc_hi[BMCTS_MAX_CORRIDORS][BMCTS_MAX_CORRIDOR_PTS];
};

InnerRects find_inner_rects(const bmcts_scenario& scn);

}  // namespace bmcts
