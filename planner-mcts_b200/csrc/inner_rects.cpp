// inner_rects.cpp -- see inner_rects.hpp
#include "inner_rects.hpp"

#include <algorithm>
#include <cmath>
#include <vector>

namespace bmcts {

namespace {

struct Poly {
  std::vector<double> x, y;
};

bool inside(const Poly& p, double px, double py) {
  bool in = false;
  const size_t n = p.x.size();
  for (size_t i = 0, j = n - 1; i < n; j = i++)
    if ((p.y[i] > py) != (p.y[j] > py)) {
      const double xc = p.x[i] + (py - p.y[i]) * (p.x[j] - p.x[i]) / (p.y[j] - p.y[i]);
      if (px < xc) in = !in;
    }
  return in;
}

// does segment (ax,ay)-(bx,by) touch the closed rectangle?  (Liang-Barsky clipping)
bool segment_hits_rect(double ax, double ay, double bx, double by, double x0, double x1, double y0,
                       double y1) {
  double t0 = 0.0, t1 = 1.0;
  const double dx = bx - ax, dy = by - ay;
  const double p[4] = {-dx, dx, -dy, dy};
  const double q[4] = {ax - x0, x1 - ax, ay - y0, y1 - ay};
  for (int i = 0; i < 4; ++i) {
    if (p[i] == 0.0) {
      if (q[i] < 0.0) return false;
    } else {
      const double r = q[i] / p[i];
      if (p[i] < 0.0) {
        if (r > t1) return false;
        if (r > t0) t0 = r;
      } else {
        if (r < t0) return false;
        if (r < t1) t1 = r;
      }
    }
  }
  return true;
}

struct Rect {
  double x0, x1, y0, y1;
  double area() const { return (x1 - x0) * (y1 - y0); }
};

double overlap(const Rect& a, const Rect& b) {
  const double w = std::min(a.x1, b.x1) - std::max(a.x0, b.x0);
  const double h = std::min(a.y1, b.y1) - std::max(a.y0, b.y0);
  return (w > 0.0 && h > 0.0) ? w * h : 0.0;
}

// distance from a point to the polygon boundary
double boundary_distance(const Poly& p, double px, double py) {
  double best = 1.0e300;
  const size_t n = p.x.size();
  for (size_t i = 0, j = n - 1; i < n; j = i++) {
    const double ex = p.x[i] - p.x[j], ey = p.y[i] - p.y[j];
    const double l2 = ex * ex + ey * ey;
    double u = l2 > 0.0 ? ((px - p.x[j]) * ex + (py - p.y[j]) * ey) / l2 : 0.0;
    u = std::min(1.0, std::max(0.0, u));
    const double dx = px - (p.x[j] + u * ex), dy = py - (p.y[j] + u * ey);
    best = std::min(best, std::sqrt(dx * dx + dy * dy));
  }
  return best;
}

// safe arc intervals of the lane centre lines (see InnerRects::arc_lo)
void find_safe_arcs(const bmcts_scenario& scn, const Poly& p, double margin, InnerRects* out) {
  for (int c = 0; c < BMCTS_MAX_CORRIDORS; ++c)
    for (int k = 0; k < BMCTS_MAX_CORRIDOR_PTS; ++k) {
      out->arc_lo[c][k] = 1.0e30f;
      out->arc_hi[c][k] = -1.0e30f;
    }
  for (int c = 0; c < scn.n_corridors; ++c) {
    const bmcts_corridor& cor = scn.corridors[c];
    for (int k = 0; k + 1 < cor.n_pts; ++k) {
      const double len = cor.seg_len[k];
      const int steps = (int)std::min(4000.0, std::max(8.0, std::ceil(len / 0.05)));
      const double h = len / steps;
      // the distance to the boundary is 1-Lipschitz along the segment: samples that keep
      // margin + h guarantee the margin in between
      int best_a = 0, best_b = -1, run_a = 0;
      bool in_run = false;
      for (int i = 0; i <= steps + 1; ++i) {
        bool ok = false;
        if (i <= steps) {
          const double u = h * i;
          const double x = cor.px[k] + u * cor.tx[k], y = cor.py[k] + u * cor.ty[k];
          ok = inside(p, x, y) && boundary_distance(p, x, y) >= margin + h;
        }
        if (ok && !in_run) {
          in_run = true;
          run_a = i;
        } else if (!ok && in_run) {
          in_run = false;
          if (i - 1 - run_a > best_b - best_a) {
            best_a = run_a;
            best_b = i - 1;
          }
        }
      }
      if (best_b >= best_a) {
        out->arc_lo[c][k] = std::nextafter((float)(h * best_a), 3.0e38f);
        out->arc_hi[c][k] = std::nextafter((float)(h * best_b), -3.0e38f);
      }
    }
  }
}

}  // namespace

InnerRects find_inner_rects(const bmcts_scenario& scn) {
  InnerRects out{};
  for (int c = 0; c < BMCTS_MAX_CORRIDORS; ++c)
    for (int k = 0; k < BMCTS_MAX_CORRIDOR_PTS; ++k) {
      out.arc_lo[c][k] = 1.0e30f;
      out.arc_hi[c][k] = -1.0e30f;
    }
  const int n = scn.n_road_pts;
  if (n < 3) return out;
  Poly p;
  double maxabs = 0.0;
  for (int i = 0; i < n; ++i) {
    p.x.push_back(scn.road_x[i]);
    p.y.push_back(scn.road_y[i]);
    maxabs = std::max({maxabs, std::fabs(p.x[i]), std::fabs(p.y[i])});
  }
  find_safe_arcs(scn, p, 0.02 + 1.0e-5 * maxabs, &out);
  std::vector<double> xs = p.x, ys = p.y;
  std::sort(xs.begin(), xs.end());
  xs.erase(std::unique(xs.begin(), xs.end()), xs.end());
  std::sort(ys.begin(), ys.end());
  ys.erase(std::unique(ys.begin(), ys.end()), ys.end());
  if (xs.size() * ys.size() > 256) return out;  // curved road: no rectangles
  const double margin = 0.02 + 1.0e-5 * maxabs;  // FP32 noise of the scan is ~1e-7 * maxabs
  std::vector<Rect> good;
  for (size_t i = 0; i < xs.size(); ++i)
    for (size_t j = i + 1; j < xs.size(); ++j)
      for (size_t k = 0; k < ys.size(); ++k)
        for (size_t l = k + 1; l < ys.size(); ++l) {
          // the polygon must contain the rectangle grown back by the margin
          const Rect g{xs[i] + 0.5 * margin, xs[j] - 0.5 * margin, ys[k] + 0.5 * margin,
                       ys[l] - 0.5 * margin};
          const Rect r{xs[i] + margin, xs[j] - margin, ys[k] + margin, ys[l] - margin};
          if (!(r.x1 > r.x0 && r.y1 > r.y0)) continue;
          if (!inside(p, 0.5 * (g.x0 + g.x1), 0.5 * (g.y0 + g.y1))) continue;
          bool clear = true;
          for (int e = 0, f = n - 1; e < n && clear; f = e++)
            clear = !segment_hits_rect(p.x[f], p.y[f], p.x[e], p.y[e], g.x0, g.x1, g.y0, g.y1);
          if (clear) good.push_back(r);
        }
  if (good.empty()) return out;
  size_t b0 = 0;
  for (size_t i = 1; i < good.size(); ++i)
    if (good[i].area() > good[b0].area()) b0 = i;
  auto put = [&](const Rect& r) {
    float* d = out.rect[out.n++];
    // round inwards so that the float rectangle is inside the double one
    d[0] = std::nextafter((float)r.x0, 3.0e38f);
    d[1] = std::nextafter((float)r.x1, -3.0e38f);
    d[2] = std::nextafter((float)r.y0, 3.0e38f);
    d[3] = std::nextafter((float)r.y1, -3.0e38f);
  };
  put(good[b0]);
  size_t b1 = good.size();
  double best_gain = 0.0;
  for (size_t i = 0; i < good.size(); ++i) {
    const double gain = good[i].area() - overlap(good[i], good[b0]);
    if (gain > best_gain) {
      best_gain = gain;
      b1 = i;
    }
  }
  if (b1 < good.size() && best_gain > 0.05 * good[b0].area()) put(good[b1]);
  return out;
}

}  // namespace bmcts
