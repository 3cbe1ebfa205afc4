// risk_calculation.cpp -- scenario-risk tooling behind include/bmcts_risk_c.h
// (reference: bark_mcts/models/behavior/risk_calculation/, see the header for line cites).
#include <cmath>
#include <vector>

#include "bmcts_risk_c.h"

extern "C" {

double bmcts_risk_region_area(const double* lo, const double* hi, int n_dims) {
  double area = 1.0;
  for (int d = 0; d < n_dims; ++d)
    if (hi[d] > lo[d]) area *= hi[d] - lo[d];
  return area;
}

long bmcts_risk_partition(const double* lo, const double* hi, int n_dims, int num_partitions,
                          double* out_lo, double* out_hi, long max_cells) {
  if (!lo || !hi || n_dims < 1 || n_dims > BMCTS_RISK_MAX_DIMS || num_partitions < 1) return -1;
  long cells = 1;
  for (int d = 0; d < n_dims; ++d) {
    cells *= num_partitions;
    if (cells > max_cells) return -1;
  }
  if (!out_lo || !out_hi) return cells;
  for (long c = 0; c < cells; ++c) {
    long rem = c;
    for (int d = n_dims - 1; d >= 0; --d) {
      const long i = rem % num_partitions;
      rem /= num_partitions;
      const double width = (hi[d] - lo[d]) / num_partitions;  // part_width (:15)
      out_lo[c * n_dims + d] = lo[d] + width * (double)i;
      out_hi[c * n_dims + d] = lo[d] + width * (double)(i + 1);
    }
  }
  return cells;
}

double bmcts_risk_linear_integral(const double* a, const double* b, const double* lo,
                                  const double* hi, int n_dims) {
  double mean = 0.0;
  for (int d = 0; d < n_dims; ++d) mean += ((a[d] * lo[d] + b[d]) + (a[d] * hi[d] + b[d])) / 2.0;
  return mean / (double)n_dims * bmcts_risk_region_area(lo, hi, n_dims);
}

int bmcts_risk_normalization_constant(const double* lo, const double* hi, int n_dims,
                                      int num_partitions, bmcts_region_integral_fn prior,
                                      void* prior_user, bmcts_region_integral_fn templ,
                                      void* templ_user, double* out_constant) {
  if (!prior || !templ || !out_constant) return 1;
  const long max_cells = 1L << 26;
  const long cells = bmcts_risk_partition(lo, hi, n_dims, num_partitions, nullptr, nullptr, max_cells);
  if (cells < 0) return 1;
  double sum = 0.0;
  std::vector<double> clo(n_dims), chi(n_dims);
  for (long c = 0; c < cells; ++c) {
    long rem = c;
    for (int d = n_dims - 1; d >= 0; --d) {
      const long i = rem % num_partitions;
      rem /= num_partitions;
      const double width = (hi[d] - lo[d]) / num_partitions;
      clo[d] = lo[d] + width * (double)i;
      chi[d] = lo[d] + width * (double)(i + 1);
    }
    const double p = prior(clo.data(), chi.data(), n_dims, prior_user);
    const double t = templ(clo.data(), chi.data(), n_dims, templ_user);
    if (!(p > 0.0) || !(t > 0.0)) return 1;  // BARK_EXPECT_TRUE (:23-24)
    // correct for the double ds of the product (:26)
    sum += (p * t) / bmcts_risk_region_area(clo.data(), chi.data(), n_dims);
  }
  *out_constant = 1.0 / sum;
  return 0;
}

}  // extern "C"
