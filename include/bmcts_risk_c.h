/*
 * bmcts_risk_c.h -- C ABI of the scenario-risk tooling: the offline
 * computation that turns prior knowledge about the behaviour space into the risk bound of
 * BehaviorUCTRiskConstraint.  Restates bark_mcts/models/behavior/risk_calculation/ of the
 * reference (paths below are relative to that directory): regions are axis-aligned boxes over
 * named dimensions (here: dimension index order), knowledge functions are given through their
 * region integral.  Host-only; no GPU involved.
 */
#ifndef BMCTS_RISK_C_H_
#define BMCTS_RISK_C_H_

#ifdef __cplusplus
extern "C" {
#endif

#define BMCTS_RISK_MAX_DIMS 8

/* KnowledgeFunctionDefinition::CalculateIntegral(region)
 * (knowledge_function_definition/knowledge_function_definition.hpp:28) */
typedef double (*bmcts_region_integral_fn)(const double* lo, const double* hi, int n_dims,
                                           void* user);

/* CalculateRegionBoundariesArea (prior_knowledge_region.hpp:17-25): product of the positive
 * extents */
double bmcts_risk_region_area(const double* lo, const double* hi, int n_dims);

/* PriorKnowledgeRegion::Partition (prior_knowledge_region.cpp:10-53): num_partitions equal
 * cells per dimension, cartesian product; cell c of the output has its bounds at
 * out_lo/out_hi[c * n_dims + d], the last dimension varying fastest.  Returns the number of
 * cells (num_partitions ^ n_dims) or -1 if it exceeds max_cells / an argument is invalid. */
long bmcts_risk_partition(const double* lo, const double* hi, int n_dims, int num_partitions,
                          double* out_lo, double* out_hi, long max_cells);

/* LinearKnowledgeFunctionDefinition::CalculateIntegral
 * (knowledge_function_definition/linear_knowledge_function_definition.hpp:32-40): mean over
 * the dimensions of the mid-point value of a*x + b, times the region area */
double bmcts_risk_linear_integral(const double* a, const double* b, const double* lo,
                                  const double* hi, int n_dims);

/* PriorKnowledgeFunction::CalculateScenarioRiskFunction (prior_knowledge_function.cpp:14-30):
 * the constant c such that c * sum over the partition cells of
 * prior(cell) * template(cell) / area(cell) equals 1.  Both integrals must be positive on
 * every cell (the reference aborts otherwise): returns 1 in that case, 0 on success. */
int bmcts_risk_normalization_constant(const double* lo, const double* hi, int n_dims,
                                      int num_partitions, bmcts_region_integral_fn prior,
                                      void* prior_user, bmcts_region_integral_fn templ,
                                      void* templ_user, double* out_constant);

#ifdef __cplusplus
}
#endif
#endif /* BMCTS_RISK_C_H_ */
