/*
 * bmcts_oracle.h -- CPU oracle of the MCTS leaf-evaluation path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under planner-mcts_b200/ may include,
 * link or execute this; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker or the
 * timed CPU baseline.
 *
 * PARITY STATUS: "parity unpinned" at the BARK / mamcts boundary.  The
 * arithmetic of the reference's path lives in two un-vendored dependencies that
 * are absent from /root/reference and cannot be built here:
 *   BARK   github.com/bark-simulator/bark  @ 0c8f1b52bffcc639ac04ab944d41653f356eb1c8
 *   mamcts github.com/juloberno/mamcts     @ b3e65e7dcd508e43fd77eb75d89b89b3ea608fff
 * (util/deps.bzl:19-24,53-58).  The parts of the path that ARE in the
 * reference (flags -> reward / cost / terminal, cooperative mixing, prediction
 * time span, likelihood histogram) are restated line by line and pinned by the
 * reference's own known-answer tests (tests/test_oracle_kat.py); the BARK
 * internals (IDM, single-track, geometry, safe distance) restate the published
 * algorithms as fixed in DESIGN.md "Normative step spec".  Those are checked
 * against sources outside this file (tests/test_independent_spec.py): a float64
 * numpy restatement with different algorithms for every predicate, a brute-force
 * simulation of the safe-distance rule, Treiber's IDM equilibrium, Random123's
 * Philox vectors -- but not against BARK's own numbers, which do not exist here.
 */
#ifndef BMCTS_ORACLE_H_
#define BMCTS_ORACLE_H_

#include "bmcts_c.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Philox4x32-10 (Salmon et al., SC'11); ctr[4], key[2] -> out[4] */
void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* MctsStateBase::calculate_prediction_time_span (mcts_state_base.hpp:238-241) */
float oracle_prediction_time_span(const bmcts_config* cfg, unsigned depth);
/* Euler steps of one macro action of duration Dt (BehaviorMotionPrimitives::IntegrationTimeDelta) */
int oracle_ego_substeps(const bmcts_config* cfg, float Dt);

/* reward_from_evaluation_results (mcts_state_base.hpp:107-118) */
float oracle_reward_from_flags(const bmcts_config* cfg, uint32_t flags, float dt);
/* ego_costs_from_evaluation_results (mcts_state_base.hpp:120-142); returns #costs */
int oracle_costs_from_flags(const bmcts_config* cfg, uint32_t flags, float dt, float cost[2]);
/* is_terminal rule of mcts_observed_world_evaluation (mcts_state_base.hpp:273-277) */
uint32_t oracle_terminal_from_flags(const bmcts_config* cfg, uint32_t flags);

/* Frenet projection of (x, y) on corridor c: s, signed lateral d (left > 0),
 * segment index, in_range (0 when beyond either end of the centre line) */
/* out[i] = f(a[i], b[i]) for the elementary function `op` (BMCTS_MATH_*) */
void oracle_math_probe(int op, const float* a, const float* b, int n, float* out);
void oracle_project(const bmcts_corridor* c, float x, float y, float* s, float* d, int* seg,
                    int* in_range);

/* One env step = MctsState*::execute (mcts_state_hypothesis.cpp:62-72,
 * mcts_state_cooperative.cpp:45-116) with explicit joint action.
 * pose[A*4], hyp_id[A], action[A], draw_id[A] as in bmcts_step_in. */
int oracle_step(const bmcts_config* cfg, const bmcts_scenario* scn, const float* pose,
                uint32_t alive, const uint8_t* hyp_id, unsigned depth, const uint8_t* action,
                const uint64_t* draw_id, uint64_t seed, uint32_t stream, float* next_pose,
                uint32_t* next_alive, float* reward, float* cost, uint32_t* flags,
                float* last_action);

/* One default-policy rollout = mamcts RandomHeuristic::calculate_heuristic_values
 * on one leaf. ret_agents / final_pose may be NULL. */
int oracle_rollout(const bmcts_config* cfg, const bmcts_scenario* scn, const float* pose,
                   uint32_t alive, const uint8_t* hyp_id, unsigned depth, uint64_t seed,
                   uint32_t stream, uint64_t rollout_id, bmcts_rollout_result* res,
                   float* ret_agents, float* final_pose);

/* Batched forms with the same argument structs as the C ABI (host pointers
 * only); n_threads > 1 splits the batch over pthreads (CPU baseline). */
int oracle_rollout_batch(const bmcts_config* cfg, const bmcts_scenario* scn,
                         const bmcts_leaf_batch* in, const bmcts_rollout_out* out, int n_threads);
int oracle_step_batch(const bmcts_config* cfg, const bmcts_scenario* scn,
                      const bmcts_step_in* in, const bmcts_step_out* out);
int oracle_expand_rollout_batch(const bmcts_config* cfg, const bmcts_scenario* scn,
                                const bmcts_step_in* in, const bmcts_step_out* step_out,
                                uint64_t first_rollout_id, const bmcts_rollout_out* rollout_out,
                                int n_threads);
/* BehaviorHypothesisIDM::GetProbability (hypothesis/idm/hypothesis_idm.cpp:40-126) */
int oracle_likelihood_batch(const bmcts_config* cfg, const bmcts_scenario* scn,
                            const bmcts_likelihood_in* in, float* prob);

#ifdef __cplusplus
}
#endif
#endif
