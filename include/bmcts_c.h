/*
 * bmcts_c.h -- C ABI of the B200 rollout engine (libbmcts.so).
 *
 * This is the drop-in boundary for the MCTS leaf-evaluation path of
 * bark-simulator/planner-mcts.  Every entry point names
 * the reference interface it replaces (paths relative to
 * /root/reference/bark_mcts/models/behavior/).  Plain pointers and sizes only;
 * no C++ / torch types cross this boundary; errors are status codes
 * (the reference aborts with glog LOG(FATAL) -- mcts_state_hypothesis.cpp:65,130).
 *
 * Data model (POD restatement of what the reference reads from BARK's
 * ObservedWorld, see DESIGN.md "Normative step spec"):
 *   - agent slot 0 is the ego agent (mamcts ego_agent_idx = 0), slots 1..A-1 are
 *     the other agents in ascending BARK AgentId order
 *     (mcts_state_base.hpp:218-226);
 *   - a pose record is 4 floats {x, y, theta, v} (BARK state [t,x,y,theta,v]
 *     without the time column, tests/test_helpers.cpp:71-72);
 *   - lane corridors are centre-line polylines with a half width; the road
 *     corridor (drivable area) and the goal are polygons.
 */
#ifndef BMCTS_C_H_
#define BMCTS_C_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BMCTS_ABI_VERSION 2

#define BMCTS_MAX_AGENTS 32
#define BMCTS_MAX_CORRIDORS 4
#define BMCTS_MAX_CORRIDOR_PTS 16
#define BMCTS_MAX_ROAD_PTS 32
#define BMCTS_MAX_GOAL_PTS 8
#define BMCTS_MAX_EXTRA_GOALS 3
#define BMCTS_MAX_HYPOTHESES 64
#define BMCTS_MAX_EGO_ACTIONS 16
#define BMCTS_NUM_IDM_PARAMS 6
#define BMCTS_MAX_DEPTH 2048

/*
 * Counter-based RNG contract (Philox4x32-10).  Every random number of the path
 * is a pure function of (seed, stream, id, agent, domain, index), so a rollout
 * is reproducible whatever the launch geometry or the order of evaluation:
 *   key     = { (uint32)seed, (uint32)(seed >> 32) ^ stream }
 *   counter = { (uint32)id, (uint32)(id >> 32), agent | domain << 16, index }
 *   ROLLOUT_PARAMS : id = rollout id, index = 2*step + block; block 0 -> u0..u3
 *                    (headway, spacing, max-acc, desired-vel), block 1 -> u4, u5
 *                    (comfortable braking, coolness)
 *   ROLLOUT_ACTION : id = rollout id, index = step; word 0 -> macro-action index
 *   TREE_PARAMS    : id = draw_id of bmcts_step_in, index = block
 *   LIKELIHOOD     : id = sample | world << 32, index = 2*hypothesis + block
 * This replaces one std::mt19937 per BARK distribution object
 * (hypothesis/idm/hypothesis_idm.hpp:43 ChangeSeed).
 */
#define BMCTS_RNG_ROLLOUT_PARAMS 0u
#define BMCTS_RNG_ROLLOUT_ACTION 1u
#define BMCTS_RNG_TREE_PARAMS 2u
#define BMCTS_RNG_LIKELIHOOD 3u

/* status codes */
enum {
  BMCTS_OK = 0,
  BMCTS_ERR_INVALID_ARG = 1,
  BMCTS_ERR_CUDA = 2,
  BMCTS_ERR_NO_SCENARIO = 3,
  BMCTS_ERR_UNSUPPORTED = 4,
  BMCTS_ERR_NO_DEVICE = 5,
  BMCTS_ERR_COMM = 6
};

/* EvaluationResults bits, member order of mcts_state_base.hpp:98-104 */
enum {
  BMCTS_F_COLLISION_OTHER = 1u << 0,
  BMCTS_F_STATIC_SAFE_DIST = 1u << 1,
  BMCTS_F_DYNAMIC_SAFE_DIST = 1u << 2,
  BMCTS_F_COLLISION_DRIVABLE = 1u << 3,
  BMCTS_F_GOAL_REACHED = 1u << 4,
  BMCTS_F_OUT_OF_MAP = 1u << 5,
  BMCTS_F_TERMINAL = 1u << 6
};

/* which MctsState the engine restates */
enum {
  BMCTS_STATE_HYPOTHESIS = 0,      /* MctsStateHypothesis<void>  (mcts_state_hypothesis.hpp) */
  BMCTS_STATE_RISK_CONSTRAINT = 1, /* MctsStateRiskConstraint    (mcts_state_risk_constraint.cpp:67-81) */
  BMCTS_STATE_COOPERATIVE = 2      /* MctsStateCooperative       (mcts_state_cooperative.cpp:56-116) */
};

/* ego macro actions (BARK BehaviorMPMacroActions primitives) */
enum {
  BMCTS_ACT_CONST_ACC_STAY_LANE = 0,
  BMCTS_ACT_CHANGE_LEFT = 1,
  BMCTS_ACT_CHANGE_RIGHT = 2
};

/* bmcts_step_in.action[slot 0] of an element that is NOT to be expanded: the rollout starts from
 * the given state itself (a non-terminal node at Mcts::MaxSearchDepth: mamcts iterate() runs the
 * heuristic on it).  Its step outputs are left untouched.  Only with a rollout. */
#define BMCTS_ACTION_NONE 0xFFu

enum { BMCTS_GOAL_NONE = 0, BMCTS_GOAL_POLYGON = 1, BMCTS_GOAL_FRENET_LIMITS = 2 };

enum { BMCTS_MEM_HOST = 0, BMCTS_MEM_DEVICE = 1 };

/* IDM parameter order of BehaviorIDMStochastic::SampleParameters
 * (tests/behavior_uct_hypothesis_test.cc:78-92) */
enum {
  BMCTS_IDM_HEADWAY = 0,
  BMCTS_IDM_SPACING = 1,
  BMCTS_IDM_MAX_ACC = 2,
  BMCTS_IDM_DESIRED_VEL = 3,
  BMCTS_IDM_COMFT_BRAKING = 4,
  BMCTS_IDM_COOLNESS = 5
};

/*
 * Flat POD of the parameter keys the path honours
 * (param_config/mcts_state_parameters_from_param_server.hpp:19-51,
 *  param_config/mcts_parameters_from_param_server.hpp:45,55-58,
 *  tests/data/nheuristic_test.json:152-199,281-311).
 */
typedef struct bmcts_config {
  /* Mcts::State::* */
  float goal_reward;               /* GoalReward 100 */
  float collision_reward;          /* CollisionReward -1000 */
  float safe_dist_violated_reward; /* SafeDistViolatedReward -0.1 */
  float out_of_drivable_reward;    /* DrivableCollisionReward 0 */
  float goal_cost;                 /* GoalCost -100 */
  float collision_cost;            /* CollisionCost 1000 */
  float safe_dist_violated_cost;   /* SafeDistViolatedCost 0.1 */
  float out_of_drivable_cost;      /* DrivableCollisionCost 0 */
  float cooperation_factor;        /* CooperationFactor 0.2 */
  float step_reward;               /* StepReward 0 */
  float prediction_k;              /* PredictionK 0.5 */
  float prediction_alpha;          /* PredictionAlpha 0 */
  float normalization_tau;         /* NormalizationTau 0.2 (read, unused by the reference) */
  int32_t split_safe_dist_collision; /* SplitSafeDistCollision false */
  int32_t chance_costs;              /* ChanceCosts false */
  /* Mcts::State::EvaluationParameters::* */
  int32_t add_safe_dist;
  int32_t static_safe_dist_is_terminal;
  int32_t dynamic_safe_dist_is_terminal;
  int32_t out_of_drivable_is_terminal;
  /* Mcts::State::EvaluatorParams::* */
  float drivable_lat_safe_dist;    /* EvaluatorSafeDistDrivableArea::LateralSafeDist 0.5 */
  float drivable_lon_safe_dist;    /* ...::LongitudinalSafeDist 0.5 */
  float static_lat_safe_dist;      /* EvaluatorStaticSafeDist::LateralSafeDist 1.5 */
  float static_lon_safe_dist;      /* ...::LongitudinalSafeDist 1.5 */
  float dyn_max_other_decel;       /* EvaluatorDynamicSafeDist::MaxOtherDecceleration 5 */
  float dyn_max_ego_decel;         /* ...::MaxEgoDecceleration 5 */
  float dyn_reaction_time_others;  /* ...::ReactionTimeOthers 0.8 */
  float dyn_reaction_time_ego;     /* ...::ReactionTimeEgo 0.2 */
  int32_t dyn_to_rear;             /* ...::ToRear true */
  /* Mcts::DiscountFactor, Mcts::RandomHeuristic::MaxNumIterations */
  float discount_factor;           /* 0.9 */
  int32_t max_rollout_steps;       /* 1000 in the reference; benchmarks also use 20 */
  /* EgoBehavior (BehaviorMPMacroActions on SingleTrackModel) */
  float wheel_base;                /* DynamicModel::wheel_base 2.7 */
  float delta_max;                 /* DynamicModel::delta_max 0.2 (must be <= pi/4) */
  float crosstrack_gain;           /* BehaviorIDMLaneTracking::CrosstrackErrorGain 1 */
  float lon_acc_min;               /* DynamicModel::LonAccelerationMin -8 */
  float lon_acc_max;               /* DynamicModel::LonAccelerationMax 4 */
  float ego_max_velocity;          /* BehaviorIDMClassic::MaxVelocity 50 */
  int32_t num_substeps;            /* NumTrajectoryTimePoints - 1 = 10 (IDM sub-steps of the others) */
  int32_t state_kind;              /* BMCTS_STATE_* */
  /* macro-action integration (tests/data/nheuristic_test.json:153-155,183-196) */
  int32_t limit_steering_rate;     /* BehaviorIDMLaneTracking::LimitSteeringRate true: the steering
                                      angle is kept inside the lateral-acceleration limits,
                                      v^2 tan(delta) / wheel_base in [LatAccMin, LatAccMax] */
  float lat_acc_max;               /* DynamicModel::LatAccMax 4 */
  float lat_acc_min;               /* DynamicModel::LatAccMin -4 */
  float ego_integration_dt;        /* BehaviorMotionPrimitives::IntegrationTimeDelta 0.02: a macro
                                      action is ceil(dt / this) Euler steps; <= 0: num_substeps */
} bmcts_config;

/* One lane corridor: centre line + half width.  The *derived* arrays are
 * filled by bmcts_scenario_finalize() and are inputs to kernel and oracle. */
typedef struct bmcts_corridor {
  int32_t n_pts; /* 2..BMCTS_MAX_CORRIDOR_PTS */
  int32_t left;  /* index of the corridor to the left, -1 if none */
  int32_t right; /* index of the corridor to the right, -1 if none */
  float half_width;
  float px[BMCTS_MAX_CORRIDOR_PTS];
  float py[BMCTS_MAX_CORRIDOR_PTS];
  /* derived, per segment k = 0..n_pts-2 */
  float tx[BMCTS_MAX_CORRIDOR_PTS];      /* unit tangent */
  float ty[BMCTS_MAX_CORRIDOR_PTS];
  float seg_len[BMCTS_MAX_CORRIDOR_PTS];
  float s0[BMCTS_MAX_CORRIDOR_PTS];      /* arc length at the segment start */
  float angle[BMCTS_MAX_CORRIDOR_PTS];   /* tangent angle */
  float length;                          /* total arc length */
  int32_t pad_[3];
} bmcts_corridor;

/* One behaviour hypothesis = parameter region of BehaviorIDMStochastic
 * (hypothesis/idm/hypothesis_idm.hpp:26-50); lo == hi is a FixedValue. */
typedef struct bmcts_hypothesis {
  float lo[BMCTS_NUM_IDM_PARAMS];
  float hi[BMCTS_NUM_IDM_PARAMS];
  float acc_lower_bound; /* BehaviorIDMClassic::AccelerationLowerBound -5 */
  float acc_upper_bound; /* ...::AccelerationUpperBound 8 */
  float min_velocity;    /* ...::MinVelocity 0 */
  float max_velocity;    /* ...::MaxVelocity 50 */
  int32_t exponent;      /* ...::Exponent 4 */
  int32_t pad_[3];
} bmcts_hypothesis;

typedef struct bmcts_ego_action {
  int32_t type;       /* BMCTS_ACT_* */
  float acceleration; /* AccelerationInputs entry; 0 for lane changes */
} bmcts_ego_action;

typedef struct bmcts_goal {
  int32_t kind;  /* BMCTS_GOAL_* */
  int32_t n_pts; /* polygon goal (GoalDefinitionPolygon) */
  float px[BMCTS_MAX_GOAL_PTS];
  float py[BMCTS_MAX_GOAL_PTS];
  /* GoalDefinitionStateLimitsFrenet (tests/test_helpers.cpp:67-69) */
  int32_t corridor;
  float max_lat_left, max_lat_right;
  float max_angle_left, max_angle_right;
  float vel_lo, vel_hi;
  int32_t pad_;
} bmcts_goal;

typedef struct bmcts_scenario {
  int32_t n_corridors;
  int32_t n_road_pts;
  int32_t n_agents;     /* A, incl. ego */
  int32_t n_hypotheses; /* H (0 allowed for the cooperative state) */
  int32_t n_ego_actions;
  int32_t finalized;    /* set by bmcts_scenario_finalize */
  int32_t pad_[2];
  bmcts_corridor corridors[BMCTS_MAX_CORRIDORS];
  float road_x[BMCTS_MAX_ROAD_PTS]; /* road-corridor polygon (drivable area), open ring */
  float road_y[BMCTS_MAX_ROAD_PTS];
  bmcts_goal goal;
  float agent_length[BMCTS_MAX_AGENTS]; /* rectangle shapes; BARK CarRectangle = 4 x 2 */
  float agent_width[BMCTS_MAX_AGENTS];
  float agent_rear[BMCTS_MAX_AGENTS];   /* rear edge -> reference point, 1.25 */
  /* Cooperative state: every agent is evaluated from its own perspective AGAINST ITS OWN GOAL
   * (mcts_state_cooperative.cpp:76-84: ObservedWorld(predicted_world, agent)).  agent_goal[a] = 0:
   * `goal` above; k > 0: extra_goals[k - 1].  Slot 0 (the ego) is always 0; the hypothesis /
   * risk states only evaluate the ego and ignore the table. */
  int32_t agent_goal[BMCTS_MAX_AGENTS];
  bmcts_goal extra_goals[BMCTS_MAX_EXTRA_GOALS];
  bmcts_hypothesis hypotheses[BMCTS_MAX_HYPOTHESES];
  bmcts_ego_action ego_actions[BMCTS_MAX_EGO_ACTIONS];
  int32_t pad2_[4]; /* sizeof % 16 == 0: the kernels stage this struct with 16-byte copies */
} bmcts_scenario;

/* per-rollout scalar results, one 32-byte record */
typedef struct bmcts_rollout_result {
  float ret_ego;        /* discounted ego return  (RandomHeuristic accum. reward) */
  float cost0;          /* discounted ego cost[0] (envelope risk when split)     */
  float cost1;          /* discounted ego cost[1] (collision hazard when split)  */
  float step_len;       /* sum of executed prediction time spans                 */
  uint32_t flags;       /* bits 0-7: EvaluationResults of the last step; bits 8-15: OR over all steps */
  int32_t n_steps;      /* executed env steps = terminal step index              */
  int32_t agent_steps;  /* sum over steps of agents alive at step start (metric) */
  uint32_t final_alive; /* alive mask after the last step                         */
} bmcts_rollout_result;

/*
 * Batch layout: every per-agent array is AGENT-MAJOR struct-of-arrays, i.e. one
 * contiguous array per agent slot indexed by the leaf / rollout number:
 *     pose[(a*n + k)*4 + f], hyp_id[a*n + k], action[a*n + k], draw_id[a*n + k],
 *     ret_agents[a*n + k], reward[a*n + k], last_action[a*n + k], cost[c*n + k]
 * so that lanes holding consecutive rollouts read the records of one agent slot from
 * consecutive addresses (16-byte float4 poses), and so that a chunk of rollouts [k0, k1) of a
 * host batch is one strided 2-D copy per array.
 */

/* A batch of leaf states for the default-policy rollout.
 * Replaces the input of mamcts RandomHeuristic::calculate_heuristic_values
 * (interface shape: mcts_heuristic/mcts_neural_heuristic.hpp:67-69): the leaf
 * MctsState (world poses + currently sampled hypothesis per agent). */
typedef struct bmcts_leaf_batch {
  int32_t n_leaves;
  int32_t memory;            /* BMCTS_MEM_HOST / BMCTS_MEM_DEVICE for ALL pointers below */
  const float* pose;         /* [A][n][4]  ([n_states][A][4], STATE-major, in the unique-state form) */
  const uint32_t* alive;     /* [n] bit a = agent slot a present  ([n_states]) */
  const uint8_t* hyp_id;     /* [A][n] sampled hypothesis per agent (slot 0 ignored) */
  const uint16_t* depth;     /* [n] depth_ of the leaf state (root = 1), NULL = all 1  ([n_states]) */
  uint64_t seed;             /* Mcts::RandomSeed */
  uint32_t stream;           /* root-parallel stream (rank / tree index) */
  uint32_t pad_;
  uint64_t first_rollout_id; /* rollout k uses Philox counter id first_rollout_id + k */
  /* Unique-state form (the waves of a search revisit few distinct states,
   * under freshly sampled hypotheses): pose / alive / depth hold n_states distinct states and
   * rollout k starts from state state_index[k].  The state table is STATE-major -- record s is the
   * A poses pose[(s*A + a)*4 ..], the way a search keeps the poses of a node -- so that a gather
   * in any order reads whole records.  n_states == 0: one state per leaf, agent-major arrays. */
  int32_t n_states;
  int32_t pad2_;
  const uint32_t* state_index; /* [n], each < n_states; NULL when n_states == 0 */
} bmcts_leaf_batch;

typedef struct bmcts_rollout_out {
  bmcts_rollout_result* result; /* [n] */
  float* ret_agents;            /* [A][n] per-agent returns (cooperative), may be NULL */
  float* final_pose;            /* [A][n][4] may be NULL */
} bmcts_rollout_out;

/* A batch of single env steps with explicit joint actions (tree expansion).
 * Replaces MctsStateHypothesis<T>::execute (mcts_state_hypothesis.cpp:62-72) /
 * MctsStateCooperative::execute (mcts_state_cooperative.cpp:45-54) together with
 * plan_action_current_hypothesis (mcts_state_hypothesis.cpp:104-114): another
 * agent's "action" is the IDM parameter draw identified by draw_id, so
 * replaying the same draw_id replays the stored trajectory
 * (action_store/behavior_action_store.cpp:16-51). */
typedef struct bmcts_step_in {
  int32_t n;
  int32_t memory;
  const float* pose;        /* [A][n][4] */
  const uint32_t* alive;    /* [n] */
  const uint8_t* hyp_id;    /* [A][n] */
  const uint16_t* depth;    /* [n] or NULL */
  const uint8_t* action;    /* [A][n] macro-action index (slot 0; all slots when cooperative) */
  const uint64_t* draw_id;  /* [A][n] Philox counter id of the parameter draw of each other agent */
  uint64_t seed;
  uint32_t stream;
  uint32_t pad_;
} bmcts_step_in;

typedef struct bmcts_step_out {
  float* next_pose;      /* [A][n][4] */
  uint32_t* next_alive;  /* [n] */
  float* reward;         /* [A][n] rewards[] of execute(); only slot 0 non-zero unless cooperative */
  float* cost;           /* [2][n] ego_cost[] */
  uint32_t* flags;       /* [n] EvaluationResults bits */
  float* last_action;    /* [A][n] longitudinal acceleration each agent applied (Continuous1DAction) */
} bmcts_step_out;

/* A batch of hypothesis-likelihood queries.
 * Replaces BehaviorHypothesisIDM::GetProbability (hypothesis/idm/hypothesis_idm.cpp:40-126):
 * for world w, agent a, hypothesis h: P(action | h) from a histogram of n_samples
 * sampled IDM accelerations. */
typedef struct bmcts_likelihood_in {
  int32_t n_worlds;
  int32_t memory;
  const float* pose;      /* [A][n_worlds][4] */
  const uint32_t* alive;  /* [n_worlds] */
  const float* action;    /* [A][n_worlds] observed longitudinal acceleration per agent */
  int32_t n_samples;      /* BehaviorHypothesisIDM::NumSamples 10000 */
  int32_t n_buckets;      /* ::NumBuckets 100 */
  float buckets_lower;    /* ::BucketsLowerBound -10 */
  float buckets_upper;    /* ::BucketsUpperBound 10 */
  uint64_t seed;
  uint32_t stream;
  uint32_t pad_;
} bmcts_likelihood_in;

typedef struct bmcts_engine bmcts_engine; /* opaque, one per GPU stream */

/* ---- lifecycle ------------------------------------------------------- */
int bmcts_abi_version(void);
/* reference defaults of every key (param_config/ headers) */
void bmcts_config_default(bmcts_config* cfg);
/* validates and fills the derived corridor fields; host-only, no GPU needed */
int bmcts_scenario_finalize(bmcts_scenario* scn);
int bmcts_config_validate(const bmcts_config* cfg);
/* sizeof of the ABI structs in declaration order (0 = bmcts_config ... 11 =
 * bmcts_likelihood_in), for bindings to verify their mirror of this header */
size_t bmcts_struct_size(int which);
/* pinned host memory for BMCTS_MEM_HOST batches (cudaHostAlloc / cudaFreeHost) */
int bmcts_alloc_host(size_t bytes, void** out);
void bmcts_free_host(void* p);
/* fails with BMCTS_ERR_NO_DEVICE when no CUDA device is present: there is no CPU fallback */
int bmcts_create(int device, const bmcts_config* cfg, bmcts_engine** out);
void bmcts_destroy(bmcts_engine* e);
int bmcts_set_config(bmcts_engine* e, const bmcts_config* cfg);
int bmcts_set_scenario(bmcts_engine* e, const bmcts_scenario* scn);

/* ---- the hot path ---------------------------------------------------- */
/* replaces mamcts RandomHeuristic::calculate_heuristic_values over a wave of leaves */
int bmcts_rollout_batch(bmcts_engine* e, const bmcts_leaf_batch* in, const bmcts_rollout_out* out);
/* replaces MctsState*::execute + plan_action_current_hypothesis over a wave of expansions */
int bmcts_step_batch(bmcts_engine* e, const bmcts_step_in* in, const bmcts_step_out* out);
/* fused expansion + rollout: step with the explicit joint action, then roll out from
 * the child unless it is terminal.  rollout results exclude the expansion step. */
int bmcts_expand_rollout_batch(bmcts_engine* e, const bmcts_step_in* in,
                               const bmcts_step_out* step_out, uint64_t first_rollout_id,
                               const bmcts_rollout_out* rollout_out);
/* The same call in two halves, for a caller that drives several engines from one thread (the
 * planners' root-parallel trees): _begin packs the host buffers, enqueues upload, kernel and
 * download on the engine's stream and returns; _end waits for them and fills the output arrays
 * given to _begin (which must stay valid and untouched in between).  Between the two only
 * bmcts_last_error / bmcts_launch_count may be called on this engine.  Host buffers only. */
int bmcts_expand_rollout_batch_begin(bmcts_engine* e, const bmcts_step_in* in,
                                     const bmcts_step_out* step_out, uint64_t first_rollout_id,
                                     const bmcts_rollout_out* rollout_out);
int bmcts_expand_rollout_batch_end(bmcts_engine* e);
/* replaces BehaviorHypothesisIDM::GetProbability for all (world, agent, hypothesis):
 * prob[(a*H + h)*n_worlds + w] */
int bmcts_likelihood_batch(bmcts_engine* e, const bmcts_likelihood_in* in, float* prob);

/* ---- multi-GPU: the one collective of the path -------------------------- */
/* Root-parallel search (Mcts::NumParallelMcts): every rank searches its own
 * trees under its own Philox stream and the root statistics are summed at the end of a Plan().
 * These calls replace mamcts merge_node_statistics across processes
 * (mcts_statistics/mcts_neural_cost_constrained_statistic.hpp:157-163,219-222) with one
 * ncclAllReduce(sum, float64) over NVLink; NCCL is loaded at run time (libnccl.so.2). */
#define BMCTS_COMM_ID_BYTES 128
typedef struct bmcts_comm bmcts_comm; /* opaque, one per rank */
/* rank 0 makes the id (ncclGetUniqueId) and hands it to the other ranks by its own means */
int bmcts_comm_unique_id(void* id /* [BMCTS_COMM_ID_BYTES] */);
/* collective over all ranks (ncclCommInitRank): one rank per GPU */
int bmcts_comm_create(int device, int nranks, int rank, const void* id, bmcts_comm** out);
void bmcts_comm_destroy(bmcts_comm* c);
int bmcts_comm_nranks(const bmcts_comm* c);
/* in-place sum over the ranks of n host doubles (SURVEY.md Appendix E bmcts_allreduce_root) */
int bmcts_allreduce_root(bmcts_comm* c, double* stats, size_t n);
const char* bmcts_comm_last_error(const bmcts_comm* c);

/* ---- self test ---------------------------------------------------------- */
/* Evaluates one elementary function of include/bmcts_math.h on the device for n host inputs:
 * out[i] = f(a[i], b[i]).  The flags of the path are threshold tests on FP32 geometry, so the
 * device and the CPU restatement must agree on these functions bit for bit, special values
 * included (tests/test_parity_gpu.py compares with the same header compiled by gcc). */
enum {
  BMCTS_MATH_SIN = 0, BMCTS_MATH_COS, BMCTS_MATH_ATAN2, BMCTS_MATH_TAN_SMALL, BMCTS_MATH_DIV,
  BMCTS_MATH_SQRT, BMCTS_MATH_MIN, BMCTS_MATH_MAX, BMCTS_MATH_NORM_ANGLE, BMCTS_MATH_FMA_AAB,
  BMCTS_MATH_POW4, BMCTS_MATH_NUM_OPS
};
int bmcts_math_probe(bmcts_engine* e, int op, const float* a, const float* b, int n, float* out);

/* all batch calls are stream-ordered and asynchronous for BMCTS_MEM_DEVICE; with
 * BMCTS_MEM_HOST they return after the results are in the caller's buffers. */
int bmcts_sync(bmcts_engine* e);
/* device time of the last batch call's kernel(s) in milliseconds (CUDA events on the engine stream) */
int bmcts_last_kernel_ms(bmcts_engine* e, float* ms);
/* the two event records around every launch are on by default; a caller that never asks for
 * bmcts_last_kernel_ms (the planner's tree search: one launch per MCTS wave, many host threads
 * sharing the driver) switches them off.  bmcts_last_kernel_ms then fails with INVALID_ARG. */
int bmcts_set_kernel_timing(bmcts_engine* e, int on);
/* how the last rollout / step kernel was launched: lanes per rollout (1 ... 32), rollout warps
 * per block, and whether it was the "ahead" variant the engine picks for batches too small to
 * fill the GPU (0 / 1).  Zeros before the first launch.  Any pointer may be null. */
int bmcts_last_launch_geometry(const bmcts_engine* e, int* lanes_per_rollout, int* warps_per_block,
                               int* ahead);
/* number of kernels this engine has launched since creation */
uint64_t bmcts_launch_count(const bmcts_engine* e);
/* the engine's CUDA stream as an opaque pointer (cudaStream_t) */
void* bmcts_stream(bmcts_engine* e);
const char* bmcts_last_error(const bmcts_engine* e);
const char* bmcts_status_string(int status);

#ifdef __cplusplus
}
#endif

#endif /* BMCTS_C_H_ */
