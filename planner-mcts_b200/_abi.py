"""ctypes mirror of include/bmcts_c.h (the C ABI of libbmcts.so).

The tests reuse these structure definitions for their CPU checker, which takes
the identical PODs.
"""
import ctypes as C
import os

MAX_AGENTS = 32
MAX_CORRIDORS = 4
MAX_CORRIDOR_PTS = 16
MAX_ROAD_PTS = 32
MAX_GOAL_PTS = 8
MAX_EXTRA_GOALS = 3
MAX_HYPOTHESES = 64
MAX_EGO_ACTIONS = 16
NUM_IDM_PARAMS = 6
MAX_DEPTH = 2048

OK, ERR_INVALID_ARG, ERR_CUDA, ERR_NO_SCENARIO, ERR_UNSUPPORTED, ERR_NO_DEVICE, ERR_COMM = range(7)

F_COLLISION_OTHER = 1 << 0
F_STATIC_SAFE_DIST = 1 << 1
F_DYNAMIC_SAFE_DIST = 1 << 2
F_COLLISION_DRIVABLE = 1 << 3
F_GOAL_REACHED = 1 << 4
F_OUT_OF_MAP = 1 << 5
F_TERMINAL = 1 << 6

STATE_HYPOTHESIS, STATE_RISK_CONSTRAINT, STATE_COOPERATIVE = 0, 1, 2
ACT_CONST_ACC_STAY_LANE, ACT_CHANGE_LEFT, ACT_CHANGE_RIGHT = 0, 1, 2
GOAL_NONE, GOAL_POLYGON, GOAL_FRENET_LIMITS = 0, 1, 2
MEM_HOST, MEM_DEVICE = 0, 1
ACTION_NONE = 0xFF   # bmcts_step_in.action[slot 0]: no expansion step for this element

f32, i32, u32, u64 = C.c_float, C.c_int32, C.c_uint32, C.c_uint64


class Config(C.Structure):
    _fields_ = [
        ("goal_reward", f32), ("collision_reward", f32), ("safe_dist_violated_reward", f32),
        ("out_of_drivable_reward", f32), ("goal_cost", f32), ("collision_cost", f32),
        ("safe_dist_violated_cost", f32), ("out_of_drivable_cost", f32),
        ("cooperation_factor", f32), ("step_reward", f32), ("prediction_k", f32),
        ("prediction_alpha", f32), ("normalization_tau", f32),
        ("split_safe_dist_collision", i32), ("chance_costs", i32),
        ("add_safe_dist", i32), ("static_safe_dist_is_terminal", i32),
        ("dynamic_safe_dist_is_terminal", i32), ("out_of_drivable_is_terminal", i32),
        ("drivable_lat_safe_dist", f32), ("drivable_lon_safe_dist", f32),
        ("static_lat_safe_dist", f32), ("static_lon_safe_dist", f32),
        ("dyn_max_other_decel", f32), ("dyn_max_ego_decel", f32),
        ("dyn_reaction_time_others", f32), ("dyn_reaction_time_ego", f32), ("dyn_to_rear", i32),
        ("discount_factor", f32), ("max_rollout_steps", i32),
        ("wheel_base", f32), ("delta_max", f32), ("crosstrack_gain", f32),
        ("lon_acc_min", f32), ("lon_acc_max", f32), ("ego_max_velocity", f32),
        ("num_substeps", i32), ("state_kind", i32),
        ("limit_steering_rate", i32), ("lat_acc_max", f32), ("lat_acc_min", f32),
        ("ego_integration_dt", f32),
    ]


class Corridor(C.Structure):
    _fields_ = [
        ("n_pts", i32), ("left", i32), ("right", i32), ("half_width", f32),
        ("px", f32 * MAX_CORRIDOR_PTS), ("py", f32 * MAX_CORRIDOR_PTS),
        ("tx", f32 * MAX_CORRIDOR_PTS), ("ty", f32 * MAX_CORRIDOR_PTS),
        ("seg_len", f32 * MAX_CORRIDOR_PTS), ("s0", f32 * MAX_CORRIDOR_PTS),
        ("angle", f32 * MAX_CORRIDOR_PTS), ("length", f32), ("pad_", i32 * 3),
    ]


class Hypothesis(C.Structure):
    _fields_ = [
        ("lo", f32 * NUM_IDM_PARAMS), ("hi", f32 * NUM_IDM_PARAMS),
        ("acc_lower_bound", f32), ("acc_upper_bound", f32),
        ("min_velocity", f32), ("max_velocity", f32), ("exponent", i32), ("pad_", i32 * 3),
    ]


class EgoAction(C.Structure):
    _fields_ = [("type", i32), ("acceleration", f32)]


class Goal(C.Structure):
    _fields_ = [
        ("kind", i32), ("n_pts", i32), ("px", f32 * MAX_GOAL_PTS), ("py", f32 * MAX_GOAL_PTS),
        ("corridor", i32), ("max_lat_left", f32), ("max_lat_right", f32),
        ("max_angle_left", f32), ("max_angle_right", f32), ("vel_lo", f32), ("vel_hi", f32),
        ("pad_", i32),
    ]


class Scenario(C.Structure):
    _fields_ = [
        ("n_corridors", i32), ("n_road_pts", i32), ("n_agents", i32), ("n_hypotheses", i32),
        ("n_ego_actions", i32), ("finalized", i32), ("pad_", i32 * 2),
        ("corridors", Corridor * MAX_CORRIDORS),
        ("road_x", f32 * MAX_ROAD_PTS), ("road_y", f32 * MAX_ROAD_PTS),
        ("goal", Goal),
        ("agent_length", f32 * MAX_AGENTS), ("agent_width", f32 * MAX_AGENTS),
        ("agent_rear", f32 * MAX_AGENTS),
        ("agent_goal", i32 * MAX_AGENTS), ("extra_goals", Goal * MAX_EXTRA_GOALS),
        ("hypotheses", Hypothesis * MAX_HYPOTHESES),
        ("ego_actions", EgoAction * MAX_EGO_ACTIONS), ("pad2_", i32 * 4),
    ]


class RolloutResult(C.Structure):
    _fields_ = [
        ("ret_ego", f32), ("cost0", f32), ("cost1", f32), ("step_len", f32),
        ("flags", u32), ("n_steps", i32), ("agent_steps", i32), ("final_alive", u32),
    ]


class LeafBatch(C.Structure):
    _fields_ = [
        ("n_leaves", i32), ("memory", i32), ("pose", C.c_void_p), ("alive", C.c_void_p),
        ("hyp_id", C.c_void_p), ("depth", C.c_void_p), ("seed", u64), ("stream", u32),
        ("pad_", u32), ("first_rollout_id", u64),
        ("n_states", i32), ("pad2_", i32), ("state_index", C.c_void_p),
    ]


class RolloutOut(C.Structure):
    _fields_ = [("result", C.c_void_p), ("ret_agents", C.c_void_p), ("final_pose", C.c_void_p)]


class StepIn(C.Structure):
    _fields_ = [
        ("n", i32), ("memory", i32), ("pose", C.c_void_p), ("alive", C.c_void_p),
        ("hyp_id", C.c_void_p), ("depth", C.c_void_p), ("action", C.c_void_p),
        ("draw_id", C.c_void_p), ("seed", u64), ("stream", u32), ("pad_", u32),
    ]


class StepOut(C.Structure):
    _fields_ = [
        ("next_pose", C.c_void_p), ("next_alive", C.c_void_p), ("reward", C.c_void_p),
        ("cost", C.c_void_p), ("flags", C.c_void_p), ("last_action", C.c_void_p),
    ]


class LikelihoodIn(C.Structure):
    _fields_ = [
        ("n_worlds", i32), ("memory", i32), ("pose", C.c_void_p), ("alive", C.c_void_p),
        ("action", C.c_void_p), ("n_samples", i32), ("n_buckets", i32),
        ("buckets_lower", f32), ("buckets_upper", f32), ("seed", u64), ("stream", u32),
        ("pad_", u32),
    ]


STRUCTS_IN_HEADER_ORDER = [Config, Corridor, Hypothesis, EgoAction, Goal, Scenario, RolloutResult,
                           LeafBatch, RolloutOut, StepIn, StepOut, LikelihoodIn]

# every symbol include/bmcts_c.h declares
EXPORTED_SYMBOLS = [
    "bmcts_abi_version", "bmcts_config_default", "bmcts_scenario_finalize",
    "bmcts_config_validate", "bmcts_struct_size", "bmcts_alloc_host", "bmcts_free_host",
    "bmcts_create", "bmcts_destroy", "bmcts_set_config", "bmcts_set_scenario",
    "bmcts_rollout_batch", "bmcts_step_batch", "bmcts_expand_rollout_batch",
    "bmcts_expand_rollout_batch_begin", "bmcts_expand_rollout_batch_end",
    "bmcts_likelihood_batch", "bmcts_sync", "bmcts_last_kernel_ms", "bmcts_set_kernel_timing", "bmcts_last_launch_geometry",
    "bmcts_launch_count",
    "bmcts_stream", "bmcts_last_error", "bmcts_status_string", "bmcts_math_probe",
    "bmcts_comm_unique_id", "bmcts_comm_create", "bmcts_comm_destroy", "bmcts_comm_nranks",
    "bmcts_allreduce_root", "bmcts_comm_last_error",
]
COMM_ID_BYTES = 128
MATH_OPS = ["sin", "cos", "atan2", "tan_small", "div", "sqrt", "min", "max", "norm_angle", "fma_aab",
            "pow4"]

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libbmcts.so")

_lib = None


class BmctsError(RuntimeError):
    def __init__(self, status, message=""):
        self.status = status
        super().__init__(f"bmcts status {status}: {message}")


def load_library(path=None):
    """Load libbmcts.so.  Raises if the CUDA extension was not built: the product
    path has no CPU fallback."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise BmctsError(ERR_NO_DEVICE, f"{path} is missing: build it with "
                         "`python planner-mcts_b200/build.py` (there is no CPU fallback)")
    lib = C.CDLL(path)
    P = C.POINTER
    lib.bmcts_abi_version.restype = C.c_int
    lib.bmcts_config_default.argtypes = [P(Config)]
    lib.bmcts_config_default.restype = None
    lib.bmcts_scenario_finalize.argtypes = [P(Scenario)]
    lib.bmcts_config_validate.argtypes = [P(Config)]
    lib.bmcts_struct_size.argtypes = [C.c_int]
    lib.bmcts_struct_size.restype = C.c_size_t
    lib.bmcts_alloc_host.argtypes = [C.c_size_t, P(C.c_void_p)]
    lib.bmcts_free_host.argtypes = [C.c_void_p]
    lib.bmcts_free_host.restype = None
    lib.bmcts_create.argtypes = [C.c_int, P(Config), P(C.c_void_p)]
    lib.bmcts_destroy.argtypes = [C.c_void_p]
    lib.bmcts_destroy.restype = None
    lib.bmcts_set_config.argtypes = [C.c_void_p, P(Config)]
    lib.bmcts_set_scenario.argtypes = [C.c_void_p, P(Scenario)]
    lib.bmcts_rollout_batch.argtypes = [C.c_void_p, P(LeafBatch), P(RolloutOut)]
    lib.bmcts_step_batch.argtypes = [C.c_void_p, P(StepIn), P(StepOut)]
    lib.bmcts_expand_rollout_batch.argtypes = [C.c_void_p, P(StepIn), P(StepOut), u64,
                                               P(RolloutOut)]
    lib.bmcts_expand_rollout_batch_begin.argtypes = [C.c_void_p, P(StepIn), P(StepOut), u64,
                                                     P(RolloutOut)]
    lib.bmcts_expand_rollout_batch_end.argtypes = [C.c_void_p]
    lib.bmcts_likelihood_batch.argtypes = [C.c_void_p, P(LikelihoodIn), C.c_void_p]
    lib.bmcts_sync.argtypes = [C.c_void_p]
    lib.bmcts_last_kernel_ms.argtypes = [C.c_void_p, P(C.c_float)]
    lib.bmcts_set_kernel_timing.argtypes = [C.c_void_p, C.c_int]
    lib.bmcts_last_launch_geometry.argtypes = [C.c_void_p, P(C.c_int), P(C.c_int), P(C.c_int)]
    lib.bmcts_launch_count.argtypes = [C.c_void_p]
    lib.bmcts_launch_count.restype = C.c_uint64
    lib.bmcts_stream.argtypes = [C.c_void_p]
    lib.bmcts_stream.restype = C.c_void_p
    lib.bmcts_last_error.argtypes = [C.c_void_p]
    lib.bmcts_last_error.restype = C.c_char_p
    lib.bmcts_status_string.argtypes = [C.c_int]
    lib.bmcts_status_string.restype = C.c_char_p
    lib.bmcts_comm_unique_id.argtypes = [C.c_void_p]
    lib.bmcts_comm_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, P(C.c_void_p)]
    lib.bmcts_comm_destroy.argtypes = [C.c_void_p]
    lib.bmcts_comm_destroy.restype = None
    lib.bmcts_comm_nranks.argtypes = [C.c_void_p]
    lib.bmcts_allreduce_root.argtypes = [C.c_void_p, P(C.c_double), C.c_size_t]
    lib.bmcts_comm_last_error.argtypes = [C.c_void_p]
    lib.bmcts_comm_last_error.restype = C.c_char_p
    if path == LIB_PATH:
        _lib = lib
    return lib


def default_config(lib=None):
    lib = lib or load_library()
    cfg = Config()
    lib.bmcts_config_default(C.byref(cfg))
    return cfg
