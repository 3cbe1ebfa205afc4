"""Synthetic scenarios and leaf-state batches for the BASELINE.json configs.

SURVEY.md section 8d: 2-lane straight highway (lane centres y = -1.75 / -4.75 as in
/root/reference/bark_mcts/models/behavior/tests/test_helpers.cpp:43-44) and a merge
(ramp tapering into the main lane over 150 m); 4 m x 2 m rectangle cars with the
reference point 1.25 m ahead of the rear edge; initial speeds U[5, 15] m/s, gaps
U[8, 40] m; hypothesis sets built by equal-width splits of the parameter ranges
(hypothesis/behavior_space/behavior_space.py:102-106).
"""
import numpy as np

from . import _abi

CAR_LENGTH, CAR_WIDTH, CAR_REAR = 4.0, 2.0, 1.25


def _fill(arr, values):
    for i, v in enumerate(values):
        arr[i] = v


def make_corridor(points, half_width, left=-1, right=-1):
    c = _abi.Corridor()
    pts = np.asarray(points, dtype=np.float32)
    c.n_pts = len(pts)
    c.left, c.right, c.half_width = left, right, half_width
    _fill(c.px, pts[:, 0])
    _fill(c.py, pts[:, 1])
    return c


def make_hypothesis(headway=(1.0, 3.0), spacing=(2.0, 2.0), max_acc=(1.7, 1.7),
                    desired_vel=(15.0, 15.0), comft_braking=(1.67, 1.67), coolness=(0.0, 0.0),
                    acc_lower=-5.0, acc_upper=8.0, min_vel=0.0, max_vel=50.0, exponent=4):
    """One BehaviorHypothesisIDM region; (lo, lo) is a FixedValue distribution
    (defaults: tests/data/nheuristic_test.json:6-23)."""
    h = _abi.Hypothesis()
    ranges = [headway, spacing, max_acc, desired_vel, comft_braking, coolness]
    _fill(h.lo, [r[0] for r in ranges])
    _fill(h.hi, [r[1] for r in ranges])
    h.acc_lower_bound, h.acc_upper_bound = acc_lower, acc_upper
    h.min_velocity, h.max_velocity, h.exponent = min_vel, max_vel, exponent
    return h


def split_range(lo, hi, n):
    """create_hypothesis_set_fixed_split: equal-width partitions
    (hypothesis/behavior_space/behavior_space.py:102-106)."""
    w = (hi - lo) / n
    return [(lo + i * w, lo + (i + 1) * w) for i in range(n)]


def hypothesis_set(n_headway=1, n_desired_vel=1, headway=(0.5, 3.5), desired_vel=(8.0, 16.0), **kw):
    hyps = []
    vel_parts = split_range(*desired_vel, n_desired_vel) if n_desired_vel > 1 else [(15.0, 15.0)]
    for hw in split_range(*headway, n_headway):
        for dv in vel_parts:
            hyps.append(make_hypothesis(headway=hw, desired_vel=dv, **kw))
    return hyps


def ego_actions(acc_inputs=(0.0, 1.0, 4.0, -1.0, -8.0), lane_changes=True):
    """BehaviorMacroActionsFromParamServer: AccelerationInputs + lane changes
    (tests/behavior_uct_hypothesis_test.cc:112)."""
    acts = [(_abi.ACT_CONST_ACC_STAY_LANE, a) for a in acc_inputs]
    if lane_changes:
        acts += [(_abi.ACT_CHANGE_LEFT, 0.0), (_abi.ACT_CHANGE_RIGHT, 0.0)]
    return acts


def build_scenario(corridors, road, n_agents, hypotheses, actions, goal=None, shapes=None,
                   agent_goals=None):
    """agent_goals: {agent slot: _abi.Goal} for agents that do not share the ego's goal
    (cooperative state; at most _abi.MAX_EXTRA_GOALS distinct definitions)"""
    lib = _abi.load_library()
    s = _abi.Scenario()
    s.n_corridors = len(corridors)
    for i, c in enumerate(corridors):
        s.corridors[i] = c
    road = np.asarray(road, dtype=np.float32)
    s.n_road_pts = len(road)
    _fill(s.road_x, road[:, 0])
    _fill(s.road_y, road[:, 1])
    s.n_agents = n_agents
    for a in range(n_agents):
        ln, wd, rr = shapes[a] if shapes else (CAR_LENGTH, CAR_WIDTH, CAR_REAR)
        s.agent_length[a], s.agent_width[a], s.agent_rear[a] = ln, wd, rr
    s.n_hypotheses = len(hypotheses)
    for i, h in enumerate(hypotheses):
        s.hypotheses[i] = h
    s.n_ego_actions = len(actions)
    for i, (t, acc) in enumerate(actions):
        s.ego_actions[i].type, s.ego_actions[i].acceleration = t, acc
    if goal is not None:
        s.goal = goal
    extra = []
    for a, g in (agent_goals or {}).items():
        key = bytes(g)
        if key not in extra:
            extra.append(key)
            s.extra_goals[len(extra) - 1] = g
        s.agent_goal[a] = 1 + extra.index(key)
    st = lib.bmcts_scenario_finalize(s)
    if st != _abi.OK:
        raise _abi.BmctsError(st, "bmcts_scenario_finalize: " +
                              lib.bmcts_status_string(st).decode())
    return s


def polygon_goal(points):
    g = _abi.Goal()
    pts = np.asarray(points, dtype=np.float32)
    g.kind, g.n_pts = _abi.GOAL_POLYGON, len(pts)
    _fill(g.px, pts[:, 0])
    _fill(g.py, pts[:, 1])
    return g


def frenet_goal(corridor, max_lat=(0.2, 0.2), max_angle=(0.08, 0.08), vel=(10.0, 20.0)):
    """GoalDefinitionStateLimitsFrenet (tests/test_helpers.cpp:67-69)."""
    g = _abi.Goal()
    g.kind, g.corridor = _abi.GOAL_FRENET_LIMITS, corridor
    g.max_lat_left, g.max_lat_right = max_lat
    g.max_angle_left, g.max_angle_right = max_angle
    g.vel_lo, g.vel_hi = vel
    return g


# ----------------------------------------------------------------------------- maps

def highway_geometry(length=300.0):
    """MakeXodrMapOneRoadTwoLanes-like: left lane centre y=-1.75, right lane y=-4.75."""
    left = make_corridor([(0.0, -1.75), (length, -1.75)], 1.75, left=-1, right=1)
    right = make_corridor([(0.0, -4.75), (length, -4.75)], 1.25, left=0, right=-1)
    road = [(0.0, 0.0), (length, 0.0), (length, -6.0), (0.0, -6.0)]
    return [left, right], road


def merge_geometry(length=400.0, n_main_lanes=1, ramp_start=100.0, taper=150.0, lane_w=3.5):
    """Main road (n_main_lanes lanes, lane 0 rightmost at y=0) plus a ramp on the
    right that runs parallel up to ramp_start and tapers into main lane 0 over
    `taper` metres; the ramp corridor continues along main lane 0 (a BARK lane
    corridor follows the route through the merge)."""
    hw = lane_w / 2
    cors = []
    for i in range(n_main_lanes):
        y = i * lane_w
        cors.append(make_corridor([(0.0, y), (length, y)], hw,
                                  left=(i + 1 if i + 1 < n_main_lanes else -1),
                                  right=(i - 1 if i > 0 else n_main_lanes)))
    ramp_idx = n_main_lanes
    cors.append(make_corridor([(0.0, -lane_w), (ramp_start, -lane_w), (ramp_start + taper, 0.0),
                               (length, 0.0)], hw, left=0, right=-1))
    top = (n_main_lanes - 1) * lane_w + hw
    road = [(0.0, top), (length, top), (length, -hw), (ramp_start + taper, -hw),
            (ramp_start, -lane_w - hw), (0.0, -lane_w - hw)]
    return cors, road, ramp_idx


def curved_geometry(radius=200.0, sweep_deg=90.0, n_lanes=4, lane_w=3.5):
    """A quarter circle at the structure limits: BMCTS_MAX_CORRIDORS lanes of
    BMCTS_MAX_CORRIDOR_PTS points each (15 chords) and a road polygon of
    BMCTS_MAX_ROAD_PTS points.  Lane 0 is the innermost; the road bends to the left.  No
    axis-aligned rectangle fits this polygon, so every drivable-area test takes the scan or the
    safe arcs of the centre lines."""
    n = _abi.MAX_CORRIDOR_PTS
    ang = np.radians(-90.0 + np.linspace(0.0, sweep_deg, n))
    cors = []
    for i in range(n_lanes):
        r = radius + i * lane_w
        # left of the driving direction on a left bend = towards the centre = the lower index
        cors.append(make_corridor(list(zip(r * np.cos(ang), r * np.sin(ang))), lane_w / 2,
                                  left=(i - 1 if i > 0 else -1),
                                  right=(i + 1 if i + 1 < n_lanes else -1)))
    r_in, r_out = radius - lane_w / 2, radius + (n_lanes - 0.5) * lane_w
    m = _abi.MAX_ROAD_PTS // 2
    edge = np.radians(-90.0 + np.linspace(0.0, sweep_deg, m))
    road = list(zip(r_out * np.cos(edge), r_out * np.sin(edge))) + \
        list(zip(r_in * np.cos(edge[::-1]), r_in * np.sin(edge[::-1])))
    return cors, road


# ------------------------------------------------------------------- pose helpers

def pose_on_corridor(c, s):
    """(x, y, theta) at arc length s (numpy float64 -> callers cast to float32)."""
    n = c.n_pts
    px = np.array(c.px[:n], dtype=np.float64)
    py = np.array(c.py[:n], dtype=np.float64)
    seg = np.hypot(np.diff(px), np.diff(py))
    s0 = np.concatenate([[0.0], np.cumsum(seg)])
    s = np.clip(np.asarray(s, dtype=np.float64), 0.0, s0[-1] - 1e-3)
    k = np.clip(np.searchsorted(s0, s, side="right") - 1, 0, n - 2)
    t = (s - s0[k]) / seg[k]
    x = px[k] + t * (px[k + 1] - px[k])
    y = py[k] + t * (py[k + 1] - py[k])
    th = np.arctan2(py[k + 1] - py[k], px[k + 1] - px[k])
    return x, y, th


def sample_leaves(scn, n, lane_of_agent, s_base, rng, s_jitter=3.0, v_range=(5.0, 15.0),
                  n_hyp=None, dead_prob=0.0):
    """n leaf states around a base placement: agent a drives corridor lane_of_agent[a]
    near arc length s_base[a].  Returns agent-major arrays (pose[A,n,4], alive[n],
    hyp_id[A,n])."""
    A = scn.n_agents
    pose = np.zeros((A, n, 4), np.float32)
    for a in range(A):
        c = scn.corridors[lane_of_agent[a]]
        s = s_base[a] + rng.uniform(-s_jitter, s_jitter, n)
        x, y, th = pose_on_corridor(c, s)
        pose[a, :, 0], pose[a, :, 1], pose[a, :, 2] = x, y, th
        pose[a, :, 3] = rng.uniform(v_range[0], v_range[1], n)
    alive = np.full(n, (1 << A) - 1, np.uint32)
    if dead_prob > 0:
        for a in range(1, A):
            dead = rng.random(n) < dead_prob
            alive[dead] &= np.uint32(~(1 << a) & 0xFFFFFFFF)
    H = scn.n_hypotheses if n_hyp is None else n_hyp
    hyp = rng.integers(0, max(H, 1), size=(A, n)).astype(np.uint8)
    return pose, alive, hyp


def place_agents(n_agents, lanes, rng, s_start=20.0, gap=(8.0, 40.0), ego_lane=None, ego_s=60.0):
    """Sequential placement along each lane with bumper gaps U[gap]."""
    lane_of, s_of = [ego_lane if ego_lane is not None else lanes[0]], [ego_s]
    cursor = {l: s_start for l in lanes}
    for i in range(1, n_agents):
        l = lanes[(i - 1) % len(lanes)]
        s = cursor[l]
        if l == lane_of[0] and abs(s - ego_s) < CAR_LENGTH + gap[0]:
            s = ego_s + CAR_LENGTH + gap[0] + rng.uniform(0, 4)
        lane_of.append(l)
        s_of.append(s)
        cursor[l] = s + CAR_LENGTH + rng.uniform(*gap) + 2 * 3.0  # leave room for the leaf jitter
    return lane_of, s_of


# ------------------------------------------------------------- BASELINE.json configs

def _base_config(state_kind, horizon, ego_dt):
    cfg = _abi.default_config()
    cfg.state_kind = state_kind
    cfg.max_rollout_steps = horizon
    cfg.ego_integration_dt = ego_dt
    return cfg


def workload(name, horizon=20, seed=1000, ego_dt=0.05):
    """Returns (config, scenario, meta) for one of the BASELINE.json configs:
    'c1_highway_mdp', 'c2_merge_sbg', 'c3_merge_risk', 'c4_coop', 'c5_rsbg' (+ the test
    geometry 'x_curved_max').  ego_dt is the Euler step of the ego's macro action: 0.05 s gives
    the K_ego = 10 steps per 0.5 s that SURVEY.md section 8d's FLOP formula is stated for; the
    reference default (BehaviorMotionPrimitives::IntegrationTimeDelta) is 0.02 s = 25 steps."""
    rng = np.random.default_rng(seed)
    _cfg = lambda kind, hz: _base_config(kind, hz, ego_dt)
    if name == "c1_highway_mdp":
        cors, road = highway_geometry(300.0)
        A, hyps = 4, hypothesis_set(1, 1, headway=(1.0, 3.0))
        cfg = _cfg(_abi.STATE_HYPOTHESIS, horizon)
        goal = polygon_goal([(250.0, 0.0), (300.0, 0.0), (300.0, -3.5), (250.0, -3.5)])
        lanes, ego_lane, length = [0, 1], 0, 300.0
    elif name in ("c2_merge_sbg", "c3_merge_risk", "c5_rsbg"):
        n_main = 1 if name == "c2_merge_sbg" else 2
        length = 400.0 if name != "c5_rsbg" else 600.0
        cors, road, ramp = merge_geometry(length, n_main_lanes=n_main)
        if name == "c2_merge_sbg":
            A, hyps = 10, hypothesis_set(8, 1)
            cfg = _cfg(_abi.STATE_HYPOTHESIS, horizon)
        elif name == "c3_merge_risk":
            A, hyps = 16, hypothesis_set(8, 1)
            cfg = _cfg(_abi.STATE_RISK_CONSTRAINT, horizon)
        else:
            A, hyps = 32, hypothesis_set(8, 8)
            cfg = _cfg(_abi.STATE_RISK_CONSTRAINT, horizon)
        if name != "c2_merge_sbg":
            cfg.add_safe_dist, cfg.split_safe_dist_collision = 1, 1
        goal = polygon_goal([(length - 60.0, 1.75), (length, 1.75), (length, -1.75),
                             (length - 60.0, -1.75)])
        lanes, ego_lane = list(range(n_main + 1)), ramp
    elif name == "c4_coop":
        cors, road = highway_geometry(500.0)
        A, hyps = 20, []
        cfg = _cfg(_abi.STATE_COOPERATIVE, horizon)
        goal = polygon_goal([(440.0, 0.0), (500.0, 0.0), (500.0, -3.5), (440.0, -3.5)])
        lanes, ego_lane, length = [0, 1], 1, 500.0
    elif name == "x_curved_max":
        # not a BASELINE.json config: geometry at the structure limits for the parity tests
        cors, road = curved_geometry()
        A, hyps = 12, hypothesis_set(2, 2)
        cfg = _cfg(_abi.STATE_RISK_CONSTRAINT, horizon)
        cfg.add_safe_dist, cfg.split_safe_dist_collision = 1, 1
        goal = frenet_goal(2, max_lat=(0.6, 0.6), max_angle=(0.15, 0.15), vel=(3.0, 30.0))
        lanes, ego_lane, length = [0, 1, 2, 3], 1, 300.0
    else:
        raise ValueError(f"unknown workload {name}")
    scn = build_scenario(cors, road, A, hyps, ego_actions(), goal)
    lane_of, s_of = place_agents(A, lanes, rng, ego_lane=ego_lane)
    meta = {"name": name, "A": A, "H": len(hyps), "lane_of": lane_of, "s_of": s_of,
            "length": length, "horizon": horizon}
    return cfg, scn, meta


def leaves_for(scn, meta, n, seed=1):
    rng = np.random.default_rng(seed)
    return sample_leaves(scn, n, meta["lane_of"], meta["s_of"], rng)
