"""Independent float64 restatement of the normative step spec.

TEST INFRASTRUCTURE.  Written from the prose of DESIGN.md section 3 and the cited reference
lines, NOT from oracle/bmcts_oracle.c: different language, float64 instead of FP32, numpy /
libm elementary functions instead of include/bmcts_math.h, and different algorithms wherever
the spec states a predicate rather than a procedure:

  * point in polygon      : winding number            (oracle / kernels: even-odd crossing test)
  * rectangles overlap    : Sutherland-Hodgman clip, area of the intersection > 0
                                                      (oracle / kernels: separating axes)
  * Frenet projection     : all segments at once, numpy argmin
  * safe distance         : the four cases of Rizaldi, Immler, Althoff (2016), "A formally
                            verified checker of the safe distance traffic rules", section 4,
                            written from the paper's case split
  * Philox4x32-10         : Python integers, from Salmon et al. (SC'11) / the Random123 spec

tests/test_independent_spec.py compares the oracle (and through it the CUDA path) against this
file: values within 1e-5 / 1e-4, flags exactly whenever no predicate is within a margin of its
threshold (FP32 and float64 may legitimately disagree there; `margin` reports the distance).

Paths cited are relative to /root/reference/bark_mcts/models/behavior/.
"""
import math

import numpy as np

F_COLLISION, F_STATIC, F_DYNAMIC, F_DRIVABLE, F_GOAL, F_OOM, F_TERMINAL = (1 << i for i in range(7))
RNG_ROLLOUT_PARAMS, RNG_ROLLOUT_ACTION, RNG_TREE_PARAMS, RNG_LIKELIHOOD = 0, 1, 2, 3
M32 = 0xFFFFFFFF


# ------------------------------------------------------------------ Philox4x32-10
def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0 = 0xD2511F53 * c0
        p1 = 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & M32, p1 & M32, ((p0 >> 32) ^ c3 ^ k1) & M32, p0 & M32
        k0 = (k0 + 0x9E3779B9) & M32
        k1 = (k1 + 0xBB67AE85) & M32
    return c0, c1, c2, c3


def draw(seed, stream, ident, agent, domain, index):
    """include/bmcts_c.h 'Counter-based RNG contract'"""
    key = (seed & M32, ((seed >> 32) & M32) ^ stream)
    ctr = (ident & M32, (ident >> 32) & M32, agent | (domain << 16), index)
    return philox4x32_10(ctr, key)


def u01(bits):
    return (bits >> 8) / 16777216.0


# ------------------------------------------------------------------ scenario view
class Scenario:
    """plain-Python view of a finalized bmcts_scenario + bmcts_config (ctypes structures)"""

    def __init__(self, cfg, scn):
        self.cfg = cfg
        self.A = scn.n_agents
        self.corridors = []
        for c in range(scn.n_corridors):
            k = scn.corridors[c]
            pts = np.array([[k.px[i], k.py[i]] for i in range(k.n_pts)], dtype=np.float64)
            self.corridors.append({"pts": pts, "half_width": float(k.half_width), "left": k.left,
                                   "right": k.right})
        self.road = np.array([[scn.road_x[i], scn.road_y[i]] for i in range(scn.n_road_pts)], np.float64)
        # goal per agent slot: the ego's, or (cooperative state) the agent's own
        self.goals = []
        for a in range(self.A):
            g = scn.extra_goals[scn.agent_goal[a] - 1] if scn.agent_goal[a] > 0 else scn.goal
            self.goals.append((g.kind, np.array([[g.px[i], g.py[i]] for i in range(g.n_pts)], np.float64), g))
        self.length = [float(scn.agent_length[a]) for a in range(self.A)]
        self.width = [float(scn.agent_width[a]) for a in range(self.A)]
        self.rear = [float(scn.agent_rear[a]) for a in range(self.A)]
        self.hyps = [scn.hypotheses[h] for h in range(scn.n_hypotheses)]
        self.actions = [(scn.ego_actions[i].type, float(scn.ego_actions[i].acceleration))
                        for i in range(scn.n_ego_actions)]
        for c in self.corridors:   # derived geometry, recomputed in float64
            d = np.diff(c["pts"], axis=0)
            c["seg_len"] = np.hypot(d[:, 0], d[:, 1])
            c["tan"] = d / c["seg_len"][:, None]
            c["s0"] = np.concatenate([[0.0], np.cumsum(c["seg_len"])[:-1]])
            c["angle"] = np.arctan2(d[:, 1], d[:, 0])
            c["total"] = float(c["seg_len"].sum())


def wrap(a):
    """angle in [-pi, pi)"""
    return (a + math.pi) % (2.0 * math.pi) - math.pi


def project(corr, x, y):
    """spec item 2: nearest point on the centre line; returns s, signed lateral offset d (left
    positive, measured to the LINE of the nearest segment), segment index, in_range"""
    r = np.array([x, y]) - corr["pts"][:-1]
    u = np.einsum("ij,ij->i", r, corr["tan"])
    uc = np.clip(u, 0.0, corr["seg_len"])
    e = r - uc[:, None] * corr["tan"]
    k = int(np.argmin(np.einsum("ij,ij->i", e, e)))       # first minimum wins
    d = corr["tan"][k, 0] * r[k, 1] - corr["tan"][k, 1] * r[k, 0]
    n = len(uc)
    in_range = not ((k == 0 and u[k] < 0.0) or (k == n - 1 and u[k] > corr["seg_len"][k]))
    return corr["s0"][k] + uc[k], d, k, in_range


def winding_inside(poly, x, y):
    """winding number != 0 (for the simple polygons of a road corridor: same set as even-odd)"""
    wn = 0
    n = len(poly)
    for i in range(n):
        x0, y0 = poly[i]
        x1, y1 = poly[(i + 1) % n]
        left = (x1 - x0) * (y - y0) - (x - x0) * (y1 - y0)
        if y0 <= y:
            if y1 > y and left > 0:
                wn += 1
        elif y1 <= y and left < 0:
            wn -= 1
    return wn != 0


def polygon_margin(poly, x, y):
    """distance of a point to the polygon's boundary (how far a containment test is from flipping)"""
    best = math.inf
    n = len(poly)
    for i in range(n):
        a, b = poly[i], poly[(i + 1) % n]
        ab = b - a
        t = np.clip(np.dot(np.array([x, y]) - a, ab) / np.dot(ab, ab), 0.0, 1.0)
        best = min(best, float(np.hypot(*(a + t * ab - np.array([x, y])))))
    return best


def rectangle(sc, a, x, y, th, lat=0.0, lon=0.0):
    """corners (counter-clockwise) of agent a's shape inflated by lon at both ends and lat at both
    sides; the reference point sits `rear` ahead of the rear edge on the centre axis"""
    front = sc.length[a] - sc.rear[a] + lon
    back = sc.rear[a] + lon
    hw = 0.5 * sc.width[a] + lat
    c, s = math.cos(th), math.sin(th)
    loc = [(front, hw), (-back, hw), (-back, -hw), (front, -hw)]
    return np.array([[x + lx * c - ly * s, y + lx * s + ly * c] for lx, ly in loc])


def clip_area(subject, clipper):
    """area of the intersection of two convex counter-clockwise polygons (Sutherland-Hodgman)"""
    out = [tuple(p) for p in subject]
    n = len(clipper)
    for i in range(n):
        a, b = clipper[i], clipper[(i + 1) % n]
        inp, out = out, []
        if not inp:
            break

        def side(p):
            return (b[0] - a[0]) * (p[1] - a[1]) - (b[1] - a[1]) * (p[0] - a[0])
        for j in range(len(inp)):
            p, q = inp[j], inp[(j + 1) % len(inp)]
            sp, sq = side(p), side(q)
            if sp >= 0:
                out.append(p)
            if (sp >= 0) != (sq >= 0):
                t = sp / (sp - sq)
                out.append((p[0] + t * (q[0] - p[0]), p[1] + t * (q[1] - p[1])))
    if len(out) < 3:
        return 0.0
    xs, ys = np.array([p[0] for p in out]), np.array([p[1] for p in out])
    return 0.5 * abs(float(np.dot(xs, np.roll(ys, -1)) - np.dot(ys, np.roll(xs, -1))))


def rect_gap(r1, r2):
    """separation of two convex polygons (> 0 apart, < 0 penetration depth), by the separating
    axes -- used ONLY as the margin of the overlap predicate, not as the predicate"""
    best = -math.inf
    for poly, other in ((r1, r2), (r2, r1)):
        n = len(poly)
        for i in range(n):
            e = poly[(i + 1) % n] - poly[i]
            nrm = np.array([e[1], -e[0]]) / np.hypot(*e)      # outward normal of a ccw polygon
            best = max(best, float(np.min((other - poly[i]) @ nrm)))
    return best


# ------------------------------------------------------------------ safe distance
def safe_distance_ok(v_front, v_rear, dist, a_rear, a_front, delta):
    """Rizaldi, Immler, Althoff: the rear vehicle reacts after `delta` seconds and then brakes with
    a_rear < 0, the front vehicle brakes with a_front < 0 from t = 0.  True = `dist` is safe.
    Cases (paper section 4): the front vehicle has stopped before the reaction time is over or
    not; and whether the rear vehicle stops behind the stopping point of the front vehicle or
    the braking trajectories must be compared in motion (a_front > a_rear, the rear vehicle still
    faster at the end of its reaction time, and it stops earlier than the front vehicle would)."""
    if dist < 0.0:
        return False
    stop_front = -v_front / a_front
    stop_dist_rear = v_rear * delta - v_rear * v_rear / (2.0 * a_rear)     # reaction + braking
    stop_dist_front = -v_front * v_front / (2.0 * a_front)
    if dist > stop_dist_rear:
        return True                                    # rear stops before the front's position now
    front_moving_at_delta = delta <= stop_front
    travelled_front = v_front * delta + 0.5 * a_front * delta * delta
    if front_moving_at_delta and dist > stop_dist_rear - travelled_front:
        return True
    if front_moving_at_delta:
        v_front_delta = v_front + a_front * delta
        t_stop_rear = -v_rear / a_rear
        t_stop_front_after = -v_front_delta / a_front
        if a_front > a_rear and v_front_delta < v_rear and t_stop_rear < t_stop_front_after:
            rel = v_front_delta - v_rear
            return dist > rel * rel / (2.0 * (a_front - a_rear)) - travelled_front + v_rear * delta
    return dist > stop_dist_rear - stop_dist_front


# ------------------------------------------------------------------ world
class World:
    def __init__(self, sc, pose, alive_mask):
        self.sc = sc
        self.x = np.array(pose[:, 0], dtype=np.float64)
        self.y = np.array(pose[:, 1], dtype=np.float64)
        self.th = np.array(pose[:, 2], dtype=np.float64)
        self.v = np.array(pose[:, 3], dtype=np.float64)
        self.alive = [bool((int(alive_mask) >> a) & 1) for a in range(sc.A)]
        self.frenet()

    def copy(self):
        w = World.__new__(World)
        w.sc = self.sc
        w.x, w.y, w.th, w.v = self.x.copy(), self.y.copy(), self.th.copy(), self.v.copy()
        w.alive = list(self.alive)
        return w

    def frenet(self):
        sc = self.sc
        C = len(sc.corridors)
        self.s = np.zeros((sc.A, C))
        self.d = np.zeros((sc.A, C))
        self.seg = np.zeros((sc.A, C), dtype=int)
        self.inr = np.zeros((sc.A, C), dtype=bool)
        self.cur = [0] * sc.A
        for a in range(sc.A):
            if not self.alive[a]:
                continue
            for c in range(C):
                self.s[a, c], self.d[a, c], self.seg[a, c], self.inr[a, c] = project(
                    sc.corridors[c], self.x[a], self.y[a])
            cand = [(abs(self.d[a, c]), c) for c in range(C) if self.inr[a, c]]
            self.cur[a] = min(cand)[1] if cand else 0

    def inside(self, a, c):
        return self.inr[a, c] and abs(self.d[a, c]) <= self.sc.corridors[c]["half_width"]

    def mask(self):
        return sum(1 << a for a in range(self.sc.A) if self.alive[a])


def prediction_time_span(cfg, depth):
    return float(np.float32(float(cfg.prediction_k) * float(depth) ** float(cfg.prediction_alpha)))


def idm_parameters(hyp, r0, r1):
    u = [u01(r0[0]), u01(r0[1]), u01(r0[2]), u01(r0[3]), u01(r1[0])]
    return [float(hyp.lo[m]) + u[m] * (float(hyp.hi[m]) - float(hyp.lo[m])) for m in range(5)]


def idm_step(w, i, hyp, params, Dt):
    """spec item 4: BehaviorIDMStochastic::Plan of agent i against the pre-step world w"""
    sc, cfg = w.sc, w.sc.cfg
    T, s0, a_max, v0, b = params
    c = w.cur[i]
    corr = sc.corridors[c]
    ahead = [(w.s[j, c] - w.s[i, c], j) for j in range(sc.A)
             if j != i and w.alive[j] and w.inside(j, c) and w.s[j, c] - w.s[i, c] > 0.0]
    lead = min(ahead) if ahead else None     # nearest; lowest slot on equal distances
    nsub = cfg.num_substeps
    dt = Dt / nsub
    v, trav, acc = w.v[i], 0.0, 0.0
    if lead:
        gap = lead[0] - ((sc.length[i] - sc.rear[i]) + sc.rear[lead[1]])
        v_l = w.v[lead[1]]
    for _ in range(nsub):
        free = 1.0 - (v / max(v0, 1e-3)) ** hyp.exponent
        if lead:
            s_star = s0 + max(0.0, v * T + v * (v - v_l) / (2.0 * math.sqrt(max(a_max * b, 1e-6))))
            acc = a_max * (free - (s_star / max(gap, 1e-3)) ** 2)
        else:
            acc = a_max * free
        acc = min(max(acc, float(hyp.acc_lower_bound)), float(hyp.acc_upper_bound))
        ds = max(0.0, v * dt + 0.5 * acc * dt * dt)
        v = min(max(v + acc * dt, float(hyp.min_velocity)), float(hyp.max_velocity))
        if lead:
            gap += v_l * dt - ds
        trav += ds
    s_new = min(w.s[i, c] + trav, corr["total"])
    k = int(np.searchsorted(corr["s0"], s_new, side="right") - 1)
    k = min(max(k, 0), len(corr["s0"]) - 1)
    p = corr["pts"][k] + (s_new - corr["s0"][k]) * corr["tan"][k]
    return p[0], p[1], corr["angle"][k], v, acc


def ego_substeps(cfg, Dt):
    if not cfg.ego_integration_dt > 0.0:
        return cfg.num_substeps
    return int(min(max(math.ceil(Dt / float(cfg.ego_integration_dt) - 1e-3), 1), 1024))


def macro_step(w, i, action, Dt):
    """spec item 5: single-track model under the macro action, Stanley law on the front axle,
    steering limited by delta_max and (LimitSteeringRate) by the lateral acceleration limits"""
    sc, cfg = w.sc, w.sc.cfg
    kind, acc = sc.actions[action]
    acc = min(max(acc, float(cfg.lon_acc_min)), float(cfg.lon_acc_max))
    tgt = w.cur[i]
    if kind == 1 and sc.corridors[tgt]["left"] >= 0:
        tgt = sc.corridors[tgt]["left"]
    elif kind == 2 and sc.corridors[tgt]["right"] >= 0:
        tgt = sc.corridors[tgt]["right"]
    corr = sc.corridors[tgt]
    L = float(cfg.wheel_base)
    n = ego_substeps(cfg, Dt)
    dt = Dt / n
    x, y, th, v = w.x[i], w.y[i], w.th[i], w.v[i]
    for _ in range(n):
        fx, fy = x + L * math.cos(th), y + L * math.sin(th)
        _, d, seg, _ = project(corr, fx, fy)
        delta = wrap(corr["angle"][seg] - th) + math.atan2(-float(cfg.crosstrack_gain) * d, v)
        delta = min(max(delta, -float(cfg.delta_max)), float(cfg.delta_max))
        tn = math.tan(delta)
        if cfg.limit_steering_rate:
            r = L / max(v * v, 1e-6)
            tn = min(max(tn, float(cfg.lat_acc_min) * r), float(cfg.lat_acc_max) * r)
        x, y = x + dt * v * math.cos(th), y + dt * v * math.sin(th)
        th = th + dt * v * tn / L
        v = min(max(v + dt * acc, 0.0), float(cfg.ego_max_velocity))
    return x, y, wrap(th), v, acc


def evaluate(w, e, margins):
    """mcts_state_base.hpp:248-279 from agent e's perspective; appends to `margins` how far every
    geometric predicate is from its threshold"""
    sc, cfg = w.sc, w.sc.cfg
    if not w.alive[e]:
        return F_OOM | F_TERMINAL
    f = 0
    body = rectangle(sc, e, w.x[e], w.y[e], w.th[e])
    infl = rectangle(sc, e, w.x[e], w.y[e], w.th[e], float(cfg.drivable_lat_safe_dist),
                     float(cfg.drivable_lon_safe_dist))
    for p in infl:
        margins.append(polygon_margin(sc.road, p[0], p[1]))
    if not all(winding_inside(sc.road, p[0], p[1]) for p in infl):
        f |= F_DRIVABLE
    others = [j for j in range(sc.A) if j != e and w.alive[j]]
    rects = {j: rectangle(sc, j, w.x[j], w.y[j], w.th[j]) for j in others}
    for j in others:
        margins.append(abs(rect_gap(body, rects[j])))
        if clip_area(body, rects[j]) > 0.0:
            f |= F_COLLISION
    goal_kind, goal_poly, goal = sc.goals[e]
    if goal_kind == 1:
        for p in body:
            margins.append(polygon_margin(goal_poly, p[0], p[1]))
        if all(winding_inside(goal_poly, p[0], p[1]) for p in body):
            f |= F_GOAL
    elif goal_kind == 2:
        g = goal
        c = g.corridor
        ang = wrap(w.th[e] - sc.corridors[c]["angle"][w.seg[e, c]])
        d = w.d[e, c]
        tests = [g.max_lat_left - d, g.max_lat_right + d, g.max_angle_left - ang,
                 g.max_angle_right + ang, w.v[e] - g.vel_lo, g.vel_hi - w.v[e]]
        margins.extend(abs(t) for t in tests)
        if w.inr[e, c] and all(t >= 0.0 for t in tests):
            f |= F_GOAL
    if cfg.add_safe_dist:
        env = rectangle(sc, e, w.x[e], w.y[e], w.th[e], float(cfg.static_lat_safe_dist),
                        float(cfg.static_lon_safe_dist))
        for j in others:
            margins.append(abs(rect_gap(env, rects[j])))
            if clip_area(env, rects[j]) > 0.0:
                f |= F_STATIC
        c = w.cur[e]
        rel = [(w.s[j, c] - w.s[e, c], j) for j in others if w.inside(j, c)]
        for j in others:   # corridor membership is itself a threshold
            if w.inr[j, c]:
                margins.append(abs(abs(w.d[j, c]) - sc.corridors[c]["half_width"]))
        ahead = [(ds, j) for ds, j in rel if ds > 0.0]
        behind = [(-ds, j) for ds, j in rel if ds < 0.0]
        safe = True
        if ahead:
            ds, j = min(ahead)
            dist = ds - ((sc.length[e] - sc.rear[e]) + sc.rear[j])
            safe &= safe_distance_ok(w.v[j], w.v[e], dist, -float(cfg.dyn_max_ego_decel),
                                     -float(cfg.dyn_max_other_decel), float(cfg.dyn_reaction_time_ego))
        if behind and cfg.dyn_to_rear:
            ds, j = min(behind)
            dist = ds - ((sc.length[j] - sc.rear[j]) + sc.rear[e])
            safe &= safe_distance_ok(w.v[e], w.v[j], dist, -float(cfg.dyn_max_other_decel),
                                     -float(cfg.dyn_max_ego_decel), float(cfg.dyn_reaction_time_others))
        if not safe:
            f |= F_DYNAMIC
    terminal = f & (F_COLLISION | F_GOAL) or (f & F_DRIVABLE and cfg.out_of_drivable_is_terminal) or \
        (f & F_STATIC and cfg.static_safe_dist_is_terminal) or \
        (f & F_DYNAMIC and cfg.dynamic_safe_dist_is_terminal)
    return f | (F_TERMINAL if terminal else 0)


def reward_of(cfg, f, Dt):
    """mcts_state_base.hpp:107-118"""
    b = lambda m: 1.0 if f & m else 0.0
    sd = (b(F_DYNAMIC) if cfg.static_safe_dist_is_terminal else max(b(F_DYNAMIC), b(F_STATIC))) \
        * cfg.safe_dist_violated_reward * Dt
    col = max(b(F_COLLISION), b(F_STATIC) if cfg.static_safe_dist_is_terminal else 0.0) * cfg.collision_reward \
        + max(b(F_DRIVABLE), b(F_OOM)) * cfg.out_of_drivable_reward
    return max(col, cfg.collision_reward) + (sd if cfg.add_safe_dist else 0.0) \
        + b(F_GOAL) * cfg.goal_reward + cfg.step_reward


def costs_of(cfg, f, Dt):
    """mcts_state_base.hpp:120-142"""
    b = lambda m: 1.0 if f & m else 0.0
    sdc = (b(F_DYNAMIC) if cfg.static_safe_dist_is_terminal else max(b(F_DYNAMIC), b(F_STATIC))) \
        * cfg.safe_dist_violated_cost * Dt
    hit = max(b(F_COLLISION), b(F_STATIC) if cfg.static_safe_dist_is_terminal else 0.0,
              b(F_DYNAMIC) if cfg.dynamic_safe_dist_is_terminal else 0.0)
    tot = hit * cfg.collision_cost + max(b(F_DRIVABLE), b(F_OOM)) * cfg.out_of_drivable_cost \
        + (sdc if cfg.add_safe_dist and not cfg.split_safe_dist_collision else 0.0) + b(F_GOAL) * cfg.goal_cost
    if cfg.split_safe_dist_collision:
        return (min(sdc, 1.0) if cfg.chance_costs else sdc), min(tot, cfg.collision_cost) * Dt
    return (min(tot, 1.0) if cfg.chance_costs else tot), 0.0


def step(w, hyp_id, actions, draws, depth, coop):
    """one execute(): returns the post-step world, ego flags, rewards[A], costs, applied
    accelerations, the prediction time span and the predicate margins.
    draws[i] = (r0, r1): the two Philox blocks of other agent i's parameter draw."""
    sc, cfg = w.sc, w.sc.cfg
    Dt = prediction_time_span(cfg, depth)
    nw = w.copy()
    last = [0.0] * sc.A
    for i in range(sc.A):
        if not w.alive[i]:
            continue
        if i == 0 or coop:
            nw.x[i], nw.y[i], nw.th[i], nw.v[i], last[i] = macro_step(w, i, actions[i], Dt)
        else:
            hyp = sc.hyps[hyp_id[i]]
            nw.x[i], nw.y[i], nw.th[i], nw.v[i], last[i] = idm_step(
                w, i, hyp, idm_parameters(hyp, *draws[i]), Dt)
    margins = []
    for i in range(sc.A):
        if nw.alive[i]:
            margins.append(polygon_margin(sc.road, nw.x[i], nw.y[i]))
            if not winding_inside(sc.road, nw.x[i], nw.y[i]):
                nw.alive[i] = False
    nw.frenet()
    f = evaluate(nw, 0, margins)
    rewards = [0.0] * sc.A
    if not coop:
        rewards[0] = reward_of(cfg, f, Dt)
        c0, c1 = costs_of(cfg, f, Dt)
    else:
        # mcts_state_cooperative.cpp:70-116, including the divisor that grows over agents of
        # this state that left the map (the reference's other_agents_rewards[idx] inserts zeros)
        own = {0: reward_of(cfg, f, Dt)}
        for j in range(1, sc.A):
            if nw.alive[j]:
                own[j] = reward_of(cfg, evaluate(nw, j, margins), Dt)
        total = sum(own.values())
        cf = float(cfg.cooperation_factor)
        n_map = len(own) - 1
        for j in range(1, sc.A):
            if not w.alive[j]:
                continue
            if j not in own:
                n_map += 1
            r = own.get(j, 0.0)
            rewards[j] = ((1.0 - cf) * r + cf * (total - r)) / (n_map + 1)
        rewards[0] = ((1.0 - cf) * own[0] + cf * (total - own[0])) / (n_map + 1)
        c0, c1 = -rewards[0], 0.0
    return nw, f, rewards, (c0, c1), last, Dt, margins


def rollout(sc, pose, alive, hyp_id, depth, seed, stream, rollout_id):
    """spec item 9 (mamcts RandomHeuristic): returns dict with the accumulated values, the OR of
    flags, step count and the smallest predicate margin met on the way"""
    cfg = sc.cfg
    coop = cfg.state_kind == 2
    w = World(sc, np.asarray(pose, dtype=np.float64), alive)
    disc = float(cfg.discount_factor)
    R = C0 = C1 = L = 0.0
    Ra = [0.0] * sc.A
    f_any = f_last = 0
    n = agent_steps = 0
    margin = math.inf
    while n < cfg.max_rollout_steps:
        actions = [0] * sc.A
        for i in range(sc.A):
            if (i == 0 or coop) and w.alive[i]:
                r = draw(seed, stream, rollout_id, i, RNG_ROLLOUT_ACTION, n)
                actions[i] = (r[0] * len(sc.actions)) >> 32
        draws = {i: (draw(seed, stream, rollout_id, i, RNG_ROLLOUT_PARAMS, 2 * n),
                     draw(seed, stream, rollout_id, i, RNG_ROLLOUT_PARAMS, 2 * n + 1))
                 for i in range(1, sc.A) if not coop}
        agent_steps += sum(w.alive)
        w, f, rw, (c0, c1), _, Dt, m = step(w, hyp_id, actions, draws, depth, coop)
        margin = min([margin] + m)
        R += disc * rw[0]
        C0 += disc * c0
        C1 += disc * c1
        for i in range(sc.A):
            Ra[i] += disc * rw[i]
        L += Dt
        disc *= float(cfg.discount_factor)
        depth += 1
        n += 1
        f_last = f
        f_any |= f
        if f & F_TERMINAL:
            break
    return {"ret_ego": R, "cost0": C0, "cost1": C1, "step_len": L, "flags": f_last | (f_any << 8),
            "n_steps": n, "agent_steps": agent_steps, "final_alive": w.mask(), "ret_agents": Ra,
            "final_pose": np.stack([w.x, w.y, w.th, w.v], axis=1), "margin": margin}
