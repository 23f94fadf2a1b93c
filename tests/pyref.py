"""Second, independent restatement of the reference rules in plain Python (TEST INFRASTRUCTURE ONLY).

Written directly from the Julia sources (reference src/agent_actions.jl:24-336, src/model_actions.jl:15-32) with
dict/list data structures shaped like Agents.jl's, for SMALL cases only.  It shares no code with the C++ oracle
(oracle/abm_oracle.cpp): the two restatements pin each other, and the micro-scenario tests derive their expected
values from this file's primitives (Philox draws, samplers) rather than from the oracle.
"""
import math
import struct

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    c = list(ctr)
    k = list(key)
    for _ in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & MASK, (p0 >> 32) ^ c[3] ^ k[1], p0 & MASK]
        k = [(k[0] + W0) & MASK, (k[1] + W1) & MASK]
    return c


class Rng:
    """Injected AbstractRNG: counts rand(UInt64) calls since reset(tick, id); include/abm.h randomness contract."""

    def __init__(self, seed):
        self.seed, self.tick, self.id, self.n = seed, 0, 0, 0

    def reset(self, tick, aid):
        self.tick, self.id, self.n = tick, aid, 0

    def u64(self):
        o = philox4x32_10([(self.n >> 1) & MASK, self.tick & MASK, self.id & MASK, (self.id >> 32) & MASK],
                          [self.seed & MASK, (self.seed >> 32) & MASK])
        h = self.n & 1
        self.n += 1
        return o[2 * h] | (o[2 * h + 1] << 32)

    def f64(self):  # Julia Random: CloseOpen12 from the low 52 bits, minus 1
        bits = 0x3FF0000000000000 | (self.u64() & 0x000FFFFFFFFFFFFF)
        return struct.unpack("<d", struct.pack("<Q", bits))[0] - 1.0

    def range1(self, s):  # rand(rng, 1:s), SamplerRangeNDL
        m = self.u64() * s
        lo = m & 0xFFFFFFFFFFFFFFFF
        if lo < s:
            t = ((1 << 64) - s) % s
            while lo < t:
                m = self.u64() * s
                lo = m & 0xFFFFFFFFFFFFFFFF
        return (m >> 64) + 1


OFF8 = [(-1, -1), (0, -1), (1, -1), (-1, 0), (1, 0), (-1, 1), (0, 1), (1, 1)]


def mod1(a, n):
    return (a - 1) % n + 1


class Model:
    def __init__(self, nx, ny, periodic, walk_type, seed=321, eat=True, reproduce=True, walkmap_regrowth=False,
                 regrowth_time=30, visual_distance=5, counter=50, prob_random_walk=0.9, delta_energy=4.0,
                 movement_cost=1.0, reproduction_prob=0.001, attractor=None, ifempty=True, astar_metric=0, astar_penalty=0, scheduler=0):
        self.nx, self.ny, self.periodic, self.walk_type = nx, ny, bool(periodic), walk_type
        self.eat, self.reproduce, self.walkmap_regrowth = eat, reproduce, walkmap_regrowth
        self.regrowth_time, self.r, self.counter = regrowth_time, visual_distance, counter
        self.prob, self.de, self.cost, self.rp = prob_random_walk, delta_energy, movement_cost, reproduction_prob
        self.attractor = attractor if attractor is not None else (nx / 2, ny / 2)
        self.ifempty = ifempty
        self.scheduler = scheduler  # 0 = Schedulers.by_id, 1 = Schedulers.randomly (include/abm.h, ABM_SCHED_RANDOMLY)
        self.seed = seed
        self.rng = Rng(seed)
        self.agents = {}   # id -> [x, y, energy]
        self.cells = {}    # (x, y) -> [ids] in insertion order
        self.grown = {}    # (x, y) -> bool
        self.countdown = {}
        self.walk = {}
        self.behav, self.behav_counter, self.maxid, self.tick = 0, 0, 0, 0
        self.metric, self.penalty = astar_metric, bool(astar_penalty)
        self.elev = {}
        self.paths = {}    # pathfinder.agent_paths :: Dict{Int, Path}

    # ---- set-up
    def set_grid(self, fg, cd, wm=None, el=None):
        for y in range(1, self.ny + 1):
            for x in range(1, self.nx + 1):
                self.grown[(x, y)] = bool(fg[y - 1][x - 1])
                self.countdown[(x, y)] = int(cd[y - 1][x - 1])
                self.walk[(x, y)] = True if wm is None else bool(wm[y - 1][x - 1])
                self.elev[(x, y)] = 0 if el is None else int(el[y - 1][x - 1])
                self.cells[(x, y)] = []

    def set_agents(self, ids, xs, ys, es):
        for i, x, y, e in sorted(zip(ids, xs, ys, es)):
            self.agents[int(i)] = [int(x), int(y), float(e)]
            self.cells[(int(x), int(y))].append(int(i))
        self.maxid = max([int(i) for i in ids], default=0)

    # ---- Agents.jl primitives
    def move(self, aid, pos):
        a = self.agents[aid]
        self.cells[(a[0], a[1])].remove(aid)
        a[0], a[1] = pos
        self.cells[pos].append(aid)

    def target(self, pos, d):
        x, y = pos[0] + d[0], pos[1] + d[1]
        if self.periodic:
            return mod1(x, self.nx), mod1(y, self.ny)
        return min(max(x, 1), self.nx), min(max(y, 1), self.ny)

    def walk_dir(self, aid, d, ifempty):
        a = self.agents[aid]
        t = self.target((a[0], a[1]), d)
        if not ifempty or not self.cells[t]:
            self.move(aid, t)

    def randomwalk(self, aid):
        self.walk_dir(aid, OFF8[self.rng.range1(8) - 1], self.ifempty)

    def nearby_positions(self, pos):
        out = []
        for d in OFF8:
            x, y = pos[0] + d[0], pos[1] + d[1]
            if self.periodic:
                out.append((mod1(x, self.nx), mod1(y, self.ny)))
            elif 1 <= x <= self.nx and 1 <= y <= self.ny:
                out.append((x, y))
        return out

    def eucl(self, a, b):
        dx, dy = abs(a[0] - b[0]), abs(a[1] - b[1])
        if self.periodic:
            dx, dy = min(dx, self.nx - dx), min(dy, self.ny - dy)
        return math.sqrt(dx * dx + dy * dy)

    # ---- rules
    def to_attractor(self, aid):
        a = self.agents[aid]
        dx, dy = self.attractor[0] - a[0], self.attractor[1] - a[1]
        if dx == 0 and dy == 0:
            d = (0, 0)
        else:
            inv = 1.0 / math.sqrt(dx * dx + dy * dy)  # LinearAlgebra.normalize: rmul!(a, inv(norm(a)))
            d = (round(dx * inv), round(dy * inv))  # Python round = ties-to-even, as Julia round(Int, x)
        self.walk_dir(aid, d, True)

    def gregarious(self, aid):
        a = self.agents[aid]
        pos = (a[0], a[1])
        possible = self.nearby_positions(pos)
        nb = []
        for by in range(-self.r, self.r + 1):
            for bx in range(-self.r, self.r + 1):
                x, y = pos[0] + bx, pos[1] + by
                if self.periodic:
                    x, y = mod1(x, self.nx), mod1(y, self.ny)
                elif not (1 <= x <= self.nx and 1 <= y <= self.ny):
                    continue
                nb += [j for j in self.cells[(x, y)] if j != aid]
        if nb:
            best, cpos = math.inf, None
            for j in nb:
                p = (self.agents[j][0], self.agents[j][1])
                d = self.eucl(p, pos)
                if d < best:
                    best, cpos = d, p
            w = [math.exp(-self.eucl(p, cpos) / 3) for p in possible]
            s1 = w[0]
            for v in w[1:]:
                s1 += v
            p = [v / s1 for v in w]
            s2 = p[0]
            for v in p[1:]:
                s2 += v
            t = self.rng.f64() * s2
            i, cw = 0, p[0]
            while cw < t and i + 1 < len(p):
                i += 1
                cw += p[i]
            self.move(aid, possible[i])
        else:
            self.move(aid, possible[self.rng.range1(8) - 1])

    def walkmap_walk(self, aid):
        a = self.agents[aid]
        ok = [p for p in self.nearby_positions((a[0], a[1])) if self.walk[p]]
        self.move(aid, ok[self.rng.range1(len(ok)) - 1])

    # ---- Agents.Pathfinding 5.14 (astar.jl / metrics.jl / pathfinding_utils.jl), written from the Julia semantics:
    # a heap of (f, pos) TUPLES, lazy insertion, strict improvement, stale entries re-expanded exactly as upstream does
    def delta_cost(self, a, b):
        dx, dy = abs(a[0] - b[0]), abs(a[1] - b[1])
        if self.periodic:
            dx, dy = min(dx, self.nx - dx), min(dy, self.ny - dy)
        if self.metric == 1:                       # MaxDistance
            c = max(dx, dy)
        else:                                      # DirectDistance, direction_costs = [10, 14]
            lo, hi = sorted((dx, dy))
            c = 14 * lo + 10 * (hi - lo)
        if self.penalty:                           # PenaltyMap(pmap, base): base + |pmap[from] - pmap[to]|
            c += abs(self.elev[a] - self.elev[b])
        return c

    def find_path(self, src, dst):
        import heapq
        INF = 2 ** 62
        g = {src: 0}
        parent = {}
        closed = set()
        heap = [(self.delta_cost(src, dst), src)]  # f = round(Int, g + (1 + admissibility) * h), admissibility = 0
        self.expansions = 0
        while heap:
            _, cur = heapq.heappop(heap)
            if cur == dst:
                break
            closed.add(cur)
            self.expansions += 1
            for d in OFF8:                         # Moore offsets in product order, x fastest
                x, y = cur[0] + d[0], cur[1] + d[1]
                if self.periodic:
                    nb = (mod1(x, self.nx), mod1(y, self.ny))
                elif 1 <= x <= self.nx and 1 <= y <= self.ny:
                    nb = (x, y)
                else:
                    continue
                if not self.walk[nb] or nb in closed:
                    continue
                ng = g[cur] + self.delta_cost(cur, nb)
                if ng < g.get(nb, INF):
                    parent[nb] = cur
                    g[nb] = ng
                    heapq.heappush(heap, (ng + self.delta_cost(nb, dst), nb))
        path, cur = [], dst
        while cur in parent:
            path.insert(0, cur)
            cur = parent[cur]
        if cur != src:
            return None, -1
        return path, g[dst]

    def plan_route(self, aid, dest):
        a = self.agents[aid]
        path, _ = self.find_path((a[0], a[1]), dest)
        if path is None:                           # isnothing(path) && return: an older route stays
            return
        self.paths[aid] = path

    def move_along_route(self, aid):
        path = self.paths.get(aid)
        if not path:
            return
        self.move(aid, path.pop(0))                # unconditional move_agent!, no ifempty

    def random_walkable(self):
        while True:                                # random_position(model) until the walkmap accepts it
            pos = (self.rng.range1(self.nx), self.rng.range1(self.ny))
            if self.walk[pos]:
                return pos

    def alternated(self, aid):
        if self.behav_counter == self.counter:
            self.behav = self.rng.range1(3)
            if self.behav == 2:
                self.plan_route(aid, self.random_walkable())
        if 0 < self.behav_counter <= self.counter:
            if self.behav == 1:
                self.randomwalk(aid)
            elif self.behav == 2:
                self.move_along_route(aid)
            elif self.behav == 3:
                a = self.agents[aid]
                self.move(aid, (a[0], a[1]))
            self.behav_counter -= 1
        if self.behav_counter == 0:
            self.behav_counter = self.counter

    def agent_step(self, aid):
        wt = self.walk_type
        if wt == 0:
            self.randomwalk(aid)
        else:
            prob = self.prob if wt in (1, 2) else 0
            if self.rng.f64() < prob:
                self.randomwalk(aid)
            else:
                {1: self.to_attractor, 2: self.gregarious, 3: self.alternated, 4: self.walkmap_walk}[wt](aid)
        a = self.agents[aid]
        a[2] -= self.cost
        if a[2] < 0:
            self.cells[(a[0], a[1])].remove(aid)
            del self.agents[aid]
        if self.eat and self.grown[(a[0], a[1])]:
            a[2] += self.de
            self.grown[(a[0], a[1])] = False
        if self.reproduce and self.rng.f64() <= self.rp:
            a[2] /= 2
            self.maxid += 1
            self.agents[self.maxid] = [a[0], a[1], a[2]]
            self.cells[(a[0], a[1])].append(self.maxid)

    def model_step(self):
        for p in self.grown:
            if (not self.walkmap_regrowth or self.walk[p]) and not self.grown[p]:
                if self.countdown[p] <= 0:
                    self.grown[p] = True
                    self.countdown[p] = self.regrowth_time
                else:
                    self.countdown[p] -= 1

    def schedule(self):
        """ids in the order the scheduler steps them this tick: sorted (by_id), or by the Feistel order key (randomly)."""
        ids = sorted(self.agents)
        if not self.scheduler:
            return ids
        hb = 1
        while hb < 16 and (1 << (2 * hb)) < self.maxid + 1:
            hb += 1
        k = philox4x32_10([0, self.tick & MASK, MASK, MASK], [self.seed & MASK, (self.seed >> 32) & MASK])
        m = (1 << hb) - 1

        def mix(z):
            z = (z * 0x9E3779B1) & MASK
            z ^= z >> 15
            z = (z * 0x85EBCA77) & MASK
            z ^= z >> 13
            return z

        def key(i):
            left, right = i >> hb, i & m
            for r in range(4):
                left, right = right, left ^ (mix(right ^ k[r]) & m)
            return (left << hb) | right

        return sorted(ids, key=key)

    def step(self, n=1):
        for _ in range(n):
            for aid in self.schedule():
                if aid in self.agents:
                    self.rng.reset(self.tick, aid)
                    self.agent_step(aid)
            self.model_step()
            self.tick += 1

    def snapshot(self):
        import numpy as np
        ids = np.array(sorted(self.agents), np.int64)
        fg = np.zeros((self.ny, self.nx), np.uint8)
        cd = np.zeros((self.ny, self.nx), np.int64)
        for (x, y), g in self.grown.items():
            fg[y - 1, x - 1] = g
            cd[y - 1, x - 1] = self.countdown[(x, y)]
        return dict(ids=ids, x=np.array([self.agents[i][0] for i in ids], np.int64),
                    y=np.array([self.agents[i][1] for i in ids], np.int64),
                    energy=np.array([self.agents[i][2] for i in ids], np.float64), fully_grown=fg, countdown=cd,
                    tick=self.tick, behav=self.behav, behav_counter=self.behav_counter, next_id=self.maxid + 1)
