"""Host-side executable specification of MSSPE_SELECT_PARTITIONED (csrc/select_part.cu).  TEST INFRASTRUCTURE.

The greedy loop of od-msspe/src/main.rs:331-406 factorises over partitions.  A k-mer whose postings all lie in ONE
partition p ("single-partition list" -- the rule in a pre-aligned alignment) only ever changes the live counts of k-mers
of partition-p segments, and its partition_tie_score (main.rs:261-283) is 0.0 + 1/(partition_coverage[p] + 1) whichever
of its postings are live.  So every partition runs its OWN greedy sequence (key: frequency, then smaller word),
independently of all the others ("unit" = one partition of one direction), and the reference's global selection order
is the merge of those sequences by (frequency desc, partition_coverage asc, word asc), cut by the stop rules of
main.rs:353-390.  Lists that span several partitions couple units: they are never candidates inside a unit; after
every merge each of them is checked against the merged order ("would it have beaten the winner of iteration t?", with
the live count of iteration t taken from the cover times of its postings and the exact sequential f32 score).  The
earliest such iteration t* is an EXTERNAL winner: everything before t* is final, the units it touches are rolled back
to their state at t* and re-run, the rest is re-merged.  This file is the algorithm in plain Python, step for step as
the CUDA kernels do it (rounds of extend -> merge -> verify -> finalise), so that tests/test_partitioned_model.py can
prove the decomposition against the oracle on the CPU, including degenerate inputs where every list is multi-partition.
"""
from __future__ import annotations

import struct

import numpy as np

NO = np.uint64(0xFFFFFFFFFFFFFFFF)
INF = 1 << 60


def f32(x):
    return struct.unpack("f", struct.pack("f", x))[0]


def f32_bits(x):
    return struct.unpack("I", struct.pack("f", x))[0]


class PartitionedSelect:
    def __init__(self, slots: np.ndarray, part: np.ndarray, max_iter: int, mms: int, chunk0: int | None = None, chunk: int = 4):
        G, s = slots.shape if slots.ndim == 2 else (0, 0)
        self.G, self.s, self.max_iter, self.mms = G, s, max_iter, mms
        self.part = [int(p) for p in part[:G]]
        valid = slots != NO
        seg_ids = np.repeat(np.arange(G, dtype=np.int64), s).reshape(G, s) if G else np.zeros((0, 0), np.int64)
        cf, sf = slots[valid], seg_ids[valid]
        order = np.lexsort((sf, cf))
        cs, post = cf[order], sf[order]
        self.codes, start = np.unique(cs, return_index=True)
        self.off = [int(x) for x in start] + [len(post)]
        self.post = [int(x) for x in post]
        D = len(self.codes)
        self.fwd = np.full((G, s), -1, np.int64)
        if G:
            self.fwd[valid] = np.searchsorted(self.codes, slots[valid])
        self.U = (max(self.part) + 1) if G else 0
        self.list_part = []
        for c in range(D):
            ps = {self.part[g] for g in self.post[self.off[c]:self.off[c + 1]]}
            self.list_part.append(ps.pop() if len(ps) == 1 else -1)
        self.unit_codes = [[] for _ in range(self.U)]
        self.multi = []
        for c in range(D):
            (self.unit_codes[self.list_part[c]] if self.list_part[c] >= 0 else self.multi).append(c)
        self.usegs = [[] for _ in range(self.U)]
        for g in range(G):
            self.usegs[self.part[g]].append(g)
        # state
        self.token = [None] * G                 # None = live; ('e', u, r) covered by entry r of unit u; ('x', j) by external winner j
        self.pfreq = [0] * D
        self.ulive = [0] * self.U
        self.entries = [[] for _ in range(self.U)]   # dict(freq, cid, tied, live_before, pos)
        self.rfin = [0] * self.U
        self.finished = [False] * self.U
        self.need_recount = [True] * self.U
        self.want_extend = [True] * self.U
        self.ext_cov = [0] * self.U
        self.ext_time = []
        self.out = []                            # final winners: (code, freq, n_tied, score)
        self.t_final = 0
        self.evals = 0
        self.iterations = 0
        self.rounds = 0
        self.rollbacks = 0
        self.chunk0 = chunk0 if chunk0 is not None else max(4, -(-3 * max_iter // (2 * max(1, self.U))))
        self.chunk = chunk

    # ---- unit kernel -----------------------------------------------------------------------------------------------
    def extend(self, u, nsteps):
        if self.need_recount[u]:
            for g in self.usegs[u]:
                tk = self.token[g]
                if tk is not None and tk[0] == "e" and tk[2] >= self.rfin[u]:
                    self.token[g] = None
            for c in self.unit_codes[u]:
                self.pfreq[c] = 0
            self.ulive[u] = 0
            for g in self.usegs[u]:
                if self.token[g] is None:
                    for c in self.fwd[g]:
                        if c >= 0:
                            self.ulive[u] += 1
                            if self.list_part[c] >= 0:
                                self.pfreq[c] += 1
            self.need_recount[u] = False
        for _ in range(nsteps):
            if self.finished[u]:
                break
            best_f, best_c = 0, -1
            for c in self.unit_codes[u]:
                if self.pfreq[c] > best_f:
                    best_f, best_c = self.pfreq[c], c
            if best_f < 2:                      # freq == 1 stops before the push (main.rs:354-360); 0 = None
                self.finished[u] = True
                break
            tied = sum(1 for c in self.unit_codes[u] if self.pfreq[c] == best_f)
            r = len(self.entries[u])
            self.entries[u].append(dict(freq=best_f, cid=best_c, tied=tied, live_before=self.ulive[u], pos=INF))
            for g in self.post[self.off[best_c]:self.off[best_c + 1]]:
                if self.token[g] is None:
                    self.token[g] = ("e", u, r)
                    for c in self.fwd[g]:
                        if c >= 0:
                            self.ulive[u] -= 1
                            if self.list_part[c] >= 0:
                                self.pfreq[c] -= 1
            if best_f < self.mms:               # pushed, then break (main.rs:387-390)
                self.finished[u] = True
                break

    # ---- helpers ---------------------------------------------------------------------------------------------------
    def time_of(self, g):
        tk = self.token[g]
        if tk is None:
            return INF
        if tk[0] == "x":
            return self.ext_time[tk[1]]
        return self.entries[tk[1]][tk[2]]["pos"]

    def cov_at(self, p, t):
        return self.ext_cov[p] + sum(1 for e in self.entries[p] if e["pos"] < t)

    def unit_state_at(self, q, t):
        """(head entry or None, live records) of unit q at global iteration t."""
        for e in self.entries[q][self.rfin[q]:]:
            if e["pos"] >= t:
                return e, e["live_before"]
        return None, self.ulive[q]

    def multi_score(self, c, t):
        seen, score = set(), f32(0.0)
        for g in self.post[self.off[c]:self.off[c + 1]]:
            if self.time_of(g) < t:
                continue
            p = self.part[g]
            if p not in seen:
                seen.add(p)
                score = f32(score + f32(1.0 / f32(f32(self.cov_at(p, t)) + 1.0)))
        return score

    # ---- one round -------------------------------------------------------------------------------------------------
    def round(self):
        self.rounds += 1
        first = self.rounds == 1
        for u in range(self.U):
            if self.want_extend[u] and not self.finished[u] or self.need_recount[u]:
                self.extend(u, self.chunk0 if first else self.chunk)
            self.want_extend[u] = False
        # merge
        nonfinal = [(u, r) for u in range(self.U) for r in range(self.rfin[u], len(self.entries[u]))]
        key = lambda ur: (-self.entries[ur[0]][ur[1]]["freq"], ur[1] + self.ext_cov[ur[0]], self.entries[ur[0]][ur[1]]["cid"])  # noqa: E731
        nonfinal.sort(key=key)
        for i, (u, r) in enumerate(nonfinal):
            self.entries[u][r]["pos"] = self.t_final + i
        H = INF
        for u in range(self.U):
            if not self.finished[u]:
                last = self.entries[u][-1]["pos"] + 1 if len(self.entries[u]) > self.rfin[u] else self.t_final
                H = min(H, last)
        cutbound, terminal = self.max_iter, False
        for i, (u, r) in enumerate(nonfinal):
            if self.entries[u][r]["freq"] < self.mms:
                cutbound = min(cutbound, self.t_final + i + 1)
                break
        else:
            if self.t_final + len(nonfinal) < self.max_iter:
                cutbound, terminal = self.t_final + len(nonfinal), True      # valid only if every unit is finished (H == INF)
        V = min(H, cutbound)
        by_pos = {self.t_final + i: ur for i, ur in enumerate(nonfinal)}
        # verify the multi-partition lists over [t_final, V) (+ the terminal iteration)
        do_terminal = terminal and H == INF
        t_hi = V + (1 if do_terminal else 0)
        viol = None                                # (t, cnt, score, -cid)
        multi_ties = {}
        for c in self.multi:
            times = [self.time_of(g) for g in self.post[self.off[c]:self.off[c + 1]]]
            for t in range(self.t_final, t_hi):
                if viol is not None and t > viol[0]:
                    break
                cnt = sum(1 for x in times if x >= t)
                if t == V and do_terminal:
                    beats = cnt >= 2
                    sc = self.multi_score(c, t) if beats else None
                else:
                    u, r = by_pos[t]
                    e = self.entries[u][r]
                    if cnt < e["freq"]:
                        continue
                    sc = self.multi_score(c, t)
                    wsc = f32(1.0 / f32(f32(r + self.ext_cov[u]) + 1.0))
                    if cnt == e["freq"]:
                        multi_ties[t] = multi_ties.get(t, 0) + 1
                    beats = cnt > e["freq"] or sc > wsc or (sc == wsc and c < e["cid"])
                if beats:
                    cand = (t, cnt, sc, -c)
                    if viol is None or t < viol[0] or (t == viol[0] and cand[1:] > viol[1:]):
                        viol = cand
                    break
        t_new = viol[0] if viol else V
        # everything before t_new becomes final
        for t in range(self.t_final, t_new + (1 if viol else 0)):
            heads = [self.unit_state_at(q, t) for q in range(self.U)]
            self.evals += sum(h[1] for h in heads)
            self.iterations += 1
            if t < t_new:
                u, r = by_pos[t]
                e = self.entries[u][r]
                tied = sum(h[0]["tied"] for h in heads if h[0] is not None and h[0]["freq"] == e["freq"]) + multi_ties.get(t, 0)
                self.out.append((int(self.codes[e["cid"]]), e["freq"], tied, f32(1.0 / f32(f32(r + self.ext_cov[u]) + 1.0))))
            else:                               # the external winner's iteration
                _, cnt, sc, negc = viol
                c = -negc
                tied = sum(h[0]["tied"] for h in heads if h[0] is not None and h[0]["freq"] == cnt)
                for c2 in self.multi:
                    if sum(1 for g in self.post[self.off[c2]:self.off[c2 + 1]] if self.time_of(g) >= t) == cnt:
                        tied += 1
                self.out.append((int(self.codes[c]), cnt, tied, sc))
        for u in range(self.U):
            while self.rfin[u] < len(self.entries[u]) and self.entries[u][self.rfin[u]]["pos"] < t_new:
                self.rfin[u] += 1
        if viol:
            self.rollbacks += 1
            t, cnt, sc, negc = viol
            c = -negc
            j = len(self.ext_time)
            self.ext_time.append(t)
            touched, affected = set(), set()
            for g in self.post[self.off[c]:self.off[c + 1]]:
                touched.add(self.part[g])
                if self.time_of(g) >= t:
                    self.token[g] = ("x", j)
                    affected.add(self.part[g])
            for p in touched:
                self.ext_cov[p] += 1
            for u in affected:
                del self.entries[u][self.rfin[u]:]
                self.need_recount[u] = True
                self.finished[u] = False
                self.want_extend[u] = True
            self.t_final = t + 1
            if cnt < self.mms or self.t_final >= self.max_iter:
                return True
            return False
        self.t_final = t_new
        if V == cutbound and (not terminal or H == INF):
            if do_terminal:                     # the call that found freq == 1 / None still counted its evals
                self.evals += sum(self.ulive)
                self.iterations += 1
            return True
        for u in range(self.U):
            if not self.finished[u]:
                last = self.entries[u][-1]["pos"] + 1 if len(self.entries[u]) > self.rfin[u] else self.t_final
                if last < (self.max_iter if terminal else cutbound):   # a terminal "cut" is only the end of what is known
                    self.want_extend[u] = True
        return False

    def run(self):
        if self.max_iter == 0:
            return self
        while not self.round():
            assert self.rounds < 100000
        return self


def select(slots, part, max_iter, mms, **kw):
    m = PartitionedSelect(slots, part, max_iter, mms, **kw).run()
    return dict(codes=[o[0] for o in m.out], freqs=[o[1] for o in m.out], n_tied=[o[2] for o in m.out],
                score_bits=[f32_bits(o[3]) for o in m.out], evals=m.evals, iterations=m.iterations, rounds=m.rounds,
                rollbacks=m.rollbacks)


# ---- the multi-GPU loop (csrc/select_dist.cu) -----------------------------------------------------------------------------
class DistributedPartitionedSelect:
    """One rank of the column-sharded loop.  `comm.all_gather(obj)` returns the list of every rank's obj (rank order); all
    decisions are taken on gathered data, identically on every rank -- the protocol of select_dist.cu: (1) all-gather of
    the units' not-yet-final entries, replicated merge; (2) per round one exchange carrying each rank's best local
    external winner, the cover-time histograms of the cross-rank lists and, per partition of such a list, its last cover
    time (+ the first live genome at a queried iteration: the f32 tie score is order-sensitive from three partitions on)."""

    def __init__(self, slots, part, n_part_local, comm, rank, world, max_iter, mms, kmax=8, wmax=16, chunk0=4, chunk=2):
        self.L = PartitionedSelect(slots, part, max_iter, mms)
        self.comm, self.rank, self.world, self.max_iter, self.mms = comm, rank, world, max_iter, mms
        self.kmax, self.wmax_cap, self.wmax, self.chunk0, self.chunk = kmax, wmax, wmax, chunk0, chunk
        L = self.L
        self.n_part_local = n_part_local
        self.P = L.U                                            # partitions (uniform alignment: genome = segment // P)
        sizes = comm.all_gather(L.U)
        self.UM = max(max(sizes), 1)
        self.U_pad = self.UM * world
        # words that occur on several ranks leave the units' candidate lists
        all_codes = comm.all_gather([int(c) for c in L.codes])
        seen = {}
        for r, cs in enumerate(all_codes):
            for c in cs:
                seen[c] = seen.get(c, 0) + 1
        self.cross = sorted(c for c, n in seen.items() if n >= 2)
        self.xlocal = {}
        code_to_cid = {int(c): i for i, c in enumerate(L.codes)}
        for code in self.cross:
            if code in code_to_cid:
                cid = code_to_cid[code]
                self.xlocal[code] = cid
                if L.list_part[cid] >= 0:
                    L.unit_codes[L.list_part[cid]].remove(cid)
                    L.list_part[cid] = -1
                    L.multi.append(cid)
        self.local_multi = [c for c in L.multi if int(L.codes[c]) not in self.xlocal]
        self.xparts = {}
        for code in self.cross:
            mine = {self.gu(L.part[g]) for g in L.post[L.off[self.xlocal[code]]:L.off[self.xlocal[code] + 1]]} if code in self.xlocal else set()
            self.xparts[code] = set().union(*comm.all_gather(mine))
        self.ext_cov = [0] * self.U_pad                        # replicated
        self.t_final, self.tq = 0, None
        self.out, self.evals, self.iterations, self.rounds, self.rollbacks, self.asks = [], 0, 0, 0, 0, 0

    def gu(self, u_local):
        return self.rank * self.UM + u_local

    def cov_view(self, gu, t):
        e = self.table[gu]
        return self.ext_cov[gu] + e["rfin"] + sum(1 for p in e["pos"] if p < t)

    def round(self):
        L = self.L
        self.rounds += 1
        for u in range(L.U):
            if (L.want_extend[u] and not L.finished[u]) or L.need_recount[u]:
                n = self.chunk0 if self.rounds == 1 else self.chunk
                room = self.kmax - (len(L.entries[u]) - L.rfin[u]) if not L.need_recount[u] else self.kmax
                L.extend(u, max(0, min(n, room)))
            L.want_extend[u] = False
        # (1) the units' not-yet-final entries -> replicated table
        pack = []
        for u in range(self.UM):
            if u < L.U:
                ents = [(e["freq"], int(L.codes[e["cid"]]), e["tied"], e["live_before"]) for e in L.entries[u][L.rfin[u]:]]
                pack.append(dict(rfin=L.rfin[u], ents=ents, finished=L.finished[u], ulive=L.ulive[u], pos=[]))
            else:
                pack.append(dict(rfin=0, ents=[], finished=True, ulive=0, pos=[]))
        self.table = [e for blk in self.comm.all_gather(pack) for e in blk]
        nonfinal = [(gu, i) for gu in range(self.U_pad) for i in range(len(self.table[gu]["ents"]))]
        key = lambda x: (-self.table[x[0]]["ents"][x[1]][0], self.table[x[0]]["rfin"] + x[1] + self.ext_cov[x[0]], self.table[x[0]]["ents"][x[1]][1])  # noqa: E731
        nonfinal.sort(key=key)
        for gu in range(self.U_pad):
            self.table[gu]["pos"] = [None] * len(self.table[gu]["ents"])
        for i, (gu, idx) in enumerate(nonfinal):
            self.table[gu]["pos"][idx] = self.t_final + i
        for u in range(L.U):                                    # positions back into this rank's entries (cover tokens refer to them)
            for idx, p in enumerate(self.table[self.gu(u)]["pos"]):
                L.entries[u][L.rfin[u] + idx]["pos"] = p
        H = INF
        for gu in range(self.U_pad):
            e = self.table[gu]
            if not e["finished"]:
                H = min(H, e["pos"][-1] + 1 if e["pos"] else self.t_final)
        cutbound, terminal = self.max_iter, False
        for i, (gu, idx) in enumerate(nonfinal):
            if self.table[gu]["ents"][idx][0] < self.mms:
                cutbound = min(cutbound, self.t_final + i + 1)
                break
        else:
            if self.t_final + len(nonfinal) < self.max_iter:
                cutbound, terminal = self.t_final + len(nonfinal), True
        V = min(H, cutbound)
        clipped = V - self.t_final > self.wmax
        if clipped:
            V = self.t_final + self.wmax
        do_terminal = terminal and H == INF and not clipped
        t_hi = V + (1 if do_terminal else 0)
        win = {}
        for i in range(V - self.t_final):
            gu, idx = nonfinal[i]
            f, code, tied, lb = self.table[gu]["ents"][idx]
            win[self.t_final + i] = (f, self.table[gu]["rfin"] + idx + self.ext_cov[gu], code)

        def heads_at(t):
            tot, tied_by_f = 0, {}
            for gu in range(self.U_pad):
                e = self.table[gu]
                for idx, p in enumerate(e["pos"]):
                    if p >= t:
                        tot += e["ents"][idx][3]
                        tied_by_f[e["ents"][idx][0]] = tied_by_f.get(e["ents"][idx][0], 0) + e["ents"][idx][2]
                        break
                else:
                    tot += e["ulive"]
            return tot, tied_by_f

        # (2a) this rank's lists that lie inside it: exactly the one-GPU check
        L.t_final = self.t_final
        local_best, mt_local, local_viol = None, {}, []
        for c in self.local_multi:
            times = [L.time_of(g) for g in L.post[L.off[c]:L.off[c + 1]]]
            for t in range(self.t_final, t_hi):
                if local_best is not None and t > local_best[0]:
                    break
                cnt = sum(1 for x in times if x >= t)
                if t == V and do_terminal:
                    beats = cnt >= 2
                else:
                    f, cov, wcode = win[t]
                    if cnt < f:
                        continue
                    if cnt == f:
                        mt_local[t] = mt_local.get(t, 0) + 1
                sc = self._local_score(c, t)
                if not (t == V and do_terminal):
                    wsc = f32(1.0 / f32(f32(cov) + 1.0))
                    beats = cnt > f or sc > wsc or (sc == wsc and int(L.codes[c]) < wcode)
                if beats:
                    cand = (t, cnt, sc, -int(L.codes[c]))
                    local_viol.append(cand)
                    if local_best is None or t < local_best[0] or (t == local_best[0] and cand[1:] > local_best[1:]):
                        local_best = cand
                    break
        n_same = sum(1 for v in local_viol if local_best is not None and v[0] == local_best[0] and v[1] == local_best[1])
        lb_parts = None
        if local_best is not None:
            cid = {int(L.codes[c]): c for c in self.local_multi}[-local_best[3]]
            lb_parts = {self.gu(L.part[g]) for g in L.post[L.off[cid]:L.off[cid + 1]]}
        # (2b) cross-rank lists: local histogram, live count, per-partition last cover time (+ first live genome at tq)
        xinfo = {}
        for code, cid in self.xlocal.items():
            hist, l0, parts = {}, 0, {}
            for g in L.post[L.off[cid]:L.off[cid + 1]]:
                tm = L.time_of(g)
                if tm >= self.t_final:
                    l0 += 1
                    if tm < t_hi:
                        hist[tm] = hist.get(tm, 0) + 1
                pr = parts.setdefault(self.gu(L.part[g]), [0, INF])
                pr[0] = max(pr[0], tm)
                if self.tq is not None and tm >= self.tq:
                    pr[1] = min(pr[1], g // self.P)
            xinfo[code] = (l0, hist, parts)
        gathered = self.comm.all_gather((local_best, (lb_parts, n_same), xinfo, mt_local))
        # decide, identically everywhere
        mt, cands, tq_next = {}, [], None
        for lb, lbp, xi, mtl in gathered:
            for t, n in mtl.items():
                mt[t] = mt.get(t, 0) + n
            if lb is not None:
                cands.append((lb[0], lb[1], lb[2], lb[3], ("local", lbp[0]), lbp[1]))
        for code in self.cross:
            l0, hist, parts = 0, {}, {}
            for lb, lbp, xi, mtl in gathered:
                if code in xi:
                    l0 += xi[code][0]
                    for t, n in xi[code][1].items():
                        hist[t] = hist.get(t, 0) + n
                    parts.update(xi[code][2])
            covered = 0
            for t in range(self.t_final, t_hi):
                cnt = l0 - covered
                covered += hist.get(t, 0)
                term = t == V and do_terminal
                f = 1 if term else win[t][0]
                if cnt < f or (term and cnt < 2):
                    continue
                strict = term or cnt > f
                if not strict:
                    mt[t] = mt.get(t, 0) + 1
                live = [(gu, v[1]) for gu, v in parts.items() if v[0] >= t]
                if len(live) > 2:
                    if t != self.tq:
                        tq_next = t if tq_next is None else min(tq_next, t)
                        break
                    live.sort(key=lambda x: (x[1], x[0]))
                sc = f32(0.0)
                for gu, _ in live:
                    sc = f32(sc + f32(1.0 / f32(f32(self.cov_view(gu, t)) + 1.0)))
                wins = strict
                if not wins:
                    wsc = f32(1.0 / f32(f32(win[t][1]) + 1.0))
                    wins = sc > wsc or (sc == wsc and code < win[t][2])
                if wins:
                    cands.append((t, cnt, sc, -code, ("cross", code), 1))
                    break
        tv = min((c[0] for c in cands), default=None)
        ask = tq_next is not None and (tv is None or tq_next <= tv)
        if ask:
            tv = None
        t_new = tq_next if ask else (tv if tv is not None else V)
        for t in range(self.t_final, t_new + (1 if tv is not None else 0)):
            tot, tied_by_f = heads_at(t)
            self.evals += tot
            self.iterations += 1
            if t < t_new:
                f, cov, code = win[t]
                self.out.append((code, f, tied_by_f.get(f, 0) + mt.get(t, 0), f32(1.0 / f32(f32(cov) + 1.0))))
            else:
                at = [c for c in cands if c[0] == tv]
                best = max(at, key=lambda c: (c[1], c[2], c[3]))
                cnt = best[1]
                tie_case = t in win and win[t][0] == cnt
                n_tied = tied_by_f.get(cnt, 0) + mt.get(t, 0) if tie_case else sum(c[5] for c in at if c[1] == cnt)
                self.out.append((-best[3], cnt, n_tied, best[2]))
        for u in range(L.U):
            while L.rfin[u] < len(L.entries[u]) and L.entries[u][L.rfin[u]]["pos"] < t_new:
                L.rfin[u] += 1
        self.tq = tq_next if ask else None
        if tv is not None:
            self.rollbacks += 1
            kind, what = best[4]
            touched = what if kind == "local" else self.xparts[what]
            for gu in touched:
                self.ext_cov[gu] += 1
                if gu // self.UM == self.rank:
                    L.ext_cov[gu - self.rank * self.UM] += 1
            code = -best[3]
            cid = {int(L.codes[c]): c for c in L.multi}.get(code)
            j = len(L.ext_time)
            L.ext_time.append(tv)
            if cid is not None:
                affected = set()
                for g in L.post[L.off[cid]:L.off[cid + 1]]:
                    if L.time_of(g) >= tv:
                        L.token[g] = ("x", j)
                        affected.add(L.part[g])
                for u in affected:
                    del L.entries[u][L.rfin[u]:]
                    L.need_recount[u] = True
                    L.finished[u] = False
                    L.want_extend[u] = True
            self.t_final = tv + 1
            self.wmax = self.wmax_cap
            return best[1] < self.mms or self.t_final >= self.max_iter
        self.t_final = t_new
        if ask:
            self.asks += 1
            return False
        done = V == cutbound and (not terminal or H == INF)
        if done:
            if do_terminal:
                tot, _ = heads_at(t_new)
                self.evals += tot
                self.iterations += 1
            return True
        if not clipped:
            for u in range(L.U):
                if not L.finished[u]:
                    last = L.entries[u][-1]["pos"] + 1 if len(L.entries[u]) > L.rfin[u] else self.t_final
                    if last < (self.max_iter if terminal else cutbound):
                        L.want_extend[u] = True
        return False

    def _local_score(self, c, t):
        L = self.L
        seen, score = set(), f32(0.0)
        for g in L.post[L.off[c]:L.off[c + 1]]:
            if L.time_of(g) < t:
                continue
            p = L.part[g]
            if p not in seen:
                seen.add(p)
                score = f32(score + f32(1.0 / f32(f32(self.cov_view(self.gu(p), t)) + 1.0)))
        return score

    def run(self):
        if self.max_iter == 0:
            return self
        while not self.round():
            assert self.rounds < 100000
        return self
