"""TEST INFRASTRUCTURE: host restatement of the device-resident search (csrc/tree_kernel.cu), i.e. of the reference's
ParallelRussianDollMCTS (mcts.py:14-194) with this repo's stated substitutions (counter-based random draws, nodes
keyed by a hash of what a FEN holds, waves of `workers` simulations that select in worker order and back up together).
Plain Python ints / floats and dicts.  What it takes from the device is only what other tests already pin against
the oracle: the ordered legal successors with their categorize_moves weights (cdq_expand_weighted) and the leaf
values (cdq_leaf_eval); selection, sampling, Board.push bookkeeping (halfmove clock, irreversibility, repetition
counts, rep3 bit), hashing and backup are restated here."""
import math

import numpy as np
import torch

M64 = (1 << 64) - 1
GOLD = 0x9E3779B97F4A7C15
REP3 = 1 << 16


def mix64(z):
    z = (z + 0x9E3779B97F4A7C15) & M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def mulhi(z, n):
    return (int(z) * int(n)) >> 64


class Restatement:
    def __init__(self, cdq, net, workers, samples_per_level, exploration_weight=1.0, training_progress=0.0, max_moves=200):
        self.cdq, self.net = cdq, net
        self.ev = cdq.LeafEvaluator(net)
        self.W, self.D, self.width = workers, len(samples_per_level), samples_per_level[0]
        self.ew, self.alpha, self.max_moves = exploration_weight, 1.0 + training_progress, max_moves
        self._kids = {}

    # ------------------------------------------------------------------ device-provided, oracle-pinned ingredients
    def successors(self, rec):
        key = rec.tobytes()
        if key not in self._kids:
            cdq, L = self.cdq, self.cdq._lib.lib()
            d = cdq.board_utils.to_device(rec.reshape(1))
            kids = torch.empty((1, 224, 9), dtype=torch.int64, device=d.device)
            counts = torch.empty(1, dtype=torch.int32, device=d.device)
            wts = torch.empty((1, 224), dtype=torch.uint8, device=d.device)
            flags = torch.empty(1, dtype=torch.int32, device=d.device)
            st = cdq._lib.stream_ptr()
            cdq._lib.check(L.cdq_expand_weighted(d.data_ptr(), 1, 224, kids.data_ptr(), counts.data_ptr(), wts.data_ptr(), st))
            cdq._lib.check(L.cdq_eval_terms(d.data_ptr(), 1, 0, None, None, flags.data_ptr(), st))
            c = int(counts.item())
            recs = kids[0, :c].cpu().numpy().view(np.uint64).reshape(-1).view(cdq.BOARD_DTYPE).copy()
            self._kids[key] = (recs, [int(x) for x in wts[0, :c].cpu().numpy()], int(flags.item()))
        return self._kids[key]

    # ------------------------------------------------------------------ restated bookkeeping
    def ep_legal(self, rec):
        ep = (int(rec["meta"]) >> 8) & 127
        if ep <= 1:
            return False
        kids, _, _ = self.successors(rec)
        pawns = int(rec["bb"][0])
        for k in kids:
            f, t, _ = self.cdq.mcts.move_from_records(rec, k)
            if (pawns >> f) & 1 and t == ep - 1 and (f & 7) != (t & 7):
                return True
        return False

    def tkey(self, rec, ep_legal):
        h = 0x243F6A8885A308D3
        bb = [int(x) for x in rec["bb"]]
        for x in bb[:6] + [bb[6], bb[7]]:
            h = mix64(h ^ x)
        meta = int(rec["meta"])
        ep = ((meta >> 8) & 127) if ep_legal else 0
        return mix64(h ^ (meta & 0x1F) ^ (ep << 8))

    @staticmethod
    def node_key(tkey, hmc, ply):
        return mix64(tkey ^ mix64(((hmc & 0xFFFFFFFF) << 32) | (ply & 0xFFFFFFFF) | 0xA5A5000000000000))

    def push(self, parent, child, hmc, parent_ep_legal, combined_keys, valid_from):
        """-> (child record with its rep3 bit, hmc, valid_from, key, ep_legal, reps); combined_keys = keys of every
        position before the child (game history, root, path), in order."""
        pb, cb = [int(x) for x in parent["bb"]], [int(x) for x in child["bb"]]
        f, t, _ = self.cdq.mcts.move_from_records(parent, child)
        occ = pb[6] | pb[7]
        is_ep = (pb[0] >> f) & 1 and not (occ >> t) & 1 and (f & 7) != (t & 7)
        zeroing = bool((pb[0] >> f) & 1) or bool((occ >> t) & 1) or bool(is_ep)
        hmc2 = 0 if zeroing else hmc + 1
        irreversible = zeroing or ((int(parent["meta"]) ^ int(child["meta"])) & 0x1E) != 0 or parent_ep_legal
        epl = self.ep_legal(child)
        key = self.tkey(child, epl)
        cur = len(combined_keys)
        vf = cur if irreversible else valid_from
        reps = 1 + sum(1 for i in range(vf, cur) if combined_keys[i] == key)
        out = child.copy()
        out["meta"] = (int(child["meta"]) & ~REP3) | (REP3 if reps >= 3 else 0)
        return out, hmc2, vf, key, epl, reps

    def expand(self, rec, hmc, seed):
        kids, wts, flags = self.successors(rec)
        if len(kids) == 0 or (flags & 0x40) or hmc >= 150:
            return []
        n = len(kids)
        if n <= self.width:
            return list(range(n))
        w = list(wts)
        picked = []
        for k in range(self.width):
            z = mix64((seed + (k + 1) * GOLD) & M64)
            r = mulhi(z, sum(w))
            acc = 0
            for i in range(n):
                acc += w[i]
                if acc > r:
                    picked.append(i)
                    w[i] = 0
                    break
        return picked

    # ------------------------------------------------------------------ one search of one game
    def search(self, root, seed, iterations, hmc=0, ply=0, history=()):
        """-> dict(root_children (records), N, Q, best (record or None), node_count, leaves (records of the last wave),
        leaf_depth, evaluated)."""
        cdq = self.cdq
        root = root.copy()
        search_seed = mix64((seed + ply * GOLD) & M64)
        hist_keys = [self.tkey(h, self.ep_legal(h)) for h in history]
        root_epl = self.ep_legal(root)
        root_key = self.tkey(root, root_epl)
        nodes = {}                                     # node key -> dict(kids, N, Q, fresh)
        evaluated = 0
        last = None
        for wave, first in enumerate(range(0, iterations, self.W)):
            size = min(self.W, iterations - first)
            sims = [dict(rec=root.copy(), hmc=hmc, depth=0, vf=0, keys=[root_key], epl=root_epl, alive=True, path=[]) for _ in range(size)]
            for level in range(self.D):
                for w, s in enumerate(sims):           # worker order: expansion order and claims of fresh children
                    if not s["alive"]:
                        continue
                    nk = self.node_key(s["keys"][s["depth"]], s["hmc"], ply + s["depth"])
                    if nk not in nodes:
                        recs = self.successors(s["rec"])[0]
                        idx = self.expand(s["rec"], s["hmc"], mix64(search_seed ^ nk))
                        nodes[nk] = dict(kids=[recs[i] for i in idx], N=[0] * len(idx), Q=[0.0] * len(idx), fresh=list(range(len(idx))))
                    s["node"] = nk
                for w, s in enumerate(sims):
                    if not s["alive"]:
                        continue
                    nd = nodes[s["node"]]
                    if not nd["kids"]:
                        s["alive"] = False
                        continue
                    z = mix64(search_seed ^ mix64((wave << 32) | (w << 8) | s["depth"] | 0x5E1EC70000000000))
                    if nd["fresh"]:
                        c = nd["fresh"].pop(mulhi(z, len(nd["fresh"])))          # fresh stays sorted: the k-th set bit
                    else:
                        total = sum(nd["N"])
                        if total == 0:
                            c = mulhi(z, len(nd["kids"]))
                        else:
                            beta = (ply + s["depth"]) / self.max_moves if self.max_moves > 0 else 0.5
                            eff = self.ew * (1.0 - 0.5 * self.alpha * beta)
                            lt = math.log(total + 1e-10)
                            ucb = [q / (n + 1e-10) + eff * math.sqrt(lt / (n + 1e-10)) for q, n in zip(nd["Q"], nd["N"])]
                            c = int(np.argmax(ucb))
                    s["path"].append((s["node"], c))
                    s["move"] = c
                for s in sims:
                    if not s["alive"] or "move" not in s:
                        continue
                    nd = nodes[s["node"]]
                    child = nd["kids"][s.pop("move")]
                    rec, h2, vf, key, epl, reps = self.push(s["rec"], child, s["hmc"], s["epl"], hist_keys + s["keys"], s["vf"])
                    s["rec"], s["hmc"], s["vf"], s["epl"] = rec, h2, vf, epl
                    s["keys"].append(key)
                    s["depth"] += 1
                    s["alive"] = reps < 5 and s["depth"] < self.D
            leaves = np.array([s["rec"] for s in sims], dtype=cdq.BOARD_DTYPE)
            values = self.ev.evaluate_device(leaves)["leaf"].cpu().numpy()
            evaluated += size
            for s, v in zip(sims, values):
                value = float(v)
                for nk, c in reversed(s["path"]):
                    nd = nodes[nk]
                    nd["N"][c] += 1
                    nd["Q"][c] += (value - nd["Q"][c]) / nd["N"][c]
                    value = -value
            last = (leaves, [s["depth"] for s in sims], values)
        rn = nodes[self.node_key(root_key, hmc, ply)]
        best = None
        if rn["kids"]:
            b = int(np.argmax(rn["N"])) if max(rn["N"]) > 0 else mulhi(mix64(search_seed ^ 0xBE57C401CE), len(rn["kids"]))
            best = rn["kids"][b]
        return dict(root_children=rn["kids"], N=rn["N"], Q=rn["Q"], best=best, node_count=len(nodes), leaves=last[0],
                    leaf_depth=last[1], leaf_values=last[2], evaluated=evaluated)
