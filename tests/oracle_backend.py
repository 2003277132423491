"""A CPU stand-in for metacoag_b200._backend.GpuBackend, built on the pinned oracle.

TEST INFRASTRUCTURE: it lets the host logic of the drop-in modules (marker-gene matching,
heap loops, dict staging) run in the `-m "not gpu"` suite.  The package itself never
imports it -- without a GPU the product raises.
"""
import numpy as np

from metacoag_b200 import _backend
from oracle import oracle


def ref_bfs(start, limit, labels, indptr, indices):
    """label_prop_utils.py:24-100 traversal restated on CSR, in pop order."""
    queue, visited, depth, hits = [start], set(), {start: 0}, []
    while queue:
        v = queue.pop(0)
        visited.add(v)
        if labels[v] >= 0 and len(visited) > 1:
            hits.append((v, int(labels[v]), depth[v]))
        else:
            for nb in indices[indptr[v]:indptr[v + 1]]:
                nb = int(nb)
                if nb not in visited:
                    depth[nb] = depth[v] + 1
                    if depth[nb] > limit:
                        continue
                    queue.append(nb)
    return hits


class OracleBackend:
    def _tables(self, profiles, coverages):
        prof, cov, have = _backend.dense_tables(profiles, coverages)
        self._have = have
        return prof, cov

    def reset(self):
        pass

    def tetramer_scan(self, blob, offsets, packed=None):
        counts, freqs = oracle.count_kmers_batch(blob, offsets)
        return counts.astype(np.uint32), freqs

    def bin_profiles(self, profiles, coverages, members, seg_offsets):
        prof, cov = self._tables(profiles, coverages)
        return (oracle.segment_mean(prof, members, seg_offsets),
                oracle.segment_mean(cov, members, seg_offsets))

    def score_members(self, profiles, coverages, queries, members, seg_offsets, rule):
        prof, cov = self._tables(profiles, coverages)
        for c in list(queries) + list(members):
            if not self._have[c]:
                raise KeyError(int(c))
        return oracle.score_members(prof, cov, queries, members, seg_offsets, rule)

    def score_member_items(self, profiles, coverages, item_query, item_bin, members, seg_offsets,
                           rule):
        queries = sorted(set(int(q) for q in item_query))
        row = {q: i for i, q in enumerate(queries)}
        dense = self.score_members(profiles, coverages, np.asarray(queries, dtype=np.int64),
                                   members, seg_offsets, rule)
        return np.array([dense[row[int(q)], int(b)] for q, b in zip(item_query, item_bin)])

    def frontier(self, indptr, indices, version, labels):
        self._graph = (np.asarray(indptr), np.asarray(indices))
        self._labels = np.array(labels, dtype=np.int32)

    def frontier_patch(self, ids, bins):
        for i, b in zip(ids, bins):
            self._labels[int(i)] = int(b)

    def frontier_walks(self, starts, depth_limit):
        indptr, indices = self._graph
        lists = [ref_bfs(int(s), depth_limit, self._labels, indptr, indices) for s in starts]
        count = np.array([len(h) for h in lists], dtype=np.int32)
        first = np.concatenate([[0], np.cumsum(count)[:-1]]).astype(np.int64)
        flat = [t for h in lists for t in h]
        return first, count, np.array(flat, dtype=np.int32).reshape(-1, 3)

    def score_centroids(self, profiles, coverages, queries, cen_prof, cen_cov):
        prof, cov = self._tables(profiles, coverages)
        arg, w, _ = oracle.score_centroids(prof, cov, cen_prof, cen_cov, queries=queries)
        return arg, w

    def pair_matrix(self, prof, cov):
        n = prof.shape[0]
        pc = np.zeros((n, n))
        lp = np.zeros((n, n))
        for i in range(n):
            for j in range(n):
                pc[i, j] = oracle.comp_prob(oracle.euclidean(prof[i], prof[j]))
                lp[i, j] = oracle.pair_lp(prof[i], prof[j], cov[i], cov[j])[1]
        return pc, lp

    def bfs_candidates(self, indptr, indices, labels, starts, depth_limit):
        hits = [ref_bfs(int(s), depth_limit, labels, indptr, indices) for s in starts]
        width = max([len(h) for h in hits] + [1])
        node = np.zeros((len(hits), width), dtype=np.int32)
        bin_ = np.zeros_like(node)
        depth = np.zeros_like(node)
        for i, h in enumerate(hits):
            for j, (v, b, d) in enumerate(h):
                node[i, j], bin_[i, j], depth[i, j] = v, b, d
        return node, bin_, depth, np.array([len(h) for h in hits], dtype=np.int32)

    def connected_components(self, indptr, indices):
        n = len(indptr) - 1
        labels = np.full(n, -1, dtype=np.int32)
        for s in range(n):
            if labels[s] >= 0:
                continue
            labels[s] = s
            stack = [s]
            while stack:
                v = stack.pop()
                for u in indices[indptr[v]:indptr[v + 1]]:
                    if labels[u] < 0:
                        labels[u] = s
                        stack.append(int(u))
        return labels

    def path_lengths(self, indptr, indices, sources, targets):
        out = np.zeros((len(sources), len(targets)), dtype=np.int32)
        for i, s in enumerate(sources):
            dist = {int(s): 1}
            frontier = [int(s)]
            while frontier:
                nxt = []
                for v in frontier:
                    for u in indices[indptr[v]:indptr[v + 1]]:
                        u = int(u)
                        if u not in dist:
                            dist[u] = dist[v] + 1
                            nxt.append(u)
                frontier = nxt
            for j, t in enumerate(targets):
                out[i, j] = dist.get(int(t), 0)
        return out

    def normpdf(self, x, mean, sd):
        return np.array([oracle.normpdf(v, mean, sd) for v in np.atleast_1d(x)])

    def tetramer_distance(self, u, v):
        u, v = np.atleast_2d(u), np.atleast_2d(v)
        return np.array([oracle.euclidean(a, b) for a, b in zip(u, v)])

    def comp_probability(self, dist):
        out = np.array([oracle.comp_prob(d) for d in np.atleast_1d(dist)])
        if np.isnan(out).any():
            raise ZeroDivisionError("float division by zero")
        return out

    def cov_probability(self, cov1, cov2):
        a, b = np.atleast_2d(cov1), np.atleast_2d(cov2)
        return np.array([oracle.cov_prob(x, y) for x, y in zip(a, b)])
