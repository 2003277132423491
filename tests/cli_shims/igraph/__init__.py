"""TEST INFRASTRUCTURE: the part of python-igraph's ``Graph`` that MetaCoAG v1.1.5 touches
(metacoag:429-469, 1080-1149; graph_utils.py:399-421; label_prop_utils.py:93,134;
matching_utils.py:192,286), for containers without python-igraph.
Undirected; after ``simplify`` the neighbours of a vertex come in ascending id order, which is
igraph's adjacency order for ``mode="ALL"``."""


class Graph:
    class _VS:
        def __init__(self, g):
            self.g = g

        def __getitem__(self, i):
            return self.g._vattr.setdefault(i, {})

        def __len__(self):
            return self.g.n

    def __init__(self):
        self.n = 0
        self.adj = []
        self._vattr = {}
        self.vs = Graph._VS(self)

    def vcount(self):
        return self.n

    def ecount(self):
        return sum(len(a) for a in self.adj) // 2

    def add_vertices(self, n):
        self.adj.extend([] for _ in range(n))
        self.n += n

    def add_edges(self, edges):
        for a, b in edges:
            self.add_edge(a, b)

    def add_edge(self, a, b):
        self.adj[a].append(b)
        self.adj[b].append(a)

    def simplify(self, multiple=True, loops=True, combine_edges=None):
        for v in range(self.n):
            self.adj[v] = sorted(set(w for w in self.adj[v] if w != v))

    @property
    def es(self):
        return [(a, b) for a in range(self.n) for b in self.adj[a] if a < b]

    def neighbors(self, v, mode="ALL"):
        return list(self.adj[v])

    def get_shortest_paths(self, v, to=None):
        prev = {v: None}
        frontier = [v]
        while frontier and to not in prev:
            nxt = []
            for a in frontier:
                for b in self.adj[a]:
                    if b not in prev:
                        prev[b] = a
                        nxt.append(b)
            frontier = nxt
        if to not in prev:
            return [[]]
        path = [to]
        while prev[path[-1]] is not None:
            path.append(prev[path[-1]])
        return [path[::-1]]

    def maximal_cliques(self):
        import networkx as nx

        g = nx.Graph()
        g.add_nodes_from(range(self.n))
        g.add_edges_from(self.es)
        return [tuple(sorted(c)) for c in nx.find_cliques(g)]
