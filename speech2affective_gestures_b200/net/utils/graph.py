"""Constant ST-GCN adjacency tensors (host side, constructor time only).

Restates the behaviour of the reference's `Graph(num_nodes, links, strategy='spatial', max_hop=2)`
(net/utils/graph.py:26-103, helpers :106-129): hop distances capped at `max_hop` (pairs further
apart are +inf), a column-normalised hop<=max_hop adjacency, and the "spatial configuration"
partition around centre node 0: for each hop h, edges j->i are split by comparing the (capped) hop
distance of j and i to the centre: equal -> root, j further -> "closer" set, else "further" set;
slices are [root_0, (root+close)_1, further_1, (root+close)_2, further_2, ...].
"""
from collections import deque

import numpy as np


def hop_distance(num_nodes, edges, max_hop):
    adj = [[] for _ in range(num_nodes)]
    for a, b in edges:
        if a != b:
            adj[a].append(b)
            adj[b].append(a)
    dist = np.full((num_nodes, num_nodes), np.inf)
    for s in range(num_nodes):
        dist[s, s] = 0
        q = deque([s])
        while q:
            u = q.popleft()
            if dist[s, u] >= max_hop:
                continue
            for v in adj[u]:
                if dist[s, v] == np.inf:
                    dist[s, v] = dist[s, u] + 1
                    q.append(v)
    return dist


class Graph:
    def __init__(self, num_nodes, neighbor_links, strategy="spatial", layout="openpose", max_hop=1, dilation=1):
        if strategy != "spatial":
            raise ValueError("only the 'spatial' strategy is on the hot path")
        self.num_nodes = num_nodes
        self.max_hop = max_hop
        self.dilation = dilation
        self.center = 0
        self.edges = [(i, i) for i in range(num_nodes)] + list(neighbor_links)
        self.hop_dis = hop_distance(num_nodes, self.edges, max_hop)
        self.A = self._spatial()

    def _spatial(self):
        n = self.num_nodes
        hops = range(0, self.max_hop + 1, self.dilation)
        reach = np.zeros((n, n))
        for h in hops:
            reach[self.hop_dis == h] = 1
        col = reach.sum(0)
        norm = reach / np.where(col > 0, col, 1)[None, :]
        norm[:, col == 0] = 0
        dc = self.hop_dis[:, self.center]
        out = []
        for h in hops:
            sel = self.hop_dis == h                      # sel[j, i]
            dj, di = dc[:, None], dc[None, :]
            root = np.where(sel & (dj == di), norm, 0.0)
            close = np.where(sel & (dj > di), norm, 0.0)
            further = np.where(sel & ~(dj == di) & ~(dj > di), norm, 0.0)
            if h == 0:
                out.append(root)
            else:
                out.append(root + close)
                out.append(further)
        return np.stack(out)
