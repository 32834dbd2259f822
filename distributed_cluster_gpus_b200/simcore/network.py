"""WAN graph (reference: simcore/network.py:7-62).

The graph is static, so the batched engine asks it once per (ingress, DC) pair while flattening the
scenario (spec.py) instead of once per arrival as the reference does (SIM:487).
"""
import heapq
import math
from dataclasses import dataclass
from typing import Dict, List, Tuple


@dataclass(frozen=True)
class Ingress:
    name: str
    region: str


@dataclass
class Edge:
    to: str
    latency_ms: float
    capacity_gbps: float = math.inf
    cost_per_GB: float = 0.0


class Graph:
    """Directed graph; nodes are ingress / DC names."""

    def __init__(self):
        self.adj: Dict[str, List[Edge]] = {}

    def add_edge(self, u: str, v: str, latency_ms: float, capacity_gbps: float = math.inf,
                 cost_per_GB: float = 0.0):
        self.adj.setdefault(u, []).append(Edge(v, latency_ms, capacity_gbps, cost_per_GB))

    def shortest_path_latency(self, src: str, dst: str) -> Tuple[float, List[str], float, float]:
        """Dijkstra on latency -> (latency_s, path, bottleneck_Gbps or 0.0 if unbounded, sum cost/GB)."""
        best_ms: Dict[str, float] = {src: 0.0}
        came_from: Dict[str, Tuple[str, Edge]] = {}
        frontier: List[Tuple[float, str]] = [(0.0, src)]
        while frontier:
            d_ms, node = heapq.heappop(frontier)
            if node == dst:
                break
            if d_ms > best_ms.get(node, math.inf):
                continue
            for edge in self.adj.get(node, ()):
                cand = d_ms + edge.latency_ms
                if cand < best_ms.get(edge.to, math.inf):
                    best_ms[edge.to] = cand
                    came_from[edge.to] = (node, edge)
                    heapq.heappush(frontier, (cand, edge.to))
        if dst not in best_ms:
            return math.inf, [], 0.0, math.inf
        hops = [dst]
        narrowest = math.inf
        cost = 0.0
        node = dst
        while node != src:
            prev, edge = came_from[node]
            hops.append(prev)
            narrowest = min(narrowest, edge.capacity_gbps)
            cost += edge.cost_per_GB
            node = prev
        hops.reverse()
        return best_ms[dst] / 1000.0, hops, (0.0 if narrowest is math.inf else narrowest), cost
