"""Array-based A* over the GPU-produced frame record - host code, as the north star keeps A* on the host.

Same search as the reference's `PathFinder.find_path` (PathFinder.py:119-186) on the graph of
`FrameProcessor._create_graph` (FrameProcessor.py:184-207), but fed from the record's arrays instead of pydantic `Grid`
objects and a dict-of-lists graph: nodes are `(x, y)` pixel coordinates (the reference's keys - they also break ties
in the heap), a cell's penalty and its neighbours come from `FrameRecord.penalty / occ / lookup_row`.

Behaviour reproduced on purpose (tests/test_pathfinder_port.py compares paths and costs with the reference itself):
  * cost of a step = distance * (1 + 0.5 * penalty(neighbour) + 1.5 * angle_penalty), angle_penalty = 0 up to 30 degrees
    else (angle / 90) ** 1.5, the angle being the maximum over a 7-cell sliding window of the path so far (:49-99);
  * `angle_cache` outlives a search and stores RADIANS while a fresh computation returns DEGREES (:97-99) - so the
    same window is judged differently the second time it is seen; the cache is shared by every search of a process;
  * no decrease-key: a node already in the open heap keeps its old priority (:181-184);
  * edges exist towards every cell of `grid_lookup`, empty ones included (FrameProcessor.py:203), and a row that is
    listed twice (the duplicate-row quirk of `_extract_grid_information`) contributes its edges twice;
  * the penalty of a neighbour is the one of the `grid_lookup` cell: `None` (0) for empty cells and for rows that are
    only reachable through the lookup.
"""
from __future__ import annotations

from heapq import heappop, heappush

import numpy as np


class ArrayPathFinder:
    """One instance per process keeps the reference's `angle_cache` semantics (PathFinder is a singleton there)."""

    def __init__(self):
        self.angle_cache: dict = {}

    @staticmethod
    def graph_arrays(rec, gs: int):
        """-> (adjacency {(x, y): [((nx, ny), distance), ...]} built on demand, penalty-of-lookup-cell function).

        Mirrors FrameProcessor._create_graph on the record: list rows in order, non-empty cells only, neighbours in the
        order right, left, down, up, an edge whenever `grid_lookup` has the neighbour.  The edge list of a cell is made
        when the search first expands it (a search touches a fraction of the cells)."""
        R, C, x0 = rec.R, rec.C, rec.x0
        lookup_row = rec.lookup_row
        n_lr = len(lookup_row)
        # penalties of the record rows (list rows, then orphan rows: never scored -> None -> 0)
        pen_rows = rec.penalty
        occ = rec.occ
        x_end = x0 + C * gs
        dist = np.sqrt(gs ** 2)                   # np.sqrt((x - nx) ** 2 + (y - ny) ** 2) of every edge (FrameProcessor.py:205)
        rows_of_y: dict = {}                      # a y listed twice (duplicate-row quirk) contributes its edges twice
        for k in range(R):
            rows_of_y.setdefault(int(rec.rows_y[k]), []).append(k)

        def has_row(y):
            ly = y // gs
            return y >= 0 and y % gs == 0 and ly < n_lr and lookup_row[ly] >= 0

        def lookup_penalty(x, y):
            row = int(lookup_row[y // gs])
            if row >= R:
                return 0                      # orphan row: Grid.penalty is None
            p = pen_rows[row][(x - x0) // gs]
            return 0 if p != p else float(p)  # NaN = empty cell = None

        class LazyGraph(dict):
            def get(self, node, default=()):
                edges = dict.get(self, node)
                if edges is None:
                    x, y = node
                    edges = []
                    c = (x - x0) // gs
                    if 0 <= c < C and (x - x0) % gs == 0:
                        for k in rows_of_y.get(y, ()):
                            if not (occ[k][c] & 1):
                                continue
                            for nx, ny in ((x + gs, y), (x - gs, y), (x, y + gs), (x, y - gs)):
                                in_lookup = (x0 <= nx < x_end and has_row(ny)) if ny != y else (x0 <= nx < x_end)
                                if in_lookup:
                                    edges.append(((nx, ny), dist))
                    self[node] = edges
                return edges if edges else default

        return LazyGraph(), lookup_penalty

    def _window_angle(self, nodes):
        """One window of `_angle_between_grids` (PathFinder.py:62-99): `nodes` = the 7 cells path[i-3 .. i+3] ->
        (value as the reference would append it now, value on every later visit), both None when the window is skipped.
        A fresh computation appends DEGREES and caches RADIANS; a cached key appends the cached radians."""
        prev_vector = (nodes[3][0] - nodes[0][0], nodes[3][1] - nodes[0][1])
        next_vector = (nodes[6][0] - nodes[4][0], nodes[6][1] - nodes[4][1])
        key = (prev_vector, next_vector)
        cached = self.angle_cache.get(key)
        if cached is not None:
            return cached, cached
        dot_product = prev_vector[0] * next_vector[0] + prev_vector[1] * next_vector[1]
        magnitude_prev = (prev_vector[0] ** 2 + prev_vector[1] ** 2) ** 0.5
        magnitude_next = (next_vector[0] ** 2 + next_vector[1] ** 2) ** 0.5
        if magnitude_prev == 0 or magnitude_next == 0:
            return None, None
        angle = np.arccos(np.clip(dot_product / (magnitude_prev * magnitude_next), -1.0, 1.0))
        self.angle_cache[key] = angle
        return np.degrees(angle), angle

    # PathFinder.py:119-186
    def find_path(self, graph: dict, lookup_penalty, start: tuple, end: tuple):
        """start / end: (x, y) of the start and end cells -> ([(x, y), ...] from start to end, total cost), ([], inf) if
        there is no path.

        The reference rebuilds the path to `current` for every neighbour and re-evaluates every 7-cell window of it
        (`_angle_between_grids(path_so_far + [neighbour], 7)` - the neighbour itself never enters a window).  The same
        values are produced incrementally here: the ancestors of an expanded node are closed, so their chains are
        frozen; every window but the last one of the path to `current` was evaluated (and cached, in radians) when the
        parent was expanded; the first evaluation at `current` sees the last window fresh (degrees) or cached, every
        later one finds all windows in the cache."""
        open_set: list = []
        open_nodes: set = set()                   # the nodes in open_set (a node is pushed only when absent: :181-184)
        closed_set: set = set()
        came_from: dict = {}
        depth: dict = {start: 1}                  # cells on the path start .. node
        rad_max: dict = {}                        # node -> max of the cached (radian) values over the windows of its path, or None
        g_score: dict = {start: 0}
        f_score: dict = {start: abs(start[0] - end[0]) + abs(start[1] - end[1])}
        heappush(open_set, (f_score[start], start))
        open_nodes.add(start)
        while open_set:
            current = heappop(open_set)[1]
            open_nodes.discard(current)
            if current == end:
                path = []
                node = end
                total_cost = g_score[node]
                while node in came_from:
                    path.append(node)
                    node = came_from[node]
                path.append(start)
                path.reverse()
                return path, total_cost
            closed_set.add(current)
            L = depth[current]
            first_eval = True
            later_value = 0
            for neighbour, distance in graph.get(current, ()):
                if neighbour in closed_set:
                    continue
                if first_eval:
                    first_eval = False
                    # windows exist from 7 cells on (len(path) = L + 1 >= 8); one new window per step
                    if L >= 7:
                        nodes = [current]
                        node = current
                        for _ in range(6):
                            node = came_from[node]
                            nodes.append(node)
                        nodes.reverse()
                        now, later = self._window_angle(nodes)
                        prefix = rad_max.get(came_from[current])
                        if now is None:
                            best_now = best_later = prefix
                        elif prefix is None:
                            best_now, best_later = now, later
                        else:
                            # max() keeps the first of equal values; values compare as numbers either way
                            best_now = max(prefix, now)
                            best_later = max(prefix, later)
                        rad_max[current] = best_later
                        avg_angle_change = 0 if best_now is None else best_now
                        later_value = 0 if best_later is None else best_later
                    else:
                        rad_max[current] = None
                        avg_angle_change = 0
                        later_value = 0
                else:
                    avg_angle_change = later_value
                angle_penalty = 0 if avg_angle_change <= 30 else (avg_angle_change / 90) ** 1.5
                penalty_multiplier = 1 + (0.5 * (lookup_penalty(*neighbour) or 0)) + angle_penalty * 1.5
                tentative = g_score[current] + (distance * penalty_multiplier)
                if neighbour not in g_score or tentative < g_score[neighbour]:
                    came_from[neighbour] = current
                    depth[neighbour] = L + 1
                    g_score[neighbour] = tentative
                    f_score[neighbour] = tentative + abs(neighbour[0] - end[0]) + abs(neighbour[1] - end[1])
                    if neighbour not in open_nodes:
                        heappush(open_set, (f_score[neighbour], neighbour))
                        open_nodes.add(neighbour)
        return [], float("inf")

    def find_paths(self, rec, gs: int):
        """All paths of a frame record: from the record's start cell to the end cell of every peak
        (FrameProcessor._find_paths :230-251, before the similarity filter) -> [(cells [(x, y)...], cost) or None]."""
        if rec.R == 0 or rec.start[0] < 0:
            return []
        graph, pen = self.graph_arrays(rec, gs)
        sx, sy = rec.x0 + rec.start[1] * gs, int(rec.rows_y[rec.start[0]])
        out = []
        for k, c in rec.goals:
            end = (rec.x0 + int(c) * gs, int(rec.rows_y[int(k)]))
            path, cost = self.find_path(graph, pen, (sx, sy), end)
            out.append((path, cost) if path else None)
        return out


path_finder = ArrayPathFinder()
