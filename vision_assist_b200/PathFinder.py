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

    # PathFinder.py:49-99
    def _angle_between_grids(self, path, segment_size: int):
        if len(path) < segment_size:
            return 0
        angles = []
        half = segment_size // 2
        for i in range(half, len(path) - half - 1):
            prev_points = path[i - half:i + 1]
            next_points = path[i + 1:i + half + 1]
            prev_vector = (prev_points[-1][0] - prev_points[0][0], prev_points[-1][1] - prev_points[0][1])
            next_vector = (next_points[-1][0] - next_points[0][0], next_points[-1][1] - next_points[0][1])
            key = (tuple(prev_vector), tuple(next_vector))
            if key in self.angle_cache:
                angles.append(self.angle_cache[key])
                continue
            dot_product = prev_vector[0] * next_vector[0] + prev_vector[1] * next_vector[1]
            magnitude_prev = (prev_vector[0] ** 2 + prev_vector[1] ** 2) ** 0.5
            magnitude_next = (next_vector[0] ** 2 + next_vector[1] ** 2) ** 0.5
            if magnitude_prev == 0 or magnitude_next == 0:
                continue
            angle = np.arccos(np.clip(dot_product / (magnitude_prev * magnitude_next), -1.0, 1.0))
            angles.append(np.degrees(angle))
            self.angle_cache[key] = angle
        return max(angles) if angles else 0

    @staticmethod
    def graph_arrays(rec, gs: int):
        """-> (adjacency dict {(x, y): [((nx, ny), distance), ...]}, penalty-of-lookup-cell function).

        Mirrors FrameProcessor._create_graph on the record: list rows in order, non-empty cells only, neighbours in the
        order right, left, down, up, an edge whenever `grid_lookup` has the neighbour."""
        R, C, x0 = rec.R, rec.C, rec.x0
        lookup_row = rec.lookup_row
        n_lr = len(lookup_row)
        # penalties of the record rows (list rows, then orphan rows: never scored -> None -> 0)
        pen_rows = rec.penalty

        def has_row(y):
            ly = y // gs
            return y >= 0 and y % gs == 0 and ly < n_lr and lookup_row[ly] >= 0

        def lookup_penalty(x, y):
            row = int(lookup_row[y // gs])
            if row >= R:
                return 0                      # orphan row: Grid.penalty is None
            p = pen_rows[row][(x - x0) // gs]
            return 0 if p != p else float(p)  # NaN = empty cell = None

        graph: dict = {}
        occ = rec.occ
        for k in range(R):
            y = int(rec.rows_y[k])
            for c in range(C):
                if not (occ[k][c] & 1):
                    continue
                x = x0 + c * gs
                edges = graph.setdefault((x, y), [])
                for nx, ny in ((x + gs, y), (x - gs, y), (x, y + gs), (x, y - gs)):
                    in_lookup = (x0 <= nx < x0 + C * gs and has_row(ny)) if ny != y else (x0 <= nx < x0 + C * gs)
                    if in_lookup:
                        edges.append(((nx, ny), np.sqrt((x - nx) ** 2 + (y - ny) ** 2)))
        return graph, lookup_penalty

    # PathFinder.py:119-186
    def find_path(self, graph: dict, lookup_penalty, start: tuple, end: tuple):
        """start / end: (x, y) of the start and end cells -> ([(x, y), ...] from start to end, total cost), ([], inf) if
        there is no path."""
        open_set: list = []
        closed_set: set = set()
        came_from: dict = {}
        g_score: dict = {start: 0}
        f_score: dict = {start: abs(start[0] - end[0]) + abs(start[1] - end[1])}
        heappush(open_set, (f_score[start], start))
        while open_set:
            current = heappop(open_set)[1]
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
            for neighbour, distance in graph.get(current, ()):
                if neighbour in closed_set:
                    continue
                path_so_far = [current]
                previous = current
                while previous in came_from:
                    previous = came_from[previous]
                    path_so_far.append(previous)
                path_so_far.reverse()
                avg_angle_change = self._angle_between_grids(path_so_far + [neighbour], 7)
                angle_penalty = 0 if avg_angle_change <= 30 else (avg_angle_change / 90) ** 1.5
                penalty_multiplier = 1 + (0.5 * (lookup_penalty(*neighbour) or 0)) + angle_penalty * 1.5
                tentative = g_score[current] + (distance * penalty_multiplier)
                if neighbour not in g_score or tentative < g_score[neighbour]:
                    came_from[neighbour] = current
                    g_score[neighbour] = tentative
                    f_score[neighbour] = tentative + abs(neighbour[0] - end[0]) + abs(neighbour[1] - end[1])
                    if not any(coords == neighbour for _, coords in open_set):
                        heappush(open_set, (f_score[neighbour], neighbour))
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
