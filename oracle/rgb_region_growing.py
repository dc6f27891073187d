"""Oracle for pcl::RegionGrowingRGB::extract (PCL 1.7, segmentation/impl/region_growing_rgb.hpp + region_growing.hpp [upstream]) as
color_growing_segmentation configures it (src/segmentation.cpp:161-216).

TEST INFRASTRUCTURE ONLY.  Pure-Python loops over dict / list containers, written statement by statement after the published PCL
algorithm (std::queue grow, std::priority_queue segment neighbours, vector-of-vector merge) -- deliberately not sharing structure with
the product's C++ (pointcloudcomparator_b200/csrc/pcc_consumers.cu: flat arrays, stamps, sorted candidate lists).  Small cases only.

Parity status: UNPINNED against PCL itself (PCL 1.7 is not in /root/reference and cannot be built here); the restatement follows the
published source as recalled, including its quirks: growRegion walks only the first neighbour_number_ (30) of the 100 neighbours,
colour differences are squared sums of unsigned channel differences, mean segment colours are truncated, ">" rejects in the point test
and "<" accepts in the region test, regions below min size fold into the region of their nearest neighbouring segment.
"""
from __future__ import annotations

import heapq
from collections import deque

import numpy as np

FLT_MAX = float(np.finfo(np.float32).max)


def _colour(rgba, i):
    v = int(rgba[i])
    return ((v >> 16) & 255, (v >> 8) & 255, v & 255)


def _diff(a, b):
    # calculateColorimetricalDifference: three squared channel differences accumulated in a float
    d = np.float32(0.0)
    for x, y in zip(a, b):
        d = np.float32(d + np.float32((x - y) * (x - y)))
    return float(d)


def extract(neighbours, sqr_distances, rgba, distance_threshold=10.0, point_color_threshold=6.0, region_color_threshold=5.0,
            neighbour_number=30, region_neighbour_number=None, min_size=200, max_size=2**31 - 1):
    nb = np.asarray(neighbours); nd = np.asarray(sqr_distances, np.float32); rgba = np.asarray(rgba, np.uint32)
    n, k = nb.shape
    region_neighbour_number = k if region_neighbour_number is None else region_neighbour_number
    dist_thr = float(np.float32(distance_threshold) * np.float32(distance_threshold))
    p2p = float(np.float32(point_color_threshold) * np.float32(point_color_threshold))
    r2r = float(np.float32(region_color_threshold) * np.float32(region_color_threshold))
    rows = [[int(j) for j in nb[i] if j >= 0] for i in range(n)]

    # --- applySmoothRegionGrowingAlgorithm (normal_flag_ = false: residuals all 0, seeds in index order) + growRegion + validatePoint
    point_labels = [-1] * n
    num_pts_in_segment = []
    for seed in range(n):
        if point_labels[seed] != -1:
            continue
        segment = len(num_pts_in_segment)
        seeds = deque([seed]); point_labels[seed] = segment; count = 1
        while seeds:
            curr = seeds.popleft()
            i_nghbr = 0
            while i_nghbr < neighbour_number and i_nghbr < len(rows[curr]):
                index = rows[curr][i_nghbr]; i_nghbr += 1
                if point_labels[index] != -1:
                    continue
                if _diff(_colour(rgba, curr), _colour(rgba, index)) > p2p:
                    continue
                point_labels[index] = segment; count += 1
                seeds.append(index)                                       # is_a_seed stays true: curvature / residual tests are off
        num_pts_in_segment.append(count)
    number_of_segments = len(num_pts_in_segment)
    clusters = [[] for _ in range(number_of_segments)]                     # RegionGrowing::assembleRegions
    for i in range(n):
        clusters[point_labels[i]].append(i)

    # --- findSegmentNeighbours -> findRegionsKNN
    segment_neighbours, segment_distances = [], []
    for index in range(number_of_segments):
        distances = {}
        for point_index in clusters[index]:
            for i_nghbr, other in enumerate(rows[point_index]):
                seg = point_labels[other]
                if seg != index:
                    d = float(nd[point_index, i_nghbr])
                    if distances.get(seg, FLT_MAX) > d:
                        distances[seg] = d
        heap = []                                                          # max-heap of (distance, segment) capped at region_neighbour_number
        for i_seg in range(number_of_segments):
            if i_seg in distances and distances[i_seg] < FLT_MAX:
                heapq.heappush(heap, (-distances[i_seg], -i_seg))
                if len(heap) > region_neighbour_number:
                    heapq.heappop(heap)
        nghbrs, dist = [], []
        while heap and len(nghbrs) < region_neighbour_number:
            d, s = heapq.heappop(heap)                                     # top of the std::priority_queue = largest pair
            dist.append(-d); nghbrs.append(-s)
        segment_neighbours.append(nghbrs); segment_distances.append(dist)

    # --- applyRegionMergingAlgorithm
    segment_color = [[0, 0, 0] for _ in range(number_of_segments)]
    for i in range(n):
        c = _colour(rgba, i)
        for ch in range(3):
            segment_color[point_labels[i]][ch] += c[ch]
    for s in range(number_of_segments):
        for ch in range(3):
            segment_color[s][ch] = int(np.float32(segment_color[s][ch]) / np.float32(num_pts_in_segment[s]))
    segment_labels = [-1] * number_of_segments
    num_pts_in_region, num_seg_in_region = [], []
    for i_seg in range(number_of_segments):
        if segment_labels[i_seg] == -1:
            segment_labels[i_seg] = len(num_pts_in_region)
            num_pts_in_region.append(num_pts_in_segment[i_seg]); num_seg_in_region.append(1)
        curr = segment_labels[i_seg]
        i_nghbr = 0
        while i_nghbr < region_neighbour_number and i_nghbr < len(segment_neighbours[i_seg]):
            index = segment_neighbours[i_seg][i_nghbr]
            far = segment_distances[i_seg][i_nghbr] > dist_thr
            i_nghbr += 1
            if far:
                continue
            if segment_labels[index] == -1 and _diff(segment_color[i_seg], segment_color[index]) < r2r:
                segment_labels[index] = curr
                num_pts_in_region[curr] += num_pts_in_segment[index]; num_seg_in_region[curr] += 1
    region_number = len(num_pts_in_region)
    final_segments = [[] for _ in range(region_number)]
    for i_seg in range(number_of_segments):
        final_segments[segment_labels[i_seg]].append(i_seg)
    # findRegionNeighbours (comparePair orders by distance only; a stable sort fixes the order of equal distances)
    region_neighbours = []
    for i_reg in range(region_number):
        out = []
        for seg in final_segments[i_reg]:
            for d, other in zip(segment_distances[seg], segment_neighbours[seg]):
                if d == FLT_MAX:
                    continue
                if segment_labels[other] != i_reg:
                    out.append((d, other))
        out.sort(key=lambda pr: pr[0])
        region_neighbours.append(out)
    for i_reg in range(region_number):
        if num_pts_in_region[i_reg] >= min_size:
            continue
        if not region_neighbours[i_reg] or region_neighbours[i_reg][0][0] == FLT_MAX:
            continue
        reg_index = segment_labels[region_neighbours[i_reg][0][1]]
        for seg in final_segments[i_reg]:
            final_segments[reg_index].append(seg); segment_labels[seg] = reg_index
        final_segments[i_reg] = []
        num_pts_in_region[reg_index] += num_pts_in_region[i_reg]; num_pts_in_region[i_reg] = 0
        num_seg_in_region[reg_index] += num_seg_in_region[i_reg]; num_seg_in_region[i_reg] = 0
        region_neighbours[reg_index] = [(FLT_MAX, 0) if segment_labels[s] == reg_index else (d, s) for d, s in region_neighbours[reg_index]]
        region_neighbours[reg_index] += [(d, s) for d, s in region_neighbours[i_reg] if segment_labels[s] != reg_index]
        region_neighbours[i_reg] = []
        region_neighbours[reg_index].sort(key=lambda pr: pr[0])
    # --- assembleRegions (empty regions erased) + the size filter of extract()
    labels = np.full(n, -1, np.int32)
    kept = 0
    for i_reg in range(region_number):
        sz = num_pts_in_region[i_reg]
        if sz == 0 or sz < min_size or sz > max_size:
            continue
        for seg in final_segments[i_reg]:
            for p in clusters[seg]:
                labels[p] = kept
        kept += 1
    return labels, kept
