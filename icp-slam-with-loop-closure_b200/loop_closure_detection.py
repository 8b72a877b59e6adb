"""Drop-in for the ICP-carrying parts of the reference's ``src/loop_closure_detection.py``.

* ``detect_proximity(pose_graph, lidar_points, min_dist_along_path=2, max_dist=1, err_thresh=110)``
  -- reference :11-39, same signature and side effect: accepted loop closures are added to the
  graph with ``pose_graph.add_constraint(i, j, tf)`` in the reference's order.  Candidate
  generation (:12-25) runs on the GPU without the S x S matrix, ICP for every candidate is one
  batched launch, and the greedy ``points_used`` loop (:26-39) is replayed on the results -- ICP is
  a pure function of its arguments, so the outcome is the reference's.
* ``add_image_match_constraints(pose_graph, good_matches, lidar_points, image_rate=1,
  icp_err_thresh=30)`` -- the ICP block of ``detect_images_direct_similarity`` (:134-159): takes the
  ``good_matches`` the reference's ORB/matcher front end produces (that front end is image
  processing and stays with the reference) and adds the accepted constraints.

Both return the list of (i, j, tf) they added.  No CPU fallback.
"""
from __future__ import annotations

from . import callers


def detect_proximity(pose_graph, lidar_points, min_dist_along_path=2, max_dist=1, err_thresh=110,
                     device=None):
    loops, _ = callers.proximity_loop_closures(pose_graph.poses, lidar_points, min_dist_along_path,
                                               max_dist, err_thresh, device=device)
    for i, j, tf in loops:
        pose_graph.add_constraint(i, j, tf)                     # src/loop_closure_detection.py:36
    return loops


def add_image_match_constraints(pose_graph, good_matches, lidar_points, image_rate=1, icp_err_thresh=30,
                                device=None):
    loops, _ = callers.image_match_loop_closures(good_matches, lidar_points, image_rate, icp_err_thresh,
                                                 device=device)
    for i, j, tf in loops:
        pose_graph.add_constraint(i, j, tf)                     # :155-156
    return loops
