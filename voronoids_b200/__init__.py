"""voronoids_b200 -- B200-native drop-in for the parallel incremental Delaunay path of kazewong/Voronoids.

Python surface mirrors the reference's PyO3 module (/root/reference/src/lib.rs:12-134):

    tree = voronoids_b200.delaunay(points)       # lib.rs:104-125
    tree.max_simplex_id, tree.vertices, tree.simplices   # lib.rs:66-101  (PyDelauanyTree, sic)
    tree.vertices[i].point / .simplex            # PyVertex   lib.rs:12-29
    tree.simplices[j].vertices / .center / .radius / .neighbors   # PySimplex  lib.rs:31-60

    voronoids_b200.scheduler.make_queue / find_placement         # scheduler.rs:6-55
    voronoids_b200.geometry.circumsphere / in_sphere / bounding_sphere   # geometry.rs

plus what the reference leaves implicit: tree.edges() (canonical Delaunay graph) and tree.check_delaunay().
All compute runs in hand-written sm_100a CUDA behind the C ABI of include/voronoids_b200.h; the library is loaded
lazily so that importing the package (and pointgen) works on a machine without the built extension.
"""
from .api import DelaunayTree, PyDelauanyTree, PySimplex, PyVertex, delaunay, delaunay_batch, delaunay_batch_stream  # noqa: F401
from . import geometry, scheduler  # noqa: F401

__all__ = ["delaunay", "delaunay_batch", "delaunay_batch_stream", "DelaunayTree", "PyDelauanyTree", "PySimplex", "PyVertex", "geometry", "scheduler"]
