"""`import voronoids` -- the module name of the reference's PyO3 extension (/root/reference/src/lib.rs:127-134,
`#[pymodule] fn voronoids`), served by the B200 engine: a user of kazewong/Voronoids keeps

    import voronoids
    tree = voronoids.delaunay(points)
    tree.simplices[9].vertices, tree.vertices[8].point, tree.max_simplex_id

unchanged.  Everything lives in voronoids_b200 (hand-written sm_100a CUDA behind include/voronoids_b200.h)."""
from voronoids_b200 import *  # noqa: F401,F403
from voronoids_b200 import __all__, geometry, scheduler  # noqa: F401
