/*
 * voronoids_b200.h -- C ABI of the B200-native Delaunay engine (libvoronoids_b200.so).
 *
 * Drop-in boundary for the parallel randomized-incremental insertion path of kazewong/Voronoids.  The reference
 * exposes that path as a Rust rlib (`pub mod delaunay_tree/geometry/scheduler`, /root/reference/src/lib.rs:3-5)
 * and as a PyO3 module (/root/reference/src/lib.rs:104-134); it has no C FFI of its own, so each entry point
 * below names the Rust item it replaces.  A Rust `build.rs` + `extern "C"` crate, a cgo stub or Python ctypes bind
 * these symbols directly.
 *
 * Conventions
 *   - opaque handle, caller-owned flat buffers, two-phase size queries (pass NULL to get the count)
 *   - every function returns a vor_status; nothing unwinds across the ABI; vor_last_error() gives the text
 *   - points are row-major n x dim float64 (the reference's Vec<[f64; N]>)
 *   - *_device variants take CUDA device pointers (inputs already resident in HBM)
 *   - one handle <-> one CUDA stream; calls on one handle must be serialised by the caller (&mut self)
 *   - there is NO CPU fallback: without a CUDA device every compute call fails with VOR_ERR_CUDA
 *
 * Vertex ids in exported simplices follow the reference's sequential numbering: 0..dim = super-simplex vertices
 * (delaunay_tree.rs:395-406), dim+1..2*dim+1 = the ghost copies (never part of an exported simplex), input point i
 * = 2*(dim+1) + i  (delaunay_tree.rs:173-174 with lib.rs:107,118).  Simplex ids are engine specific (the
 * reference's depend on its insertion order and are not reproducible by any parallel construction).
 */
#ifndef VORONOIDS_B200_H
#define VORONOIDS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum vor_status {
    VOR_OK = 0,
    VOR_ERR_NO_CONFLICT = 1,     /* reference: panic!("No simplex found ...")  delaunay_tree.rs:47-54 */
    VOR_ERR_DEGENERATE = 2,      /* reference: LU .unwrap() panic             geometry.rs:49 */
    VOR_ERR_DUPLICATE_POINT = 3, /* duplicate input points were dropped (undefined behaviour in the reference) */
    VOR_ERR_CUDA = 4,
    VOR_ERR_OOM = 5,
    VOR_ERR_CAPACITY = 6,        /* a conflict region exceeded the overflow scratch, or the tree would exceed its 2^29
                                  * simplex slots (~19M 3D / ~85M 2D points per tree, all sets and inserts together;
                                  * checked before the insert changes anything) */
    VOR_ERR_RANGE = 7,           /* coordinate dynamic range beyond the exact-arithmetic limb budget */
    VOR_ERR_OUTSIDE = 8,         /* a point lies outside the super simplex built by vor_tree_create */
    VOR_ERR_INTERNAL = 9,
    VOR_ERR_ARG = 10
} vor_status;

typedef struct vor_tree vor_tree; /* opaque; replaces DelaunayTree<N,M>  delaunay_tree.rs:24-30 */

typedef enum vor_insert_mode {
    VOR_INSERT_SINGLE = 0,   /* loop of TreeUpdate::new + insert_point     delaunay_tree.rs:710-739, :125-211 */
    VOR_INSERT_PARALLEL = 1  /* add_points_to_tree                          delaunay_tree.rs:336-386 */
} vor_insert_mode;
/* Both modes run the same device rounds (the triangulation is unique); the mode is kept for API parity. */

/* ---- construction ------------------------------------------------------------------------------------------- */

/* DelaunayTree::<3,4>::new / ::<2,3>::new (delaunay_tree.rs:390, :545): bounding sphere of ALL n points
 * (geometry.rs:99-142), 10x super simplex, nothing inserted yet.  device = CUDA ordinal. */
vor_status vor_tree_create(int dim, const double *points, size_t n, int device, vor_tree **out);
vor_status vor_tree_create_device(int dim, const double *d_points, size_t n, int device, void *cuda_stream, vor_tree **out);
/* batch of independent point sets in one store (BASELINE.json config 5): set s = points[set_offsets[s] .. set_offsets[s+1]) */
vor_status vor_tree_create_batch(int dim, const double *points, const int64_t *set_offsets, size_t n_sets, int device, vor_tree **out);
/* the same across several devices of one node: contiguous blocks of ceil(n_sets / n_dev) sets per
 * device, one host thread per device, no exchange; trees[d] is a batch tree holding sets [shard[d], shard[d+1]) (NULL when
 * the block is empty); shard has n_dev + 1 entries.  A device may be listed more than once. */
vor_status vor_delaunay_batch(int dim, const double *points, const int64_t *set_offsets, size_t n_sets, const int *devices, size_t n_dev,
                              vor_tree **trees, int64_t *shard);
vor_status vor_tree_create_batch_device(int dim, const double *d_points, const int64_t *set_offsets, size_t n_sets, int device,
                                        void *cuda_stream, vor_tree **out);

/* Streaming batch driver (BASELINE.json configs[4]: 8,192 independent sets x 100k points; the caller-side pattern of
 * examples/parallel_insert.rs:56-78): the sets are triangulated in chunks of at most chunk_sets sets / chunk_points points
 * (0 = defaults: 128 sets, what one 2^29-slot store holds), one batch tree per chunk, stores recycled; with host input
 * (points_on_device == 0) the copy of the next chunk overlaps the current chunk's rounds.  One device per call: run one
 * call per device (thread or process) on contiguous blocks of sets -- the sets share nothing.
 * Per set: n_edges[s], checksums[s] (order-independent 64-bit checksum of the set-local canonical edge list, the same
 * function as vor_tree_edges_device); either may be NULL.  cb (optional) is called once per chunk with the chunk's
 * canonical edge list on the host: indices are local to the chunk's first point (first_point = its offset in `points`);
 * the list is only valid during the call. */
typedef void (*vor_chunk_cb)(void *user, size_t first_set, size_t n_sets, int64_t first_point, const uint32_t *edges, size_t n_edges);
vor_status vor_delaunay_batch_stream(int dim, const double *points, int points_on_device, const int64_t *set_offsets, size_t n_sets, int device,
                                     size_t chunk_sets, size_t chunk_points, uint64_t *n_edges, uint64_t *checksums, vor_chunk_cb cb, void *user);
void vor_tree_destroy(vor_tree *t);

/* insert n more points; input index of points[i] = (points inserted so far) + i.
 * For a batch tree, set_offsets (n_sets+1 entries into `points`) is required; pass NULL for a single set. */
vor_status vor_tree_insert(vor_tree *t, const double *points, size_t n, vor_insert_mode mode);
vor_status vor_tree_insert_device(vor_tree *t, const double *d_points, size_t n, vor_insert_mode mode);
vor_status vor_tree_insert_batch(vor_tree *t, const double *points, const int64_t *set_offsets);
vor_status vor_tree_insert_batch_device(vor_tree *t, const double *d_points, const int64_t *set_offsets);

/* voronoids.delaunay(points) (lib.rs:104-125): create + insert everything (3D in the reference; 2D accepted here) */
vor_status vor_delaunay(int dim, const double *points, size_t n, int device, vor_tree **out);

/* ---- queries -------------------------------------------------------------------------------------------------- */

/* vertices.len(), live simplices, max_simplex_id (delaunay_tree.rs:26-29).  n_vertices counts the 2*(dim+1)
 * super+ghost vertices like the reference; max_simplex_id = (dim+1) + simplices created so far. */
vor_status vor_tree_counts(vor_tree *t, uint64_t *n_vertices, uint64_t *n_simplices, uint64_t *max_simplex_id);

/* Delaunay graph: sorted unique {lo,hi} input-index pairs, u32 little endian.
 * edges == NULL: only *n_edges is written. */
vor_status vor_tree_edges(vor_tree *t, uint32_t *edges, size_t cap, size_t *n_edges);
/* same list in a host block of the library's caching host allocator, owned by the CALLER afterwards (release it with
 * vor_host_free): a recycled block is already resident, so the copy skips the first-touch page faults that dominate
 * vor_tree_edges into a fresh buffer; VOR_PINNED_RESULTS=1 makes the blocks page-locked (direct DMA).  This is what
 * voronoids_b200.delaunay(pts).edges() wraps as a numpy array without a copy; it has no counterpart in the reference,
 * whose getters copy the maps element by element (src/lib.rs:73-101). */
vor_status vor_tree_edges_host(vor_tree *t, uint32_t **edges, size_t *n_edges);
vor_status vor_host_free(void *block);
/* same list left on the device (pointer valid until the next insert/destroy) + an order-independent checksum */
vor_status vor_tree_edges_device(vor_tree *t, const uint32_t **d_edges, size_t *n_edges, uint64_t *checksum);

/* live simplices: vertices [n x (dim+1)] (reference ids), neighbours [n x (dim+1)] (index into this export, slot k is
 * opposite vertex k, -1 = hull facet / ghost), centers [n x dim] and radii [n] as geometry.rs computes them.
 * Any array may be NULL.  Replaces the PyDelauanyTree.simplices getter (lib.rs:87-101). */
vor_status vor_tree_export_simplices(vor_tree *t, int32_t *vertices, int32_t *neighbors, double *centers, double *radii, size_t cap,
                                     size_t *n_simplices);

/* DelaunayTree::locate (delaunay_tree.rs:33-58) for n query points of a single-set tree: the conflict region of each
 * point (simplices whose open circumsphere contains it) as indices into the current vor_tree_export_simplices order,
 * `cap` slots per query (unsorted; the reference sorts by its own ids).  counts[i] = region size; 0 = the point
 * coincides with a vertex (the reference panics, delaunay_tree.rs:47-54); -1 = more than cap simplices. */
vor_status vor_tree_locate(vor_tree *t, const double *points, size_t n, int32_t *out_ids, size_t cap, int32_t *counts);

/* vertices in reference id order -- 0..dim super vertices, dim+1..2dim+1 their ghost copies (delaunay_tree.rs:407-412),
 * then input point i at 2(dim+1) + i -- coords [n_vertices x dim] and the incident live simplices of each vertex
 * (Vertex.simplex, delaunay_tree.rs:20-24) as CSR: simp_off [n_vertices + 1], simps [n_incidences] indices into
 * vor_tree_export_simplices, each row sorted (the reference's lists are unordered).  Any array may be NULL (two-phase
 * size query).  Replaces the PyDelauanyTree.vertices getter (lib.rs:73-85). */
vor_status vor_tree_export_vertices(vor_tree *t, double *coords, int64_t *simp_off, int32_t *simps, size_t cap, size_t *n_vertices,
                                    size_t *n_incidences);

/* scheduler::make_queue (src/scheduler.rs:6-28): footprint of every query point against the current tree = sorted
 * unique neighbours-of-neighbours of its conflict region, as indices into vor_tree_export_simplices.  CSR output:
 * offsets[n+1], ids[offsets[n]].  ids == NULL: only offsets and *total are written.  The reference's ghost simplices do
 * not exist here: hull facets contribute nothing (footprints differ only where the 2-ring reaches the super simplex). */
vor_status vor_make_queue(vor_tree *t, const double *points, size_t n, int64_t *offsets, int32_t *ids, size_t cap, size_t *total);
/* scheduler::find_placement (src/scheduler.rs:30-55): 1-based greedy round of every queue entry, in queue order,
 * computed on `device`.  An empty footprint is VOR_ERR_NO_CONFLICT (the reference's .max().unwrap() panics). */
vor_status vor_find_placement(const int64_t *offsets, const int32_t *ids, size_t n, uint64_t *placement, int device);

/* check_delaunay (delaunay_tree.rs:512-541): *ok = 1 iff every live simplex is positively oriented, adjacency is
 * symmetric and every interior facet is locally Delaunay (equivalent to the brute-force empty-sphere test).
 * fail_counts (optional, 6 entries): orientation, dead neighbour, asymmetric, facet mismatch, not Delaunay, and
 * stored circumsphere filter (the cached centre/radius of delaunay_tree.rs:11-16, here a certified filter) contradicting
 * the exact predicate on a simplex's own vertices or on the opposite vertices of its neighbours. */
vor_status vor_tree_check_delaunay(vor_tree *t, int *ok, int32_t *fail_counts);
/* ---- slab mode: ONE triangulation spread over several GPUs (driver: voronoids_b200/slab.py) ----------
 * Every slab must build the same super simplex as a single-GPU run of the whole set (delaunay_tree.rs:392-406 from
 * geometry.rs:99-142): the slabs combine their local bounds (min / max), then their counts of points that fail the
 * strict in_sphere test of the global half-diagonal sphere (sum), and bootstrap from both.  After inserting its own
 * points, the shared coarse sample and a halo, a slab asks whether every simplex around a point it OWNS is a simplex
 * of the global triangulation: the part of its circumsphere inside the data box must stay inside the range of `axis`
 * in which the tree holds every global point -- that range, plus a shell of depth `shell` under the lateral faces of the
 * data box (the empty cap of a hull simplex's nearly flat sphere is thin but wide).  n_uncertified == 0 is the
 * certificate; otherwise need[0..1] is the range along the axis the halo would have to cover. */
vor_status vor_slab_local_bounds(int dim, const double *d_points, size_t n, int device, double *lo, double *hi);
vor_status vor_slab_count_outside(int dim, const double *d_points, size_t n, int device, const double *lo, const double *hi, uint64_t *count);
vor_status vor_tree_create_bounds(int dim, const double *lo, const double *hi, uint64_t outside, size_t capacity_hint, int device, void *cuda_stream,
                                  vor_tree **out);
vor_status vor_tree_certify_slab(vor_tree *t, const uint8_t *owned, size_t n_owned, int axis, double range_lo, double range_hi, double shell,
                                 uint64_t *n_uncertified, double *need);
/* The same pass, also handing out the uncertified simplices themselves: verts = up to cap x (dim+1) x dim vertex coordinates
 * (order of the mesh record: positively oriented), reach = cap x 2 extents along the axis.  For the certificate that is not a
 * ball: a sliver on the hull has a circumsphere of 1e3..1e5 box widths that f64 cannot bound tightly (or at all); its owner asks
 * the peers whether any of THEIR points lies strictly inside it (vor_points_in_spheres, exact predicate) -- none anywhere means the
 * simplex is a simplex of the global triangulation.  inside[j] = number of the n device-resident points strictly inside the
 * circumsphere of simplex j (k simplices of (dim+1) x dim doubles, positively oriented, on the host). */
vor_status vor_tree_uncertified_slab(vor_tree *t, const uint8_t *owned, size_t n_owned, int axis, double range_lo, double range_hi, double shell,
                                     double *verts, double *reach, size_t cap, uint64_t *n_uncertified, double *need);
vor_status vor_points_in_spheres(int dim, const double *d_points, size_t n, const double *simplices, size_t k, int device, uint64_t *inside);

/* this slab's part of the global canonical edge list (host block of the library's allocator, release with vor_host_free):
 * edges of the tree whose endpoint with the lower GLOBAL index is owned by the slab, as sorted (lo, hi) global index pairs.
 * global_index / owned: one entry per inserted point, in insertion order.  The parts of all slabs are disjoint; their
 * sorted union is the canonical edge list of the whole set. */
vor_status vor_tree_edges_slab(vor_tree *t, const int64_t *global_index, const uint8_t *owned, size_t n, uint32_t **edges, size_t *n_edges);

/* TEST HOOK for the checker above (the reference's check_delaunay is never fed a broken mesh either,
 * tests/test_delaunay_tree.rs:37): damages one interior simplex of a finished tree so that fail counter `kind`
 * (0..5, order of fail_counts) must fire.  The tree is unusable for further inserts afterwards. */
vor_status vor_debug_corrupt(vor_tree *t, int kind);

/* bootstrap data: super-simplex vertices [(dim+1) x dim] per set, bounding-sphere centre [dim] and 10x radius */
vor_status vor_tree_super_simplex(vor_tree *t, size_t set, double *super_vertices, double *center, double *radius);

/* engine statistics: [rounds, attempts, winners, owner_resets, compactions, stages,
 *                     walk_steps W, in-sphere tests E, killed K, created C, exact_calls, exact_zero, duplicates, simplex_slots,
 *                     attempts that lost during the flood, in-sphere tests of the attempts that completed,
 *                     conflict tests the cached-sphere filter left to the determinant, points handed to the exact twin,
 *                     attempt slots launched (all rounds)] */
#define VOR_N_STATS 19
vor_status vor_tree_stats(vor_tree *t, uint64_t *stats);

/* with option "profile": CUDA-event milliseconds per kernel class [attempt, check, retri, setup] followed by the
 * launch counts of the same four classes (8 doubles) */
vor_status vor_tree_profile(vor_tree *t, double *out8);

/* ---- geometry (reference: pub mod geometry) ---------------------------------------------------------------- */

/* circumsphere (geometry.rs:58-87): n simplices, verts [n x (dim+1) x dim] -> centers [n x dim], radii [n] */
vor_status vor_circumsphere(int dim, const double *verts, size_t n, double *centers, double *radii, int device);
/* in_sphere (geometry.rs:91-97): out[i] = |c_i - p_i|^2 < r_i^2 (float test, strict) */
vor_status vor_in_sphere(int dim, const double *p, const double *c, const double *r, size_t n, int32_t *out, int device);
/* bounding_sphere (geometry.rs:99-142) */
vor_status vor_bounding_sphere(int dim, const double *points, size_t n, double *center, double *radius, int device);
/* exact predicates (no reference counterpart): kind 0 orient2d [3 pts], 1 orient3d [4 pts],
 * 2 incircle [4 pts], 3 insphere [5 pts]; rows are packed points; out = sign.  n_exact (optional) = calls that
 * needed the exact path. */
vor_status vor_predicates(int kind, const double *rows, size_t n, int32_t *out, uint64_t *n_exact, int device);
/* The cached circumsphere of Simplex{center, radius} (delaunay_tree.rs:11-16) as this engine stores it: a CERTIFIED
 * filter in front of the exact in-sphere predicate (float centre relative to `origin`, inner / outer squared radii).
 * rows = dim+1 simplex vertices followed by one query point; `reach` bounds |p - origin|_1 of every query;
 * out[i] = +1 certainly strictly inside, -1 certainly not strictly inside, 0 undecided (the engine then evaluates
 * the determinant).  blocks (optional, 5 floats per row) = cx, cy, cz, rin2, rout2. */
vor_status vor_sphere_filter(int dim, const double *origin, double reach, const double *rows, size_t n, int32_t *out, float *blocks, int device);

/* ---- misc ------------------------------------------------------------------------------------------------------ */
const char *vor_last_error(void);          /* thread-local text of the last failure */
uint64_t vor_kernel_launches(void);        /* kernels of this library launched so far in this process */
void vor_release_memory(void);             /* return the caching device allocator's free blocks to the driver */
int vor_set_option(const char *name, double value); /* engine options for trees created afterwards */
void vor_tree_set_stream(vor_tree *t, void *cuda_stream);

#ifdef __cplusplus
}
#endif
#endif
