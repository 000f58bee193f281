//! `extern "C"` declarations of include/voronoids_b200.h (the subset the safe layer uses plus the batch / export entries).
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_int, c_void};

#[repr(C)]
pub struct vor_tree {
    _private: [u8; 0],
}
pub type vor_status = c_int;
pub const VOR_OK: vor_status = 0;
pub const VOR_ERR_DUPLICATE_POINT: vor_status = 3; // points were dropped, the mesh stays valid
pub const VOR_INSERT_SINGLE: c_int = 0; // loop of TreeUpdate::new + insert_point (delaunay_tree.rs:125-211)
pub const VOR_INSERT_PARALLEL: c_int = 1; // add_points_to_tree (delaunay_tree.rs:336-386)

extern "C" {
    pub fn vor_tree_create(dim: c_int, points: *const f64, n: usize, device: c_int, out: *mut *mut vor_tree) -> vor_status;
    pub fn vor_tree_create_device(dim: c_int, d_points: *const f64, n: usize, device: c_int, cuda_stream: *mut c_void, out: *mut *mut vor_tree) -> vor_status;
    pub fn vor_tree_insert(t: *mut vor_tree, points: *const f64, n: usize, mode: c_int) -> vor_status;
    pub fn vor_tree_insert_device(t: *mut vor_tree, d_points: *const f64, n: usize, mode: c_int) -> vor_status;
    pub fn vor_delaunay(dim: c_int, points: *const f64, n: usize, device: c_int, out: *mut *mut vor_tree) -> vor_status;
    pub fn vor_tree_counts(t: *mut vor_tree, n_vertices: *mut u64, n_simplices: *mut u64, max_simplex_id: *mut u64) -> vor_status;
    pub fn vor_tree_edges(t: *mut vor_tree, edges: *mut u32, cap: usize, n_edges: *mut usize) -> vor_status;
    pub fn vor_tree_edges_host(t: *mut vor_tree, edges: *mut *mut u32, n_edges: *mut usize) -> vor_status;
    pub fn vor_host_free(block: *mut c_void) -> vor_status;
    pub fn vor_tree_locate(t: *mut vor_tree, points: *const f64, n: usize, out_ids: *mut i32, cap: usize, counts: *mut i32) -> vor_status;
    pub fn vor_tree_export_simplices(t: *mut vor_tree, vertices: *mut i32, neighbors: *mut i32, centers: *mut f64, radii: *mut f64, cap: usize, n: *mut usize) -> vor_status;
    pub fn vor_tree_export_vertices(t: *mut vor_tree, coords: *mut f64, simp_off: *mut i64, simps: *mut i32, cap: usize, n_vertices: *mut usize, n_incidences: *mut usize) -> vor_status;
    pub fn vor_tree_check_delaunay(t: *mut vor_tree, ok: *mut c_int, fail_counts: *mut i32) -> vor_status;
    pub fn vor_delaunay_batch_stream(dim: c_int, points: *const f64, points_on_device: c_int, set_offsets: *const i64, n_sets: usize, device: c_int,
                                     chunk_sets: usize, chunk_points: usize, n_edges: *mut u64, checksums: *mut u64,
                                     cb: Option<extern "C" fn(*mut c_void, usize, usize, i64, *const u32, usize)>, user: *mut c_void) -> vor_status;
    pub fn vor_make_queue(t: *mut vor_tree, points: *const f64, n: usize, offsets: *mut i64, ids: *mut i32, cap: usize, total: *mut usize) -> vor_status;
    pub fn vor_find_placement(offsets: *const i64, ids: *const i32, n: usize, placement: *mut u64, device: c_int) -> vor_status;
    pub fn vor_circumsphere(dim: c_int, verts: *const f64, n: usize, centers: *mut f64, radii: *mut f64, device: c_int) -> vor_status;
    pub fn vor_in_sphere(dim: c_int, p: *const f64, c: *const f64, r: *const f64, n: usize, out: *mut i32, device: c_int) -> vor_status;
    pub fn vor_bounding_sphere(dim: c_int, points: *const f64, n: usize, center: *mut f64, radius: *mut f64, device: c_int) -> vor_status;
    pub fn vor_tree_destroy(t: *mut vor_tree);
    pub fn vor_last_error() -> *const c_char;
}
