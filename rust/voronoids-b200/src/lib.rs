//! Safe layer with the reference's names over the C ABI (the Rust rlib surface of kazewong/Voronoids).
//! NOT BUILT HERE (no Rust toolchain in the build image) -- see Cargo.toml.  What it replaces:
//!   DelaunayTree::<3,4>::new / ::<2,3>::new   src/delaunay_tree.rs:390-510, :545-640
//!   add_points_to_tree                        src/delaunay_tree.rs:336-386
//!   TreeUpdate::new + insert_point            src/delaunay_tree.rs:697-740, :125-211
//!   locate                                    src/delaunay_tree.rs:33-75
//!   check_delaunay                            src/delaunay_tree.rs:512-541
//!   geometry::{circumsphere, in_sphere, bounding_sphere}   src/geometry.rs:2-142
//! Panics mirror the reference's panics (delaunay_tree.rs:53, geometry.rs:49): a non-zero status other than "duplicate points
//! dropped" panics with the library's error text.
pub mod ffi;

use std::ffi::CStr;

fn last_error() -> String {
    unsafe { CStr::from_ptr(ffi::vor_last_error()).to_string_lossy().into_owned() }
}
fn check(st: ffi::vor_status) {
    assert!(st == ffi::VOR_OK || st == ffi::VOR_ERR_DUPLICATE_POINT, "voronoids_b200: {}", last_error());
}

/// Edge list handed out in a block of the library's caching (page-locked) host allocator; returned to it on drop.
pub struct Edges {
    ptr: *mut u32,
    n: usize,
}
impl Edges {
    /// sorted unique (lo, hi) input-index pairs
    pub fn as_pairs(&self) -> &[[u32; 2]] {
        unsafe { std::slice::from_raw_parts(self.ptr as *const [u32; 2], self.n) }
    }
}
impl Drop for Edges {
    fn drop(&mut self) {
        unsafe { ffi::vor_host_free(self.ptr as *mut core::ffi::c_void) };
    }
}

pub struct DelaunayTree<const N: usize, const M: usize> {
    h: *mut ffi::vor_tree,
}
// one handle <-> one CUDA stream: &self calls may run concurrently, mutation needs &mut self (as in the reference)
unsafe impl<const N: usize, const M: usize> Send for DelaunayTree<N, M> {}

impl<const N: usize, const M: usize> DelaunayTree<N, M> {
    /// bounding sphere + super simplex from ALL the points the tree will see; nothing inserted yet
    pub fn new(vertices: Vec<[f64; N]>) -> Self {
        assert!(M == N + 1 && (N == 2 || N == 3));
        let mut h = std::ptr::null_mut();
        check(unsafe { ffi::vor_tree_create(N as i32, vertices.as_ptr() as *const f64, vertices.len(), 0, &mut h) });
        DelaunayTree { h }
    }
    /// the round-parallel path
    pub fn add_points_to_tree(&mut self, vertices: Vec<[f64; N]>) {
        check(unsafe { ffi::vor_tree_insert(self.h, vertices.as_ptr() as *const f64, vertices.len(), ffi::VOR_INSERT_PARALLEL) });
    }
    /// TreeUpdate::new(id, p, &tree) followed by insert_point(&update)
    pub fn insert_point(&mut self, vertex: [f64; N]) {
        check(unsafe { ffi::vor_tree_insert(self.h, vertex.as_ptr(), 1, ffi::VOR_INSERT_SINGLE) });
    }
    /// conflict region of p: indices (export order of vor_tree_export_simplices) of the simplices whose open circumsphere
    /// contains p, ascending; empty if p coincides with a vertex (the reference panics there, delaunay_tree.rs:47-54)
    pub fn locate(&self, p: [f64; N]) -> Vec<usize> {
        let mut count = 0i32;
        let mut ids = vec![0i32; 256];
        loop {
            check(unsafe { ffi::vor_tree_locate(self.h, p.as_ptr(), 1, ids.as_mut_ptr(), ids.len(), &mut count) });
            if count >= 0 { break; }           // -1: more than `cap` simplices, ask again with room
            let n = ids.len() * 4;
            ids.resize(n, 0);
        }
        ids.truncate(count as usize);
        let mut out: Vec<usize> = ids.into_iter().map(|x| x as usize).collect();
        out.sort_unstable();
        out
    }
    pub fn max_simplex_id(&self) -> usize {
        let mut m = 0u64;
        check(unsafe { ffi::vor_tree_counts(self.h, std::ptr::null_mut(), std::ptr::null_mut(), &mut m) });
        m as usize
    }
    pub fn check_delaunay(&self) -> bool {
        let (mut ok, mut fails) = (0i32, [0i32; 6]);
        check(unsafe { ffi::vor_tree_check_delaunay(self.h, &mut ok, fails.as_mut_ptr()) });
        ok != 0
    }
    /// the Delaunay graph as sorted unique (lo, hi) input-index pairs
    pub fn edges(&self) -> Edges {
        let (mut ptr, mut n) = (std::ptr::null_mut(), 0usize);
        check(unsafe { ffi::vor_tree_edges_host(self.h, &mut ptr, &mut n) });
        Edges { ptr, n }
    }
}
impl<const N: usize, const M: usize> Drop for DelaunayTree<N, M> {
    fn drop(&mut self) {
        unsafe { ffi::vor_tree_destroy(self.h) }
    }
}

/// `voronoids.delaunay(points)` of src/lib.rs:104-125: tree + every point inserted
pub fn delaunay(points: Vec<[f64; 3]>) -> DelaunayTree<3, 4> {
    let mut h = std::ptr::null_mut();
    check(unsafe { ffi::vor_delaunay(3, points.as_ptr() as *const f64, points.len(), 0, &mut h) });
    DelaunayTree { h }
}

pub mod geometry {
    use super::{check, ffi};
    pub fn circumsphere<const N: usize, const M: usize>(vertices: [[f64; N]; M]) -> ([f64; N], f64) {
        let (mut c, mut r) = ([0.0; N], 0.0);
        check(unsafe { ffi::vor_circumsphere(N as i32, vertices.as_ptr() as *const f64, 1, c.as_mut_ptr(), &mut r, 0) });
        (c, r)
    }
    pub fn in_sphere<const N: usize>(vertex: [f64; N], center: [f64; N], radius: f64) -> bool {
        let mut out = 0i32;
        check(unsafe { ffi::vor_in_sphere(N as i32, vertex.as_ptr(), center.as_ptr(), &radius, 1, &mut out, 0) });
        out != 0
    }
    pub fn bounding_sphere<const N: usize>(vertices: Vec<[f64; N]>) -> ([f64; N], f64) {
        let (mut c, mut r) = ([0.0; N], 0.0);
        check(unsafe { ffi::vor_bounding_sphere(N as i32, vertices.as_ptr() as *const f64, vertices.len(), c.as_mut_ptr(), &mut r, 0) });
        (c, r)
    }
}
