//! Forwarding layer: `parallel-hnsw`'s public types over the C ABI of `include/phnsw.h`.
//!
//! Mirrors the crate's names and argument meaning (paths relative to the reference crate):
//! `SearchParameters` / `BuildParameters` (src/parameters.rs:3-64), `VectorId` (src/types.rs:3-14),
//! `AbstractVector` (src/types.rs:41-75), `ProgressMonitor` (src/progress.rs:12-29),
//! `BigComparator` (src/bigvec.rs:36-57), `Hnsw::{generate, search, search_upto, knn,
//! threshold_nn, improve_index, promote_at_layer, extend_layer, discover_unreachable_vectors,
//! stochastic_recall, serialize, deserialize}` (src/lib.rs:653-1699), `PqBuildParameters`
//! (src/parameters.rs:66-71), `HnswQuantizer::{quantize, reconstruct}` and `QuantizedHnsw::{new,
//! search, vector_count, quantizer, serialize, deserialize}` (src/pq.rs:61-82, 120-477), plus the
//! device-only ADC view (`AdcIndex`, no crate analogue).
//!
//! NOT COMPILED in this repository's environment (no Rust toolchain); `ffi.rs` is generated from
//! the header and checked by `tests/test_rust_shim.py`, this file is reviewed by hand against it.
pub mod ffi;

use ffi::*;
use rayon::prelude::*;
use std::ffi::{CStr, CString};
use std::os::raw::{c_char, c_int, c_void};
use std::path::Path;
use std::ptr;
use std::sync::Arc;

#[derive(Clone, Copy, Debug, PartialEq, Eq, PartialOrd, Ord, Hash)]
pub struct VectorId(pub usize);

pub enum AbstractVector<'a> {
    Stored(VectorId),
    Unstored(&'a [f32]),
}

pub type SearchParameters = phnsw_search_params;
pub type OptimizationParameters = phnsw_optimization_params;
pub type BuildParameters = phnsw_build_params;

impl Default for phnsw_search_params {
    fn default() -> Self {
        let mut sp = phnsw_search_params { number_of_candidates: 0, upper_layer_candidate_count: 0, probe_depth: 0 };
        unsafe { phnsw_default_search_params(&mut sp) };
        sp
    }
}

impl Default for phnsw_build_params {
    fn default() -> Self {
        let mut bp = std::mem::MaybeUninit::<phnsw_build_params>::uninit();
        unsafe {
            phnsw_default_build_params(bp.as_mut_ptr());
            bp.assume_init()
        }
    }
}

/// src/serialize.rs:11-19 plus the conditions the crate reports by panicking
#[derive(Debug)]
pub enum Error {
    Io(String),
    Serde(String),
    IndexNotFound,
    Interrupted,
    Other(c_int, String),
}

fn check(rc: c_int) -> Result<(), Error> {
    if rc == PHNSW_OK {
        return Ok(());
    }
    let msg = unsafe { CStr::from_ptr(phnsw_last_error()) }.to_string_lossy().into_owned();
    Err(match rc {
        PHNSW_ERR_IO => Error::Io(msg),
        PHNSW_ERR_FORMAT => Error::Serde(msg),
        PHNSW_ERR_NOT_FOUND => Error::IndexNotFound,
        PHNSW_ERR_INTERRUPTED => Error::Interrupted,
        _ => Error::Other(rc, msg),
    })
}

/// the crate panics where the library returns a status (src/lib.rs:181, 261; src/types.rs:86)
fn check_or_panic(rc: c_int) {
    if let Err(e) = check(rc) {
        panic!("phnsw: {e:?}");
    }
}

/// src/progress.rs:12-29
pub struct Interrupt;
pub trait ProgressMonitor: Send {
    fn alive(&mut self) -> Result<(), Interrupt>;
    fn update(&mut self, phase: &str, fraction: f64) -> Result<(), Interrupt>;
}
impl ProgressMonitor for () {
    fn alive(&mut self) -> Result<(), Interrupt> { Ok(()) }
    fn update(&mut self, _phase: &str, _fraction: f64) -> Result<(), Interrupt> { Ok(()) }
}

unsafe extern "C" fn progress_trampoline(user: *mut c_void, phase: *const c_char, fraction: f64) -> c_int {
    let monitor = &mut *(user as *mut &mut dyn ProgressMonitor);
    let phase = CStr::from_ptr(phase).to_string_lossy();
    match monitor.alive().and_then(|_| monitor.update(&phase, fraction)) {
        Ok(()) => 0,
        Err(Interrupt) => 1,
    }
}

struct StoreHandle(*mut phnsw_store);
unsafe impl Send for StoreHandle {}
unsafe impl Sync for StoreHandle {}
impl Drop for StoreHandle {
    fn drop(&mut self) { unsafe { phnsw_store_destroy(self.0) } }
}

/// src/bigvec.rs:36-57: the comparator is a device-resident store; `metric` picks the crate's
/// distance body (0 bigvec.rs:47-53, 1 lib.rs:1985-1991, 2 lib.rs:2431-2437, 3 pq.rs:481-497)
#[derive(Clone)]
pub struct BigComparator {
    store: Arc<StoreHandle>,
    dim: usize,
}

impl BigComparator {
    pub fn new(data: &[Vec<f32>], metric: c_int, device: c_int) -> Result<Self, Error> {
        let dim = data.first().map_or(0, |v| v.len());
        let flat: Vec<f32> = data.iter().flat_map(|v| v.iter().copied()).collect();
        let mut s = ptr::null_mut();
        check(unsafe { phnsw_store_create(metric, dim as u64, data.len() as u64, flat.as_ptr(), device, &mut s) })?;
        Ok(Self { store: Arc::new(StoreHandle(s)), dim })
    }
    pub fn len(&self) -> usize { unsafe { phnsw_store_len(self.store.0) as usize } }
    /// Comparator::compare_vec(Stored(a), Stored(b)) for a batch of pairs (src/lib.rs:69-73)
    pub fn compare_vec(&self, a: &[VectorId], b: &[VectorId]) -> Vec<f32> {
        assert_eq!(a.len(), b.len());
        let (a, b): (Vec<u64>, Vec<u64>) = (a.iter().map(|v| v.0 as u64).collect(), b.iter().map(|v| v.0 as u64).collect());
        let mut out = vec![0f32; a.len()];
        check_or_panic(unsafe { phnsw_store_compare(self.store.0, a.as_ptr(), b.as_ptr(), a.len() as u64, out.as_mut_ptr()) });
        out
    }
    /// Comparator::lookup (src/lib.rs:55-59) for a batch of stored ids
    pub fn lookup(&self, ids: &[VectorId]) -> Vec<Vec<f32>> {
        let raw: Vec<u64> = ids.iter().map(|v| v.0 as u64).collect();
        let mut out = vec![0f32; ids.len() * self.dim];
        check_or_panic(unsafe { phnsw_store_get_rows(self.store.0, raw.as_ptr(), raw.len() as u64, out.as_mut_ptr()) });
        out.chunks(self.dim.max(1)).map(|c| c.to_vec()).collect()
    }
}

struct IndexHandle(*mut phnsw_index);
unsafe impl Send for IndexHandle {}
unsafe impl Sync for IndexHandle {}
impl Drop for IndexHandle {
    fn drop(&mut self) { unsafe { phnsw_index_destroy(self.0) } }
}

pub struct Hnsw {
    comparator: BigComparator,
    index: IndexHandle,
    seed: u64,
}

impl Hnsw {
    /// src/lib.rs:825-893; `seed` replaces the crate's thread_rng
    pub fn generate(c: BigComparator, vs: Vec<VectorId>, bp: BuildParameters, seed: u64,
                    mut progress: &mut dyn ProgressMonitor) -> Result<Self, Error> {
        let ids: Vec<u64> = vs.iter().map(|v| v.0 as u64).collect();
        let mut ix = ptr::null_mut();
        // improve = 1: improve_index after every layer, promote_at_layer live, as the crate
        check(unsafe {
            phnsw_generate_with(c.store.0, ids.as_ptr(), ids.len() as u64, &bp, seed, 1, Some(progress_trampoline),
                                &mut progress as *mut &mut dyn ProgressMonitor as *mut c_void, &mut ix)
        })?;
        Ok(Self { comparator: c, index: IndexHandle(ix), seed })
    }

    pub fn layer_count(&self) -> usize { unsafe { phnsw_index_layer_count(self.index.0) as usize } }
    pub fn vector_count(&self) -> usize { unsafe { phnsw_index_vector_count(self.index.0) as usize } }
    pub fn entry_vector(&self) -> VectorId { VectorId(unsafe { phnsw_index_entry_vector(self.index.0) } as usize) }
    pub fn comparator(&self) -> &BigComparator { &self.comparator }

    /// src/lib.rs:654-665 for a batch: one launch; `upto_layer_from_top` = 0 searches every layer
    pub fn search_batch(&self, vs: &[AbstractVector], sp: SearchParameters, upto_layer_from_top: usize)
                        -> Vec<Vec<(VectorId, f32)>> {
        let nq = vs.len();
        let ef = sp.number_of_candidates as usize;
        let stored = vs.iter().all(|v| matches!(v, AbstractVector::Stored(_)));
        let mut ids: Vec<u64> = Vec::new();
        let mut rows: Vec<f32> = Vec::new();
        if stored {
            ids = vs.iter().map(|v| match v { AbstractVector::Stored(id) => id.0 as u64, _ => unreachable!() }).collect();
        } else {
            // mixed batches: stored queries are fetched as rows first
            for v in vs {
                match v {
                    AbstractVector::Unstored(r) => rows.extend_from_slice(r),
                    AbstractVector::Stored(id) => {
                        let mut r = vec![0f32; self.comparator.dim];
                        let id = id.0 as u64;
                        check_or_panic(unsafe { phnsw_store_get_rows(self.comparator.store.0, &id, 1, r.as_mut_ptr()) });
                        rows.extend_from_slice(&r);
                    }
                }
            }
        }
        let (mut out_ids, mut out_ds, mut cnt) = (vec![0u64; nq * ef], vec![0f32; nq * ef], vec![0u32; nq]);
        check_or_panic(unsafe {
            phnsw_search_batch(self.index.0, if stored { ptr::null() } else { rows.as_ptr() },
                               if stored { ids.as_ptr() } else { ptr::null() }, nq as u64, &sp,
                               upto_layer_from_top as u64, ptr::null(), ef as u64, out_ids.as_mut_ptr(),
                               out_ds.as_mut_ptr(), cnt.as_mut_ptr(), ptr::null_mut(), ptr::null_mut())
        });
        (0..nq).map(|q| (0..cnt[q] as usize).map(|i| (VectorId(out_ids[q * ef + i] as usize), out_ds[q * ef + i])).collect()).collect()
    }
    pub fn search(&self, v: AbstractVector, sp: SearchParameters) -> Vec<(VectorId, f32)> {
        self.search_batch(&[v], sp, 0).pop().unwrap()
    }
    pub fn search_upto(&self, v: AbstractVector, sp: SearchParameters, upto_layer_from_top: usize) -> Vec<(VectorId, f32)> {
        self.search_batch(&[v], sp, upto_layer_from_top).pop().unwrap()
    }

    /// `layers.last().nodes`: the VectorId of every bottom-layer node, in node order
    fn bottom_layer_nodes(&self) -> Vec<u64> {
        let (mut nc, mut m) = (0u64, 0u64);
        let bottom = self.layer_count() as u64 - 1;
        check_or_panic(unsafe { phnsw_index_layer_info(self.index.0, bottom, &mut nc, &mut m) });
        let mut nodes = vec![0u64; nc as usize];
        let mut nb = vec![0u64; (nc * m) as usize];
        check_or_panic(unsafe { phnsw_index_export_layer(self.index.0, bottom, nodes.as_mut_ptr(), nb.as_mut_ptr()) });
        nodes
    }

    /// src/lib.rs:905-928: rows follow bottom-layer node order
    pub fn knn(&self, k: usize, probe_depth: usize) -> impl ParallelIterator<Item = (VectorId, Vec<(VectorId, f32)>)> {
        let n = self.vector_count();
        let (mut ids, mut ds, mut cnt) = (vec![0u64; n * k], vec![0f32; n * k], vec![0u32; n]);
        check_or_panic(unsafe { phnsw_knn(self.index.0, k as u64, probe_depth as u64, ids.as_mut_ptr(), ds.as_mut_ptr(), cnt.as_mut_ptr()) });
        let (mut nc, mut m) = (0u64, 0u64);
        let bottom = self.layer_count() as u64 - 1;
        check_or_panic(unsafe { phnsw_index_layer_info(self.index.0, bottom, &mut nc, &mut m) });
        let mut nodes = vec![0u64; nc as usize];
        let mut nb = vec![0u64; (nc * m) as usize];
        check_or_panic(unsafe { phnsw_index_export_layer(self.index.0, bottom, nodes.as_mut_ptr(), nb.as_mut_ptr()) });
        (0..n).into_par_iter().map(move |i| {
            let row = (0..cnt[i] as usize).map(|j| (VectorId(ids[i * k + j] as usize), ds[i * k + j])).collect();
            (VectorId(nodes[i] as usize), row)
        })
    }

    /// src/lib.rs:930-962
    pub fn threshold_nn(&self, threshold: f32, probe_depth: usize, initial_search_depth: usize)
                        -> Vec<(VectorId, Vec<(VectorId, f32)>)> {
        let n = self.vector_count();
        let (mut off, mut ids, mut ds) = (ptr::null_mut::<u64>(), ptr::null_mut::<u64>(), ptr::null_mut::<f32>());
        check_or_panic(unsafe {
            phnsw_threshold_nn(self.index.0, threshold, probe_depth as u64, initial_search_depth as u64, &mut off, &mut ids, &mut ds)
        });
        // rows follow bottom-layer node order: row i belongs to `layer.nodes[i]` (src/lib.rs:939-960),
        // which is not `i` for an index built over a subset of the store
        let nodes = self.bottom_layer_nodes();
        let out = unsafe {
            let off = std::slice::from_raw_parts(off, n + 1);
            (0..n).map(|i| {
                let (a, b) = (off[i] as usize, off[i + 1] as usize);
                (VectorId(nodes[i] as usize), (a..b).map(|j| (VectorId(*ids.add(j) as usize), *ds.add(j))).collect())
            }).collect()
        };
        unsafe {
            phnsw_free(off as *mut c_void);
            phnsw_free(ids as *mut c_void);
            phnsw_free(ds as *mut c_void);
        }
        out
    }

    /// src/lib.rs:1664-1685
    pub fn improve_index(&mut self, bp: BuildParameters, mut progress: &mut dyn ProgressMonitor) -> f32 {
        let mut recall = 0f32;
        check_or_panic(unsafe {
            phnsw_improve_index(self.index.0, &bp, Some(progress_trampoline),
                                &mut progress as *mut &mut dyn ProgressMonitor as *mut c_void, &mut recall)
        });
        recall
    }
    /// src/lib.rs:1273-1427
    pub fn promote_at_layer(&mut self, layer_from_top: usize, bp: BuildParameters, mut progress: &mut dyn ProgressMonitor) -> bool {
        let mut promoted: c_int = 0;
        check_or_panic(unsafe {
            phnsw_promote_at_layer(self.index.0, layer_from_top as u64, &bp, Some(progress_trampoline),
                                   &mut progress as *mut &mut dyn ProgressMonitor as *mut c_void, &mut promoted)
        });
        promoted != 0
    }
    /// src/lib.rs:1039-1068 (`layer_id` counts from the bottom)
    pub fn extend_layer(&mut self, layer_id: usize, vecs: Vec<VectorId>) {
        let ids: Vec<u64> = vecs.iter().map(|v| v.0 as u64).collect();
        let from_top = (self.layer_count() - layer_id - 1) as u64;
        check_or_panic(unsafe { phnsw_extend_layer(self.index.0, from_top, ids.as_ptr(), ids.len() as u64) });
    }
    /// src/lib.rs:1002-1037
    pub fn discover_unreachable_vectors(&self, layer_id_from_top: usize, sp: SearchParameters) -> Vec<VectorId> {
        let (mut p, mut n) = (ptr::null_mut::<u64>(), 0u64);
        check_or_panic(unsafe { phnsw_discover_unreachable(self.index.0, layer_id_from_top as u64, &sp, &mut p, &mut n) });
        let out = (0..n as usize).map(|i| VectorId(unsafe { *p.add(i) } as usize)).collect();
        unsafe { phnsw_free(p as *mut c_void) };
        out
    }
    /// Layer::node_distances, src/lib.rs:425-489 -> (hops, index_sum) per NodeId of the layer;
    /// usize::MAX marks a node the walk never reached
    pub fn node_distances(&self, layer_from_top: usize, supers: &[VectorId]) -> (Vec<usize>, Vec<usize>) {
        let (mut nc, mut m) = (0u64, 0u64);
        check_or_panic(unsafe { phnsw_index_layer_info(self.index.0, layer_from_top as u64, &mut nc, &mut m) });
        let sup: Vec<u64> = supers.iter().map(|v| v.0 as u64).collect();
        let (mut hops, mut isum) = (vec![0u64; nc as usize], vec![0u64; nc as usize]);
        check_or_panic(unsafe {
            phnsw_node_distances(self.index.0, layer_from_top as u64, sup.as_ptr(), sup.len() as u64,
                                 hops.as_mut_ptr(), isum.as_mut_ptr())
        });
        (hops.into_iter().map(|x| x as usize).collect(), isum.into_iter().map(|x| x as usize).collect())
    }
    /// Layer::discover_nodes_to_promote, src/lib.rs:510-536
    pub fn discover_nodes_to_promote(&self, layer_from_top: usize, supers: &[VectorId]) -> Vec<usize> {
        let sup: Vec<u64> = supers.iter().map(|v| v.0 as u64).collect();
        let (mut p, mut n) = (ptr::null_mut::<u64>(), 0u64);
        check_or_panic(unsafe {
            phnsw_discover_nodes_to_promote(self.index.0, layer_from_top as u64, sup.as_ptr(), sup.len() as u64, &mut p, &mut n)
        });
        let out = (0..n as usize).map(|i| unsafe { *p.add(i) } as usize).collect();
        if n > 0 { unsafe { phnsw_free(p as *mut c_void) }; }
        out
    }
    /// Layer::reachables_from, src/lib.rs:491-508 -> (NodeId, index distance) in discovery order
    pub fn reachables_from(&self, layer_from_top: usize, node: usize, check: &[usize]) -> Vec<(usize, usize)> {
        let chk: Vec<u64> = check.iter().map(|&v| v as u64).collect();
        let (mut on, mut od, mut n) = (vec![0u64; chk.len() + 1], vec![0u64; chk.len() + 1], 0u64);
        check_or_panic(unsafe {
            phnsw_reachables_from(self.index.0, layer_from_top as u64, node as u64, chk.as_ptr(), chk.len() as u64,
                                  on.as_mut_ptr(), od.as_mut_ptr(), &mut n)
        });
        (0..n as usize).map(|i| (on[i] as usize, od[i] as usize)).collect()
    }
    /// src/lib.rs:1501-1505
    pub fn stochastic_recall(&self, op: OptimizationParameters) -> f32 {
        let mut r = 0f32;
        check_or_panic(unsafe { phnsw_stochastic_recall(self.index.0, &op, &mut r) });
        r
    }

    /// src/lib.rs:1688-1699, src/serialize.rs:33-209 (the crate's own directory layout)
    pub fn serialize<P: AsRef<Path>>(&self, path: P) -> Result<(), Error> {
        let dir = CString::new(path.as_ref().to_string_lossy().as_bytes()).map_err(|e| Error::Io(e.to_string()))?;
        check(unsafe { phnsw_index_save(self.index.0, dir.as_ptr()) })
    }
    pub fn deserialize<P: AsRef<Path>>(path: P, device: c_int) -> Result<Self, Error> {
        let dir = CString::new(path.as_ref().to_string_lossy().as_bytes()).map_err(|e| Error::Io(e.to_string()))?;
        let (mut s, mut ix) = (ptr::null_mut(), ptr::null_mut());
        check(unsafe { phnsw_index_load(dir.as_ptr(), device, &mut s, &mut ix) })?;
        let dim = unsafe { phnsw_store_dim(s) } as usize;
        Ok(Self { comparator: BigComparator { store: Arc::new(StoreHandle(s)), dim }, index: IndexHandle(ix), seed: 0 })
    }
}

// ------------------------------------------------------------------------------------------
// Product quantisation: src/pq.rs.  The crate's const generics (SIZE, CENTROID_SIZE,
// QUANTIZED_SIZE) become run-time sizes read back from the handle; codes are `u16` per
// sub-vector as in src/pq.rs:20.
pub type PqBuildParameters = phnsw_pq_build_params;

impl Default for phnsw_pq_build_params {
    fn default() -> Self {
        let mut bp = std::mem::MaybeUninit::<phnsw_pq_build_params>::uninit();
        unsafe {
            phnsw_default_pq_build_params(bp.as_mut_ptr());
            bp.assume_init()
        }
    }
}

struct PqHandle(*mut phnsw_pq);
unsafe impl Send for PqHandle {}
unsafe impl Sync for PqHandle {}
impl Drop for PqHandle {
    fn drop(&mut self) { unsafe { phnsw_pq_destroy(self.0) } }
}

/// src/pq.rs:61-82 `HnswQuantizer`: a borrowed view of the quantizer half of a `QuantizedHnsw`
pub struct HnswQuantizer<'a> {
    pq: &'a PqHandle,
}

impl<'a> HnswQuantizer<'a> {
    pub fn centroid_count(&self) -> usize { unsafe { phnsw_pq_centroid_count(self.pq.0) as usize } }
    pub fn centroid_size(&self) -> usize { unsafe { phnsw_pq_centroid_size(self.pq.0) as usize } }
    pub fn quantized_size(&self) -> usize { unsafe { phnsw_pq_quantized_size(self.pq.0) as usize } }
    /// `Quantizer::quantize` (src/pq.rs:133-146) for a batch: one code per sub-vector, each the
    /// nearest centroid found by an HNSW search over the centroid index
    pub fn quantize(&self, vecs: &[Vec<f32>]) -> Vec<Vec<u16>> {
        let q = self.quantized_size();
        let flat: Vec<f32> = vecs.iter().flat_map(|v| v.iter().copied()).collect();
        let mut codes = vec![0u16; vecs.len() * q];
        check_or_panic(unsafe { phnsw_pq_quantize(self.pq.0, flat.as_ptr(), vecs.len() as u64, codes.as_mut_ptr()) });
        codes.chunks(q.max(1)).map(|c| c.to_vec()).collect()
    }
    /// `Quantizer::reconstruct` (src/pq.rs:148-152): the concatenated centroids
    pub fn reconstruct(&self, codes: &[Vec<u16>]) -> Vec<Vec<f32>> {
        let size = self.quantized_size() * self.centroid_size();
        let flat: Vec<u16> = codes.iter().flat_map(|c| c.iter().copied()).collect();
        let mut out = vec![0f32; codes.len() * size];
        check_or_panic(unsafe { phnsw_pq_reconstruct(self.pq.0, flat.as_ptr(), codes.len() as u64, out.as_mut_ptr()) });
        out.chunks(size.max(1)).map(|c| c.to_vec()).collect()
    }
}

/// src/pq.rs:120-131 `QuantizedHnsw`: centroid index + quantizer + index over the codes + the
/// full-precision comparator used for the re-rank
pub struct QuantizedHnsw {
    pq: PqHandle,
    comparator: BigComparator,
}

impl QuantizedHnsw {
    /// src/pq.rs:287-344.  `centroid_metric` / `quantized_metric` select the distance bodies the
    /// crate's tests plug in as CentroidComparator / QuantizedComparator (2 = src/pq.rs:499-505,
    /// 3 = src/pq.rs:481-497).
    pub fn new(number_of_centroids: usize, centroid_size: usize, comparator: BigComparator, centroid_metric: c_int,
               quantized_metric: c_int, bp: PqBuildParameters, seed: u64, mut progress: &mut dyn ProgressMonitor) -> Self {
        let mut pq = ptr::null_mut();
        check_or_panic(unsafe {
            phnsw_pq_build(comparator.store.0, number_of_centroids as u64, centroid_size as u64, centroid_metric,
                           quantized_metric, &bp, seed, Some(progress_trampoline),
                           &mut progress as *mut &mut dyn ProgressMonitor as *mut c_void, &mut pq)
        });
        Self { pq: PqHandle(pq), comparator }
    }
    pub fn vector_count(&self) -> usize { self.comparator.len() }
    pub fn quantizer(&self) -> HnswQuantizer<'_> { HnswQuantizer { pq: &self.pq } }
    pub fn full_comparator(&self) -> &BigComparator { &self.comparator }
    /// the stored codes, one row of `quantized_size` u16 per vector (the QuantizedComparator's data)
    pub fn codes(&self) -> Vec<Vec<u16>> {
        let q = self.quantizer().quantized_size();
        let mut codes = vec![0u16; self.vector_count() * q];
        check_or_panic(unsafe { phnsw_pq_codes(self.pq.0, codes.as_mut_ptr()) });
        codes.chunks(q.max(1)).map(|c| c.to_vec()).collect()
    }
    /// src/pq.rs:346-364 for a batch: quantise, search the code graph, re-rank every hit with the
    /// full comparator, sort by (distance, id)
    pub fn search_batch(&self, vs: &[AbstractVector], sp: SearchParameters) -> Vec<Vec<(VectorId, f32)>> {
        let nq = vs.len();
        let ef = sp.number_of_candidates as usize;
        let (mut out_ids, mut out_ds, mut cnt) = (vec![0u64; nq * ef], vec![0f32; nq * ef], vec![0u32; nq]);
        let stored = vs.iter().all(|v| matches!(v, AbstractVector::Stored(_)));
        if stored {
            let ids: Vec<u64> = vs.iter().map(|v| match v { AbstractVector::Stored(i) => i.0 as u64, _ => unreachable!() }).collect();
            check_or_panic(unsafe {
                phnsw_pq_search_batch(self.pq.0, ptr::null(), ids.as_ptr(), nq as u64, &sp, ef as u64,
                                      out_ids.as_mut_ptr(), out_ds.as_mut_ptr(), cnt.as_mut_ptr())
            });
        } else {
            // a mixed batch is looked up on the host first (Comparator::lookup_abstract, src/lib.rs:61-67)
            let flat: Vec<f32> = vs.iter().flat_map(|v| match v {
                AbstractVector::Unstored(x) => x.to_vec(),
                AbstractVector::Stored(i) => self.comparator.lookup(&[*i]).pop().unwrap(),
            }).collect();
            check_or_panic(unsafe {
                phnsw_pq_search_batch(self.pq.0, flat.as_ptr(), ptr::null(), nq as u64, &sp, ef as u64,
                                      out_ids.as_mut_ptr(), out_ds.as_mut_ptr(), cnt.as_mut_ptr())
            });
        }
        (0..nq).map(|q| (0..cnt[q] as usize).map(|i| (VectorId(out_ids[q * ef + i] as usize), out_ds[q * ef + i])).collect()).collect()
    }
    pub fn search(&self, v: AbstractVector, sp: SearchParameters) -> Vec<(VectorId, f32)> {
        self.search_batch(&[v], sp).pop().unwrap()
    }
    /// src/pq.rs:433-476: `quantizer/`, `hnsw/`, `comparator/`, `pq_build_parameters.json`
    pub fn serialize<P: AsRef<Path>>(&self, path: P) -> Result<(), Error> {
        let dir = CString::new(path.as_ref().to_string_lossy().as_bytes()).map_err(|e| Error::Io(e.to_string()))?;
        check(unsafe { phnsw_pq_save(self.pq.0, dir.as_ptr()) })
    }
    pub fn deserialize<P: AsRef<Path>>(path: P, device: c_int) -> Result<Self, Error> {
        let dir = CString::new(path.as_ref().to_string_lossy().as_bytes()).map_err(|e| Error::Io(e.to_string()))?;
        let (mut s, mut pq) = (ptr::null_mut(), ptr::null_mut());
        check(unsafe { phnsw_pq_load(dir.as_ptr(), device, &mut s, &mut pq) })?;
        let dim = unsafe { phnsw_store_dim(s) } as usize;
        Ok(Self { pq: PqHandle(pq), comparator: BigComparator { store: Arc::new(StoreHandle(s)), dim } })
    }
}

/// Device-only asymmetric-distance view (no crate analogue; BASELINE.json north_star kernel 2):
/// u8 codes + one k-means codebook, searched with per-query tables in shared memory and
/// re-ranked exactly (the second half of src/pq.rs:346-364) in one call.
pub struct AdcIndex {
    codes: Arc<StoreHandle>,
    index: IndexHandle,
    full: BigComparator,
}

impl AdcIndex {
    /// `graph`: an index built over `full` (its layers are reused over the codes)
    pub fn new(graph: &Hnsw, full: BigComparator, number_of_centroids: usize, centroid_size: usize,
               kmeans_iters: usize, seed: u64, quantised_tables: bool) -> Result<Self, Error> {
        let mut codebook = vec![0f32; number_of_centroids * centroid_size];
        let mut k = 0u64;
        check(unsafe {
            phnsw_pq8_train(full.store.0, number_of_centroids as u64, centroid_size as u64, kmeans_iters as u64, seed,
                            codebook.as_mut_ptr(), &mut k)
        })?;
        let mut s = ptr::null_mut();
        check(unsafe { phnsw_pq8_store_create(full.store.0, codebook.as_ptr(), k, centroid_size as u64, &mut s) })?;
        let codes = Arc::new(StoreHandle(s));
        check(unsafe { phnsw_pq8_store_set_adc_table(codes.0, if quantised_tables { 1 } else { 0 }) })?;
        let mut ix = ptr::null_mut();
        check(unsafe { phnsw_index_rebind(graph.index.0, codes.0, &mut ix) })?;
        Ok(Self { codes, index: IndexHandle(ix), full })
    }
    /// ADC walk + exact re-rank of the first `rerank_k` hits (0 = all candidates), ascending (d, id)
    pub fn search_batch(&self, queries: &[Vec<f32>], sp: SearchParameters, rerank_k: usize, k: usize) -> Vec<Vec<(VectorId, f32)>> {
        let nq = queries.len();
        let flat: Vec<f32> = queries.iter().flat_map(|v| v.iter().copied()).collect();
        let (mut out_ids, mut out_ds, mut cnt) = (vec![0u64; nq * k], vec![0f32; nq * k], vec![0u32; nq]);
        check_or_panic(unsafe {
            phnsw_pq8_search_batch(self.index.0, self.full.store.0, flat.as_ptr(), nq as u64, &sp, rerank_k as u64, k as u64,
                                   out_ids.as_mut_ptr(), out_ds.as_mut_ptr(), cnt.as_mut_ptr())
        });
        (0..nq).map(|q| (0..cnt[q] as usize).map(|i| (VectorId(out_ids[q * k + i] as usize), out_ds[q * k + i])).collect()).collect()
    }
}

// ------------------------------------------------------------------------------------------
// Several GPUs (no analogue in the crate: one index, rayon in one process): one sub-index per
// rank, the exchange inside the library (csrc/sharded.cu).  All buffers are device pointers and
// every call is asynchronous on `stream`.
pub struct CommHandle(*mut phnsw_comm);
unsafe impl Send for CommHandle {}
impl Drop for CommHandle {
    fn drop(&mut self) { unsafe { phnsw_comm_destroy(self.0) } }
}

pub struct ShardedHnsw {
    comm: CommHandle,
    shard: Hnsw,
    id_offset: u64,
}

impl ShardedHnsw {
    /// rank 0 creates the id and hands it to the other ranks over the host's own channel
    pub fn unique_id() -> Result<[u8; 128], Error> {
        let mut id = [0u8; 128];
        check(unsafe { phnsw_comm_unique_id(id.as_mut_ptr() as *mut c_void, id.len() as u64) })?;
        Ok(id)
    }
    pub fn new(shard: Hnsw, id_offset: u64, nranks: c_int, rank: c_int, unique_id: &[u8; 128], device: c_int) -> Result<Self, Error> {
        let mut c = ptr::null_mut();
        check(unsafe { phnsw_comm_init(nranks, rank, unique_id.as_ptr() as *const c_void, device, &mut c) })?;
        Ok(Self { comm: CommHandle(c), shard, id_offset })
    }
    /// broadcast (root >= 0) -> shard search -> one all-gather -> merge, all on `stream`
    pub fn search_batch_device(&self, queries_device: *mut f32, nq: usize, sp: SearchParameters, k: usize, root: c_int,
                               out_ids_device: *mut u64, out_dists_device: *mut f32, stream: *mut c_void) -> Result<(), Error> {
        check(unsafe {
            phnsw_search_batch_sharded(self.comm.0, self.shard.index.0, ptr::null(), queries_device, nq as u64, &sp, 0,
                                       k as u64, self.id_offset, root, out_ids_device, out_dists_device, stream)
        })
    }
    /// the pipelined form: results are complete on `stream` after `flush`
    pub fn search_batch_queued(&self, queries_device: *const f32, nq: usize, sp: SearchParameters, k: usize,
                               out_ids_device: *mut u64, out_dists_device: *mut f32, stream: *mut c_void) -> Result<(), Error> {
        check(unsafe {
            phnsw_search_batch_sharded_queued(self.comm.0, self.shard.index.0, queries_device, nq as u64, &sp, k as u64,
                                              self.id_offset, out_ids_device, out_dists_device, stream)
        })
    }
    pub fn flush(&self, stream: *mut c_void) -> Result<(), Error> {
        check(unsafe { phnsw_comm_flush(self.comm.0, stream) })
    }
}

