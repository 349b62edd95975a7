//! src/semantic/gpu.rs — FFI over include/sema_b200.h.  SOURCE ONLY: not compiled in this repo.
use std::os::raw::{c_char, c_int};

#[repr(C)]
pub struct SemaIndex {
    _private: [u8; 0],
}

pub const SEMA_METRIC_COSINE: c_int = 0;
pub const SEMA_METRIC_L2: c_int = 1;

#[link(name = "sema_b200")]
extern "C" {
    fn sema_index_create(device: c_int, dim: u32, capacity_rows: u64, metric: c_int, out: *mut *mut SemaIndex) -> c_int;
    fn sema_index_destroy(idx: *mut SemaIndex) -> c_int;
    fn sema_index_append(idx: *mut SemaIndex, rows: *const f32, n: u64, valid: *const u8, normalize: c_int,
                         first_row: *mut u64) -> c_int;
    fn sema_index_tombstone(idx: *mut SemaIndex, rows: *const u64, n: u64) -> c_int;
    fn sema_index_compact(idx: *mut SemaIndex, new_row_of_old: *mut u64, n_live: *mut u64) -> c_int;
    fn sema_index_search(idx: *mut SemaIndex, q: *const f32, k: u32, row_ids: *mut u64, scores: *mut f32,
                         n_found: *mut u32) -> c_int;
    fn sema_index_search_submit(idx: *mut SemaIndex, q: *const f32, k: u32, ticket: *mut u64) -> c_int;
    fn sema_index_search_collect(idx: *mut SemaIndex, ticket: u64, row_ids: *mut u64, scores: *mut f32,
                                 n_found: *mut u32) -> c_int;
    fn sema_index_search_batch(idx: *mut SemaIndex, q: *const f32, nq: u32, k: u32, row_ids: *mut u64,
                               scores: *mut f32, n_found: *mut u32) -> c_int;
    fn sema_mean_pool(idx: *mut SemaIndex, tokens: *const f32, mask: *const f32, n: u64, seq_len: u32,
                      skip_masked: c_int, out: *mut f32) -> c_int;
    fn sema_index_size(idx: *const SemaIndex) -> u64;
    fn sema_index_save(idx: *mut SemaIndex, path: *const c_char) -> c_int;
    fn sema_index_load(path: *const c_char, device: c_int, capacity_rows: u64, out: *mut *mut SemaIndex) -> c_int;
    fn sema_index_set_row_base(idx: *mut SemaIndex, row_base: u64) -> c_int;
    fn sema_device_count() -> c_int;
    fn sema_shard_group_create_local(shards: *const *mut SemaIndex, n_shards: u32, out: *mut *mut SemaShardGroup) -> c_int;
    fn sema_shard_group_search(g: *mut SemaShardGroup, q: *const f32, k: u32, row_ids: *mut u64, scores: *mut f32,
                               n_found: *mut u32) -> c_int;
    fn sema_shard_group_destroy(g: *mut SemaShardGroup) -> c_int;
    fn sema_last_error() -> *const c_char;
}

#[repr(C)]
pub struct SemaShardGroup {
    _private: [u8; 0],
}

fn check(rc: c_int) -> anyhow::Result<()> {
    if rc == 0 {
        return Ok(());
    }
    let msg = unsafe { std::ffi::CStr::from_ptr(sema_last_error()) }.to_string_lossy().into_owned();
    Err(anyhow::anyhow!("sema_b200 error {rc}: {msg}"))
}

/// The chunk-embedding matrix resident in one GPU's HBM.
pub struct GpuIndex {
    raw: *mut SemaIndex,
    dim: usize,
}

// every entry point re-binds the CUDA device; `&mut self` gives the one-caller-at-a-time rule
unsafe impl Send for GpuIndex {}

impl GpuIndex {
    pub fn new(device: i32, dim: usize, capacity_rows: u64) -> anyhow::Result<Self> {
        let mut raw = std::ptr::null_mut();
        check(unsafe { sema_index_create(device, dim as u32, capacity_rows, SEMA_METRIC_COSINE, &mut raw) })?;
        Ok(Self { raw, dim })
    }

    /// rows: n x dim contiguous f32 (the Arrow values buffer); valid: one byte per row, empty = all valid.
    pub fn append(&mut self, rows: &[f32], valid: &[u8], normalize: bool) -> anyhow::Result<u64> {
        let n = (rows.len() / self.dim) as u64;
        let mut first = 0u64;
        let v = if valid.is_empty() { std::ptr::null() } else { valid.as_ptr() };
        check(unsafe { sema_index_append(self.raw, rows.as_ptr(), n, v, normalize as c_int, &mut first) })?;
        Ok(first)
    }

    /// Ranked (row, cosine) pairs, best first — replaces nearest_to(q).limit(limit).
    pub fn search(&mut self, q: &[f32], limit: usize) -> anyhow::Result<Vec<(u64, f32)>> {
        anyhow::ensure!(q.len() == self.dim, "query has {} dims, index has {}", q.len(), self.dim);
        let (mut ids, mut sc, mut nf) = (vec![0u64; limit.max(1)], vec![0f32; limit.max(1)], 0u32);
        check(unsafe { sema_index_search(self.raw, q.as_ptr(), limit as u32, ids.as_mut_ptr(), sc.as_mut_ptr(), &mut nf) })?;
        Ok(ids.into_iter().zip(sc).take(nf as usize).collect())
    }

    /// Asynchronous search: enqueue now (q is consumed before this returns), `collect` later.  Up to 8
    /// tickets may be outstanding; consecutive scans overlap on the device.
    pub fn submit(&mut self, q: &[f32], limit: usize) -> anyhow::Result<u64> {
        anyhow::ensure!(q.len() == self.dim, "query has {} dims, index has {}", q.len(), self.dim);
        let mut ticket = 0u64;
        check(unsafe { sema_index_search_submit(self.raw, q.as_ptr(), limit as u32, &mut ticket) })?;
        Ok(ticket)
    }

    pub fn collect(&mut self, ticket: u64, limit: usize) -> anyhow::Result<Vec<(u64, f32)>> {
        let (mut ids, mut sc, mut nf) = (vec![0u64; limit.max(1)], vec![0f32; limit.max(1)], 0u32);
        check(unsafe { sema_index_search_collect(self.raw, ticket, ids.as_mut_ptr(), sc.as_mut_ptr(), &mut nf) })?;
        Ok(ids.into_iter().zip(sc).take(nf as usize).collect())
    }

    /// nq queries at once (tensor-core path when the shape allows); result[i] = ranked hits of query i.
    pub fn search_batch(&mut self, queries: &[f32], limit: usize) -> anyhow::Result<Vec<Vec<(u64, f32)>>> {
        let nq = queries.len() / self.dim;
        let (mut ids, mut sc, mut nf) = (vec![0u64; nq * limit], vec![0f32; nq * limit], vec![0u32; nq]);
        check(unsafe {
            sema_index_search_batch(self.raw, queries.as_ptr(), nq as u32, limit as u32, ids.as_mut_ptr(),
                                    sc.as_mut_ptr(), nf.as_mut_ptr())
        })?;
        Ok((0..nq)
            .map(|i| (0..nf[i] as usize).map(|j| (ids[i * limit + j], sc[i * limit + j])).collect())
            .collect())
    }

    /// Drop-in for `mean_pool` (src/semantic/embeddings.rs:61-91) over n texts: tokens is the ONNX
    /// last_hidden_state (n x seq_len x dim), mask the f32 attention mask (n x seq_len).  The GPU
    /// keeps the reference's summation order, so the result is bit-identical to the CPU function.
    pub fn mean_pool(&mut self, tokens: &[f32], mask: &[f32], seq_len: usize) -> anyhow::Result<Vec<f32>> {
        let n = mask.len() / seq_len;
        anyhow::ensure!(tokens.len() == n * seq_len * self.dim, "tokens / mask shape mismatch");
        let mut out = vec![0f32; n * self.dim];
        check(unsafe {
            sema_mean_pool(self.raw, tokens.as_ptr(), mask.as_ptr(), n as u64, seq_len as u32, 1, out.as_mut_ptr())
        })?;
        Ok(out)
    }

    pub fn tombstone(&mut self, rows: &[u64]) -> anyhow::Result<()> {
        check(unsafe { sema_index_tombstone(self.raw, rows.as_ptr(), rows.len() as u64) })
    }

    pub fn len(&self) -> u64 {
        unsafe { sema_index_size(self.raw) }
    }
}

impl Drop for GpuIndex {
    fn drop(&mut self) {
        unsafe { sema_index_destroy(self.raw) };
    }
}

/// The same index spread over every GPU of the box behind ONE handle (Sema is one process with one
/// StorageManager: src/main.rs:9, src/storage/mod.rs:13-16).  Rows are dealt out in contiguous ranges of
/// `capacity_rows / n_gpus`; a search is one host call: every GPU scans its range, the shards exchange their
/// top-k over NVLink inside the scan kernel, the call returns the global top-k (limit <= 128).
pub struct MultiGpuIndex {
    shards: Vec<GpuIndex>,
    group: *mut SemaShardGroup,
    per: u64,
    dim: usize,
}

unsafe impl Send for MultiGpuIndex {}

impl MultiGpuIndex {
    pub fn new(dim: usize, capacity_rows: u64) -> anyhow::Result<Self> {
        let n = unsafe { sema_device_count() }.max(1) as u64;
        let per = (capacity_rows + n - 1) / n;
        let mut shards = Vec::new();
        for g in 0..n {
            let shard = GpuIndex::new(g as i32, dim, per)?;
            check(unsafe { sema_index_set_row_base(shard.raw, g * per) })?;
            shards.push(shard);
        }
        let raws: Vec<*mut SemaIndex> = shards.iter().map(|s| s.raw).collect();
        let mut group = std::ptr::null_mut();
        check(unsafe { sema_shard_group_create_local(raws.as_ptr(), raws.len() as u32, &mut group) })?;
        Ok(Self { shards, group, per, dim })
    }

    /// Appends fill the ranges in table order; returns the global id of the first appended row.
    pub fn append(&mut self, rows: &[f32], valid: &[u8], normalize: bool) -> anyhow::Result<u64> {
        let n = rows.len() / self.dim;
        let (mut done, mut first_global) = (0usize, None);
        for (g, shard) in self.shards.iter_mut().enumerate() {
            if done == n {
                break;
            }
            let room = (self.per - shard.len()) as usize;
            let m = room.min(n - done);
            if m == 0 {
                continue;
            }
            let v = if valid.is_empty() { &valid[..] } else { &valid[done..done + m] };
            let local = shard.append(&rows[done * self.dim..(done + m) * self.dim], v, normalize)?;
            first_global.get_or_insert(g as u64 * self.per + local);
            done += m;
        }
        anyhow::ensure!(done == n, "the index is full");
        Ok(first_global.unwrap_or(0))
    }

    /// Ranked (global row, cosine) pairs, best first.
    pub fn search(&mut self, q: &[f32], limit: usize) -> anyhow::Result<Vec<(u64, f32)>> {
        anyhow::ensure!(q.len() == self.dim, "query has {} dims, index has {}", q.len(), self.dim);
        let (mut ids, mut sc, mut nf) = (vec![0u64; limit.max(1)], vec![0f32; limit.max(1)], 0u32);
        check(unsafe { sema_shard_group_search(self.group, q.as_ptr(), limit as u32, ids.as_mut_ptr(), sc.as_mut_ptr(), &mut nf) })?;
        Ok(ids.into_iter().zip(sc).take(nf as usize).collect())
    }
}

impl Drop for MultiGpuIndex {
    fn drop(&mut self) {
        unsafe { sema_shard_group_destroy(self.group) };   // before the shards it refers to
    }
}
