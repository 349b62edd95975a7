//! src/storage/gpu_uploader.rs — index uploader + search seam.  SOURCE ONLY: not compiled in this repo.
//!
//! Call sites in the reference:
//!   * `LanceIndexer::index_chunks`, right after `FixedSizeListArray::from_iter_primitive(vectors, 384)`
//!     (src/storage/lance_indexer.rs:75-76): `uploader.upload(&vector_array, chunks)?;`
//!   * `LanceIndexer::search`, replacing `table.query().nearest_to(q)?.limit(limit).execute()`
//!     (src/storage/lance_indexer.rs:121-141): `return uploader.search(&query_embedding, limit);`
//!   * `LanceIndexer::remove_file_chunks` (src/storage/lance_indexer.rs:234-250):
//!     `uploader.remove_file(file_path)?;`
use arrow_array::{cast::AsArray, types::Float32Type, Array, FixedSizeListArray};
use std::path::Path;

use crate::semantic::gpu::GpuIndex;
use crate::types::Chunk;

pub struct GpuUploader {
    index: GpuIndex,
    rows: Vec<Option<Chunk>>, // row -> Chunk (None after remove_file)
}

impl GpuUploader {
    pub fn new(device: i32, capacity_rows: u64) -> anyhow::Result<Self> {
        Ok(Self { index: GpuIndex::new(device, 384, capacity_rows)?, rows: Vec::new() })
    }

    /// The vector half of index_chunks: the column's values buffer goes to the GPU as is.
    pub fn upload(&mut self, vectors: &FixedSizeListArray, chunks: &[Chunk]) -> anyhow::Result<()> {
        let values = vectors.values().as_primitive::<Float32Type>().values();
        let valid: Vec<u8> = (0..vectors.len()).map(|i| vectors.is_valid(i) as u8).collect();
        // rows are already L2-normalised by mean_pool (src/semantic/embeddings.rs:83-88)
        let first = self.index.append(values, &valid, false)?;
        debug_assert_eq!(first as usize, self.rows.len());
        self.rows.extend(chunks.iter().cloned().map(Some));
        Ok(())
    }

    /// Vec<(Chunk, f32)> in rank order with the real cosine (src/storage/mod.rs:123 attached 1.0).
    pub fn search(&mut self, query_embedding: &[f32], limit: usize) -> anyhow::Result<Vec<(Chunk, f32)>> {
        Ok(self
            .index
            .search(query_embedding, limit)?
            .into_iter()
            .filter_map(|(row, score)| self.rows[row as usize].clone().map(|c| (c, score)))
            .collect())
    }

    pub fn remove_file(&mut self, file_path: &Path) -> anyhow::Result<()> {
        let dead: Vec<u64> = self
            .rows
            .iter()
            .enumerate()
            .filter(|(_, c)| c.as_ref().map_or(false, |c| c.file_path == file_path))
            .map(|(i, _)| i as u64)
            .collect();
        self.index.tombstone(&dead)?;
        for r in dead {
            self.rows[r as usize] = None;
        }
        Ok(())
    }
}
