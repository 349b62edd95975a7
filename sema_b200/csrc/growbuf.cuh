// growbuf.cuh — a device buffer that grows in place: one virtual-address reservation of the
// maximum size, physical memory mapped chunk by chunk as it is needed (CUDA virtual memory
// management: cuMemAddressReserve / cuMemCreate / cuMemMap / cuMemSetAccess).
//
// Why: the reference's table grows without a declared size (table.add appends a fragment,
// src/storage/lance_indexer.rs:92-95).  A growable index keeps the matrix contiguous — what every
// scan kernel relies on — without the realloc-and-copy (and the transient 2x footprint) a plain
// cudaMalloc would need, and without committing HBM for rows that do not exist yet.
// The driver entry points are fetched with cudaGetDriverEntryPoint, so the library keeps linking
// against the static runtime only and still loads on a machine without a driver.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

namespace sema_impl {

struct GrowBuf {
    CUdeviceptr base = 0;
    size_t reserved = 0;      // bytes of address space
    size_t committed = 0;     // bytes backed by physical memory (a prefix)
    size_t chunk = 0;         // bytes mapped per step (multiple of the allocation granularity)
    int device = 0;
    std::vector<CUmemGenericAllocationHandle> handles;
    bool active() const { return base != 0; }
};

// reserve address space for max_bytes (rounded up to chunks of ~chunk_hint bytes); nothing is committed
int growbuf_reserve(GrowBuf &b, int device, size_t max_bytes, size_t chunk_hint);
// make [0, bytes) usable; maps further chunks as needed (bytes <= reserved)
int growbuf_commit(GrowBuf &b, size_t bytes);
void growbuf_free(GrowBuf &b);

}  // namespace sema_impl
