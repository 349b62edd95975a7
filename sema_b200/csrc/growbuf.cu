// growbuf.cu — see growbuf.cuh.
#include "growbuf.cuh"
#include "index_impl.cuh"

namespace sema_impl {

namespace {

struct Driver {
    CUresult (*getGranularity)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags) = nullptr;
    CUresult (*addressReserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*addressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*create)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long) = nullptr;
    CUresult (*release)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*unmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*setAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t) = nullptr;
    bool ok = false;
};

template <typename F>
bool entry(const char *name, F &fn)
{
    void *p = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess || !p) {
        cudaGetLastError();
        return false;
    }
    fn = reinterpret_cast<F>(p);
    return true;
}

const Driver &driver()
{
    static Driver d = [] {
        Driver r;
        r.ok = entry("cuMemGetAllocationGranularity", r.getGranularity) && entry("cuMemAddressReserve", r.addressReserve) &&
               entry("cuMemAddressFree", r.addressFree) && entry("cuMemCreate", r.create) && entry("cuMemRelease", r.release) &&
               entry("cuMemMap", r.map) && entry("cuMemUnmap", r.unmap) && entry("cuMemSetAccess", r.setAccess);
        return r;
    }();
    return d;
}

CUmemAllocationProp prop_for(int device)
{
    CUmemAllocationProp p = {};
    p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    p.location.id = device;
    return p;
}

}  // namespace

int growbuf_reserve(GrowBuf &b, int device, size_t max_bytes, size_t chunk_hint)
{
    const Driver &d = driver();
    if (!d.ok) return fail(SEMA_ERR_UNSUPPORTED, "the driver does not expose the virtual memory management entry points");
    const CUmemAllocationProp p = prop_for(device);
    size_t gran = 0;
    CUresult r = d.getGranularity(&gran, &p, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED);
    if (r != CUDA_SUCCESS || gran == 0) return fail(SEMA_ERR_CUDA, "cuMemGetAllocationGranularity failed (%d)", (int)r);
    size_t chunk = ((chunk_hint ? chunk_hint : gran) + gran - 1) / gran * gran;
    if (max_bytes < 1) max_bytes = 1;
    const size_t reserved = (max_bytes + chunk - 1) / chunk * chunk;
    CUdeviceptr base = 0;
    r = d.addressReserve(&base, reserved, 0, 0, 0);
    if (r != CUDA_SUCCESS) return fail(SEMA_ERR_NOMEM, "cuMemAddressReserve of %zu bytes failed (%d)", reserved, (int)r);
    b.base = base;
    b.reserved = reserved;
    b.committed = 0;
    b.chunk = chunk;
    b.device = device;
    return SEMA_OK;
}

int growbuf_commit(GrowBuf &b, size_t bytes)
{
    if (bytes <= b.committed) return SEMA_OK;
    if (bytes > b.reserved) return fail(SEMA_ERR_CAPACITY, "growable buffer: %zu bytes exceed the %zu reserved", bytes, b.reserved);
    const Driver &d = driver();
    const CUmemAllocationProp p = prop_for(b.device);
    CUmemAccessDesc acc = {};
    acc.location = p.location;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    while (b.committed < bytes) {
        CUmemGenericAllocationHandle h;
        CUresult r = d.create(&h, b.chunk, &p, 0);
        if (r != CUDA_SUCCESS)
            return fail(r == CUDA_ERROR_OUT_OF_MEMORY ? SEMA_ERR_NOMEM : SEMA_ERR_CUDA, "cuMemCreate of %zu bytes failed (%d)", b.chunk, (int)r);
        r = d.map(b.base + b.committed, b.chunk, 0, h, 0);
        if (r == CUDA_SUCCESS) r = d.setAccess(b.base + b.committed, b.chunk, &acc, 1);
        if (r != CUDA_SUCCESS) {
            d.unmap(b.base + b.committed, b.chunk);
            d.release(h);
            return fail(SEMA_ERR_CUDA, "cuMemMap / cuMemSetAccess failed (%d)", (int)r);
        }
        b.handles.push_back(h);
        b.committed += b.chunk;
    }
    return SEMA_OK;
}

void growbuf_free(GrowBuf &b)
{
    if (!b.base) return;
    const Driver &d = driver();
    if (b.committed) d.unmap(b.base, b.committed);
    for (CUmemGenericAllocationHandle h : b.handles) d.release(h);
    d.addressFree(b.base, b.reserved);
    b = GrowBuf();
}

}  // namespace sema_impl
