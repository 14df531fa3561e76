// peer.cu -- result buffers that several processes (one per GPU of a box) can store into.
//
// The multi-GPU path shards independent units over the ranks and needs exactly one piece of data
// movement: the result rows of every rank on ONE rank ("a final gather over NVLink").
// Instead of a separate collective after the kernel, the collecting rank exports its result buffer
// through CUDA IPC; every other rank maps it into its own address space (peer access over NVLink /
// NVSwitch is enabled by the driver when the handle is opened) and simply passes the mapped pointer
// as the OUTPUT pointer of the ordinary launchers of this library.  The kernels' coalesced row stores
// then cross the link tile by tile while the arithmetic of the next rows is running: the gather is
// fused into the kernel and costs no extra pass over HBM.
//
// The buffer is allocated with cudaMalloc here (not through a caching allocator), so the IPC handle
// refers to the allocation's base and needs no offset.
#include "mpk_common.cuh"

using namespace mpk;

static_assert(sizeof(cudaIpcMemHandle_t) == MPK_PEER_HANDLE_BYTES, "handle size");

extern "C" int mpk_peer_alloc(size_t bytes, void **dev_ptr, unsigned char *handle) {
    if (!dev_ptr || !handle || bytes == 0) return fail(MPK_EINVAL, "bad peer_alloc arguments");
    *dev_ptr = nullptr;
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return fail(MPK_ECUDA, std::string("peer_alloc: cudaMalloc: ") + cudaGetErrorString(e));
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return fail(MPK_ECUDA, std::string("peer_alloc: cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
    }
    memcpy(handle, &h, sizeof h);
    *dev_ptr = p;
    return MPK_OK;
}

extern "C" int mpk_peer_free(void *dev_ptr) {
    if (!dev_ptr) return MPK_OK;
    cudaError_t e = cudaFree(dev_ptr);
    if (e != cudaSuccess) return fail(MPK_ECUDA, std::string("peer_free: ") + cudaGetErrorString(e));
    return MPK_OK;
}

extern "C" int mpk_peer_open(const unsigned char *handle, void **dev_ptr) {
    if (!dev_ptr || !handle) return fail(MPK_EINVAL, "bad peer_open arguments");
    *dev_ptr = nullptr;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    void *p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(MPK_ECUDA, std::string("peer_open: cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
    }
    *dev_ptr = p;
    return MPK_OK;
}

extern "C" int mpk_peer_close(void *dev_ptr) {
    if (!dev_ptr) return MPK_OK;
    cudaError_t e = cudaIpcCloseMemHandle(dev_ptr);
    if (e != cudaSuccess) return fail(MPK_ECUDA, std::string("peer_close: ") + cudaGetErrorString(e));
    return MPK_OK;
}
