// partition.cu -- two SM partitions per device for the pipelined blocked factorisation (cholesky.cu):
// a small one (8 SMs by default) for the latency-bound chain of diagonal-block factorisations and the rest of
// the device for the bulk products that run beside it.  Built on CUDA green contexts (driver API, CUDA >= 12.4):
// a kernel launched into a stream of a green context runs on that context's SMs only, so the chain's one-CTA
// leaf (182 KB of shared memory: it needs an EMPTY SM) never waits for a bulk CTA to retire.
// Measured on B200 (tools/prio_probe.cu, tools/green_probe.cu, profiles/r02_partition_probes.txt): with stream
// priorities alone a dependent 1-CTA kernel waits 116 us per launch behind a saturating grid of 150 us CTAs
// (waves retire together); in its own partition it waits 0 us, and the bulk grid loses 8 / 148 of its SMs.
//
// The driver entry points are fetched with cudaGetDriverEntryPoint, so libgpb200.so has no link-time
// dependency on libcuda (it must load on a CPU-only box for the symbol checks).  If anything here fails the
// handle simply has no partitions and the engine uses the single-stream recursion.
#include <cuda.h>
#include <stdlib.h>

#include "engine.cuh"

namespace gpb {

namespace {
struct DriverApi {
    CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
    CUresult (*DeviceGetDevResource)(CUdevice, CUdevResource*, CUdevResourceType) = nullptr;
    CUresult (*DevSmResourceSplitByCount)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int,
                                          unsigned int) = nullptr;
    CUresult (*DevResourceGenerateDesc)(CUdevResourceDesc*, CUdevResource*, unsigned int) = nullptr;
    CUresult (*GreenCtxCreate)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int) = nullptr;
    CUresult (*GreenCtxDestroy)(CUgreenCtx) = nullptr;
    CUresult (*GreenCtxStreamCreate)(CUstream*, CUgreenCtx, unsigned int, int) = nullptr;
    bool ok = false;
};

template <class F>
bool fetch(const char* name, F*& fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess || !p) {
        cudaGetLastError();
        return false;
    }
    fn = reinterpret_cast<F*>(p);
    return true;
}

const DriverApi& driver() {
    static DriverApi api = [] {
        DriverApi a;
        a.ok = fetch("cuDeviceGet", a.DeviceGet) && fetch("cuDeviceGetDevResource", a.DeviceGetDevResource) &&
               fetch("cuDevSmResourceSplitByCount", a.DevSmResourceSplitByCount) &&
               fetch("cuDevResourceGenerateDesc", a.DevResourceGenerateDesc) && fetch("cuGreenCtxCreate", a.GreenCtxCreate) &&
               fetch("cuGreenCtxDestroy", a.GreenCtxDestroy) && fetch("cuGreenCtxStreamCreate", a.GreenCtxStreamCreate);
        return a;
    }();
    return api;
}
}  // namespace

// Called (once per handle) when the pipelined factorisation is switched on, with the handle's device current.
// Never fails the handle.
void partitions_create(gpb_handle* h) {
    h->part_ok = false;
    const char* off = getenv("GPB_NO_PARTITIONS");
    if (off && atoi(off) != 0) return;
    const DriverApi& d = driver();
    if (!d.ok) return;
    unsigned want = 8;
    if (const char* e = getenv("GPB_CRIT_SMS")) want = (unsigned)atoi(e);
    if (want < 8 || want >= (unsigned)h->sm_count) return;
    CUdevice dev;
    if (d.DeviceGet(&dev, h->device) != CUDA_SUCCESS) return;
    CUdevResource all, small_part, rest;
    if (d.DeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return;
    unsigned groups = 1;
    if (d.DevSmResourceSplitByCount(&small_part, &groups, &all, &rest, 0, want) != CUDA_SUCCESS || groups != 1) return;
    if (rest.sm.smCount == 0) return;
    CUdevResourceDesc ds = nullptr, dr = nullptr;
    if (d.DevResourceGenerateDesc(&ds, &small_part, 1) != CUDA_SUCCESS) return;
    if (d.DevResourceGenerateDesc(&dr, &rest, 1) != CUDA_SUCCESS) return;
    CUgreenCtx gs = nullptr, gr = nullptr;
    if (d.GreenCtxCreate(&gs, ds, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return;
    if (d.GreenCtxCreate(&gr, dr, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) {
        d.GreenCtxDestroy(gs);
        return;
    }
    bool ok = true;
    CUstream s = nullptr;
    ok = ok && d.GreenCtxStreamCreate(&s, gr, CU_STREAM_NON_BLOCKING, 0) == CUDA_SUCCESS;
    h->part_bulk = reinterpret_cast<cudaStream_t>(s);
    ok = ok && d.GreenCtxStreamCreate(&s, gs, CU_STREAM_NON_BLOCKING, 0) == CUDA_SUCCESS;
    h->part_crit = reinterpret_cast<cudaStream_t>(s);
    for (int i = 0; ok && i < gpb_handle::MAX_DEPTH; ++i) {
        ok = d.GreenCtxStreamCreate(&s, gs, CU_STREAM_NON_BLOCKING, 0) == CUDA_SUCCESS;
        h->part_crit_side[i] = reinterpret_cast<cudaStream_t>(s);
    }
    h->part_ctx[0] = gs;
    h->part_ctx[1] = gr;
    h->part_crit_sms = (int)small_part.sm.smCount;
    h->part_bulk_sms = (int)rest.sm.smCount;
    if (!ok) {
        partitions_destroy(h);
        return;
    }
    h->part_ok = true;
}

void partitions_destroy(gpb_handle* h) {
    h->part_ok = false;
    if (h->part_bulk) cudaStreamDestroy(h->part_bulk);
    if (h->part_crit) cudaStreamDestroy(h->part_crit);
    h->part_bulk = h->part_crit = nullptr;
    for (int i = 0; i < gpb_handle::MAX_DEPTH; ++i) {
        if (h->part_crit_side[i]) cudaStreamDestroy(h->part_crit_side[i]);
        h->part_crit_side[i] = nullptr;
    }
    for (cudaEvent_t e : h->part_events) cudaEventDestroy(e);
    h->part_events.clear();
    const DriverApi& d = driver();
    for (int i = 0; i < 2; ++i) {
        if (h->part_ctx[i] && d.ok) d.GreenCtxDestroy(static_cast<CUgreenCtx>(h->part_ctx[i]));
        h->part_ctx[i] = nullptr;
    }
}

// i-th reusable (timing-disabled) event of the pipeline
cudaEvent_t partition_event(gpb_handle* h, size_t i) {
    while (h->part_events.size() <= i) {
        cudaEvent_t e = nullptr;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        h->part_events.push_back(e);
    }
    return h->part_events[i];
}

}  // namespace gpb
