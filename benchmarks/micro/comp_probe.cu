// Does B200 speed up zero / sparse-bitmap writes into COMPRESSIBLE memory (cuMemCreate, CU_MEM_ALLOCATION_COMP_GENERIC)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o comp_probe comp_probe.cu -lcuda && ./comp_probe
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
#define CU(x) do { CUresult e = (x); if (e != CUDA_SUCCESS) { const char *s; cuGetErrorString(e, &s); printf("%s: %s\n", #x, s); return 1; } } while (0)

__global__ void z_plain(uint4 *p, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += st) p[i] = make_uint4(0, 0, 0, 0);
}
__global__ void r_sum(const uint4 *p, size_t n, uint32_t *out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    uint32_t a = 0;
    for (; i < n; i += st) { uint4 v = p[i]; a += v.x ^ v.y ^ v.z ^ v.w; }
    if (a == 0x12345678u) *out = a;
}
__global__ void sparse_bits(uint32_t *p, size_t nwords, uint32_t seed) {  // a few random bits per 16 KB row
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nwords / 512) atomicOr(p + (i * 512 + (seed * 2654435761u + (uint32_t)i * 40503u) % 512), 1u << (i & 31));
}

int alloc(bool comp, size_t bytes, CUdeviceptr *out, size_t *padded) {
    CUmemAllocationProp prop = {};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = 0;
    if (comp) prop.allocFlags.compressionType = CU_MEM_ALLOCATION_COMP_GENERIC;
    size_t gran = 0;
    CU(cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
    size_t sz = (bytes + gran - 1) / gran * gran;
    CUmemGenericAllocationHandle h;
    CU(cuMemCreate(&h, sz, &prop, 0));
    CUmemAllocationProp got = {};
    CU(cuMemGetAllocationPropertiesFromHandle(&got, h));
    if (comp) printf("  compression granted: %d (granularity %zu)\n", (int)got.allocFlags.compressionType, gran);
    CU(cuMemAddressReserve(out, sz, 0, 0, 0));
    CU(cuMemMap(*out, sz, 0, h, 0));
    CUmemAccessDesc acc = {};
    acc.location = prop.location;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    CU(cuMemSetAccess(*out, sz, &acc, 1));
    *padded = sz;
    return 0;
}

int main() {
    CK(cudaSetDevice(0));
    CK(cudaFree(0));
    int sup = 0;
    CU(cuDeviceGetAttribute(&sup, CU_DEVICE_ATTRIBUTE_GENERIC_COMPRESSION_SUPPORTED, 0));
    printf("GENERIC_COMPRESSION_SUPPORTED = %d\n", sup);
    const size_t bytes = (size_t)4096 * 4008 * 4, NB = 8;
    uint32_t *sink; CK(cudaMalloc(&sink, 4));
    cudaStream_t s[NB];
    for (size_t i = 0; i < NB; ++i) CK(cudaStreamCreateWithFlags(&s[i], cudaStreamNonBlocking));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int comp = 0; comp <= (sup ? 1 : 0); ++comp) {
        CUdeviceptr buf[NB]; size_t padded = 0;
        for (size_t i = 0; i < NB; ++i) if (alloc(comp, bytes, &buf[i], &padded)) return 1;
        const size_t nvec = bytes / 16;
        for (int test = 0; test < 3; ++test) {
            float ms = 0;
            const int steps = 400;
            for (int rep = 0; rep < 2; ++rep) {
                CK(cudaDeviceSynchronize());
                CK(cudaEventRecord(e0, 0));
                for (size_t i = 0; i < NB; ++i) CK(cudaStreamWaitEvent(s[i], e0, 0));
                for (int k = 0; k < steps; ++k) {
                    cudaStream_t st = s[k % NB];
                    uint4 *p = (uint4 *)buf[k % NB];
                    if (test == 0) z_plain<<<148 * 16, 256, 0, st>>>(p, nvec);
                    else if (test == 1) { z_plain<<<148 * 16, 256, 0, st>>>(p, nvec); sparse_bits<<<(unsigned)((bytes / 4 / 512 + 255) / 256), 256, 0, st>>>((uint32_t *)p, bytes / 4, k); }
                    else r_sum<<<148 * 16, 256, 0, st>>>(p, nvec, sink);
                }
                cudaEvent_t ev[NB];
                for (size_t i = 0; i < NB; ++i) { cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming); cudaEventRecord(ev[i], s[i]); cudaStreamWaitEvent(0, ev[i], 0); }
                CK(cudaEventRecord(e1, 0));
                CK(cudaDeviceSynchronize());
                CK(cudaGetLastError());
                cudaEventElapsedTime(&ms, e0, e1);
                for (size_t i = 0; i < NB; ++i) cudaEventDestroy(ev[i]);
            }
            const char *names[] = {"zero fill", "zero fill + sparse bits", "read back (sparse rows)"};
            printf("%s %-26s: %7.2f us per 65.7 MB = %6.0f GB/s\n", comp ? "COMPRESSIBLE" : "plain       ", names[test], ms / steps * 1e3, bytes / (ms / steps * 1e-3) / 1e9);
        }
    }
    return 0;
}
