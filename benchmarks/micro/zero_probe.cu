// How fast can 65.7 MB of mask rows be zeroed on B200?  (the dominant traffic of the headline step)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o zero_probe zero_probe.cu && ./zero_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void z_plain(uint4 *p, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += st) p[i] = make_uint4(0, 0, 0, 0);
}
__global__ void z_cs(uint4 *p, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += st) __stcs(p + i, make_uint4(0, 0, 0, 0));
}
__global__ void z_wt(uint4 *p, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += st) __stwt(p + i, make_uint4(0, 0, 0, 0));
}
__global__ void z_v8(uint4 *p, size_t n) {  // 256-bit stores
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 2, st = (size_t)gridDim.x * blockDim.x * 2;
    for (; i + 1 < n; i += st)
        asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p + i), "r"(0) : "memory");
}
// row-wise like k_mask_top: one warp per 16,032-B row
__global__ void z_rows(uint4 *p, uint32_t rows, uint32_t nv) {
    uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= rows) return;
    uint4 *r = p + (size_t)w * nv;
    for (uint32_t i = lane; i < nv; i += 32) r[i] = make_uint4(0, 0, 0, 0);
}
// bulk async store from a zeroed shared buffer (TMA unit does the writes)
__global__ void z_bulk(uint4 *p, size_t n_chunks) {  // chunk = 16 KB
    extern __shared__ __align__(128) unsigned char sm[];
    for (int i = threadIdx.x; i < 16384 / 16; i += blockDim.x) ((uint4 *)sm)[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"((char *)p + c * 16384),
                         "r"((uint32_t)__cvta_generic_to_shared(sm)), "r"(16384) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

int main() {
    const size_t B = 4096, W = 4008, bytes = B * W * 4, NB = 8;
    uint4 *buf[NB];
    for (size_t i = 0; i < NB; ++i) CK(cudaMalloc(&buf[i], bytes + 65536));
    cudaStream_t s[NB];
    for (size_t i = 0; i < NB; ++i) CK(cudaStreamCreateWithFlags(&s[i], cudaStreamNonBlocking));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const size_t nvec = bytes / 16;
    const int steps = 400;
    CK(cudaFuncSetAttribute(z_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384));
    for (int variant = 0; variant < 9; ++variant) {
        for (int multi = 0; multi < 2; ++multi) {
            float ms = 0;
            for (int rep = 0; rep < 2; ++rep) {
                CK(cudaDeviceSynchronize());
                CK(cudaEventRecord(e0, 0));
                for (size_t i = 0; i < NB; ++i) CK(cudaStreamWaitEvent(s[i], e0, 0));
                for (int k = 0; k < steps; ++k) {
                    cudaStream_t st = multi ? s[k % NB] : s[0];
                    uint4 *p = buf[k % NB];
                    switch (variant) {
                        case 0: CK(cudaMemsetAsync(p, 0, bytes, st)); break;
                        case 1: z_plain<<<148 * 8, 256, 0, st>>>(p, nvec); break;
                        case 2: z_plain<<<148 * 32, 256, 0, st>>>(p, nvec); break;
                        case 3: z_cs<<<148 * 8, 256, 0, st>>>(p, nvec); break;
                        case 4: z_wt<<<148 * 8, 256, 0, st>>>(p, nvec); break;
                        case 5: z_v8<<<148 * 8, 256, 0, st>>>(p, nvec); break;
                        case 6: z_rows<<<(unsigned)(B * 32 / 128), 128, 0, st>>>(p, (uint32_t)B, (uint32_t)(W / 4)); break;
                        case 7: z_bulk<<<148 * 4, 128, 16384, st>>>(p, bytes / 16384); break;
                        case 8: z_plain<<<(unsigned)((nvec + 255) / 256), 256, 0, st>>>(p, nvec); break;
                    }
                }
                cudaEvent_t ev[NB];
                for (size_t i = 0; i < NB; ++i) { cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming); cudaEventRecord(ev[i], s[i]); cudaStreamWaitEvent(0, ev[i], 0); }
                CK(cudaEventRecord(e1, 0));
                CK(cudaDeviceSynchronize());
                CK(cudaGetLastError());
                cudaEventElapsedTime(&ms, e0, e1);
                for (size_t i = 0; i < NB; ++i) cudaEventDestroy(ev[i]);
            }
            const char *names[] = {"cudaMemsetAsync", "st.v4 grid 148x8", "st.v4 grid 148x32", "st.cs", "st.wt", "st.v8 (256-bit)", "warp per row", "bulk async store (TMA)", "st.v4 one elem/thread"};
            printf("%-26s %s: %7.2f us per 65.7 MB  = %6.0f GB/s\n", names[variant], multi ? "8 streams" : "1 stream ", ms / steps * 1e3, bytes / (ms / steps * 1e-3) / 1e9);
        }
    }
    return 0;
}
