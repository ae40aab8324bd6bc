#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(uint32_t* out) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = slot;
    // write: lane L (absolute 32*warp + lane), column c  <-  L * 100 + c   (columns 0..31)
    for (int half = 0; half < 2; ++half) {
        uint32_t v[16];
        for (int i = 0; i < 16; ++i) v[i] = (32 * warp + lane) * 100 + 16 * half + i;
        const uint32_t ta = tm + ((uint32_t)(32 * warp) << 16) + 16 * half;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
            ::"r"(ta), "r"(v[0]),"r"(v[1]),"r"(v[2]),"r"(v[3]),"r"(v[4]),"r"(v[5]),"r"(v[6]),"r"(v[7]),"r"(v[8]),"r"(v[9]),"r"(v[10]),"r"(v[11]),"r"(v[12]),"r"(v[13]),"r"(v[14]),"r"(v[15]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int g = 0; g < 2; ++g) {
        uint32_t r[16];
        const uint32_t ta = tm + ((uint32_t)(32 * warp + 16 * g) << 16) + 0;
        asm volatile("tcgen05.ld.sync.aligned.16x32bx2.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16], 16;"
            : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15])
            : "r"(ta));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; ++i) out[((g * 4 + warp) * 32 + lane) * 16 + i] = r[i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tm));
}
int main() {
    uint32_t* d; cudaMalloc(&d, 2 * 4 * 32 * 16 * 4);
    k<<<1, 128>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    printf("err %d %s\n", (int)e, cudaGetErrorString(e));
    static uint32_t h[2 * 4 * 32 * 16];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    for (int g = 0; g < 2; ++g) for (int w = 0; w < 4; ++w) for (int l = 0; l < 32; l += (l == 1 ? 14 : (l==15?1:(l==17?14:1)))) {
        if (!(l < 2 || l == 15 || l == 16 || l == 17 || l == 31)) continue;
        printf("g%d w%d lane%2d:", g, w, l);
        for (int i = 0; i < 16; ++i) printf(" %5u", h[((g * 4 + w) * 32 + l) * 16 + i]);
        printf("\n");
    }
    return 0;
}
