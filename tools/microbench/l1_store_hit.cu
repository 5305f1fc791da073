// Does a global store to a line that is resident in L1 keep the line valid for later loads of the same thread?
// (single thread, latencies in clock64 ticks).  usage: nvcc -arch=sm_100a -o l1_store_hit l1_store_hit.cu && ./l1_store_hit
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* a, long long* out, int pf)
{
    long long t[8];
    int acc = 0;
    int* line = a + 1024;  // some line
    if (pf) asm volatile("prefetch.global.L1 [%0];" ::"l"(line));
    for (int i = 0; i < 200; i++) acc += __float2int_rn(sqrtf((float)(i + acc)));  // give the prefetch time
    t[0] = clock64();
    int v0 = *((volatile int*)line + 0 + (acc & 0));  // volatile would bypass L1: use asm ld.ca instead
    asm volatile("ld.global.ca.s32 %0, [%1];" : "=r"(v0) : "l"(line + (acc & 0)) : "memory");
    acc += v0;
    t[1] = clock64();
    int v1;
    asm volatile("ld.global.ca.s32 %0, [%1];" : "=r"(v1) : "l"(line + 1 + (acc & 0)) : "memory");
    acc += v1;
    t[2] = clock64();
    asm volatile("st.global.s32 [%0], %1;" ::"l"(line + 2), "r"(acc) : "memory");
    int v2;
    asm volatile("ld.global.ca.s32 %0, [%1];" : "=r"(v2) : "l"(line + 3 + (acc & 0)) : "memory");
    acc += v2;
    t[3] = clock64();
    int v3;
    asm volatile("ld.global.ca.s32 %0, [%1];" : "=r"(v3) : "l"(line + 2 + (acc & 0)) : "memory");  // the stored word itself
    acc += v3;
    t[4] = clock64();
    int v4;
    asm volatile("ld.global.ca.s32 %0, [%1];" : "=r"(v4) : "l"(line + 4 + (acc & 0)) : "memory");
    acc += v4;
    t[5] = clock64();
    for (int i = 0; i < 6; i++) out[i] = t[i];
    out[7] = acc;
}
int main()
{
    int* a; long long* o;
    cudaMalloc(&a, 1 << 20); cudaMemset(a, 0, 1 << 20);
    cudaMallocManaged(&o, 64);
    for (int pf = 0; pf < 2; pf++) {
        for (int rep = 0; rep < 2; rep++) {
            k<<<1, 1>>>(a + rep * 4096 + pf * 65536, o, pf);
            cudaDeviceSynchronize();
            printf("prefetch=%d rep=%d: first load %lld | same-line load %lld | load after store (same line, other word) %lld | load of stored word %lld | another word %lld\n",
                   pf, rep, o[1] - o[0], o[2] - o[1], o[3] - o[2], o[4] - o[3], o[5] - o[4]);
        }
    }
    return 0;
}
