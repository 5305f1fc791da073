// Which "prefetch" actually puts a line into L1 on this GPU?  Pointer chase over N distinct lines by one thread, after
// (0) nothing, (1) prefetch.global.L1, (2) cp.async.ca 4 B into shared memory, (3) prefetch.global.L2, (4) a first chase
// (real L1 hits), (5) chase again after a store to every line.  Prints cycles per dependent load.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int N = 256, STRIDE = 64;  // ints: one line = 32 ints, use every other line
__global__ void k(int* a, long long* out, int mode)
{
    __shared__ int sink[N];
    if (mode == 1) for (int i = 0; i < N; i++) asm volatile("prefetch.global.L1 [%0];" ::"l"(a + i * STRIDE));
    if (mode == 3) for (int i = 0; i < N; i++) asm volatile("prefetch.global.L2 [%0];" ::"l"(a + i * STRIDE));
    if (mode == 2) {
        for (int i = 0; i < N; i++) {
            unsigned sa = (unsigned)__cvta_generic_to_shared(&sink[i]);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(a + i * STRIDE) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    int idx = 0;
    if (mode >= 4) {
        for (int i = 0; i < N; i++) { int v; asm volatile("ld.global.ca.s32 %0, [%1];" : "=r"(v) : "l"(a + idx) : "memory"); idx = v; }
        if (mode == 5) for (int i = 0; i < N; i++) asm volatile("st.global.s32 [%0], %1;" ::"l"(a + i * STRIDE + 1), "r"(i) : "memory");
    }
    // delay
    float f = 1.f + idx;
    for (int i = 0; i < 4000; i++) f = f * 1.0001f + 0.5f;
    if (mode == 2) asm volatile("cp.async.wait_group 0;" ::: "memory");
    idx = (f < 0.f) ? 1 : 0;
    long long t0 = clock64();
    for (int i = 0; i < N; i++) { int v; asm volatile("ld.global.ca.s32 %0, [%1];" : "=r"(v) : "l"(a + idx) : "memory"); idx = v; }
    long long t1 = clock64();
    out[0] = t1 - t0 + (idx & 0);
    out[1] = idx + sink[5];
}
int main()
{
    int *a, *h = new int[N * STRIDE];
    long long* o;
    cudaMallocManaged(&o, 64);
    const char* names[] = {"cold (L2/DRAM)", "after prefetch.global.L1", "after cp.async.ca to smem", "after prefetch.global.L2", "second chase (L1 hits)", "second chase after a store to every line"};
    for (int mode = 0; mode < 6; mode++) {
        cudaMalloc(&a, N * STRIDE * 4);  // fresh memory each time: nothing cached
        for (int i = 0; i < N * STRIDE; i++) h[i] = 0;
        for (int i = 0; i < N; i++) h[i * STRIDE] = ((i + 1) % N) * STRIDE;
        cudaMemcpy(a, h, N * STRIDE * 4, cudaMemcpyHostToDevice);
        k<<<1, 1>>>(a, o, mode);
        cudaDeviceSynchronize();
        printf("%-45s %6.1f cycles per dependent load\n", names[mode], (double)o[0] / N);
    }
    return 0;
}
