// microbench.cu — store-only fill bandwidth and 64-bit RED.MIN throughput on the B200 (design probes for k_clear and
// the visibility buffer).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mb/microbench tools/mb/microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int MODE> __global__ void __launch_bounds__(256) k_fill(uint4 *p, size_t n16, uint32_t v)
{
    const uint4 v4 = make_uint4(v, v, v, v);
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n16; i += (size_t)gridDim.x * 256) {
        if (MODE == 0) p[i] = v4;
        else if (MODE == 1) __stcs(p + i, v4);
        else if (MODE == 2) __stcg(p + i, v4);
        else if (MODE == 3) __stwt(p + i, v4);
    }
}
// 256-bit stores
__global__ void __launch_bounds__(256) k_fill256(uint4 *p, size_t n32, uint32_t v)
{
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n32; i += (size_t)gridDim.x * 256)
        asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" :: "l"(p + 2 * i), "r"(v) : "memory");
}
// CTA-contiguous fill: each CTA owns a contiguous block
__global__ void __launch_bounds__(256) k_fill_blocked(uint4 *p, size_t n16, uint32_t v)
{
    const uint4 v4 = make_uint4(v, v, v, v);
    const size_t per = (n16 + gridDim.x - 1) / gridDim.x;
    const size_t a = per * blockIdx.x, b = min(n16, a + per);
    for (size_t i = a + threadIdx.x; i < b; i += 256) __stcs(p + i, v4);
}
// RED.MIN.64: MODE 0: lanes consecutive keys (coalesced); 1: every thread walks its own row of `run` consecutive keys
template <int MODE> __global__ void __launch_bounds__(256) k_red(unsigned long long *vis, size_t n, int run, uint32_t salt)
{
    const size_t tid = (size_t)blockIdx.x * 256 + threadIdx.x, nt = (size_t)gridDim.x * 256;
    if (MODE == 0) {
        for (size_t i = tid; i < n; i += nt) atomicMin(&vis[i], ((unsigned long long)(uint32_t)(i * 2654435761u + salt) << 32) | 5u);
    } else {
        for (size_t r = tid; r * run < n; r += nt) {
            // rows scattered: row r starts at a pseudo-random run-aligned place
            const size_t rows = n / run;
            const size_t rr = (r * 2654435761ull) % rows;
            unsigned long long *p = vis + rr * run;
            for (int k = 0; k < run; k++) atomicMin(&p[k], ((unsigned long long)(uint32_t)(k * 2654435761u + salt + (uint32_t)r) << 32) | 5u);
        }
    }
}
// plain 64-bit stores with the same patterns (what the RED costs over a store)
template <int MODE> __global__ void __launch_bounds__(256) k_st64(unsigned long long *vis, size_t n, int run, uint32_t salt)
{
    const size_t tid = (size_t)blockIdx.x * 256 + threadIdx.x, nt = (size_t)gridDim.x * 256;
    if (MODE == 0) { for (size_t i = tid; i < n; i += nt) vis[i] = i + salt; }
    else for (size_t r = tid; r * run < n; r += nt) {
        const size_t rows = n / run; const size_t rr = (r * 2654435761ull) % rows;
        unsigned long long *p = vis + rr * run;
        for (int k = 0; k < run; k++) p[k] = k + salt;
    }
}

template <typename F> float time_it(F f, int reps)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < reps; i++) f();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main()
{
    const size_t bytes = 100ull << 20;          // one 4K frame's clear: ~100 MB
    uint4 *buf; CK(cudaMalloc(&buf, 4 * bytes));
    CK(cudaMemset(buf, 0, 4 * bytes));
    const size_t n16 = bytes / 16;
    // rotate over 4 x 100 MB so that the L2 (126 MB) cannot absorb the stores
    int rot = 0;
    auto p = [&]() { rot = (rot + 1) & 3; return buf + rot * n16; };
    printf("fill of %zu MB (rotating over 4 buffers), GB/s:\n", bytes >> 20);
    float ms = time_it([&] { cudaMemsetAsync(p(), 0x7F, bytes); }, 20);
    printf("  cudaMemsetAsync            %8.1f  (%.2f us)\n", bytes / ms * 1e-6, ms * 1e3);
    for (int g : { 148, 296, 592, 1184, 2368, 4736 }) {
        ms = time_it([&] { k_fill<0><<<g, 256>>>(p(), n16, 1); }, 20);  printf("  grid %5d st.default       %8.1f\n", g, bytes / ms * 1e-6);
        ms = time_it([&] { k_fill<1><<<g, 256>>>(p(), n16, 1); }, 20);  printf("  grid %5d st.cs            %8.1f\n", g, bytes / ms * 1e-6);
        ms = time_it([&] { k_fill<2><<<g, 256>>>(p(), n16, 1); }, 20);  printf("  grid %5d st.cg            %8.1f\n", g, bytes / ms * 1e-6);
        ms = time_it([&] { k_fill<3><<<g, 256>>>(p(), n16, 1); }, 20);  printf("  grid %5d st.wt            %8.1f\n", g, bytes / ms * 1e-6);
        ms = time_it([&] { k_fill256<<<g, 256>>>(p(), n16 / 2, 1); }, 20);  printf("  grid %5d st.v8 (256 bit)  %8.1f\n", g, bytes / ms * 1e-6);
        ms = time_it([&] { k_fill_blocked<<<g, 256>>>(p(), n16, 1); }, 20);  printf("  grid %5d blocked st.cs    %8.1f\n", g, bytes / ms * 1e-6);
    }
    // one-shot grid (a thread per 16 B)
    ms = time_it([&] { k_fill<0><<<(unsigned)((n16 + 255) / 256), 256>>>(p(), n16, 1); }, 20);  printf("  one thread per 16 B        %8.1f\n", bytes / ms * 1e-6);
    // L2-resident fill (same 32 MB again and again)
    ms = time_it([&] { k_fill<0><<<1184, 256>>>(buf, (32ull << 20) / 16, 1); }, 20);  printf("  32 MB, L2 resident         %8.1f\n", (32ull << 20) / ms * 1e-6);

    // ---- RED.MIN.64 ----
    unsigned long long *vis = reinterpret_cast<unsigned long long *>(buf);
    const size_t nk = (64ull << 20) / 8;        // 8.4 M keys = a 4K visibility buffer
    CK(cudaMemset(vis, 0xFF, nk * 8));
    printf("red.min.u64 over %zu keys (64 MB), G ops/s:\n", nk);
    uint32_t salt = 1;
    for (int g : { 592, 1184, 2368 }) {
        ms = time_it([&] { k_red<0><<<g, 256>>>(vis, nk, 1, salt++); }, 10);   printf("  grid %5d coalesced            %8.2f G/s\n", g, nk / ms * 1e-6);
        ms = time_it([&] { k_st64<0><<<g, 256>>>(vis, nk, 1, salt++); }, 10);  printf("  grid %5d coalesced plain st   %8.2f G/s\n", g, nk / ms * 1e-6);
        for (int run : { 4, 16, 32, 128 }) {
            ms = time_it([&] { k_red<1><<<g, 256>>>(vis, nk, run, salt++); }, 10);   printf("  grid %5d thread-row run %3d    %8.2f G/s\n", g, run, nk / ms * 1e-6);
            ms = time_it([&] { k_st64<1><<<g, 256>>>(vis, nk, run, salt++); }, 10);  printf("  grid %5d thread-row run %3d st %8.2f G/s\n", g, run, nk / ms * 1e-6);
        }
    }
    // sparse: 1.27 M REDs into the 8.4 M-key buffer (the truck frame), L2 mostly cold
    {
        const size_t nsp = 1270000;
        ms = time_it([&] { k_red<1><<<1184, 256>>>(vis, nk, 24, salt++); }, 1);
        cudaMemset(buf + 2 * n16, 0, 2 * bytes);    // flush L2
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a);
        k_red<1><<<1184, 256>>>(vis, nsp, 24, salt++);
        cudaEventRecord(b); cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms, a, b);
        printf("  1.27 M thread-row REDs (run 24), cold L2: %.2f us\n", ms * 1e3);
        cudaMemset(buf + 2 * n16, 0, 2 * bytes);
        cudaEventRecord(a);
        k_red<0><<<1184, 256>>>(vis, nsp, 1, salt++);
        cudaEventRecord(b); cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms, a, b);
        printf("  1.27 M coalesced REDs, cold L2:           %.2f us\n", ms * 1e3);
    }
    return 0;
}
