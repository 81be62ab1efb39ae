// randbw.cu — design microbenchmark (not part of the product library).
//
// Measures the throughput of DEPENDENT random reads at sector granularity on one B200: each
// thread runs `ILP` independent pointer-chasing chains; every hop loads `BYTES` (32/64/128)
// from a pseudo-random, BYTES-aligned address that depends on the value just loaded — the access
// pattern of FM-index backward search (one rank block per LF-step). Reports G loads/s and GB/s
// for working sets that fit L2 and that do not, at several occupancies.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o randbw randbw.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

template <int BYTES>
__device__ __forceinline__ uint64_t load_block(const uint8_t* p) {
    uint64_t a, b, c, d, acc = 0;
#pragma unroll
    for (int s = 0; s < BYTES / 32; ++s) {
        asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
                     : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p + 32 * s));
        acc += a ^ b ^ c ^ d;
    }
    return acc;
}

template <int BYTES, int ILP>
__global__ void chase(const uint8_t* buf, uint64_t nblocks, int hops, uint64_t* sink) {
    uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t st[ILP];
#pragma unroll
    for (int c = 0; c < ILP; ++c) st[c] = mix(tid * ILP + c + 1);
    for (int h = 0; h < hops; ++h) {
        uint64_t v[ILP];
#pragma unroll
        for (int c = 0; c < ILP; ++c) v[c] = load_block<BYTES>(buf + (st[c] % nblocks) * BYTES);
#pragma unroll
        for (int c = 0; c < ILP; ++c) st[c] = mix(st[c] + v[c]);
    }
    uint64_t acc = 0;
#pragma unroll
    for (int c = 0; c < ILP; ++c) acc ^= st[c];
    if (acc == 0x1234567) sink[0] = acc;
}

template <int BYTES, int ILP>
void run(const uint8_t* buf, uint64_t ws_bytes, int threads_per_sm, int hops, uint64_t* sink, int nsm) {
    uint64_t nblocks = ws_bytes / BYTES;
    int bs = 256;
    int grid = nsm * (threads_per_sm / bs);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    chase<BYTES, ILP><<<grid, bs>>>(buf, nblocks, hops / 4, sink);  // warm-up
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        chase<BYTES, ILP><<<grid, bs>>>(buf, nblocks, hops, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    double loads = (double)grid * bs * ILP * hops;
    printf("{\"bytes\": %d, \"ilp\": %d, \"threads_per_sm\": %d, \"ws_mb\": %.0f, \"ms\": %.3f, \"gloads_s\": %.2f, \"gb_s\": %.1f}\n",
           BYTES, ILP, threads_per_sm, ws_bytes / 1048576.0, best, loads / best / 1e6, loads * BYTES / best / 1e6);
    fflush(stdout);
}

int main(int argc, char** argv) {
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
    int nsm = p.multiProcessorCount;
    printf("# %s, %d SMs, L2 %.0f MB\n", p.name, nsm, p.l2CacheSize / 1048576.0);
    uint64_t big = (argc > 1 ? strtoull(argv[1], 0, 10) : 4096ull) << 20;
    // argv[2]: cudaLimitMaxL2FetchGranularity in bytes (32/64/128; 0 = leave the driver default)
    size_t gran = argc > 2 ? strtoull(argv[2], 0, 10) : 0, got = 0;
    if (gran) CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran));
    CK(cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity));
    printf("# cudaLimitMaxL2FetchGranularity = %zu (requested %zu)\n", got, gran);
    uint8_t* buf; CK(cudaMalloc(&buf, big));
    CK(cudaMemset(buf, 0x5a, big));
    uint64_t* sink; CK(cudaMalloc(&sink, 8));
    const int hops = 400;
    uint64_t sizes[3] = {32ull << 20, 512ull << 20, big};
    for (int si = 0; si < 3; ++si) {
        uint64_t ws = sizes[si];
        for (int tps : {1024, 2048}) {
            run<32, 1>(buf, ws, tps, hops, sink, nsm);
            run<32, 2>(buf, ws, tps, hops, sink, nsm);
            run<32, 4>(buf, ws, tps, hops, sink, nsm);
            run<64, 1>(buf, ws, tps, hops, sink, nsm);
            run<64, 2>(buf, ws, tps, hops, sink, nsm);
            run<128, 1>(buf, ws, tps, hops, sink, nsm);
            run<128, 2>(buf, ws, tps, hops, sink, nsm);
        }
    }
    return 0;
}
