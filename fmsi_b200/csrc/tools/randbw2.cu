// randbw2.cu — design microbenchmark (not part of the product library).
//
// Question: what does ONE dependent random 32-byte sector read cost in DRAM traffic on B200, and can
// the load instruction change it? Every thread chases a pseudo-random chain through a working set
// far larger than L2; the variants differ only in the load instruction / cache hints. Run under
//   ncu --metrics dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,\
//       l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,gpu__time_duration.sum
// to see DRAM bytes per requested sector.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o randbw2 randbw2.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

enum { V_NC_NOALLOC = 0, V_NC_L2_64, V_NC_L2_128, V_PLAIN, V_CG, V_CV, V_EVICT_FIRST, V_LU, V_B128x2, V_B64, V_NVARIANTS };
static const char *kNames[] = {"nc.L1::no_allocate.v4.u64", "nc.L1::no_allocate.L2::64B.v4.u64", "nc.L1::no_allocate.L2::128B.v4.u64",
                               "ld.global.v4.u64", "ld.global.cg.v4.u64", "ld.global.cv.v4.u64",
                               "nc.L2::cache_hint(evict_first).v4.u64", "ld.global.lu.v4.u64", "2 x nc.v2.u64 (16 B each)", "nc.u64 (8 B only)"};

template <int V>
__device__ __forceinline__ uint64_t load32(const uint8_t *p, uint64_t pol) {
    uint64_t a = 0, b = 0, c = 0, d = 0;
    if (V == V_NC_NOALLOC) asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    if (V == V_NC_L2_64) asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    if (V == V_NC_L2_128) asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    if (V == V_PLAIN) asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    if (V == V_CG) asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    if (V == V_CV) asm volatile("ld.global.cv.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    if (V == V_EVICT_FIRST) asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u64 {%0,%1,%2,%3}, [%4], %5;" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p), "l"(pol));
    if (V == V_LU) asm volatile("ld.global.lu.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    if (V == V_B128x2) {
        asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
        asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(c), "=l"(d) : "l"(p + 16));
    }
    if (V == V_B64) asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(a) : "l"(p));
    return a ^ b ^ c ^ d;
}

template <int V>
__global__ void chase(const uint8_t *buf, uint64_t nblocks, int hops, uint64_t *sink) {
    uint64_t pol = 0;
    if (V == V_EVICT_FIRST) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t st = mix(tid + 1);
    for (int h = 0; h < hops; ++h) st = mix(st + load32<V>(buf + (st % nblocks) * 32, pol));
    if (st == 0x1234567) sink[0] = st;
}

// One 128-byte line per group of 4 lanes, fetched by ONE warp-level instruction (4 x 32 B, same line).
__global__ void chase_line(const uint8_t *buf, uint64_t nlines, int hops, uint64_t *sink) {
    uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t st = mix((tid >> 2) + 1);
    const unsigned sub = threadIdx.x & 3u;
    for (int h = 0; h < hops; ++h) {
        uint64_t v = load32<V_NC_NOALLOC>(buf + (st % nlines) * 128 + sub * 32, 0);
        v ^= __shfl_xor_sync(0xffffffffu, v, 1);
        v ^= __shfl_xor_sync(0xffffffffu, v, 2);
        st = mix(st + v);
    }
    if (st == 0x1234567) sink[0] = st;
}

template <int V>
void run(const uint8_t *buf, uint64_t ws, int tps, int hops, uint64_t *sink, int nsm) {
    const int bs = 256, grid = nsm * (tps / bs);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    chase<V><<<grid, bs>>>(buf, ws / 32, hops / 4, sink);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    chase<V><<<grid, bs>>>(buf, ws / 32, hops, sink);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    double loads = (double)grid * bs * hops;
    printf("{\"variant\": \"%s\", \"threads_per_sm\": %d, \"ws_mb\": %.0f, \"ms\": %.3f, \"gsectors_s\": %.2f, \"useful_gb_s\": %.1f}\n",
           kNames[V], tps, ws / 1048576.0, ms, loads / ms / 1e6, loads * 32 / ms / 1e6);
    fflush(stdout);
}

int main(int argc, char **argv) {
    CK(cudaSetDevice(0));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int nsm = p.multiProcessorCount;
    uint64_t ws = (argc > 1 ? strtoull(argv[1], 0, 10) : 8192ull) << 20;
    const int tps = argc > 2 ? atoi(argv[2]) : 1024;
    printf("# %s, %d SMs, L2 %.0f MB, ws %.0f MB\n", p.name, nsm, p.l2CacheSize / 1048576.0, ws / 1048576.0);
    uint8_t *buf; CK(cudaMalloc(&buf, ws));
    CK(cudaMemset(buf, 0x5a, ws));
    uint64_t *sink; CK(cudaMalloc(&sink, 8));
    const int hops = 400;
    run<V_NC_NOALLOC>(buf, ws, tps, hops, sink, nsm);
    run<V_NC_L2_64>(buf, ws, tps, hops, sink, nsm);
    run<V_NC_L2_128>(buf, ws, tps, hops, sink, nsm);
    run<V_PLAIN>(buf, ws, tps, hops, sink, nsm);
    run<V_CG>(buf, ws, tps, hops, sink, nsm);
    run<V_CV>(buf, ws, tps, hops, sink, nsm);
    run<V_EVICT_FIRST>(buf, ws, tps, hops, sink, nsm);
    run<V_LU>(buf, ws, tps, hops, sink, nsm);
    run<V_B128x2>(buf, ws, tps, hops, sink, nsm);
    run<V_B64>(buf, ws, tps, hops, sink, nsm);
    {
        const int bs = 256, grid = nsm * (tps / bs);
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        chase_line<<<grid, bs>>>(buf, ws / 128, hops / 4, sink);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        chase_line<<<grid, bs>>>(buf, ws / 128, hops, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        double lines = (double)grid * bs / 4 * hops;
        printf("{\"variant\": \"128 B line per 4 lanes, one instruction\", \"threads_per_sm\": %d, \"ws_mb\": %.0f, \"ms\": %.3f, \"glines_s\": %.2f, \"useful_gb_s\": %.1f}\n",
               tps, ws / 1048576.0, ms, lines / ms / 1e6, lines * 128 / ms / 1e6);
    }
    return 0;
}
